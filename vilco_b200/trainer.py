"""One training iteration of the Moment-Query model, data-parallel over videos.

Mirrors the body of the reference's `train_one_epoch` loop (MQ/libs/utils/train_utils.py:318-357): zero_grad -> forward ->
final_loss.backward() -> clip_grad_norm_ -> optimizer.step() (-> post_train_step for the adapters' EMA copies).  The only
addition is what DistributedDataParallel does for the reference (train_cl.py wraps the model in DDP): every rank works on
its own videos and the gradients are averaged with ONE NCCL all-reduce over a flat fp32 gradient buffer that all
`param.grad` tensors are views of (so the collective is a single launch over NVLink, no per-tensor buckets).
"""
import torch
import torch.distributed as dist


class FlatGrads:
    """All trainable parameters' .grad as views into one contiguous fp32 buffer."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view(p.shape)
            o += p.numel()

    def zero(self):
        self.flat.zero_()

    def attached(self):
        """False when something (e.g. optimizer.zero_grad(set_to_none=True)) replaced the views."""
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.params)


def broadcast_parameters(model, src=0):
    """Same initial weights on every rank (what DDP's constructor does)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        with torch.no_grad():
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t.data, src)


class Trainer:
    """One iteration of train_one_epoch, data-parallel.  With a FlatAdamW optimizer and world_size > 1 the gradient
    all-reduce is bucketed and overlapped with the backward pass: the flat buffer is laid out in (approximately) reverse
    execution order, the first step records after which backward node each bucket is complete, and from the second step on
    every finished bucket is all-reduced asynchronously (NCCL stream) while the remaining backward kernels run."""
    BUCKET = 32 * 1024 * 1024      # elements (128 MB of fp32) per all-reduce bucket

    def __init__(self, model, optimizer, clip_grad_l2norm=-1.0, scheduler=None, overlap=True, grad_comm=None):
        """grad_comm: "bf16" (default; env VILCO_GRAD_COMM) sends every gradient bucket over NVLink as bf16 — half the bytes of
        the exchange, 0.56 GB per step for the MQ model — and accumulates the reduced values back into the fp32 master
        gradient buffer; "fp32" all-reduces the fp32 buffer in place (bit-identical replicas either way: every rank receives
        the same reduced values)."""
        import os
        self.grad_comm = grad_comm or os.environ.get("VILCO_GRAD_COMM", "bf16")
        assert self.grad_comm in ("bf16", "fp32")
        self._copy_stream = None
        self._staged = {}
        self.model, self.optimizer, self.scheduler = model, optimizer, scheduler
        self.clip = float(clip_grad_l2norm)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.flat = isinstance(optimizer, FlatAdamW)
        self.overlap = bool(overlap) and self.flat and self.world > 1
        self.plan = None           # (n_nodes, {node index: [(a, b), ...]}, [(a, b) launched after the last node])
        self._plan_nodes = None
        self._work = []
        if self.flat:
            model.use_flat_optimizer(optimizer)
        else:
            self.grads = FlatGrads(model.parameters())
            if hasattr(model, "_train_forward"):
                model._direct_grads = True     # the backward kernels accumulate straight into the flat .grad views

    # ---- bucketed all-reduce -------------------------------------------------------------------------------------
    def _readiness(self):
        """{id(param): node index after which its gradient is final | None (only at the very end) | -1 (dead: never written)}
        from the record of the last backward pass."""
        n_nodes, touch, fallback = self.model._last_touch
        opt = self.optimizer
        names = {id(p): k for k, p in self.model.named_parameters()}
        out = {}
        for g in opt.param_groups:
            for p in g["params"]:
                if id(p) in opt.slots:
                    o, k = opt.slots[id(p)]
                    key = names.get(id(p))
                    if key in fallback:
                        out[id(p)] = None
                    elif key in touch:
                        out[id(p)] = touch[key]
                    elif bool((opt.flat_grad[o:o + k] != 0).any()):
                        out[id(p)] = None              # written by torch autograd of the glue (mu / sigma, prompts)
                    else:
                        out[id(p)] = -1
        return out

    def _build_plan(self, ready):
        """Bucket [a, b) of the live part of the flat gradient buffer is final after node max(last touch of the parameters
        inside it); parameters that took the fallback path (permuted layouts, glue) are final at the end."""
        opt = self.optimizer
        spans = sorted((opt.slots[i][0], opt.slots[i][0] + opt.slots[i][1], r) for i, r in ready.items())
        per_node, tail = {}, []
        for lo_, hi_ in opt.live_ranges():
            a = lo_
            while a < hi_:
                b = min(hi_, a + self.BUCKET)
                rs = [r for (o, e, r) in spans if o < b and e > a]
                if rs and all(r is not None for r in rs):
                    per_node.setdefault(max(max(rs), 0), []).append((a, b))
                else:
                    tail.append((a, b))
                a = b
        return self.model._last_touch[0], per_node, tail

    def _named(self):
        tab = getattr(self.model, "_param_table", None)
        if tab is not None:
            names, plist, _ = tab()
            return zip(names, plist)
        return self.model.named_parameters()

    def _launch(self, a, b):
        g = self._gbuf[a:b]
        if self.grad_comm == "bf16":
            buf = g.to(torch.bfloat16)                       # cast on the compute stream, behind the kernels that wrote g
            self._work.append((dist.all_reduce(buf, async_op=True), g, buf))
        else:
            self._work.append((dist.all_reduce(g, async_op=True), None, None))

    def _finish_reduce(self):
        for w, g, buf in self._work:
            w.wait()                                          # orders the compute stream behind the collective
            if buf is not None:
                g.copy_(buf)                                  # fp32 master gradient <- reduced bf16 values
        self._work = []

    # ---- input prefetch -----------------------------------------------------------------------------------------------
    def stage(self, video_list):
        """Start the host -> device copy of a batch (pinned `feats`) on a side stream and return the batch with device-resident
        features; `step(staged)` waits for the copy.  Call it for batch i+1 before `step(batch i)`: the upload then overlaps
        the previous step instead of preceding the forward pass.  Two persistent sets of device buffers are used in turn (no
        allocation per step); a set is overwritten only after the step that read it has been issued and its kernels are done."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
            self._slots = [{"bufs": {}, "free": None} for _ in range(2)]
            self._slot_i = 0
        slot = self._slots[self._slot_i]
        self._slot_i ^= 1
        ev = torch.cuda.Event()
        out = []
        with torch.cuda.stream(self._copy_stream):
            if slot["free"] is not None:
                self._copy_stream.wait_event(slot["free"])
            for i, v in enumerate(video_list):
                f = v["feats"]
                if not f.is_cuda:
                    key = (i, tuple(f.shape))
                    buf = slot["bufs"].get(key)
                    if buf is None:
                        buf = slot["bufs"][key] = torch.empty(f.shape, dtype=f.dtype, device="cuda")
                    buf.copy_(f if f.is_pinned() else f.pin_memory(), non_blocking=True)
                    f = buf
                out.append({**v, "feats": f})
            ev.record(self._copy_stream)
        self._staged[id(out)] = (out, ev, slot)
        return out

    def _after_node(self, i, n_nodes):
        if n_nodes != self.plan[0]:      # the tape changed shape (e.g. model.train() <-> eval()): fall back for this step
            return
        for a, b in self.plan[1].get(i, ()):
            self._launch(a, b)

    def step(self, video_list, task_id=0, prev_out_cls_logits=None):
        """video_list = this rank's share of the global batch.  Returns the loss dict of this rank (tensors)."""
        if self.flat:
            self.optimizer.zero_grad()
            flat = self.optimizer.flat_grad
        else:
            if not self.grads.attached():
                self.grads = FlatGrads(self.model.parameters())
            self.grads.zero()
            flat = self.grads.flat
        st = self._staged.pop(id(video_list), None)
        used_slot = None
        if st is not None and st[0] is video_list:
            torch.cuda.current_stream().wait_event(st[1])     # the prefetched upload of this batch
            used_slot = st[2]
        self._gbuf = flat
        use_plan = self.overlap and self.plan is not None
        self.model._after_backward_node = self._after_node if use_plan else None
        self._work = []
        losses = self.model(video_list, task_id=task_id, prev_out_cls_logits=prev_out_cls_logits or [])
        final = losses["final_loss"]
        if self.world > 1:
            # d(mean over ranks) : scale this rank's gradient by 1 / world, the all-reduce then only sums
            final.backward(torch.full_like(final, 1.0 / self.world))
        else:
            final.backward()
        touched_dead = self.flat and self.optimizer.dead and any(
            id(p) in self.optimizer.dead for k, p in self._named() if k in self.model._last_touch[1])
        if self.world > 1:
            if use_plan and self.model._last_touch[0] == self.plan[0]:
                for a, b in self.plan[2]:
                    self._launch(a, b)
                if touched_dead:         # a parameter parked as dead received a gradient: reduce the parked regions too
                    for _, le, e, _ in self.optimizer.segments:
                        if e > le:
                            self._launch(le, e)
                self._finish_reduce()
            else:
                assert not self._work
                self._launch(0, flat.numel())
                self._finish_reduce()
            if getattr(self, "keep_grad", False):
                torch.cuda.synchronize()
                self.last_grad = flat.clone()
        if self.flat and hasattr(self.model, "_last_touch") and \
                (self._plan_nodes != self.model._last_touch[0] or touched_dead):
            # first step (or the tape changed): find the parameters the backward never writes, move them out of the updated /
            # all-reduced region, and plan the buckets over the new layout
            ready = self._readiness()
            self.optimizer.compact({i for i, r in ready.items() if r == -1})
            self._plan_nodes = self.model._last_touch[0]
            self.plan = self._build_plan(ready) if self.overlap else None
        self.model._after_backward_node = None
        if self.flat:
            self.optimizer.step(clip_grad_l2norm=self.clip)
        else:
            if self.clip > 0.0:
                torch.nn.utils.clip_grad_norm_(self.grads.params, self.clip)
            self.optimizer.step()
        if self.scheduler is not None:
            self.scheduler.step()
        if getattr(self.model, "use_adapt", False):
            self.model.post_train_step()
        if used_slot is not None:                             # the staging buffers of this batch may be overwritten after this point
            used_slot["free"] = torch.cuda.Event()
            used_slot["free"].record(torch.cuda.current_stream())
        return losses


def make_optimizer(model, optimizer_config, flat=False):
    """Parameter grouping of the reference's make_optimizer (train_utils.py:68-143): biases, LayerNorm weights, Scale /
    AffineDropPath scales and XLNet norms are not decayed, everything else is.  flat=True returns the FlatAdamW below
    (same update rule, hand-written kernels) instead of torch.optim.AdamW."""
    from .modeling.blocks import AffineDropPath, LayerNorm, MaskedConv1D, Scale
    decay, no_decay = set(), set()
    white = (torch.nn.Linear, torch.nn.Conv1d, MaskedConv1D)
    black = (LayerNorm, torch.nn.GroupNorm)
    for mn, m in model.named_modules():
        for pn, _ in m.named_parameters():
            fpn = f"{mn}.{pn}" if mn else pn
            if pn.endswith("bias"):
                no_decay.add(fpn)
            elif "xlnet" in pn and "norm" not in pn:
                decay.add(fpn)
            elif "xlnet" in pn and "norm" in pn:
                no_decay.add(fpn)
            elif pn.endswith("weight") and isinstance(m, white):
                decay.add(fpn)
            elif pn.endswith("weight") and isinstance(m, black):
                no_decay.add(fpn)
            elif pn.endswith("scale") and isinstance(m, (Scale, AffineDropPath)):
                no_decay.add(fpn)
            elif pn.endswith("rel_pe"):
                no_decay.add(fpn)
    pd = dict(model.named_parameters())
    decay, no_decay = decay & pd.keys(), (no_decay & pd.keys()) - decay
    remain = pd.keys() - (decay | no_decay)
    wd = optimizer_config["weight_decay"]
    # torch path: alphabetical like the reference.  flat path: reverse definition order ~ the order in which the backward
    # pass completes the gradients, so that the trainer's all-reduce buckets become ready early (the update rule does not
    # depend on the order)
    pos = {n: i for i, n in enumerate(pd)}
    from .engine import _pack_kind
    late = lambda n: _pack_kind(n, pd[n]) in ("conv3", "dw", "xl_t") or pd[n].dim() == 0 or pd[n].dim() == 2 and pd[n].shape[1] == 1  # noqa: E731
    # (permuted layouts, Scale scalars and mu / sigma get their gradients at the very end of the backward pass: keep them
    # together behind the parameters whose gradients the kernels write in place)
    # the GROUPS list their parameters alphabetically in both paths — optimizer.state_dict() indexes the state by position in
    # the groups, so checkpoints of FlatAdamW and torch.optim.AdamW are interchangeable; the flat path only uses the
    # backward-completion order for where a parameter LIVES in the flat buffers (`layout_order`)
    groups = [{"params": [pd[n] for n in sorted(decay)], "weight_decay": wd},
              {"params": [pd[n] for n in sorted(no_decay)], "weight_decay": 0.0},
              {"params": [pd[n] for n in sorted(remain)], "weight_decay": wd}]
    groups = [g for g in groups if g["params"]]
    layout_order = {id(pd[n]): (late(n), -pos[n]) for n in pd}
    if optimizer_config["type"] == "SGD":
        return torch.optim.SGD(groups, lr=optimizer_config["learning_rate"], momentum=optimizer_config["momentum"])
    if optimizer_config["type"] == "AdamW":
        if flat:
            return FlatAdamW(groups, lr=optimizer_config["learning_rate"], layout_order=layout_order)
        return torch.optim.AdamW(groups, lr=optimizer_config["learning_rate"])
    raise TypeError("Unsupported optimizer!")


class FlatAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics on flat buffers with hand-written kernels (csrc/optim.cu).

    All trainable parameters, their gradients and both moments live in four contiguous fp32 buffers (parameter groups are
    contiguous segments, every tensor starts at a multiple of 8 elements); `param.data` / `param.grad` are views, so the
    model, state_dict and the gradient all-reduce keep working unchanged.  One step = global-norm kernel + clip coefficient
    + one fused AdamW launch per parameter group, which also rewrites the bf16 (hi, lo) operand planes the GEMM kernels
    read (`self.planes`), so no per-tensor weight re-packing happens between iterations.
    """

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, layout_order=None):
        """layout_order: id(param) -> sort key of its position inside its group's segment of the flat buffers (default: the
        order of the group's list).  It never affects state_dict(): the state is indexed by the position in the groups."""
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._layout_order = layout_order or {}
        dev = next(p for g in self.param_groups for p in g["params"]).device
        assert dev.type == "cuda", "FlatAdamW runs on the CUDA kernels only"
        self.device = dev
        self.slots = None        # id(param) -> (offset, numel)
        self.segments = []       # (start, live_end, end, group): [start, live_end) is updated, [live_end, end) holds dead parameters
        self.dead = set()
        self.layout_version = 0
        self._scal = torch.zeros(3, device=dev)   # (unused), clip coefficient, gradient norm
        self._partials = torch.zeros(1184, device=dev)   # VILCO_CLIP_SCRATCH per-block partial sums of squares
        self.t = 0
        self.epoch = 0           # bumped on every update; the model re-derives its permuted weight copies when it changes
        self.planes = None
        self._planes_precision = None
        self._layout(set())

    def _layout(self, dead):
        """(Re)build the flat buffers: per parameter group the live parameters first, then the dead ones (parameters the
        backward pass never writes — torch.optim.AdamW skips those because their .grad stays None; here they sit behind
        `live_end` and are neither updated, decayed nor all-reduced).  Existing values / moments are carried over."""
        old_slots = self.slots
        old = (self.flat_p, self.flat_grad, self.exp_avg, self.exp_avg_sq) if old_slots is not None else None
        seen, n, slots, segments = set(), 0, {}, []
        for g in self.param_groups:
            start = n
            ps = []
            for p in g["params"]:
                if p.requires_grad and id(p) not in seen:
                    seen.add(id(p))
                    ps.append(p)
            if self._layout_order:
                rank = {id(p): i for i, p in enumerate(ps)}
                ps.sort(key=lambda p: (self._layout_order.get(id(p), (True, 0)), rank[id(p)]))
            for p in ps:
                if id(p) not in dead:
                    slots[id(p)] = (n, p.numel())
                    n = (n + p.numel() + 7) // 8 * 8
            live_end = n
            for p in ps:
                if id(p) in dead:
                    slots[id(p)] = (n, p.numel())
                    n = (n + p.numel() + 7) // 8 * 8
            segments.append((start, live_end, n, g))
        new = [torch.zeros(n, device=self.device) for _ in range(4)]
        with torch.no_grad():
            for g in self.param_groups:
                for p in g["params"]:
                    if id(p) not in slots:
                        continue
                    o, k = slots[id(p)]
                    if old is None:
                        new[0][o:o + k].copy_(p.data.reshape(-1))
                    else:
                        oo, _ = old_slots[id(p)]
                        for dst, src in zip(new, old):
                            dst[o:o + k].copy_(src[oo:oo + k])
                    p.data = new[0][o:o + k].view(p.shape)
                    p.grad = new[1][o:o + k].view(p.shape)
        self.flat_p, self.flat_grad, self.exp_avg, self.exp_avg_sq = new
        self.n, self.slots, self.segments, self.dead = n, slots, segments, set(dead)
        self.layout_version += 1
        self.refresh_planes()

    def compact(self, dead_ids):
        """Move the given parameters (ids) behind the live region of their group."""
        if set(dead_ids) != self.dead:
            self._layout(set(dead_ids))

    def live_ranges(self):
        return [(a, le) for a, le, _, _ in self.segments if le > a]

    def refresh_planes(self):
        """(PLANES, n) bf16 hi / lo copies of every parameter (needed after load_state_dict or a precision switch; the
        optimizer step keeps them current by itself)."""
        from . import ops
        with torch.no_grad():
            # two planes whenever any contraction reads split weights (ops.PLANES_HI); single-plane consumers use plane 0
            self.planes = ops.split16(self.flat_p, planes=ops.PLANES_HI)
        self._planes_precision = ops.precision()

    def plane_view(self, p, shape):
        o, k = self.slots[id(p)]
        return self.planes[:, o:o + k].view(self.planes.shape[0], *shape)

    def zero_grad(self, set_to_none=False):
        self.flat_grad.zero_()
        for g in self.param_groups:            # re-attach views something may have replaced
            for p in g["params"]:
                if id(p) in self.slots and (p.grad is None or p.grad.untyped_storage().data_ptr() != self.flat_grad.untyped_storage().data_ptr()):
                    o, k = self.slots[id(p)]
                    p.grad = self.flat_grad[o:o + k].view(p.shape)

    @torch.no_grad()
    def step(self, closure=None, clip_grad_l2norm=-1.0):
        import ctypes as C
        from . import lib as L
        from . import ops
        if ops.precision() != self._planes_precision:
            self.refresh_planes()
        self.t += 1
        st = L.stream_ptr()
        s = self._scal
        L.check(L.lib().vilco_grad_clip_coef(ops._p(self.flat_grad), ops._i64(self.n), C.c_float(float(clip_grad_l2norm)),
                                             C.c_void_p(self._partials.data_ptr()), C.c_void_p(s.data_ptr() + 4),
                                             C.c_void_p(s.data_ptr() + 8), st),
                "vilco_grad_clip_coef")
        NP = self.planes.shape[0]
        for a, b, _, g in self.segments:
            if b <= a:
                continue
            L.check(L.lib().vilco_adamw(
                C.c_void_p(self.flat_p.data_ptr() + 4 * a), C.c_void_p(self.flat_grad.data_ptr() + 4 * a),
                C.c_void_p(self.exp_avg.data_ptr() + 4 * a), C.c_void_p(self.exp_avg_sq.data_ptr() + 4 * a), ops._i64(b - a),
                C.c_float(g["lr"]), C.c_float(g["betas"][0]), C.c_float(g["betas"][1]), C.c_float(g["eps"]),
                C.c_float(g["weight_decay"]), int(self.t), C.c_void_p(s.data_ptr() + 4),
                C.c_void_p(self.planes.data_ptr() + 2 * a), ops._i64(self.n if NP == 2 else 0), st), "vilco_adamw")
        self.epoch += 1

    # ---- checkpointing in torch.optim.AdamW's own format (train_utils.save_checkpoint stores optimizer.state_dict()) ------
    def state_dict(self):
        """Same structure as torch.optim.AdamW.state_dict(): per-parameter 'step', 'exp_avg', 'exp_avg_sq' (clones of the
        flat slices) indexed by the position of the parameter in the groups, so a checkpoint written here resumes with
        torch.optim.AdamW and vice versa.  Parked (dead) parameters have no state, like parameters without a gradient."""
        state, groups, idx = {}, [], 0
        for g in self.param_groups:
            ids = []
            for p in g["params"]:
                if id(p) in self.slots and id(p) not in self.dead and self.t > 0:
                    o, k = self.slots[id(p)]
                    state[idx] = {"step": torch.tensor(float(self.t)), "exp_avg": self.exp_avg[o:o + k].view(p.shape).clone(),
                                  "exp_avg_sq": self.exp_avg_sq[o:o + k].view(p.shape).clone()}
                ids.append(idx)
                idx += 1
            groups.append({**{k_: v for k_, v in g.items() if k_ != "params"}, "params": ids})
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        idx = 0
        steps = []
        with torch.no_grad():
            for g, sg in zip(self.param_groups, sd["param_groups"]):
                for k_, v in sg.items():
                    if k_ != "params":
                        g[k_] = v
                for p in g["params"]:
                    st = sd["state"].get(idx)
                    if st is not None and id(p) in self.slots:
                        o, k = self.slots[id(p)]
                        self.exp_avg[o:o + k].copy_(st["exp_avg"].reshape(-1))
                        self.exp_avg_sq[o:o + k].copy_(st["exp_avg_sq"].reshape(-1))
                        steps.append(int(float(st["step"])))
                    idx += 1
        self.t = max(steps) if steps else 0

    def grad_norm(self):
        """global L2 norm of the last step's (un-clipped) gradient — what clip_grad_norm_ returns."""
        return self._scal[2]
