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

    def __init__(self, model, optimizer, clip_grad_l2norm=-1.0, scheduler=None, overlap=True):
        self.model, self.optimizer, self.scheduler = model, optimizer, scheduler
        self.clip = float(clip_grad_l2norm)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.flat = isinstance(optimizer, FlatAdamW)
        self.overlap = bool(overlap) and self.flat and self.world > 1
        self.plan = None           # (n_nodes, {node index: [(a, b), ...]}, [(a, b) launched after the last node])
        self._work = []
        if self.flat:
            model.use_flat_optimizer(optimizer)
        else:
            self.grads = FlatGrads(model.parameters())

    # ---- bucketed all-reduce -------------------------------------------------------------------------------------
    def _build_plan(self):
        """From the last backward's record: bucket [a, b) of the flat gradient buffer is final after node max(last touch of
        the parameters inside it); parameters that took the fallback path (permuted layouts, glue) are final at the end."""
        last = getattr(self.model, "_last_touch", None)
        if last is None:
            return None
        n_nodes, touch, fallback = last
        opt = self.optimizer
        names = {id(p): k for k, p in self.model.named_parameters()}
        ready = []                                     # (offset, end, node index; None = only final after the whole backward)
        for g in opt.param_groups:
            for p in g["params"]:
                if id(p) in opt.slots:
                    o, k = opt.slots[id(p)]
                    key = names.get(id(p))
                    if key in fallback:
                        r = None
                    elif key in touch:
                        r = touch[key]
                    elif bool((opt.flat_grad[o:o + k] != 0).any()):
                        r = None                       # written by torch autograd of the glue (mu / sigma, prompts)
                    else:
                        r = -1                         # dead parameter: its gradient is the zero it was reset to
                    ready.append((o, o + k, r))
        ready.sort()
        per_node, tail = {}, []
        a = 0
        while a < opt.n:
            b = min(opt.n, a + self.BUCKET)
            rs = [r for (o, e, r) in ready if o < b and e > a]
            if rs and all(r is not None for r in rs):
                per_node.setdefault(max(max(rs), 0), []).append((a, b))
            else:
                tail.append((a, b))
            a = b
        return n_nodes, per_node, tail

    def _launch(self, a, b):
        self._work.append(dist.all_reduce(self.optimizer.flat_grad[a:b], async_op=True))

    def _after_node(self, i):
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
        use_plan = self.overlap and self.plan is not None
        self.model._after_backward_node = self._after_node if use_plan else None
        self._work = []
        losses = self.model(video_list, task_id=task_id, prev_out_cls_logits=prev_out_cls_logits or [])
        final = losses["final_loss"]
        if self.world > 1:
            # d(mean over ranks) : scale this rank's gradient by 1 / world, the all-reduce then only sums
            final.backward(torch.full_like(final, 1.0 / self.world))
            if use_plan and getattr(self.model, "_last_touch", (None,))[0] == self.plan[0]:
                for a, b in self.plan[2]:
                    self._launch(a, b)
                for w in self._work:
                    w.wait()
            else:
                if self._work:     # the tape changed shape under an old plan: finish what was launched, redo everything
                    for w in self._work:
                        w.wait()
                    raise RuntimeError("backward structure changed while a bucket plan was active; recreate the Trainer")
                dist.all_reduce(flat)
            if getattr(self, "keep_grad", False):
                torch.cuda.synchronize()
                self.last_grad = flat.clone()
            if self.overlap and (self.plan is None or self.plan[0] != self.model._last_touch[0]):
                self.plan = self._build_plan()
        else:
            final.backward()
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
    order = (lambda names: sorted(names, key=lambda n: (late(n), -pos[n]))) if flat else sorted
    groups = [{"params": [pd[n] for n in order(decay)], "weight_decay": wd},
              {"params": [pd[n] for n in order(no_decay)], "weight_decay": 0.0},
              {"params": [pd[n] for n in order(remain)], "weight_decay": wd}]
    groups = [g for g in groups if g["params"]]
    if optimizer_config["type"] == "SGD":
        return torch.optim.SGD(groups, lr=optimizer_config["learning_rate"], momentum=optimizer_config["momentum"])
    if optimizer_config["type"] == "AdamW":
        if flat:
            return FlatAdamW(groups, lr=optimizer_config["learning_rate"])
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

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        from . import ops
        seen, n = set(), 0
        self.slots = {}          # id(param) -> (offset, numel)
        self.segments = []       # (start, end, group)
        for g in self.param_groups:
            start = n
            for p in g["params"]:
                if not p.requires_grad or id(p) in seen:
                    continue
                seen.add(id(p))
                self.slots[id(p)] = (n, p.numel())
                n = (n + p.numel() + 7) // 8 * 8
            self.segments.append((start, n, g))
        dev = next(p for g in self.param_groups for p in g["params"]).device
        assert dev.type == "cuda", "FlatAdamW runs on the CUDA kernels only"
        self.n = n
        self.flat_p = torch.zeros(n, device=dev)
        self.flat_grad = torch.zeros(n, device=dev)
        self.exp_avg = torch.zeros(n, device=dev)
        self.exp_avg_sq = torch.zeros(n, device=dev)
        self._scal = torch.zeros(3, device=dev)   # (unused), clip coefficient, gradient norm
        self._partials = torch.zeros(1184, device=dev)   # VILCO_CLIP_SCRATCH per-block partial sums of squares
        self.t = 0
        self.epoch = 0           # bumped on every update; the model re-derives its permuted weight copies when it changes
        for g in self.param_groups:
            for p in g["params"]:
                if id(p) in self.slots:
                    o, k = self.slots[id(p)]
                    self.flat_p[o:o + k].copy_(p.data.reshape(-1))
                    p.data = self.flat_p[o:o + k].view(p.shape)
                    p.grad = self.flat_grad[o:o + k].view(p.shape)
        self.planes = None
        self._planes_precision = None
        self.refresh_planes()

    def refresh_planes(self):
        """(PLANES, n) bf16 hi / lo copies of every parameter (needed after load_state_dict or a precision switch; the
        optimizer step keeps them current by itself)."""
        from . import ops
        with torch.no_grad():
            hi = self.flat_p.to(torch.bfloat16)
            self.planes = torch.stack([hi, (self.flat_p - hi.float()).to(torch.bfloat16)]) if ops.PLANES == 2 else hi.unsqueeze(0).clone()
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
        for a, b, g in self.segments:
            if b <= a:
                continue
            L.check(L.lib().vilco_adamw(
                C.c_void_p(self.flat_p.data_ptr() + 4 * a), C.c_void_p(self.flat_grad.data_ptr() + 4 * a),
                C.c_void_p(self.exp_avg.data_ptr() + 4 * a), C.c_void_p(self.exp_avg_sq.data_ptr() + 4 * a), ops._i64(b - a),
                C.c_float(g["lr"]), C.c_float(g["betas"][0]), C.c_float(g["betas"][1]), C.c_float(g["eps"]),
                C.c_float(g["weight_decay"]), int(self.t), C.c_void_p(s.data_ptr() + 4),
                C.c_void_p(self.planes.data_ptr() + 2 * a), ops._i64(self.n if NP == 2 else 0), st), "vilco_adamw")
        self.epoch += 1

    def grad_norm(self):
        """global L2 norm of the last step's (un-clipped) gradient — what clip_grad_norm_ returns."""
        return self._scal[2]
