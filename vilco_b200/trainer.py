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
    def __init__(self, model, optimizer, clip_grad_l2norm=-1.0, scheduler=None):
        self.model, self.optimizer, self.scheduler = model, optimizer, scheduler
        self.clip = float(clip_grad_l2norm)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.grads = FlatGrads(model.parameters())

    def step(self, video_list, task_id=0, prev_out_cls_logits=None):
        """video_list = this rank's share of the global batch.  Returns the loss dict of this rank (tensors)."""
        if not self.grads.attached():
            self.grads = FlatGrads(self.model.parameters())
        self.grads.zero()
        losses = self.model(video_list, task_id=task_id, prev_out_cls_logits=prev_out_cls_logits or [])
        losses["final_loss"].backward()
        if self.world > 1:
            dist.all_reduce(self.grads.flat)
            self.grads.flat.div_(self.world)
        if self.clip > 0.0:
            torch.nn.utils.clip_grad_norm_(self.grads.params, self.clip)
        self.optimizer.step()
        if self.scheduler is not None:
            self.scheduler.step()
        if getattr(self.model, "use_adapt", False):
            self.model.post_train_step()
        return losses


def make_optimizer(model, optimizer_config):
    """Parameter grouping of the reference's make_optimizer (train_utils.py:68-143): biases, LayerNorm weights, Scale /
    AffineDropPath scales and XLNet norms are not decayed, everything else is."""
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
    groups = [{"params": [pd[n] for n in sorted(decay)], "weight_decay": wd},
              {"params": [pd[n] for n in sorted(no_decay)], "weight_decay": 0.0},
              {"params": [pd[n] for n in sorted(remain)], "weight_decay": wd}]
    groups = [g for g in groups if g["params"]]
    if optimizer_config["type"] == "SGD":
        return torch.optim.SGD(groups, lr=optimizer_config["learning_rate"], momentum=optimizer_config["momentum"])
    if optimizer_config["type"] == "AdamW":
        return torch.optim.AdamW(groups, lr=optimizer_config["learning_rate"])
    raise TypeError("Unsupported optimizer!")
