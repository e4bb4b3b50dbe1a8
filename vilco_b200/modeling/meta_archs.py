"""`LocPointTransformer` meta-architecture behind the reference's surface
(MQ/libs/modeling/meta_archs.py:351-1822): same constructor kwargs, same state_dict names, same
`forward(video_list, task_id, ensemble, hidden_state, is_training, prev_out_cls_logits, get_emb, val_qilDatasetList)`
contract and the attributes the unchanged CL orchestration touches — computed by the sm_100a kernels
of libvilco_b200.so.  No PyTorch compute fallback exists: without the CUDA library / a GPU the forward raises.
"""
import ctypes as C
import math
import os

import torch
from torch import nn
from torch.nn import functional as F

from .. import engine as E
from .. import lib as L
from .. import ops
from ..utils.nms import _run as _nms_run
from .blocks import LayerNorm, MaskedConv1D, Scale
from .models import make_backbone, make_generator, make_neck, register_meta_arch


NOSYNC_STEP = os.environ.get("VILCO_NOSYNC_STEP", "1") == "1"   # 0: read the loss sums on the host between forward and backward

class BiasLayer(nn.Module):
    """BiC bias-correction layer (reference: meta_archs.py:26-36)."""

    def __init__(self):
        super().__init__()
        self.alpha = nn.Parameter(torch.ones(1, requires_grad=True))
        self.beta = nn.Parameter(torch.zeros(1, requires_grad=True))

    def forward(self, x):
        return self.alpha * x + self.beta

    def printParam(self, i):
        print(i, self.alpha.item(), self.beta.item())


class Adapter(nn.Module):
    """Temporal adapter ("pet"): Linear over the TIME axis T -> 5T -> T/2 with GELU, run in parallel to the attention of a
    branch block (reference: meta_archs.py:105-148).  Parameter container; the math runs in engine.adapter_fwd."""

    def __init__(self, embed_dim, down_sample=5, mode="parallel", scale=None, act_layer=nn.GELU, stride=1):
        super().__init__()
        assert mode == "parallel"
        hidden_dim = int(embed_dim * down_sample)
        self.layer = nn.Sequential(nn.Linear(embed_dim, hidden_dim), act_layer(), nn.Linear(hidden_dim, embed_dim // 2))
        self.mode = mode
        nn.init.kaiming_uniform_(self.layer[0].weight, a=math.sqrt(5))
        nn.init.zeros_(self.layer[0].bias)
        nn.init.zeros_(self.layer[2].weight)
        nn.init.zeros_(self.layer[2].bias)


class ModelEmaV2(nn.Module):
    """EMA copy of a module tree over its state_dict (timm.utils.model_ema.ModelEmaV2 semantics, used for the adapter
    EMA of mq_vilco: meta_archs.py:664-712)."""

    def __init__(self, model, decay=0.9999, device=None):
        super().__init__()
        import copy
        self.module = copy.deepcopy(model)
        self.module.eval()
        self.decay = decay

    @torch.no_grad()
    def update(self, model):
        for e, m in zip(self.module.state_dict().values(), model.state_dict().values()):
            e.copy_(self.decay * e + (1.0 - self.decay) * m)


class Prompt(nn.Module):
    """L2P prompt pool (reference: MQ/libs/cl_methods/prompt.py:4-137, the configuration mq_vilco uses: embedding_key
    'mean', uniform init, learnable keys, batchwise_prompt).  Selection is a handful of tiny host-driven torch ops on
    (B, 768) tensors — CL-method glue, not a hot kernel."""

    def __init__(self, length=5, embed_dim=768, pool_size=None, top_k=None, batchwise_prompt=True):
        super().__init__()
        self.length, self.embed_dim, self.pool_size, self.top_k = length, embed_dim, pool_size, top_k
        self.batchwise_prompt = batchwise_prompt
        self.prompt = nn.Parameter(torch.empty(pool_size, length, embed_dim).uniform_(-1, 1))
        self.prompt_key = nn.Parameter(torch.empty(pool_size, embed_dim).uniform_(-1, 1))

    @staticmethod
    def l2_normalize(x, dim=None, epsilon=1e-12):
        sq = torch.sum(x ** 2, dim=dim, keepdim=True)
        return x * torch.rsqrt(torch.maximum(sq, torch.tensor(epsilon, device=x.device)))

    def forward(self, x_embed, prompt_mask=None, cls_features=None):
        """x_embed (B, L, C) -> dict with 'prompted_embedding' (B, top_k*length + L, C), 'reduce_sim', ..."""
        out = {}
        dev = x_embed.device
        # the selection runs on the host in fp32: it is a (B, pool) problem whose top-k ties (every id appears once for a
        # single clip) are broken by the sort implementation — the CPU order is the one the reference's CPU path and the
        # golden vectors use
        x_mean = torch.mean(x_embed.detach().float().cpu(), dim=1)
        prompt_norm = self.l2_normalize(self.prompt_key.detach().float().cpu(), dim=1)
        x_norm = self.l2_normalize(x_mean, dim=1)
        similarity = torch.matmul(x_norm, prompt_norm.t())
        if prompt_mask is None:
            _, idx = torch.topk(similarity, k=self.top_k, dim=1)
            if self.batchwise_prompt:
                prompt_id, id_counts = torch.unique(idx, return_counts=True, sorted=True)
                if prompt_id.shape[0] < self.pool_size:
                    pad = self.pool_size - prompt_id.shape[0]
                    prompt_id = torch.cat([prompt_id, torch.full((pad,), int(torch.min(idx.flatten())))])
                    id_counts = torch.cat([id_counts, torch.full((pad,), 0)])
                _, major_idx = torch.topk(id_counts, k=self.top_k)
                idx = prompt_id[major_idx].expand(x_embed.shape[0], -1)
        else:
            idx = prompt_mask.cpu()
        similarity, idx = similarity.to(dev), idx.to(dev)
        batched_prompt = self.prompt[idx].reshape(x_embed.shape[0], -1, x_embed.shape[2])
        out["prompt_idx"], out["similarity"] = idx, similarity
        # the pull-constraint term is differentiable w.r.t. the prompt keys: recomputed on the device with autograd
        pn = self.l2_normalize(self.prompt_key, dim=1)
        xn = self.l2_normalize(torch.mean(x_embed, dim=1), dim=1)
        out["reduce_sim"] = torch.sum(pn[idx] * xn.unsqueeze(1)) / x_embed.shape[0]
        out["total_prompt_len"] = batched_prompt.shape[1]
        out["prompted_embedding"] = torch.cat([batched_prompt, x_embed], dim=1)
        return out


class MemoryBank:
    """FIFO of normalised narration embeddings used as negatives by the narration SSL term (meta_archs.py:38-60); lives on
    the device of the features it is first updated with."""

    def __init__(self, size, feature_dim):
        self.size, self.feature_dim = size, feature_dim
        self.memory = torch.randn(size, feature_dim)
        self.ptr = 0

    @torch.no_grad()
    def update(self, features):
        self.memory = self.memory.to(features.device)
        n = features.size(0)
        assert n <= self.size, "Batch size must be less than or equal to memory bank size"
        if self.ptr + n <= self.size:
            self.memory[self.ptr:self.ptr + n] = features
            self.ptr += n
        else:
            overflow = (self.ptr + n) - self.size
            self.memory[self.ptr:] = features[:self.size - self.ptr]
            self.memory[:overflow] = features[self.size - self.ptr:]
            self.ptr = overflow

    def get_all(self):
        return self.memory


class _TapeLoss(torch.autograd.Function):
    """Connects the hand-written backward pass to torch autograd.  backward(g) replays the tape scaled by g.

    Two modes: with a trainer (`model._direct_grads`) the kernels accumulate straight into the existing `.grad` buffers and
    nothing is returned.  Otherwise — the reference's own loop, possibly under DistributedDataParallel — the parameters the
    tape writes are inputs of this Function and their gradients are RETURNED, so torch's AccumulateGrad nodes (and with
    them DDP's reducer hooks, GradScaler, hooks of the EWC / MAS code, ...) see them exactly as for the reference model."""

    @staticmethod
    def forward(ctx, hook, value, run_backward, keys, *params):
        ctx.run_backward, ctx.keys, ctx.shapes = run_backward, keys, [(p.shape, p.dtype, p.device) for p in params]
        ctx.nosync = bool(getattr(run_backward, "nosync", False))
        return value.detach().clone()

    @staticmethod
    def backward(ctx, g):
        grads = ctx.run_backward(g if ctx.nosync else float(g)) or {}
        outs = []
        for k, (shape, dtype, dev) in zip(ctx.keys, ctx.shapes):
            gk = grads.get(k)
            # a listed parameter the tape did not reach (only possible in the very first step, before the set of live
            # parameters is known) gets zeros: every input of a DDP-wrapped Function must receive a gradient
            outs.append(gk if gk is not None else torch.zeros(shape, dtype=dtype, device=dev))
        return (torch.zeros_like(g), None, None, None, *outs)


class _Head(nn.Module):
    def __init__(self, input_dim, feat_dim, num_layers, kernel_size, with_ln):
        super().__init__()
        assert with_ln and kernel_size == 3 and num_layers == 3, "MQ configs: 3-layer k=3 heads with LayerNorm"
        self.head, self.norm = nn.ModuleList(), nn.ModuleList()
        for idx in range(num_layers - 1):
            self.head.append(MaskedConv1D(input_dim if idx == 0 else feat_dim, feat_dim, kernel_size, stride=1,
                                          padding=kernel_size // 2, bias=False))
            self.norm.append(LayerNorm(feat_dim))


class PtTransformerClsHead(_Head):
    """Shared classification head (reference: meta_archs.py:183-275)."""

    def __init__(self, input_dim, feat_dim, num_classes, prior_prob=0.01, num_layers=3, kernel_size=3,
                 act_layer=nn.ReLU, with_ln=False, empty_cls=(), detach_feat=False):
        super().__init__(input_dim, feat_dim, num_layers, kernel_size, with_ln)
        self.num_classes = num_classes
        self.cls_head = MaskedConv1D(feat_dim, num_classes, kernel_size, stride=1, padding=kernel_size // 2)
        torch.nn.init.constant_(self.cls_head.conv.bias, -(math.log((1 - prior_prob) / prior_prob)))
        for idx in empty_cls:
            torch.nn.init.constant_(self.cls_head.conv.bias[idx], -(math.log((1 - 1e-6) / 1e-6)))
        self.reg_params = {}

    def augment_classification(self, num_new_classes, device):
        self.cls_head.augment_classification(num_new_classes, device)
        self.num_classes += num_new_classes


class PtTransformerRegHead(_Head):
    """Shared regression head (reference: meta_archs.py:278-349)."""

    def __init__(self, input_dim, feat_dim, fpn_levels, num_layers=3, kernel_size=3, act_layer=nn.ReLU, with_ln=False,
                 num_bins=16):
        super().__init__(input_dim, feat_dim, num_layers, kernel_size, with_ln)
        self.fpn_levels = fpn_levels
        self.scale = nn.ModuleList([Scale() for _ in range(fpn_levels)])
        self.offset_head = MaskedConv1D(feat_dim, 2 * (num_bins + 1), kernel_size, stride=1, padding=kernel_size // 2)
        self.reg_params = {}


class _Cfg:
    """The engine's view of the model configuration."""
    pass


@register_meta_arch("LocPointTransformer")
class PtTransformer(nn.Module):
    def __init__(self, backbone_type, fpn_type, use_xl, backbone_arch, scale_factor, input_dim, max_seq_len,
                 max_buffer_len_factor, n_head, n_mha_win_size, embd_kernel_size, embd_dim, embd_with_ln, fpn_dim,
                 fpn_with_ln, fpn_start_level, head_dim, regression_range, head_num_layers, head_kernel_size,
                 head_with_ln, use_abs_pe, use_rel_pe, num_classes, train_cfg, test_cfg, cl_cfg, use_cross_modal,
                 n_txt_in):
        super().__init__()
        self.fpn_strides = [scale_factor ** i for i in range(fpn_start_level, backbone_arch[-1] + 1)]
        self.reg_range = regression_range
        assert len(self.fpn_strides) == len(self.reg_range)
        self.scale_factor, self.num_classes, self.max_seq_len, self.use_xl = scale_factor, num_classes, max_seq_len, use_xl
        n_levels = 1 + backbone_arch[-1]
        self.mha_win_size = [n_mha_win_size] * n_levels if isinstance(n_mha_win_size, int) else list(n_mha_win_size)
        max_div_factor = 1
        for s, w in zip(self.fpn_strides, self.mha_win_size):  # meta_archs.py:402-415
            stride = s * (w // 2) * 2 if w > 1 else s
            assert max_seq_len % stride == 0, "max_seq_len must be divisible by fpn stride and window size"
            max_div_factor = max(max_div_factor, stride)
        self.max_div_factor = max_div_factor

        tc = train_cfg
        self.train_center_sample = tc["center_sample"]
        assert self.train_center_sample in ["radius", "none"]
        self.train_center_sample_radius = tc["center_sample_radius"]
        self.train_loss_weight = tc["loss_weight"]
        self.train_cls_prior_prob = tc["cls_prior_prob"]
        self.train_dropout, self.train_droppath = tc["dropout"], tc["droppath"]
        self.train_label_smoothing = tc["label_smoothing"]
        self.t_c_alpha, self.al_loss_weight = tc["t_c_alpha"], tc["al_loss_weight"]
        assert not tc.get("use_dcn", False) and not tc.get("use_us_fpn", False)
        te = test_cfg
        self.test_pre_nms_thresh, self.test_pre_nms_topk = te["pre_nms_thresh"], te["pre_nms_topk"]
        self.test_iou_threshold, self.test_min_score = te["iou_threshold"], te["min_score"]
        self.test_max_seg_num, self.test_nms_method = te["max_seg_num"], te["nms_method"]
        assert self.test_nms_method in ["soft", "hard", "none"]
        self.test_duration_thresh, self.test_multiclass_nms = te["duration_thresh"], te["multiclass_nms"]
        self.test_nms_sigma, self.test_voting_thresh = te["nms_sigma"], te["voting_thresh"]
        self.use_cross_modal, self.n_txt_in = use_cross_modal, n_txt_in

        assert backbone_type == "convTransformer", "only the convTransformer backbone is on the MQ path"
        win = n_mha_win_size if isinstance(n_mha_win_size, int) else -1
        self.backbone = make_backbone("convTransformer", n_in=input_dim, n_embd=embd_dim, n_head=n_head,
                                      n_embd_ks=embd_kernel_size, max_len=max_seq_len, use_xl=use_xl, arch=backbone_arch,
                                      t_c_alpha=self.t_c_alpha, scale_factor=scale_factor, with_ln=embd_with_ln,
                                      attn_pdrop=0.0, proj_pdrop=self.train_dropout, path_pdrop=self.train_droppath,
                                      use_abs_pe=use_abs_pe, use_rel_pe=use_rel_pe, use_cross_modal=use_cross_modal,
                                      n_txt_in=n_txt_in, mha_win_size=-1)  # MQ's TransformerBlock ignores the window (blocks.py:497)
        del win
        if isinstance(embd_dim, (list, tuple)):
            embd_dim = sum(embd_dim)
        self.embd_dim, self.n_head, self.backbone_arch = embd_dim, n_head, tuple(backbone_arch)
        self.input_dim = input_dim[0] if isinstance(input_dim, (list, tuple)) else input_dim
        assert fpn_type in ("identity", "fpn")       # 'fpn' = FPN1D + ACConv / DenseAPP (necks.py:13-106): evaluation path only
        assert embd_dim == fpn_dim == head_dim
        self.fpn_type = fpn_type
        self.neck = make_neck(fpn_type, in_channels=[embd_dim] * n_levels, out_channel=fpn_dim,
                              scale_factor=scale_factor, start_level=fpn_start_level, with_ln=fpn_with_ln)
        self.point_generator = make_generator("point", max_seq_len=max_seq_len * max_buffer_len_factor,
                                              fpn_strides=self.fpn_strides, regression_range=self.reg_range)
        self.cls_head = PtTransformerClsHead(fpn_dim, head_dim, self.num_classes, kernel_size=head_kernel_size,
                                             prior_prob=self.train_cls_prior_prob, with_ln=head_with_ln,
                                             num_layers=head_num_layers, empty_cls=tc["head_empty_cls"])
        self.reg_head = PtTransformerRegHead(fpn_dim, head_dim, len(self.fpn_strides), kernel_size=head_kernel_size,
                                             num_layers=head_num_layers, with_ln=head_with_ln, num_bins=0)
        K = self.num_classes
        self.mu = nn.Parameter(torch.zeros(K, 1))
        self.sigma = nn.Parameter(torch.ones(K, 1))
        self.mu_reg_left = nn.Parameter(-torch.ones(K, 1) * 0.5)
        self.sigma_reg_left = nn.Parameter(torch.ones(K, 1))
        self.mu_reg_right = nn.Parameter(torch.ones(K, 1) * 0.5)
        self.sigma_reg_right = nn.Parameter(torch.ones(K, 1))
        self._ln_host, self._ln_dev = None, None
        self.loss_normalizer = tc["init_loss_norm"]
        self.loss_normalizer_momentum = 0.9
        self.reg_params = {}
        # continual-learning bookkeeping touched by the unchanged orchestration (train_cl.py / train_bic.py)
        self.cl_name = cl_cfg["name"]
        self.compute_means = self.cl_name == "icarl"
        self.exemplar_means, self.memory = [], {}
        self.adv_lambda, self.type_sampling = cl_cfg["adv_lambda"], cl_cfg["type_sampling"]
        self.n_known = 0
        self.list_bias_layers, self.list_splits = [], []
        self.prompt_pool = cl_cfg["prompt_pool"]
        self.use_prompt_mask = True
        if cl_cfg["length"] is not None and cl_cfg["pool_size"] is not None and self.prompt_pool:   # meta_archs.py:634-648
            self.prompt = Prompt(length=cl_cfg["length"], embed_dim=cl_cfg["embed_dim"], pool_size=cl_cfg["pool_size"],
                                 top_k=cl_cfg["topk"], batchwise_prompt=True)
        self.narration_ssl = cl_cfg["narration_ssl"]
        self.narration_dim = cl_cfg["narration_dim"]
        if self.narration_ssl:   # training-only branch; the parameters exist so checkpoints load (meta_archs.py:650-655)
            self.narration_encoder = nn.Linear(cl_cfg["narration_dim"], 1024)
            self.memory_bank = MemoryBank(cl_cfg["memory_size"], 1024)
        self.ssl_factor = cl_cfg["ssl_factor"]
        self.num_emas, self.ema_decay = 1, 0.999
        self.use_adapt = cl_cfg["use_adapt"]
        self.adapt_blocks = tuple(cl_cfg["adapt_blocks"]) if self.use_adapt else ()
        if self.use_adapt:       # meta_archs.py:657-700
            assert max_seq_len == 1024, "the reference sizes the temporal adapters for T = 1024 (meta_archs.py:687)"
            self.num_freeze_epochs = 10
            self.pets_emas = nn.ModuleList([])
            pets, dim = nn.ModuleList([]), 1024
            for _ in self.adapt_blocks:
                pets.append(Adapter(embed_dim=dim, down_sample=5, mode="parallel", scale="null"))
                dim //= 2
            self.pets = pets
            self.pets_emas.append(ModelEmaV2(self.pets, decay=self.ema_decay))
            self.attach_pets(self.pets)
        self._packed = None
        self._packed_key = None
        self._pe = None

    # ---- small surface used by the orchestration -------------------------------------------------
    @property
    def device(self):
        return next(self.parameters()).device

    def _param_table(self):
        """(names, parameters) cached: walking the module tree costs ~4 ms and is needed several times per step."""
        tab = getattr(self, "_ptab", None)
        if tab is None:      # reset by attach_pets / augment_classification, the only places that add or replace parameters
            named = list(self.named_parameters())
            tab = self._ptab = ([k for k, _ in named], [p_ for _, p_ in named], len(named))
        return tab

    def _canon_key(self, key):
        """tape / packed-weight key -> the name `named_parameters()` lists the parameter under.  The temporal adapters are
        registered twice (`pets.<i>.*` and `backbone.branch.<b>.adapters.attn.*`, like the reference); named_parameters()
        de-duplicates and keeps the backbone name, while the engines address them as `pets.<i>.*`."""
        if key.startswith("pets.") and self.use_adapt:
            i, rest = key[5:].split(".", 1)
            return f"backbone.branch.{self.adapt_blocks[int(i)]}.adapters.attn.{rest}"
        return key

    def _tape_key(self, name):
        """inverse of `_canon_key`"""
        if ".adapters.attn." in name and name.startswith("backbone.branch."):
            b, rest = name[len("backbone.branch."):].split(".adapters.attn.", 1)
            return f"pets.{self.adapt_blocks.index(int(b))}.{rest}"
        return name

    def attach_pets(self, pets):
        """register the adapters under backbone.branch.<b>.adapters.attn like AdapterMixin.attach_adapter (blocks.py:28-33)"""
        for i, b in enumerate(self.adapt_blocks):
            blk = self.backbone.branch[b]
            if not isinstance(getattr(blk, "adapters", None), nn.ModuleDict):
                blk.adapters = nn.ModuleDict()
            blk.adapters["attn"] = pets[i]
        self._ptab = None

    def pre_train_epoch(self, task_id=0, current_epoch=0):
        if self.use_adapt:
            for prm in self.pets.parameters():
                prm.requires_grad_(True)

    def post_train_step(self):
        if self.use_adapt:   # meta_archs.py:702-707
            for idx, ema in enumerate(reversed(self.pets_emas)):
                ema.update(self.pets if idx == 0 else self.pets_emas[idx - 1])

    def add_samples_to_mem(self, cilsettask, data, m):
        """Replay memory update after a task (train_cl.py:353; reference: meta_archs.py:972, active part :1043-1055): the new
        classes' clips are merged into `self.memory` (same class id: replaced), every class list is shuffled in place with
        Python's `random` — one shuffle per class in dictionary order, so a seeded run keeps the reference's exemplars — and
        cut to `m` clips (`'ALL'` keeps everything).  `cilsettask` is unused, as in the reference's random sampling."""
        import random
        self.memory = {**self.memory, **data}
        for class_id, videos in self.memory.items():
            random.shuffle(videos)
            self.memory[class_id] = videos if m == 'ALL' else videos[:m]
        for class_id, videos in self.memory.items():
            print('Memory... Class: {}, num videos: {}'.format(class_id, len(videos)))

    def augment_classification(self, num_new_classes, device):
        """Grow the classifier and the per-class gaussian parameters (reference: meta_archs.py:715-751)."""
        device = self.mu.device
        self.cls_head.augment_classification(num_new_classes, device)
        old = self.num_classes
        self.num_classes += num_new_classes
        for name, init in (("mu", 0.0), ("sigma", 1.0), ("mu_reg_left", -0.5), ("sigma_reg_left", 1.0),
                           ("mu_reg_right", 0.5), ("sigma_reg_right", 1.0)):
            new = nn.Parameter(torch.full((self.num_classes, 1), init, device=device))
            new.data[:old] = getattr(self, name).data
            setattr(self, name, new)
        self._packed = None
        self._ptab = None

    # ---- weights ---------------------------------------------------------------------------------
    # ---- loss normaliser (EMA of the number of positives, meta_archs.py:1425-1431): a python float for the orchestration, a
    # device scalar for the training step, synchronised lazily so that a step issues forward AND backward without waiting ----
    @property
    def loss_normalizer(self):
        if self._ln_host is None:
            self._ln_host = float(self._ln_dev)          # the one host read; happens when somebody asks (logging, checkpoints)
        return self._ln_host

    @loss_normalizer.setter
    def loss_normalizer(self, v):
        self._ln_host, self._ln_dev = float(v), None

    def _ln_device(self, dev):
        if self._ln_dev is None or self._ln_dev.device != dev:
            self._ln_dev = torch.full((), float(self._ln_host), device=dev, dtype=torch.float32)
        return self._ln_dev

    def engine_cfg(self):
        c = _Cfg()
        c.embd_dim, c.n_head, c.arch, c.scale_factor = self.embd_dim, self.n_head, self.backbone_arch, self.scale_factor
        c.use_cross_modal, c.use_xl, c.t_c_alpha, c.max_seq_len = self.use_cross_modal, self.use_xl, self.t_c_alpha, self.max_seq_len
        c.adapt_blocks = tuple(self.adapt_blocks)
        c.fpn_type = self.fpn_type
        return c

    def use_flat_optimizer(self, opt):
        """Register a trainer.FlatAdamW: its bf16 planes become the GEMM operands of this model (no re-packing per step)."""
        self._flat = opt
        self._packed = None
        self._direct_grads = True

    def packed_weights(self):
        """bf16 operand copies of the parameters, re-packed whenever a parameter changed (optimizer step, load_state_dict,
        augment_classification) or the precision mode changed.  The EMA copies of the adapters change after every training
        step (`post_train_step`) but are small: they are re-packed on their own."""
        flat = getattr(self, "_flat", None)
        names, plist, _ = self._param_table()
        ema = getattr(self, "_ema_idx", None)
        if ema is None or len(ema[0]) + len(ema[1]) != len(names):
            ema = self._ema_idx = ([i for i, k in enumerate(names) if not k.startswith("pets_emas.")],
                                   [i for i, k in enumerate(names) if k.startswith("pets_emas.")])
        key = (ops.precision(), tuple(plist[i]._version for i in ema[0]), tuple(id(p) for p in plist),
               flat.layout_version if flat is not None else 0)
        ema_key = tuple(plist[i]._version for i in ema[1])
        if self._packed is None or key != self._packed_key:
            if flat is not None:
                flat.refresh_planes()
            self._packed = E.pack_weights(self.state_dict(), self.device, flat, dict(zip(names, plist)) if flat is not None else None)
            self._packed_key = key
            self._packed_ema_key = ema_key
            self._packed_epoch = flat.epoch if flat is not None else 0
            self._pe = E.sinusoid_pe_table(self.max_seq_len, self.embd_dim, self.device)
        else:
            if flat is not None and self._packed_epoch != flat.epoch:
                E.refresh_packed(self._packed)
                self._packed_epoch = flat.epoch
            if ema_key != self._packed_ema_key:
                for i in ema[1]:
                    k, p = names[i], plist[i]
                    self._packed[k] = E._pack_one(E._pack_kind(k, p), p.detach().to(dtype=torch.float32))
                self._packed["_cache"] = {}
                self._packed_ema_key = ema_key
        return self._packed

    # ---- preprocessing (reference: meta_archs.py:1134-1221) -------------------------------------------
    def preprocessing(self, video_list, is_training=True, padding_val=0.0):
        vl = [x for x in video_list if len(x["labels"]) > 0] if is_training else list(video_list)
        feats = [x["feats"] for x in vl]
        lens = [f.shape[-1] for f in feats]
        max_len = max(lens)
        if is_training:
            assert max_len <= self.max_seq_len, "Input length must be smaller than max_seq_len during training"
            max_len = self.max_seq_len
        elif max_len <= self.max_seq_len:
            max_len = self.max_seq_len
        else:
            stride = self.max_div_factor
            max_len = (max_len + (stride - 1)) // stride * stride
        if max_len != self.max_seq_len:
            raise NotImplementedError("inputs longer than max_seq_len need the interpolated positional encoding "
                                      "(backbones.py:229-236): not built yet")
        dev = self.device
        B, Cin = len(feats), feats[0].shape[0]
        if all(f.is_cuda for f in feats):
            batched = torch.full((B, Cin, max_len), padding_val, device=dev, dtype=torch.float32)
            for i, f in enumerate(feats):
                batched[i, :, :lens[i]].copy_(f)
        elif all(f.is_pinned() for f in feats):   # pinned host buffers: straight async H2D into the padded batch
            batched = torch.full((B, Cin, max_len), padding_val, device=dev, dtype=torch.float32) \
                if min(lens) < max_len else torch.empty((B, Cin, max_len), device=dev, dtype=torch.float32)
            for i, f in enumerate(feats):
                batched[i, :, :lens[i]].copy_(f, non_blocking=True)
        else:
            stage = torch.full((B, Cin, max_len), padding_val, dtype=torch.float32).pin_memory()
            for i, f in enumerate(feats):
                stage[i, :, :lens[i]].copy_(f)
            batched = stage.to(dev, non_blocking=True)
        lens_t = torch.as_tensor(lens)
        mask = (torch.arange(max_len)[None, :] < lens_t[:, None]).float().to(dev)
        return vl, batched, mask

    def query_preprocessing(self, video_list, padding_val=0.0):
        feats = [x["prompt_feature"] for x in video_list]
        lens = [f.shape[-1] for f in feats]
        max_len = max(lens)
        B, Ct = len(feats), feats[0].shape[0]
        stage = torch.full((B, Ct, max_len), padding_val, dtype=torch.float32)
        for i, f in enumerate(feats):
            stage[i, :, :lens[i]].copy_(f)
        mask = (torch.arange(max_len)[None, :] < torch.as_tensor(lens)[:, None]).float()
        return stage.to(self.device), mask.to(self.device), torch.as_tensor(lens, dtype=torch.int32).to(self.device)

    # ---- network ---------------------------------------------------------------------------------
    @torch.no_grad()
    def _device_forward(self, batched, mask, text, tmask, tlens, is_training):
        """Everything that runs on the device between the padded input batch and the head outputs."""
        W = self.packed_weights()
        cfg = self.engine_cfg()
        x16 = ops.pack_feats(batched, planes=ops.PLANES_HI)      # the input projection reads split operands
        t16 = ops.pack_feats(text) if text is not None else None
        # evaluation: the reference runs one clip at a time, so its text is never padded; a batched evaluation must
        # therefore treat every text as un-padded (text_lens).  Training reproduces the reference's padded batch.
        trunk = E.backbone_fwd(W, cfg, x16, mask, t16, tmask, self._pe, text_lens=None if is_training else tlens,
                               trunk_only=True)
        feats, masks = E.branch_fwd(W, cfg, trunk, "pets.")
        logits, offsets, pmask, pyr = E.neck_heads_fwd(W, cfg, feats, masks)
        if not is_training and self.use_adapt:
            # EMA-adapter ensemble (meta_archs.py:854-881): the reference re-runs the whole network with the EMA copy
            # of the adapters and averages logits / offsets; only the strided branch depends on the adapters, so the
            # shared trunk (embedding, text path, stem, XLNet) is computed once here.
            for e in range(len(self.pets_emas)):
                f2, m2 = E.branch_fwd(W, cfg, trunk, f"pets_emas.{e}.module.")
                l2, o2, _, _ = E.neck_heads_fwd(W, cfg, f2, m2, pyr)
                logits, _ = ops.axpby(logits, l2, 0.5, 0.5)
                offsets, _ = ops.axpby(offsets, o2, 0.5, 0.5)
        return logits, offsets, pmask, pyr

    def _prompted_text(self, text, tlens, is_training, task_id):
        """L2P prompts prepended to the text tokens (meta_archs.py:759-780).  text (B, Ct, L) on the device.
        Returns (text', mask', lens', reduce_sim): the mask keeps the reference's quirk of being computed from the
        PRE-prompt lengths over the prompted sequence."""
        x = text.permute(0, 2, 1)
        prompt_mask = None
        if is_training:
            start, end = task_id * self.prompt.top_k, (task_id + 1) * self.prompt.top_k
            if end <= self.prompt.pool_size:
                prompt_mask = torch.arange(start, end, device=x.device).unsqueeze(0).expand(x.shape[0], -1)
        res = self.prompt(x, prompt_mask=prompt_mask, cls_features=None)
        self.total_prompt_len = res["total_prompt_len"]
        text2 = res["prompted_embedding"].permute(0, 2, 1).contiguous()
        L2 = text2.shape[-1]
        mask2 = (torch.arange(L2, device=x.device)[None, :] < tlens[:, None]).float()
        lens2 = (tlens + self.total_prompt_len).to(torch.int32)   # tensor extent of every prompted text
        return text2, mask2, lens2, res["reduce_sim"]

    @torch.no_grad()
    def _network(self, video_list, is_training, task_id=-1):
        vl, batched, mask = self.preprocessing(video_list, is_training)
        text, tmask, tlens = None, None, None
        self._reduce_sim = None
        if self.use_cross_modal:
            src = vl if is_training else video_list
            text, tmask, tlens = self.query_preprocessing(src)
            if hasattr(self, "prompt"):
                if not is_training and len(src) > 1:
                    # batchwise prompt selection + zero-padded means depend on the batch composition: keep the
                    # reference's one-clip-at-a-time result by selecting prompts per clip
                    parts = [self._prompted_text(text[i:i + 1, :, :int(tlens[i])], tlens[i:i + 1], False, task_id)
                             for i in range(len(src))]
                    L2 = max(p_[0].shape[-1] for p_ in parts)
                    text = torch.zeros(len(src), text.shape[1], L2, device=text.device)
                    tmask = torch.zeros(len(src), L2, device=text.device)
                    for i, p_ in enumerate(parts):
                        text[i, :, :p_[0].shape[-1]] = p_[0][0]
                        tmask[i, :p_[1].shape[-1]] = p_[1][0]
                    tlens = torch.cat([p_[2] for p_ in parts])
                else:
                    text, tmask, tlens, self._reduce_sim = self._prompted_text(text, tlens, is_training, task_id)
            text = text.contiguous()
        logits, offsets, pmask, pyr = self._device_forward(batched, mask.contiguous(), text, tmask, tlens, is_training)
        return vl, logits, offsets, pmask, pyr

    # ---- CUDA-graph replay of the whole evaluation step (static shapes: every clip is padded to max_seq_len) --------
    @torch.no_grad()
    def make_eval_graph(self, batch_size, text_len=128):
        """Capture pack -> backbone -> neck -> heads -> decode -> soft-NMS for `batch_size` clips into one CUDA graph.
        Returns an EvalGraph; .run(video_list) stages inputs, replays, and post-processes like forward()."""
        return EvalGraph(self, batch_size, text_len)

    def forward(self, video_list, task_id=-1, ensemble=False, hidden_state=False, is_training=True,
                prev_out_cls_logits=None, get_emb=False, val_qilDatasetList=None):
        if not is_training and not get_emb:
            assert len(video_list) >= 1
        if is_training and not get_emb and torch.is_grad_enabled():
            return self._train_forward(video_list, task_id, prev_out_cls_logits)
        vl, logits, offsets, pmask, pyr = self._network(video_list, is_training, task_id)
        logits = self._apply_bias_layers(logits)
        if get_emb:
            cls_l = [logits[:, o:o + n] for o, n in zip(pyr.off, pyr.lens)]
            off_l = [offsets[:, o:o + n] for o, n in zip(pyr.off, pyr.lens)]
            msk_l = [pmask[:, o:o + n].bool() for o, n in zip(pyr.off, pyr.lens)]
            if not is_training and self.use_adapt:
                # reference quirk: its EMA-ensemble loop re-binds fpn_masks to the un-squeezed (B,1,T_l) tensors (:864)
                msk_l = [m_.unsqueeze(1) for m_ in msk_l]
            return cls_l, off_l, msk_l
        if is_training:
            return self.losses(vl, logits, offsets, pmask, pyr, prev_out_cls_logits)
        results = self.inference(video_list, pyr, pmask, logits, offsets, cilsettask=val_qilDatasetList)
        if ensemble:
            points = self.point_generator(pyr.lens)
            cls_l = [logits[:, o:o + n] for o, n in zip(pyr.off, pyr.lens)]
            off_l = [offsets[:, o:o + n] for o, n in zip(pyr.off, pyr.lens)]
            msk_l = [pmask[:, o:o + n].bool() for o, n in zip(pyr.off, pyr.lens)]
            return video_list, points, msk_l, cls_l, off_l
        return results

    def _apply_bias_layers(self, logits):
        """BiC: BiasLayer on the class slices of the logits, at training AND inference (meta_archs.py:823-836)"""
        if self.n_known > 0 and self.cl_name == "bic" and len(self.list_splits) > 0:
            parts, lo_ = [], 0
            for i, hi_ in enumerate(self.list_splits):
                parts.append(self.list_bias_layers[i](logits[:, :, lo_:hi_]))
                lo_ = hi_
            return torch.cat(parts, dim=2).contiguous()
        return logits

    def _grad_sinks(self):
        """key -> fp32 view of the parameter's existing .grad in the kernels' packed layout, for the parameters whose
        packed layout equals the parameter layout (vectors, 1x1 convs, nn.Linear, XLNet o).  With a trainer the .grad
        tensors are persistent views of one flat buffer, so the table is rebuilt only when a .grad pointer changes."""
        names, plist, _ = self._param_table()
        sig = tuple(p.grad.data_ptr() if (p.grad is not None and p.requires_grad) else 0 for p in plist)
        if getattr(self, "_sinks_sig", None) != sig:
            sinks = {}
            for k, p in zip(names, plist):
                if p.grad is None or not p.requires_grad or not p.grad.is_contiguous() or p.grad.dtype != torch.float32:
                    continue
                k = self._tape_key(k)           # adapters: the tape addresses them as pets.<i>.*
                kind = E._pack_kind(k, p)
                if kind == "vec" and p.dim() != 2:
                    sinks[k] = p.grad.view(-1)
                elif kind == "gemm":
                    sinks[k] = p.grad.view(p.shape[0], -1)
            self._sinks, self._sinks_sig = sinks, sig
        return self._sinks

    # ---- training step: taped forward on the CUDA kernels + hand-written backward -------------------------------
    def _train_forward(self, video_list, task_id=-1, prev_out_cls_logits=None):
        """forward(is_training=True) with gradients: returns the reference's loss dict; `final_loss.backward()` runs the
        CUDA backward pass (vilco_b200/train_engine.py) and accumulates into every parameter's .grad.  mu / sigma (and
        the L2P prompt pool) receive their gradients through torch autograd of the small target-assignment glue."""
        from .. import train_engine as TE
        if self.fpn_type != "identity":
            raise NotImplementedError("fpn_type 'fpn' (FPN1D) is built for evaluation only; no MQ config trains with it")
        dev = self.device
        vl, batched, mask = self.preprocessing(video_list, True)
        text = tmask = None
        self._reduce_sim = None
        if self.use_cross_modal:
            text, tmask, tlens = self.query_preprocessing(vl)
            if hasattr(self, "prompt"):
                text, tmask, tlens, self._reduce_sim = self._prompted_text(text, tlens, True, task_id)
            text = text.contiguous()
        W = self.packed_weights()
        cfg = self.engine_cfg()
        # dropout / stochastic depth follow nn.Module.training exactly like the reference's nn.Dropout / AffineDropPath /
        # XLNet dropout (xlnet_config_*.json: 0.1); model.eval() + is_training=True gives the deterministic losses
        self._train_calls = getattr(self, "_train_calls", 0) + 1
        sinks = self._grad_sinks() if getattr(self, "_direct_grads", False) else {}
        if self.training:
            tp = TE.Tape(W, dropout=self.train_dropout, droppath=self.train_droppath,
                         xl_dropout=getattr(self, "xl_dropout", 0.1) if self.use_xl else 0.0,   # xlnet_config_*.json: 0.1
                         seed=(int(torch.initial_seed()) & 0xFFFFF) * 4096 + self._train_calls, sinks=sinks)
        else:
            tp = TE.Tape(W, sinks=sinks)
        tp.after_node = getattr(self, "_after_backward_node", None)     # set by trainer.Trainer (bucketed all-reduce)
        with torch.no_grad():
            x16 = ops.pack_feats(batched, planes=ops.PLANES_HI)
            t16 = ops.pack_feats(text.detach()) if text is not None else None
            feats, masks, tin = TE.backbone(tp, cfg, x16, mask.contiguous(), t16, tmask, self._pe)
            if tin is not None and not text.requires_grad:
                tin.const = True
            logitsV, offsetsV, pmask, pyr, fpn_lv = TE.neck_heads(tp, cfg, feats, masks)
            logits, offsets = logitsV.v, offsetsV.v
        # torch-side loss terms of the mq_vilco branches (tiny; their autograd runs inside run_backward):
        extra, extra_named, ssl_leaves = [], {}, None
        if self.training and self.narration_ssl and self.use_cross_modal:
            ssl_term, ssl_leaves = self._narration_ssl(vl, fpn_lv, masks)
            if ssl_term is not None:
                extra.append(ssl_term)
                extra_named["ssl_loss"] = ssl_term.detach()
        if self.n_known > 0 and self.cl_name == "l2p" and self._reduce_sim is not None:
            extra.append(-0.1 * self._reduce_sim)         # pull constraint of L2P (meta_archs.py:1478-1480)
        B, P, K = logits.shape
        # BiC / iCaRL (n_known > 0): bias layers on the class slices of the logits and the distillation term are torch
        # expressions on a leaf cut at the logits (meta_archs.py:823-836, 1482-1519)
        lg_leaf = lg_biased = None
        if self.n_known > 0 and self.cl_name in ("bic", "icarl"):
            lg_leaf = logits.detach().requires_grad_(True)
            with torch.enable_grad():
                lg_biased = lg_leaf
                if self.cl_name == "bic" and len(self.list_splits) > 0:
                    parts, lo_ = [], 0
                    for i, hi_ in enumerate(self.list_splits):
                        parts.append(self.list_bias_layers[i](lg_leaf[:, :, lo_:hi_]))
                        lo_ = hi_
                    lg_biased = torch.cat(parts, dim=2)
                dist = self._distill_term(lg_biased, pyr, prev_out_cls_logits)
            logits = lg_biased.detach().contiguous()
            extra.append(dist)
            extra_named["dist_loss"] = dist.detach()
        if self.train_label_smoothing > 0:
            # the fused loss kernels derive the positive set from the targets they are given; the reference takes it from the
            # UN-smoothed targets (meta_archs.py:1400) and smooths a copy — not built (no MQ config smooths: default 0.0)
            raise NotImplementedError("train_cfg.label_smoothing > 0 is not supported by the fused loss kernels")
        gt_cls, gt_off, wc, wl, wr = self._label_points(pyr, [x["segments"] for x in vl], [x["labels"] for x in vl])
        with torch.no_grad():
            gt_cls, gt_off = gt_cls.detach(), gt_off.detach()
            present = torch.zeros(B, K, pin_memory=dev.type == "cuda")
            for i, x in enumerate(vl):
                present[i, x["labels"].cpu()] = 1
            present = present.to(dev, non_blocking=True)
            sums = torch.zeros(4, device=dev)
            scratch = torch.zeros(B * K, device=dev, dtype=torch.int32)
            wcd, wld, wrd = wc.detach().contiguous(), wl.detach().contiguous(), wr.detach().contiguous()
            L.check(L.lib().vilco_mq_losses(
                ops._p(logits), ops._p(offsets), ops._p(pmask), ops._p(pyr.gap_rows), ops._p(gt_cls), ops._p(gt_off),
                ops._p(wcd), ops._p(wld), ops._p(wrd), ops._p(present), B, P, K, C.c_float(0.25), C.c_float(2.0), ops._p(sums),
                ops._p(scratch), L.stream_ptr()), "vilco_mq_losses")
            # no host round trip between forward and backward (the deep pyramid levels at the start of the backward are
            # launch-bound, so the host must already be ahead when the GPU gets there): the normaliser lives on the device.
            # The reference's adaptive regression weight (train_loss_weight <= 0) needs the loss VALUES on the host.
            nosync = self.train_loss_weight > 0 and dev.type == "cuda" and NOSYNC_STEP
            if nosync:
                norm = self.loss_normalizer_momentum * self._ln_device(dev) + \
                    (1 - self.loss_normalizer_momentum) * sums[2].clamp(min=1.0)
                self._ln_dev, self._ln_host = norm, None
                s = None
            else:
                s = sums.cpu()
                self.loss_normalizer = self.loss_normalizer_momentum * self.loss_normalizer + \
                    (1 - self.loss_normalizer_momentum) * max(float(s[2]), 1)
                norm = float(self.loss_normalizer)
            cls_loss, reg_loss = sums[0] / norm, sums[1] / norm
            al_loss = sums[3] / norm if K != 1 else torch.zeros((), device=dev)
            w_reg = self.train_loss_weight if self.train_loss_weight > 0 else float(s[0] / norm) / max(float(s[1] / norm), 0.01)
            w_al = self.al_loss_weight if K != 1 else 0.0
            final = cls_loss + reg_loss * w_reg + al_loss * w_al
            for t_ in extra:
                final = final + t_.detach()
        names_, plist_, _ = self._param_table()
        named = dict(zip(names_, plist_))
        model = self
        direct = bool(getattr(self, "_direct_grads", False))
        # parameters only the torch-side glue reaches (target-assignment Gaussians, prompt pool, narration encoder): their
        # gradients are computed inside run_backward with torch.autograd.grad — never by a nested .backward(), whose
        # AccumulateGrad hooks would fire a second time under DistributedDataParallel — and delivered like the tape's
        owned = [(k_, p_) for k_, p_ in zip(names_, plist_) if p_.requires_grad and (
            k_ in ("mu", "sigma", "mu_reg_left", "sigma_reg_left", "mu_reg_right", "sigma_reg_right")
            or k_.startswith(("prompt.", "narration_encoder.")))]
        bias_params = [q_ for bl in self.list_bias_layers for q_ in bl.parameters()] if lg_leaf is not None else []
        live = []
        if not direct:
            # autograd mode: every parameter whose gradient this Function will return (all trainable ones until the first
            # backward has shown which ones the tape reaches)
            known = getattr(self, "_live_keys", None)
            owned_names = {k_ for k_, _ in owned}
            for k_, p_ in zip(names_, plist_):
                if not p_.requires_grad or k_.startswith("pets_emas."):
                    continue
                if k_ in owned_names or known is None or k_ in known:
                    live.append((k_, p_))
        passed = {k_ for k_, _ in live}

        def run_backward(gscale):
            with torch.no_grad():
                dlogits, doffsets = torch.zeros_like(logits), torch.zeros_like(offsets)
                dwc, dwl, dwr = torch.zeros_like(wcd), torch.zeros_like(wld), torch.zeros_like(wrd)
                L.check(L.lib().vilco_mq_losses_bwd(
                    ops._p(logits), ops._p(offsets), ops._p(pmask), ops._p(pyr.gap_rows), ops._p(gt_cls), ops._p(gt_off),
                    ops._p(wcd), ops._p(wld), ops._p(wrd), ops._p(present), ops._p(scratch), B, P, K, C.c_float(0.25),
                    C.c_float(2.0), C.c_float(1.0 if nosync else norm / gscale), C.c_float(w_reg), C.c_float(w_al), ops._p(dlogits),
                    ops._p(doffsets), ops._p(dwc), ops._p(dwl), ops._p(dwr), L.stream_ptr()), "vilco_mq_losses_bwd")
                if nosync:       # every output is linear in gscale / norm: applied from the device scalars
                    fac = gscale / norm
                    for t_ in (dlogits, doffsets, dwc, dwl, dwr):
                        t_.mul_(fac)
                logitsV.g, offsetsV.g = dlogits, doffsets
                model._last_head_grads = (dlogits, doffsets, pyr)   # kept for the gradient parity tests
            owned_p = [p_ for _, p_ in owned]
            owned_g = [None] * len(owned_p)

            def glue_grads(tensors, grads, leaves):
                """d(sum grads . tensors) / d(leaves + owned parameters + bias-layer parameters) without touching any .grad"""
                ins = list(leaves) + owned_p + bias_params
                out = torch.autograd.grad(tensors, ins, grads, allow_unused=True)
                for i in range(len(owned_p)):
                    g_ = out[len(leaves) + i]
                    if g_ is not None:
                        owned_g[i] = g_ if owned_g[i] is None else owned_g[i] + g_
                for q_, g_ in zip(bias_params, out[len(leaves) + len(owned_p):]):
                    if g_ is not None:       # BiC bias layers live in a plain python list outside the module tree
                        q_.grad = g_.clone() if q_.grad is None else q_.grad + g_
                return out[:len(leaves)]

            tensors, grads = list(extra), [torch.ones_like(t_) * gscale for t_ in extra]
            if lg_leaf is not None and lg_biased is not lg_leaf:
                tensors.append(lg_biased)        # through the bias layers back to the raw logits
                grads.append(dlogits)
            if tensors:
                leaves = ([lg_leaf] if lg_leaf is not None else []) + (list(ssl_leaves) if ssl_leaves is not None else [])
                lgr = glue_grads(tensors, grads, leaves)
                if lg_leaf is not None:
                    g_lg = lgr[0]
                    if lg_biased is not lg_leaf:
                        logitsV.g = g_lg
                    elif g_lg is not None:
                        logitsV.g = dlogits + g_lg
                    lgr = lgr[1:]
                if ssl_leaves is not None:
                    for f_, g_ in zip(fpn_lv, lgr):
                        if g_ is not None:
                            tp.acc(f_, g_)
            returned = {}
            with torch.no_grad():
                tp.backward()
                model._last_touch = (tp.n_nodes, {model._canon_key(k_): v_ for k_, v_ in tp.touch.items()},
                                     {model._canon_key(k_) for k_ in tp.G.keys()})
                for key, g in tp.G.items():
                    key = model._canon_key(key)          # pets.<i>.* -> the registered (de-duplicated) parameter name
                    prm = named.get(key)
                    if prm is None or not prm.requires_grad:
                        continue
                    g = TE.unpack_grad(key, g, prm).to(prm.dtype)
                    if direct:
                        if prm.grad is None:
                            prm.grad = g.clone()
                        else:
                            prm.grad.add_(g)  # in place: .grad is a view of the trainer's flat all-reduce buffer
                    elif key in passed:
                        returned[key] = g.contiguous()
                    elif prm.grad is None:    # reached for the first time after the live set was recorded: assign directly
                        prm.grad = g.clone()
                    else:
                        prm.grad.add_(g)
                model._live_keys = {model._canon_key(k_) for k_ in tp.G.keys()} | {model._canon_key(k_) for k_ in tp.touch}
            glue = [(t_, g_) for t_, g_ in ((wc, dwc), (wl, dwl), (wr, dwr)) if t_.requires_grad]
            if tin is not None and text.requires_grad and tin.g is not None:
                glue.append((text, ops.unpack(tin.g)))
            if glue:
                glue_grads([t_ for t_, _ in glue], [g_ for _, g_ in glue], [])
            with torch.no_grad():
                for (k_, p_), g_ in zip(owned, owned_g):
                    if g_ is None:
                        continue
                    if direct or k_ not in passed:
                        if p_.grad is None:
                            p_.grad = g_.clone()
                        else:
                            p_.grad.add_(g_)
                    else:
                        returned[k_] = g_.contiguous()
            return returned

        hook = torch.zeros((), device=dev, requires_grad=True)
        run_backward.nosync = nosync
        final_t = _TapeLoss.apply(hook, final, run_backward, [k for k, _ in live], *[p_ for _, p_ in live])
        out = {"cls_loss": cls_loss, "reg_loss": reg_loss, "al_loss": al_loss, "final_loss": final_t}
        out.update(extra_named)
        return out

    def _distill_term(self, lg, pyr, prev_out_cls_logits):
        """Distillation against the previous task's recorded classification outputs (per level (T_l, K_prev) numpy arrays),
        only batch element 0, as the reference computes it — BiC: soft targets at temperature 2 weighted n_known / K
        (meta_archs.py:1482-1499); iCaRL: per-class BCE-with-logits (:1501-1519)."""
        prev = prev_out_cls_logits
        nl = len(pyr.lens)
        dist = 0
        if self.cl_name == "bic":
            alpha = self.n_known / self.cls_head.cls_head.conv.out_channels
            for i, (o, n) in enumerate(zip(pyr.off, pyr.lens)):
                p_i = torch.as_tensor(prev[i]).to(lg.device)
                logp = F.log_softmax(lg[0, o:o + n, :self.n_known] / 2, dim=1)
                dist = dist + 0.01 * alpha * (-torch.mean(torch.sum(p_i[:, :self.n_known] * logp, dim=1)))
            return dist
        bce = nn.BCEWithLogitsLoss()
        for i, (o, n) in enumerate(zip(pyr.off, pyr.lens)):
            if len(prev) != nl or len(prev) == 1:
                prev = prev[0]
            p_i = torch.as_tensor(prev[i]).to(lg.device)
            dist = dist + 0.01 * sum(bce(lg[0, o:o + n, y], p_i[:, y]) for y in range(self.n_known))
        return dist

    def _narration_ssl(self, vl, fpn_lv, masks):
        """Narration self-supervision of mq_vilco (meta_archs.py:794-811, 939-945, 1351-1372): masked-mean narration
        embedding vs the masked-mean video embedding averaged over the pyramid levels, InfoNCE against a memory bank.
        Small torch expressions on leaves cut from the tape; returns (ssl_factor * loss or None, leaves)."""
        dev = self.device
        nf = [x["narration_feats"] for x in vl]
        lens = torch.as_tensor([f.shape[-1] for f in nf])
        nb = torch.zeros(len(nf), nf[0].shape[0], int(lens.max()))
        for i, f in enumerate(nf):
            nb[i, :, :f.shape[-1]].copy_(f)
        m0 = torch.tensor([float(x["narration_mask"]) for x in vl]).to(dev)
        m1 = (torch.arange(int(lens.max()))[None, :] < lens[:, None]).unsqueeze(1).to(dev)
        leaves = [f.v.detach().requires_grad_(True) for f in fpn_lv]
        if not bool(m0.sum() > 0):
            return None, leaves
        with torch.enable_grad():
            n = self.narration_encoder(nb.to(dev).permute(0, 2, 1)).permute(0, 2, 1) * m1          # (B,1024,Ln)
            cnt = m1.sum(dim=2, dtype=torch.float)
            cnt[cnt == 0.] = 1.
            n = F.normalize(n.sum(dim=2) / cnt, dim=1)
            vfs = []
            for leaf, mk in zip(leaves, masks):                                                      # leaf (B,T_l,C), mk (B,T_l)
                c_ = mk.sum(dim=1, keepdim=True)
                c_ = torch.where(c_ == 0, torch.ones_like(c_), c_)
                vfs.append((leaf * mk[:, :, None]).sum(dim=1) / c_)
            v = F.normalize(torch.stack(vfs).mean(dim=0), dim=1)
            sel = m0.to(torch.bool)
            self.memory_bank.update(n[sel])
            t_, v_ = n[sel], v[sel]
            pos = torch.einsum("nc,nc->n", [t_, v_]).unsqueeze(-1)
            mem = self.memory_bank.get_all()
            lt = torch.cat([pos, t_ @ mem.T], dim=1) / 0.07
            lv_ = torch.cat([pos, v_ @ mem.T], dim=1) / 0.07
            lab = torch.zeros(t_.size(0), dtype=torch.long, device=dev)
            ssl = (F.cross_entropy(lt, lab) + F.cross_entropy(lv_, lab)) / 2
        return self.ssl_factor * ssl, leaves

    # ---- targets + losses (reference: meta_archs.py:1224-1344, 1374-1524) --------------------------------
    def _label_points(self, pyr, gt_segments, gt_labels):
        """label_points / label_points_single_video (meta_archs.py:1224-1344) for the whole batch at once, laid out over the
        pyramid rows: the ground truths of every video are padded to the batch maximum (padding can never be selected) so
        the assignment is ~30 batched device ops instead of a Python loop per video.  Returns gt_cls (B,P,K), gt_off (B,P,2),
        and the three Gaussian weights (B,P), which carry autograd to mu / sigma."""
        dev = self.device
        K = self.num_classes
        pts = getattr(pyr, "pts", None)
        if pts is None or pts.device != dev:
            pts = torch.zeros(pyr.P, 4)
            for l, (o, n) in enumerate(zip(pyr.off, pyr.lens)):
                pts[o:o + n] = list(self.point_generator.buffer_points)[l][:n].cpu()
            pts = pyr.pts = pts.to(dev)
        B, G = len(gt_segments), max(int(s.shape[0]) for s in gt_segments)
        pin = dev.type == "cuda"      # pinned staging + non-blocking copies: a pageable upload would stall the host mid-step
        seg_h = torch.zeros(B, G, 2, pin_memory=pin)
        lab_h = torch.zeros(B, G, dtype=torch.long, pin_memory=pin)
        val_h = torch.zeros(B, G, dtype=torch.bool, pin_memory=pin)
        for i, (sg, lb) in enumerate(zip(gt_segments, gt_labels)):
            n = sg.shape[0]
            seg_h[i, :n], lab_h[i, :n], val_h[i, :n] = sg.float().cpu(), lb.cpu(), True
            seg_h[i, n:, 1] = 1.0                      # padded segments get length 1 (never selected; avoids 0/0)
        seg, lab, valid = seg_h.to(dev, non_blocking=True), lab_h.to(dev, non_blocking=True), val_h.to(dev, non_blocking=True)
        t, stride = pts[None, :, 0, None], pts[None, :, 3, None].clamp(min=1e-9)          # (1,P,1)
        s0, s1 = seg[:, None, :, 0], seg[:, None, :, 1]                                    # (B,1,G)
        lens = (s1 - s0).expand(B, pyr.P, G)
        left, right = t - s0, s1 - t                                                       # (B,P,G)
        xrel = ((right - left) / 2.0) / (stride * lens)
        if self.train_center_sample == "radius":
            center = 0.5 * (s0 + s1)
            t_mins = center - stride * self.train_center_sample_radius
            t_maxs = center + stride * self.train_center_sample_radius
            inside = torch.minimum(t - torch.maximum(t_mins, s0), torch.minimum(t_maxs, s1) - t) > 0
        else:
            inside = torch.minimum(left, right) > 0
        maxreg = torch.maximum(left, right)
        in_range = (maxreg >= pts[None, :, 1, None]) & (maxreg <= pts[None, :, 2, None])
        lens = lens.masked_fill(~(inside & in_range & valid[:, None, :]), float("inf"))
        min_len, min_inds = lens.min(dim=2)                                                # (B,P)
        mm = ((lens <= (min_len[..., None] + 1e-3)) & (lens < float("inf"))).to(left.dtype)
        gt_cls = torch.bmm(mm, F.one_hot(lab, K).to(left.dtype)).clamp(min=0.0, max=1.0)   # (B,P,K)
        idx = min_inds[..., None]
        gt_off = torch.cat((left.gather(2, idx), right.gather(2, idx)), dim=-1) / stride   # (B,P,2)
        xs = xrel.gather(2, idx).squeeze(-1)                                               # (B,P)
        ls = lab.gather(1, min_inds)                                                       # (B,P) label of the selected gt

        oh = F.one_hot(ls, K).to(xs.dtype)        # (B,P,K): the per-point parameter lookup as a matmul (its autograd is a
                                                  # matmul too; fancy-index backward is a sort + serial accumulate per call)
        def gauss(m, s):
            return (-(xs - (oh @ m).squeeze(-1)) ** 2 / (2 * (oh @ s).squeeze(-1) ** 2)).exp()
        return (gt_cls.contiguous(), gt_off.contiguous(), gauss(self.mu, self.sigma).contiguous(),
                gauss(self.mu_reg_left, self.sigma_reg_left).contiguous(),
                gauss(self.mu_reg_right, self.sigma_reg_right).contiguous())

    @torch.no_grad()
    def losses(self, vl, logits, offsets, pmask, pyr, prev_out_cls_logits=None):
        """Forward values of cls / reg / al / final loss (reference: meta_archs.py:1374-1524) from the fused loss kernel, for
        calls under `torch.no_grad()` (`validate_loss`, train_utils.py:584-655); with gradients enabled forward() takes
        `_train_forward`, whose result carries the hand-written backward."""
        dev = self.device
        B, P, K = logits.shape
        if self.train_label_smoothing > 0:
            raise NotImplementedError("train_cfg.label_smoothing > 0 is not supported by the fused loss kernels")
        gt_cls, gt_off, wc, wl, wr = self._label_points(pyr, [x["segments"] for x in vl], [x["labels"] for x in vl])
        present = torch.zeros(B, K)
        for i, x in enumerate(vl):
            present[i, x["labels"].cpu()] = 1
        present = present.to(dev)
        sums = torch.zeros(4, device=dev)
        scratch = torch.zeros(B * K, device=dev, dtype=torch.int32)
        L.check(L.lib().vilco_mq_losses(
            ops._p(logits), ops._p(offsets), ops._p(pmask), ops._p(pyr.gap_rows), ops._p(gt_cls), ops._p(gt_off),
            ops._p(wc), ops._p(wl), ops._p(wr), ops._p(present), B, P, K, C.c_float(0.25), C.c_float(2.0), ops._p(sums),
            ops._p(scratch), L.stream_ptr()), "vilco_mq_losses")
        s = sums.cpu()
        num_pos = float(s[2])
        self.loss_normalizer = self.loss_normalizer_momentum * self.loss_normalizer + \
            (1 - self.loss_normalizer_momentum) * max(num_pos, 1)
        cls_loss = sums[0] / self.loss_normalizer
        reg_loss = sums[1] / self.loss_normalizer
        al_loss = sums[3] / self.loss_normalizer if K != 1 else torch.zeros((), device=dev)
        loss_weight = self.train_loss_weight if self.train_loss_weight > 0 else float(cls_loss) / max(float(reg_loss), 0.01)
        final = cls_loss + reg_loss * loss_weight + al_loss * self.al_loss_weight
        if self.n_known > 0 and self.cl_name == "l2p" and getattr(self, "_reduce_sim", None) is not None:
            final = final - 0.1 * self._reduce_sim.detach()       # L2P pull constraint (meta_archs.py:1478-1480)
        out = {"cls_loss": cls_loss, "reg_loss": reg_loss, "al_loss": al_loss, "final_loss": final}
        if self.n_known > 0 and self.cl_name in ("bic", "icarl"):
            # distillation against the previous task's outputs (meta_archs.py:1482-1519); forward() already applied the BiC
            # bias layers to `logits`
            out["dist_loss"] = self._distill_term(logits, pyr, prev_out_cls_logits)
            out["final_loss"] = final + out["dist_loss"]
        return out

    # ---- inference (reference: meta_archs.py:1527-1736) ---------------------------------------------
    @torch.no_grad()
    def _decode_device(self, pyr, pmask, logits, offsets):
        """decode kernel (sigmoid * mask -> threshold -> top-k per level -> segments -> duration filter, meta_archs.py:1594-1692):
        candidates per (clip, level) region of `pre_nms_topk` slots, sorted by score inside a region.
        Returns (cand_segs (B, nl*topk, 2), cand_scores (B, nl*topk), cand_labels (B, nl*topk) i32, cand_count (B, nl) i32)."""
        B, P, K = logits.shape
        dev = logits.device
        nl, topk = len(pyr.lens), int(self.test_pre_nms_topk)
        cand_segs = torch.empty(B, nl * topk, 2, device=dev)
        cand_scores = torch.empty(B, nl * topk, device=dev)
        cand_labels = torch.empty(B, nl * topk, device=dev, dtype=torch.int32)
        cand_count = torch.zeros(B, nl, device=dev, dtype=torch.int32)
        IntArr, FltArr = C.c_int * nl, C.c_float * nl
        L.check(L.lib().vilco_decode(
            ops._p(logits), ops._p(offsets), ops._p(pmask), B, P, K, nl, IntArr(*pyr.off), IntArr(*pyr.lens),
            FltArr(*[float(s) for s in self.fpn_strides]), C.c_float(self.test_pre_nms_thresh),
            C.c_float(self.test_duration_thresh), topk, ops._p(cand_segs), ops._p(cand_scores), ops._p(cand_labels),
            ops._p(cand_count), L.stream_ptr()), "vilco_decode")
        return cand_segs, cand_scores, cand_labels, cand_count

    @torch.no_grad()
    def _decode_nms_device(self, pyr, pmask, logits, offsets):
        """decode + NMS kernels; returns device tensors (segs (B,M,2), scores (B,M), labels (B,M) i64, count (B,) i32)."""
        B, P, K = logits.shape
        dev = logits.device
        nl, topk = len(pyr.lens), int(self.test_pre_nms_topk)
        cand_segs, cand_scores, cand_labels, cand_count = self._decode_device(pyr, pmask, logits, offsets)
        if self.test_nms_method == "none":   # no NMS: every decoded candidate, level-major (meta_archs.py:1711)
            cnt = cand_count.cpu()
            M = int(cnt.sum(1).max())
            segs = torch.zeros(B, max(M, 1), 2, device=dev)
            scores = torch.zeros(B, max(M, 1), device=dev)
            labels = torch.zeros(B, max(M, 1), device=dev, dtype=torch.int64)
            for b in range(B):
                o = 0
                for l in range(nl):
                    n = int(cnt[b, l])
                    segs[b, o:o + n] = cand_segs[b, l * topk:l * topk + n]
                    scores[b, o:o + n] = cand_scores[b, l * topk:l * topk + n]
                    labels[b, o:o + n] = cand_labels[b, l * topk:l * topk + n].long()
                    o += n
            return segs, scores, labels, cnt.sum(1).int().to(dev)
        method = 2 if self.test_nms_method == "soft" else 3
        res = _nms_run(cand_segs, cand_scores, cand_labels, cand_count, B, nl, topk, K, self.test_multiclass_nms, method,
                       self.test_iou_threshold, self.test_nms_sigma, self.test_min_score, self.test_max_seg_num)
        if not self.test_multiclass_nms and self.test_voting_thresh > 0:      # class-agnostic: segment voting (nms.py:174-181)
            from ..utils.nms import seg_voting
            seg_voting(res[0], res[3], cand_segs, cand_scores, cand_count, B, nl, topk, self.test_max_seg_num,
                       self.test_voting_thresh)
        return res

    @staticmethod
    def _to_results(video_list, segs, scores, labels, count):
        """device->host results to the reference's output dicts; feature grid -> seconds, clamp to [0, duration]
        (meta_archs.py:1723-1734)."""
        results = []
        for i, v in enumerate(video_list):
            k = int(count[i])
            s, sc, lb = segs[i, :k].clone(), scores[i, :k].clone(), labels[i, :k].clone()
            if k > 0:
                s = (s * v["feat_stride"] + 0.5 * v["feat_num_frames"]) / v["fps"]
                s[s <= 0.0] *= 0.0
                s[s >= v["duration"]] = s[s >= v["duration"]] * 0.0 + v["duration"]
            results.append({"video_id": v["video_id"], "segments": s, "scores": sc, "labels": lb})
        return results

    @torch.no_grad()
    def inference(self, video_list, points_or_pyr, fpn_masks, out_cls_logits, out_offsets, out_lb_logits=None,
                  out_rb_logits=None, cilsettask=None):
        """decode (sigmoid / threshold / top-k / segments) + soft-NMS on the GPU, then the conversion to seconds on the
        host exactly like the reference (meta_archs.py:1723-1727).  Accepts the concatenated-pyramid tensors produced by
        forward(); returns one dict per video with CPU tensors."""
        pyr = points_or_pyr
        if not isinstance(pyr, E.Pyramid):
            # the reference's calling form (infer_one_epoch_ensemble, train_utils.py:957-961): per-level lists as returned by
            # forward(ensemble=True) — points [T_l, 4], masks (B, T_l), logits (B, T_l, K), offsets (B, T_l, 2), e.g. averaged
            # over several models — packed back into the concatenated pyramid layout the decode kernel reads
            pyr, fpn_masks, out_cls_logits, out_offsets = self._lists_to_pyramid(fpn_masks, out_cls_logits, out_offsets)
        segs, scores, labels, count = self._decode_nms_device(pyr, fpn_masks, out_cls_logits, out_offsets)
        results = self._to_results(video_list, segs.cpu(), scores.cpu(), labels.cpu(), count.cpu())
        # iCaRL (meta_archs.py:1559-1562): while `compute_means` is set and the validation passes its task object, a clip is
        # re-scored by the nearest exemplar mean.  classify() clears the flag itself, so like in the reference this happens
        # for the first clip after the flag was raised (train_utils.py:305) and the kernel path above serves all others.
        for idx, v in enumerate(video_list):
            if cilsettask is not None and self.compute_means:
                results[idx] = self._rescored_result(v, idx, self.classify(v, cilsettask), pyr, fpn_masks, out_cls_logits,
                                                     out_offsets)
        return results

    @staticmethod
    def _lists_to_pyramid(fpn_masks, out_cls_logits, out_offsets):
        """per-level lists -> (Pyramid, pmask (B, P) fp32, logits (B, P, K) fp32, offsets (B, P, 2) fp32); gap rows stay zero
        (mask 0: never a candidate)."""
        dev = out_cls_logits[0].device
        B, K = out_cls_logits[0].shape[0], out_cls_logits[0].shape[2]
        pyr = E.Pyramid([t.shape[1] for t in out_cls_logits], dev)
        logits = torch.zeros(B, pyr.P, K, device=dev, dtype=torch.float32)
        offsets = torch.zeros(B, pyr.P, 2, device=dev, dtype=torch.float32)
        pmask = torch.zeros(B, pyr.P, device=dev, dtype=torch.float32)
        for o, n, lg, of, mk in zip(pyr.off, pyr.lens, out_cls_logits, out_offsets, fpn_masks):
            logits[:, o:o + n] = lg
            offsets[:, o:o + n] = of
            pmask[:, o:o + n] = mk.reshape(B, n).to(torch.float32)
        return pyr, pmask, logits, offsets

    def _fpn_features(self, video_list):
        """FPN outputs of `neck(backbone(...))` per level in the reference layout (B, C, T_l), fp32 — the quantity
        `classify` works on (meta_archs.py:1073-1079, 1110-1118: no prompts, no EMA ensemble)."""
        _, batched, mask = self.preprocessing(video_list, is_training=False)
        text = tmask = tlens = t16 = None
        if self.use_cross_modal:
            text, tmask, tlens = self.query_preprocessing(video_list)
            t16 = ops.pack_feats(text.contiguous())
        W, cfg = self.packed_weights(), self.engine_cfg()
        trunk = E.backbone_fwd(W, cfg, ops.pack_feats(batched, planes=ops.PLANES_HI), mask.contiguous(), t16, tmask, self._pe, text_lens=tlens,
                               trunk_only=True)
        feats, _ = E.branch_fwd(W, cfg, trunk, "pets.")
        out = []
        for l, f in enumerate(feats):
            y32, _ = ops.layernorm(f, W[f"neck.fpn_norms.{l}.weight"], W[f"neck.fpn_norms.{l}.bias"], out32=True, out16=False)
            out.append(y32.permute(0, 2, 1))
        return out

    @torch.no_grad()
    def classify(self, x, cilsettask):
        """iCaRL nearest-mean-of-exemplars distances of one clip — meta_archs.py:1061-1131.  Returns, per pyramid level, the
        (1, T_l, n_classes) squared distances between the clip's normalised FPN features and the class means of the exemplar
        memory; the means are (re)computed from `self.memory` through `cilsettask.get_dataloader` while `compute_means` is
        set.  The network passes run on the CUDA path, the reductions are `modeling/icarl.py`."""
        from . import icarl
        if self.compute_means:
            print("Computing mean of exemplars...")
            exemplar_means = [[] for _ in range(icarl.FPN_LEVELS)]
            for class_id, videos in self.memory.items():
                per_level = None
                for video_list in cilsettask.get_dataloader({class_id: videos}, sample_frame=True):
                    lv = [icarl.normalize_level(f) for f in self._fpn_features(video_list)]
                    if per_level is None:
                        per_level = [[f] for f in lv]
                    else:
                        for i, f in enumerate(lv):
                            per_level[i].append(f)
                for i, fs in enumerate(per_level):
                    exemplar_means[i].append(icarl.exemplar_mean(fs))
            self.exemplar_means = exemplar_means
            self.compute_means = False
        feats = self._fpn_features([x])
        return [icarl.nme_dists(feats[i], self.exemplar_means[i]) for i in range(icarl.FPN_LEVELS)]

    def _rescored_result(self, v, idx, dists, pyr, pmask, logits, offsets):
        """`inference_single_video` with `cls_preds_per_vid` (meta_archs.py:1625-1682) + `postprocessing` (:1695-1736) for
        one clip: candidate selection by exemplar distance (torch glue on the device), soft-NMS on the GPU kernel."""
        from . import icarl
        from ..utils.nms import batched_nms
        points = self.point_generator(pyr.lens)
        segs, scores, labels = [], [], []
        for l, (o, n) in enumerate(zip(pyr.off, pyr.lens)):
            s, sc, lb = icarl.select_candidates(logits[idx, o:o + n], offsets[idx, o:o + n], points[l].to(logits.device),
                                                pmask[idx, o:o + n], dists[l], self.num_classes, int(self.test_pre_nms_topk),
                                                self.test_duration_thresh)
            segs.append(s), scores.append(sc), labels.append(lb)
        segs, scores, labels = torch.cat(segs), torch.cat(scores), torch.cat(labels)
        if self.test_nms_method != "none":
            segs, scores, labels = batched_nms(segs, scores, labels, self.test_iou_threshold, self.test_min_score,
                                               self.test_max_seg_num, use_soft_nms=(self.test_nms_method == "soft"),
                                               multiclass=self.test_multiclass_nms, sigma=self.test_nms_sigma,
                                               voting_thresh=self.test_voting_thresh)
        n = segs.shape[0]
        return self._to_results([v], segs.cpu()[None], scores.cpu()[None], labels.cpu()[None], torch.tensor([n]))[0]


class EvalGraph:
    """One captured CUDA graph of the evaluation step for a fixed batch size (every clip padded to max_seq_len, every
    text padded to `text_len` with per-clip lengths, so shapes are static).  Inputs are written into static device
    buffers; outputs are read from static device buffers."""

    def __init__(self, model, batch_size, text_len=128):
        self.model, self.B, self.Lt = model, batch_size, text_len
        if hasattr(model, "prompt"):
            raise NotImplementedError("EvalGraph: the host-side L2P prompt selection is not capturable; use model(video_list)")
        dev = model.device
        T, Cin, Ct = model.max_seq_len, model.input_dim, model.n_txt_in
        self.feats = torch.zeros(batch_size, Cin, T, device=dev)
        self.mask = torch.ones(batch_size, T, device=dev)
        self.text = torch.zeros(batch_size, Ct, text_len, device=dev) if model.use_cross_modal else None
        self.tmask = torch.ones(batch_size, text_len, device=dev) if model.use_cross_modal else None
        self.tlens = torch.full((batch_size,), text_len, device=dev, dtype=torch.int32) if model.use_cross_modal else None
        self.launches = 0
        self._capture()
        self._stage_f = torch.zeros(batch_size, Cin, T).pin_memory()
        self._stage_t = torch.zeros(batch_size, Ct, text_len).pin_memory() if model.use_cross_modal else None

    def _weights_sig(self):
        m = self.model
        W = m.packed_weights()
        return (id(W), m._packed_key, getattr(m, "_packed_epoch", 0), getattr(m, "_packed_ema_key", None),
                m.n_known, tuple(m.list_splits), tuple(id(b) for b in m.list_bias_layers))

    def _capture(self):
        """(Re)capture the graph against the model's current packed weights."""
        self._sig = self._weights_sig()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):   # warm-up outside capture (lazy cudaFuncSetAttribute, allocator pools)
            self._step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = L.launch_count()
        with torch.cuda.graph(self.graph):
            self.out = self._step()
        self.launches = L.launch_count() - n0

    def refresh(self):
        """The captured graph reads the packed weight tensors that existed at capture time: after an optimizer step,
        load_state_dict or a precision switch some of them are new tensors, so the graph is captured again."""
        if self._weights_sig() != self._sig:
            self._capture()

    def _step(self):
        m = self.model
        logits, offsets, pmask, pyr = m._device_forward(self.feats, self.mask, self.text, self.tmask, self.tlens, False)
        logits = m._apply_bias_layers(logits)      # BiC (captured torch ops reading the live alpha / beta parameters)
        return m._decode_nms_device(pyr, pmask, logits, offsets)

    @staticmethod
    def _raw_clips(video_list):
        """the clips' features as stored on disk — `feats_raw` (T_in, C), SURVEY.md §8f-2 — when every clip carries them:
        the force_upsampling resize of the dataset (ego4d.py:644-651) then runs on the device (vilco_b200/data.py) and
        `feats` is not needed."""
        raw = [v.get("feats_raw") for v in video_list]
        if all(r is None for r in raw):
            return None
        if any(r is None for r in raw):
            raise ValueError("EvalGraph: either every clip of a batch carries 'feats_raw' or none does")
        return [r if (r.is_pinned() or r.is_cuda) else r.pin_memory() for r in raw]

    def load_inputs(self, video_list):
        """host -> static device buffers (async on the current stream)."""
        m = self.model
        assert len(video_list) == self.B
        T = m.max_seq_len
        raw = self._raw_clips(video_list)
        lens = [T] * self.B if raw is not None else [v["feats"].shape[-1] for v in video_list]
        assert max(lens) <= T
        if raw is not None:
            from .. import data
            data.resize_feats_into(raw, self.feats)
        for i, v in enumerate(video_list if raw is None else ()):
            f = v["feats"]
            if f.is_pinned() or f.is_cuda:
                if lens[i] < T:
                    self.feats[i, :, lens[i]:].zero_()
                self.feats[i, :, :lens[i]].copy_(f, non_blocking=True)
            else:
                self._stage_f[i].zero_()
                self._stage_f[i, :, :lens[i]].copy_(f)
                self.feats[i].copy_(self._stage_f[i], non_blocking=True)
        self.mask.copy_((torch.arange(T)[None, :] < torch.as_tensor(lens)[:, None]).float(), non_blocking=True)
        if m.use_cross_modal:
            tl = [v["prompt_feature"].shape[-1] for v in video_list]
            assert max(tl) <= self.Lt
            self._stage_t.zero_()
            for i, v in enumerate(video_list):
                self._stage_t[i, :, :tl[i]].copy_(v["prompt_feature"])
            self.text.copy_(self._stage_t, non_blocking=True)
            self.tmask.copy_((torch.arange(self.Lt)[None, :] < torch.as_tensor(tl)[:, None]).float(), non_blocking=True)
            self.tlens.copy_(torch.as_tensor(tl, dtype=torch.int32), non_blocking=True)

    def replay(self):
        self.graph.replay()

    # ---- pipelined evaluation: the H2D copy of batch i+1 overlaps the graph replay of batch i ------------------------
    def _host_to(self, dst, video_list, stream):
        """pinned host -> device staging set `dst` on `stream` (async)."""
        m, T = self.model, self.model.max_seq_len
        raw = self._raw_clips(video_list)
        lens = [T] * self.B if raw is not None else [v["feats"].shape[-1] for v in video_list]
        with torch.cuda.stream(stream):
            if raw is not None:
                from .. import data
                data.resize_feats_into(raw, dst["feats"])
            for i, v in enumerate(video_list if raw is None else ()):
                f = v["feats"] if (v["feats"].is_pinned() or v["feats"].is_cuda) else v["feats"].pin_memory()
                if lens[i] < T:
                    dst["feats"][i, :, lens[i]:].zero_()
                dst["feats"][i, :, :lens[i]].copy_(f, non_blocking=True)
            dst["mask_h"].copy_((torch.arange(T)[None, :] < torch.as_tensor(lens)[:, None]).float())
            dst["mask"].copy_(dst["mask_h"], non_blocking=True)
            if m.use_cross_modal:
                tl = [v["prompt_feature"].shape[-1] for v in video_list]
                dst["text_h"].zero_()
                for i, v in enumerate(video_list):
                    dst["text_h"][i, :, :tl[i]].copy_(v["prompt_feature"])
                dst["tmask_h"].copy_((torch.arange(self.Lt)[None, :] < torch.as_tensor(tl)[:, None]).float())
                dst["tlens_h"].copy_(torch.as_tensor(tl, dtype=torch.int32))
                dst["text"].copy_(dst["text_h"], non_blocking=True)
                dst["tmask"].copy_(dst["tmask_h"], non_blocking=True)
                dst["tlens"].copy_(dst["tlens_h"], non_blocking=True)

    def _new_slot(self):
        m, dev = self.model, self.model.device
        d = {"feats": torch.empty_like(self.feats), "mask": torch.empty_like(self.mask),
             "mask_h": torch.empty(self.mask.shape).pin_memory(), "ready": torch.cuda.Event(), "free": torch.cuda.Event(),
             "done": torch.cuda.Event()}
        if m.use_cross_modal:
            d.update(text=torch.empty_like(self.text), tmask=torch.empty_like(self.tmask), tlens=torch.empty_like(self.tlens),
                     text_h=torch.empty(self.text.shape).pin_memory(), tmask_h=torch.empty(self.tmask.shape).pin_memory(),
                     tlens_h=torch.empty(self.tlens.shape, dtype=torch.int32).pin_memory())
        d["out_h"] = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in self.out]
        return d

    def infer_stream(self, batches):
        """Generator over result lists, one per batch of `batches` (an iterable of video_list).  Double-buffered:
        while the graph of batch i runs, the inputs of batch i+1 are copied host->device on a side stream."""
        self.refresh()
        if not hasattr(self, "_slots"):
            self._slots = [self._new_slot(), self._new_slot()]
            self._copy_stream = torch.cuda.Stream()
        comp = torch.cuda.current_stream()
        pending = []  # (slot, video_list)
        it = iter(batches)
        k = 0

        def submit(vl, slot):
            self._copy_stream.wait_event(slot["free"])
            self._host_to(slot, vl, self._copy_stream)
            slot["ready"].record(self._copy_stream)
            comp.wait_event(slot["ready"])
            self.feats.copy_(slot["feats"], non_blocking=True)
            self.mask.copy_(slot["mask"], non_blocking=True)
            if self.model.use_cross_modal:
                self.text.copy_(slot["text"], non_blocking=True)
                self.tmask.copy_(slot["tmask"], non_blocking=True)
                self.tlens.copy_(slot["tlens"], non_blocking=True)
            slot["free"].record(comp)
            self.graph.replay()
            for h, t in zip(slot["out_h"], self.out):
                h.copy_(t, non_blocking=True)
            slot["done"].record(comp)

        for s_ in self._slots:
            s_["free"].record(comp)
        for vl in it:
            slot = self._slots[k % 2]
            if len(pending) == 2:
                ps, pvl = pending.pop(0)
                ps["done"].synchronize()
                yield PtTransformer._to_results(pvl, *ps["out_h"])
            submit(vl, slot)
            pending.append((slot, vl))
            k += 1
        for ps, pvl in pending:
            ps["done"].synchronize()
            yield PtTransformer._to_results(pvl, *ps["out_h"])

    def run(self, video_list):
        self.refresh()
        self.load_inputs(video_list)
        self.graph.replay()
        segs, scores, labels, count = self.out
        return PtTransformer._to_results(video_list, segs.cpu(), scores.cpu(), labels.cpu(), count.cpu())
