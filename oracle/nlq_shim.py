"""TEST INFRASTRUCTURE ONLY — import shim for the NLQ half of the reference (`/root/reference/NLQ`), verified
in the authoring container.  Like oracle/ref_shim.py it only papers over packages this container lacks; nothing of the
reference is modified or copied, and nothing in the product path imports this.

The NLQ package is also called `libs`, so it cannot share a process with the MQ reference: use it from its own process
(`python -m oracle.gen_golden_nlq`).

Stubs (each names the import that needs it):
* `timm.utils.model_ema.ModelEmaV2` (NLQ/libs/modeling/meta_archs.py:15), `timm.models.layers.{DropPath, to_2tuple, trunc_normal_}`
  (video_transformer.py:26) — `transformers` must be imported BEFORE `timm` is stubbed (its availability probe trips over a
  spec-less module);
* `transformers.modeling_utils.{find_pruneable_heads_and_indices, prune_linear_layer, apply_chunking_to_forward}` and
  `transformers.file_utils.add_*_docstrings` (roberta.py:36-56) moved to `transformers.pytorch_utils` / `transformers.utils`;
* `terminaltables` (NLQ/libs/utils/metrics.py:4);
* `nms_1d_cpu` = the reference's own C++ NMS built by oracle/build_ref.py (NLQ/libs/utils/nms.py:5 — same source as MQ's);
* import `libs.utils` before `libs.modeling` and run with cwd = NLQ/ (relative config paths).
"""
import copy
import os
import sys
import types

REF_NLQ = os.path.join(os.environ.get("VILCO_REFERENCE", "/root/reference"), "NLQ")
_loaded = None


def available():
    return os.path.isdir(os.path.join(REF_NLQ, "libs", "modeling"))


def load():
    """namespace with the NLQ reference's make_meta_arch / load_config / DEFAULTS / modeling package."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("NLQ reference not present at " + REF_NLQ)
    if "libs" in sys.modules and not os.path.abspath(sys.modules["libs"].__path__[0]).startswith(REF_NLQ):
        raise RuntimeError("another `libs` package (the MQ reference) is already imported in this process")
    import torch
    import transformers  # noqa: F401  before the timm stub
    from transformers import AutoModel, RobertaConfig  # noqa: F401  force the lazy modules while timm is still absent
    import transformers.file_utils as fu
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    import transformers.utils as tu

    class ModelEmaV2(torch.nn.Module):
        def __init__(self, model, decay=0.9999, device=None):
            super().__init__()
            self.module = copy.deepcopy(model).eval()
            self.decay = decay

    class DropPath(torch.nn.Module):
        def __init__(self, p=0.0):
            super().__init__()
            self.p = p

        def forward(self, x):
            return x

    for name in ("timm", "timm.utils", "timm.utils.model_ema", "timm.models", "timm.models.layers"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["timm.utils.model_ema"].ModelEmaV2 = ModelEmaV2
    tl = sys.modules["timm.models.layers"]
    tl.DropPath, tl.trunc_normal_ = DropPath, torch.nn.init.trunc_normal_
    tl.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)
    for n in ("find_pruneable_heads_and_indices", "prune_linear_layer", "apply_chunking_to_forward"):
        if not hasattr(mu, n):
            setattr(mu, n, getattr(pu, n, None) or (lambda *a, **k: None))    # head pruning is never called on the path
    for n in ("add_code_sample_docstrings", "add_start_docstrings", "add_start_docstrings_to_model_forward",
              "replace_return_docstrings"):
        if not hasattr(fu, n):
            setattr(fu, n, getattr(tu, n, None) or (lambda *a, **k: (lambda f: f)))
    tt = sys.modules.setdefault("terminaltables", types.ModuleType("terminaltables"))
    if not hasattr(tt, "AsciiTable"):
        tt.AsciiTable = object
    from . import build_ref
    sys.path[:0] = [build_ref.build_nms(), REF_NLQ]
    cwd = os.getcwd()
    os.chdir(REF_NLQ)
    try:
        import libs.utils  # noqa: F401  first (modeling <-> utils cycle)
        import libs.modeling as modeling
        from libs.core.config import DEFAULTS, load_config
    finally:
        os.chdir(cwd)
    _loaded = types.SimpleNamespace(modeling=modeling, make_meta_arch=modeling.make_meta_arch, DEFAULTS=DEFAULTS,
                                    load_config=load_config, REF_NLQ=REF_NLQ)
    return _loaded


def build_model(overrides=None, yaml_name="ego4d_nlq_v2_egovlp_1e-4.yaml"):
    """the reference's LocPointTransformer for NLQ from its own yaml (+ overrides applied to the dataset / model sections)."""
    ns = load()
    cwd = os.getcwd()
    os.chdir(REF_NLQ)
    try:
        cfg = ns.load_config(os.path.join(REF_NLQ, "configs", yaml_name), defaults=copy.deepcopy(ns.DEFAULTS))
        if overrides:
            overrides(cfg)
            for k in ("input_vid_dim", "input_txt_dim", "num_classes", "max_seq_len"):
                cfg["model"][k] = cfg["dataset"][k]
        model = ns.make_meta_arch(cfg["model_name"], **cfg["model"])
        # the model's `device` property is hard-wired to LOCAL_RANK / cuda:0 (meta_archs.py:562-567): CPU for the pin runs
        import torch
        type(model).device = property(lambda self: torch.device("cpu"))
    finally:
        os.chdir(cwd)
    return model, cfg
