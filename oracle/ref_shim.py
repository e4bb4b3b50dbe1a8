"""TEST INFRASTRUCTURE ONLY — import shim for the *Python reference* (ViLCo MQ) living at /root/reference.

Used by oracle/gen_golden.py and by the pin tests that run only in the authoring container
(`/root/reference` does not exist on the GPU box).  Nothing in the product path imports this.

The shim follows SURVEY.md Appendix B: stub `timm` (only ModelEmaV2 is needed), stub `turtle`
(stray import in MQ/libs/modeling/utils.py:25), add the helper classes transformers>=5 dropped from
`modeling_utils` (MQ/libs/modeling/modeling_xlnet_x.py:28-35), import `libs.utils` first to break the
modeling<->utils cycle, chdir to MQ/ (xlnet json is opened cwd-relative, MQ/libs/modeling/backbones.py:132),
and provide `nms_1d_cpu` built from the reference's own C++ (oracle/build_ref.py).
"""
import copy
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
# the reference's MQ tree: /root/reference in the authoring container; on the GPU box the un-modified copy that
# baseline/vendor_ref.py placed under baseline/_ref (git-ignored, travels with the gpurun snapshot) — used there only by
# bench.py's reference / baseline legs
VENDORED = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")
REF_ROOT = os.environ.get("VILCO_REFERENCE") or ("/root/reference" if os.path.isdir("/root/reference/MQ") else VENDORED)
REF_MQ = os.path.join(REF_ROOT, "MQ")


def available():
    return os.path.isdir(os.path.join(REF_MQ, "libs", "modeling"))


_loaded = None


def load():
    """Returns a namespace with the reference's `make_meta_arch`, `load_config`, `DEFAULTS`, `blocks`, ... ."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference not present at " + REF_MQ)
    import torch
    import transformers  # noqa: F401  (import before stubbing timm)
    import transformers.modeling_utils as mu
    from transformers.pytorch_utils import apply_chunking_to_forward
    for n in ["PoolerAnswerClass", "PoolerEndLogits", "PoolerStartLogits", "SequenceSummary"]:
        if not hasattr(mu, n):
            setattr(mu, n, type(n, (torch.nn.Module,), {}))
    if not hasattr(mu, "apply_chunking_to_forward"):
        mu.apply_chunking_to_forward = apply_chunking_to_forward
    t = types.ModuleType("turtle")
    t.forward = lambda *a, **k: None
    sys.modules.setdefault("turtle", t)

    class ModelEmaV2(torch.nn.Module):  # timm-equivalent EMA over the state_dict
        def __init__(self, model, decay=0.9999, device=None):
            super().__init__()
            self.module = copy.deepcopy(model).eval()
            self.decay = decay

        def update(self, model):
            with torch.no_grad():
                for e, m in zip(self.module.state_dict().values(), model.state_dict().values()):
                    e.copy_(self.decay * e + (1 - self.decay) * m)

    for name in ("timm", "timm.utils", "timm.utils.model_ema"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["timm.utils.model_ema"].ModelEmaV2 = ModelEmaV2

    from . import build_ref
    so_dir = build_ref.build_nms()
    sys.path[:0] = [so_dir, REF_MQ]
    cwd = os.getcwd()
    os.chdir(REF_MQ)
    try:
        import libs.utils  # noqa: F401  FIRST: breaks the modeling<->utils cycle
        from libs.core.config import DEFAULTS, load_config
        from libs.modeling import make_meta_arch
        import libs.modeling.blocks as blocks
        import libs.modeling.meta_archs as meta_archs
        import libs.modeling.backbones as backbones
        import libs.modeling.necks as necks
        import libs.modeling.losses as losses
        import libs.utils.nms as nms
        import nms_1d_cpu
    finally:
        os.chdir(cwd)
    ns = types.SimpleNamespace(DEFAULTS=DEFAULTS, load_config=load_config, make_meta_arch=make_meta_arch,
                               blocks=blocks, meta_archs=meta_archs, backbones=backbones, necks=necks,
                               losses=losses, nms=nms, nms_1d_cpu=nms_1d_cpu, REF_MQ=REF_MQ)
    _loaded = ns
    return ns


def build_model(cfg_overrides=None, yaml_name="mq_no_cl.yaml"):
    """Instantiate the reference PtTransformer from one of its yaml configs with overrides applied to cfg."""
    ns = load()
    cwd = os.getcwd()
    os.chdir(REF_MQ)
    try:
        cfg = ns.load_config(os.path.join(REF_MQ, "configs", yaml_name), defaults=copy.deepcopy(ns.DEFAULTS))
        if cfg_overrides:
            cfg_overrides(cfg)
            # re-derive the fields load_config copies into cfg['model'] (MQ/libs/core/config.py:189-197)
            cfg["model"]["input_dim"] = cfg["dataset"]["input_dim"]
            cfg["model"]["num_classes"] = cfg["dataset"]["num_classes"]
            cfg["model"]["max_seq_len"] = cfg["dataset"]["max_seq_len"]
        model = ns.make_meta_arch(cfg["model_name"], **cfg["model"])
    finally:
        os.chdir(cwd)
    return model, cfg
