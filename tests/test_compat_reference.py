"""Drop-in check against the UNMODIFIED reference checkout (authoring container only; skipped where /root/reference does not
exist, e.g. on the GPU box): after vilco_b200.compat.install() the reference's own `make_optimizer` (train_utils.py:68-143)
groups the parameters of OUR model exactly as it groups its own (its isinstance / name-substring rules still apply), its
config loader feeds OUR make_meta_arch, and the state_dict keys / shapes coincide."""
import os

import pytest

REF = "/root/reference/MQ"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not available")


def test_install_makes_the_reference_callers_use_vilco_b200():
    import json
    from conftest import GOLDEN
    from oracle import gen_golden as GG
    from oracle import ref_shim
    ns = ref_shim.load()                       # imports libs.utils / libs.modeling of the reference (CPU)
    import libs.modeling as lm
    import libs.utils as lu
    import libs.utils.train_utils as tu
    ref_make, ref_nms = lm.make_meta_arch, lu.batched_nms
    import libs.utils.metrics as lmet
    import libs.utils.get_retrieval_performance as lret
    saved = {(m, n): getattr(m, n) for m in (lm, lu, tu, lmet, lret, lm.meta_archs if hasattr(lm, "meta_archs") else lm)
             for n in dir(m) if n in ("ANETdetection", "evaluation_retrieval", "Moment_Retrieval",
                                      "compute_average_precision_detection","make_meta_arch", "make_backbone", "make_neck", "make_generator", "MaskedConv1D",
                                      "MaskedMHCA", "MaskedMHA", "LayerNorm", "TransformerBlock", "Scale", "AffineDropPath",
                                      "BiasLayer", "batched_nms", "XLNetModel", "XLNetLMHeadModel")}
    import vilco_b200.compat as compat
    import vilco_b200.modeling as M
    try:
        patched = compat.install()
        assert ("libs.modeling", "make_meta_arch") in patched and ("libs.utils", "batched_nms") in patched
        assert lm.make_meta_arch is M.make_meta_arch and lm.make_meta_arch is not ref_make
        assert lu.batched_nms is not ref_nms and tu.MaskedConv1D is M.MaskedConv1D and tu.LayerNorm is M.LayerNorm
        import vilco_b200.utils.metrics as E
        assert lu.ANETdetection is E.ANETdetection and tu.ANETdetection is E.ANETdetection
        assert tu.evaluation_retrieval.__module__ == "vilco_b200.utils.get_retrieval_performance"
        # the reference's config loader + OUR factory, the reference's make_optimizer on OUR model
        c = GG.small_cfg()
        cwd = os.getcwd()
        os.chdir(REF)
        try:
            import copy
            cfg = ns.load_config(os.path.join(REF, "configs", "mq_no_cl.yaml"), defaults=copy.deepcopy(ns.DEFAULTS))
            GG._override(c)(cfg)
            cfg["model"]["input_dim"] = cfg["dataset"]["input_dim"]
            cfg["model"]["num_classes"] = cfg["dataset"]["num_classes"]
            cfg["model"]["max_seq_len"] = cfg["dataset"]["max_seq_len"]
            model = lm.make_meta_arch(cfg["model_name"], **cfg["model"])
        finally:
            os.chdir(cwd)
        assert type(model).__module__.startswith("vilco_b200.")
        opt = tu.make_optimizer(model, {"type": "AdamW", "learning_rate": 1e-4, "weight_decay": 0.05, "momentum": 0.9})
        names = {id(p): k for k, p in model.named_parameters()}
        mine = [sorted(names[id(p)] for p in g["params"]) for g in opt.param_groups]
        gold = json.load(open(os.path.join(GOLDEN, "optimizer_groups.json")))["small"]
        assert mine == [g["params"] for g in gold]
    finally:
        for (m, n), v in saved.items():
            setattr(m, n, v)
