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


def test_launcher_runs_an_unmodified_style_script_on_the_mirrors(tmp_path, capsys):
    """python -m vilco_b200.run <script>: a script written like the reference's entry points (module-level
    `from libs.modeling import make_meta_arch`, `from libs.utils import ...`, argv of its own) sees the rebound names."""
    import sys
    from oracle import ref_shim
    import vilco_b200.run as launcher
    ns = ref_shim.load()                                   # this container's missing-dependency stubs (timm, turtle, ...)
    script = tmp_path / "entry_like_eval.py"
    script.write_text(
        "import sys\n"
        "from libs.modeling import make_meta_arch\n"
        "from libs.utils import batched_nms, ANETdetection\n"
        "import libs.utils.train_utils as tu\n"
        "print('ARGV', sys.argv[1:])\n"
        "print('BOUND', make_meta_arch.__module__, batched_nms.__module__, ANETdetection.__module__, tu.LayerNorm.__module__)\n"
        "print('MAIN', __name__)\n")
    import libs.modeling as lm
    import libs.utils as lu
    import libs.utils.train_utils as tu
    import libs.utils.metrics as lmet
    import libs.utils.get_retrieval_performance as lret
    mods = [m for m in (lm, lu, tu, lmet, lret, sys.modules.get("libs.modeling.models"), sys.modules.get("libs.modeling.meta_archs"),
                        sys.modules.get("libs.utils.nms"), sys.modules.get("libs.modeling.modeling_xlnet_x")) if m is not None]
    saved = [(m, dict(vars(m))) for m in mods]
    cwd, argv, path = os.getcwd(), list(sys.argv), list(sys.path)
    try:
        launcher.main(["--mq-root", ns.REF_MQ, str(script), "configs/mq_vilco.yaml", "--topk", "5"])
        out = capsys.readouterr().out
        assert "ARGV ['configs/mq_vilco.yaml', '--topk', '5']" in out and "MAIN __main__" in out
        assert "BOUND vilco_b200.modeling.models vilco_b200.utils.nms vilco_b200.utils.metrics vilco_b200.modeling.blocks" in out
        assert os.getcwd() == ns.REF_MQ
        with pytest.raises(SystemExit, match="not a ViLCo/MQ checkout"):
            launcher.main(["--mq-root", str(tmp_path), "eval.py"])
    finally:
        os.chdir(cwd)
        sys.argv[:], sys.path[:] = argv, path
        for m, d in saved:                                 # undo compat.install() for the other tests of this process
            for k, v in d.items():
                setattr(m, k, v)


def test_replay_memory_update_matches_the_reference_under_the_same_seed(capsys):
    """model.add_samples_to_mem(cilsettask, data, m) (train_cl.py:353): same exemplars as the reference's method for the same
    `random` seed, for a numeric budget and for 'ALL'."""
    import copy
    import random
    import types
    from oracle import ref_shim
    from vilco_b200.modeling.meta_archs import PtTransformer
    ns = ref_shim.load()
    ref_fn = ns.meta_archs.PtTransformer.add_samples_to_mem
    old = {3: [f"old3_{i}" for i in range(7)], 5: [f"old5_{i}" for i in range(4)]}
    new = {5: [f"new5_{i}" for i in range(6)], 8: [f"new8_{i}" for i in range(9)], 9: ["only"]}
    for m in (4, 'ALL'):
        a, b = types.SimpleNamespace(memory=copy.deepcopy(old)), types.SimpleNamespace(memory=copy.deepcopy(old))
        random.seed(11)
        ref_fn(a, None, copy.deepcopy(new), m)
        state = random.getstate()
        random.seed(11)
        PtTransformer.add_samples_to_mem(b, None, copy.deepcopy(new), m)
        assert random.getstate() == state and a.memory == b.memory, m
        assert list(b.memory) == [3, 5, 8, 9] and (m == 'ALL' or all(len(v) <= m for v in b.memory.values()))
    assert "Memory... Class: 8, num videos: 4" in capsys.readouterr().out
