"""One-call drop-in for an UNMODIFIED ViLCo/MQ checkout.

    import libs.utils, libs.modeling           # the reference packages (cwd = ViLCo/MQ), imported as train.py does
    import vilco_b200.compat
    vilco_b200.compat.install()                # from here on make_meta_arch / batched_nms are the sm_100a implementations

`install()` rebinds, in the already imported reference modules, exactly the names of the drop-in boundary:
`libs.modeling.{make_meta_arch, make_backbone, make_neck, make_generator, MaskedConv1D, MaskedMHCA, MaskedMHA, LayerNorm,
TransformerBlock, Scale, AffineDropPath, BiasLayer}`, `libs.modeling.modeling_xlnet_x.{XLNetModel, XLNetLMHeadModel}`,
`libs.utils.batched_nms` / `libs.utils.nms.batched_nms`, the evaluation tail `libs.utils.ANETdetection` /
`libs.utils.get_retrieval_performance.evaluation_retrieval`, and the copies of those names that `libs.utils.train_utils` and
`libs.modeling.meta_archs` took at import time (`from ..modeling import MaskedConv1D, ...` — the isinstance keys of
`make_optimizer`, train_utils.py:19-20, 29, 76-77, 96).  Everything else of the reference (config loader, datasets,
schedulers, the CL orchestration, the validation loops) stays in place and runs unchanged.
"""
import sys

_NAMES = ("make_meta_arch", "make_backbone", "make_neck", "make_generator", "MaskedConv1D", "MaskedMHCA", "MaskedMHA",
          "LayerNorm", "TransformerBlock", "Scale", "AffineDropPath", "BiasLayer")


def install(libs_modeling=None, libs_utils=None):
    """Rebind the reference's plugin-level names to vilco_b200.  Returns the list of (module, name) pairs that were patched."""
    from . import modeling as M
    from .modeling import modeling_xlnet_x as X
    from .utils import nms as N
    lm = libs_modeling or sys.modules.get("libs.modeling")
    lu = libs_utils or sys.modules.get("libs.utils")
    if lm is None or lu is None:
        raise ImportError("vilco_b200.compat.install(): import the reference's libs.utils and libs.modeling first "
                          "(run from the ViLCo/MQ directory)")
    patched = []

    def put(mod, name, obj):
        if mod is not None and hasattr(mod, name):
            setattr(mod, name, obj)
            patched.append((mod.__name__, name))

    for name in _NAMES:
        put(lm, name, getattr(M, name))
    put(sys.modules.get("libs.modeling.models"), "make_meta_arch", M.make_meta_arch)
    put(sys.modules.get("libs.modeling.meta_archs"), "BiasLayer", M.BiasLayer)
    xl = sys.modules.get("libs.modeling.modeling_xlnet_x")
    for name in ("XLNetModel", "XLNetLMHeadModel"):
        if hasattr(X, name):
            put(xl, name, getattr(X, name))
    put(lu, "batched_nms", N.batched_nms)
    put(sys.modules.get("libs.utils.nms"), "batched_nms", N.batched_nms)
    put(sys.modules.get("libs.modeling.meta_archs"), "batched_nms", N.batched_nms)
    tu = sys.modules.get("libs.utils.train_utils")
    for name in ("MaskedConv1D", "Scale", "AffineDropPath", "LayerNorm", "BiasLayer"):
        put(tu, name, getattr(M, name))
    for name in ("XLNetModel", "XLNetLMHeadModel"):
        if hasattr(X, name):
            put(tu, name, getattr(X, name))
    # evaluation tail: the detection-mAP evaluator eval.py / train_cl.py construct and the recall metric
    # valid_one_epoch* calls — same signatures and bit-identical results (tests/test_metrics.py), without the per-row pandas
    # walk.  The reference's metrics.py uses np.float (numpy < 1.24), so on a current numpy this is also what makes it run.
    from .utils import metrics as E
    from .utils import get_retrieval_performance as G
    for mod in (lu, sys.modules.get("libs.utils.metrics"), tu):
        put(mod, "ANETdetection", E.ANETdetection)
    put(sys.modules.get("libs.utils.metrics"), "compute_average_precision_detection", E.compute_average_precision_detection)
    for mod in (sys.modules.get("libs.utils.get_retrieval_performance"), tu):
        put(mod, "evaluation_retrieval", G.evaluation_retrieval)
    put(sys.modules.get("libs.utils.get_retrieval_performance"), "Moment_Retrieval", G.Moment_Retrieval)
    return patched
