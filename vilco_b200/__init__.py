"""vilco_b200 — B200-native (sm_100a) Moment-Query hot path of ViLCo behind the reference's own API.

    from vilco_b200.modeling import make_meta_arch      # drop-in for libs.modeling.make_meta_arch
    from vilco_b200.utils import batched_nms            # drop-in for libs.utils.batched_nms

Importing the package does not need a GPU; any compute call does (there is no CPU fallback)."""
from . import lib  # noqa: F401

__version__ = "0.1.0"
