"""Same import surface as the reference's `libs.modeling` (MQ/libs/modeling/__init__.py)."""
from .blocks import (MaskedConv1D, MaskedMHCA, LocalMaskedMHCA, MaskedMHA, LayerNorm, TransformerBlock, Scale,
                     AffineDropPath)
from .models import make_backbone, make_neck, make_meta_arch, make_generator
from . import backbones       # noqa: F401  registers "convTransformer"
from . import necks           # noqa: F401  registers "identity"
from . import loc_generators  # noqa: F401  registers "point"
from . import meta_archs      # noqa: F401  registers "LocPointTransformer"
from . import modeling_xlnet_x  # noqa: F401
from .meta_archs import BiasLayer
from . import nlq             # noqa: F401  registers "NlqLocPointTransformer" (NLQ tree, evaluation path)

__all__ = ["MaskedConv1D", "MaskedMHCA", "LocalMaskedMHCA", "MaskedMHA", "LayerNorm", "TransformerBlock", "Scale",
           "AffineDropPath", "make_backbone", "make_neck", "make_meta_arch", "make_generator", "BiasLayer"]
