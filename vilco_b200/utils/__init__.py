from .nms import batched_nms  # noqa: F401
