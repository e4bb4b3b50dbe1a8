from .nms import batched_nms  # noqa: F401
from .metrics import ANETdetection, remove_duplicate_annotations  # noqa: F401
from .get_retrieval_performance import evaluation_retrieval  # noqa: F401
from .validate import valid_one_epoch  # noqa: F401
