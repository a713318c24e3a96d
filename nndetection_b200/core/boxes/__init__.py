from .nms import nms, batched_nms  # noqa: F401
