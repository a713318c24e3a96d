from .nms import nms, batched_nms  # noqa: F401
from .ops import box_iou, generalized_box_iou, box_center_dist  # noqa: F401
from .anchors import AnchorGenerator3DS, get_anchor_generator  # noqa: F401
from .matcher import ATSSMatcher  # noqa: F401
from .sampler import HardNegativeSamplerBatched  # noqa: F401
from .coder import BoxCoderND  # noqa: F401
