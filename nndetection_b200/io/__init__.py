"""Data-format rows next to the hot path: the per-step instance -> box / semantic-target transforms."""
from .instances import FindInstances, Instances2Boxes, Instances2Segmentation, instances_to_targets  # noqa: F401
