"""Inference-side rows next to the hot path: weighted box clustering on the device."""
from .wbc import batched_wbc, wbc  # noqa: F401
