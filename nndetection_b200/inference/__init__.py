"""Inference-side rows next to the hot path: weighted box clustering and the case-level box ensembler on the device."""
from .wbc import batched_wbc, wbc  # noqa: F401
from .ensembler import (BoxEnsemblerSelective, batched_nms_ensemble, batched_nms_model, batched_wbc_ensemble,  # noqa: F401
                        batched_weighted_nms_model, wbc_nms_no_label_ensemble)
