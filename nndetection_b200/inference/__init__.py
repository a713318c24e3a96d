"""Inference-side rows next to the hot path: weighted box clustering, the case-level box ensembler and the sliding-window
predictor loop, all device-resident."""
from .wbc import batched_wbc, wbc  # noqa: F401
from .ensembler import (BoxEnsemblerSelective, batched_nms_ensemble, batched_nms_model, batched_wbc_ensemble,  # noqa: F401
                        batched_weighted_nms_model, wbc_nms_no_label_ensemble)
from .predictor import SlidingWindowPredictor, create_grid, get_tta_dims, mirror_boxes  # noqa: F401
from .helper import (get_loader_fn, get_predictor, load_all_models, load_final_model, predict_dir,  # noqa: F401
                     save_checkpoint, sweep)
from .sweeper import BoxSweeper  # noqa: F401
