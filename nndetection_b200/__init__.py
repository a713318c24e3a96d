"""nndetection_b200 -- B200-native (sm_100a) implementation of nnDetection's volumetric detection hot path.

Hand-written CUDA behind a C ABI (include/nndet_b200.h, nndetection_b200/csrc/), exposed to Python through
mirrors of the reference's own interfaces: `nndet._C.nms`, `nndet.core.boxes.*`, `nndet.arch.*`,
`nndet.core.retina.BaseRetinaNet`.  No CPU fallback, no Triton, no multi-backend dispatch.
"""
__version__ = "0.1.0"
