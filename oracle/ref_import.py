"""TEST INFRASTRUCTURE ONLY. Import shim for the *unmodified* Python reference.

Only usable in the build container (where /root/reference exists); the GPU box
never imports this.  It pre-seeds two absent third-party modules so that
`nndet.core`, `nndet.arch` and `nndet.losses` import on CPU:
  * `torch._six`   <- nndet/utils/tensor.py:12
  * `omegaconf`    <- nndet/utils/info.py:26
Used by scripts/gen_golden.py to pin the oracle and to write tests/golden/*.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("NNDET_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "nndet"))


def load():
    if not available():
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    if "torch._six" not in sys.modules:
        m = types.ModuleType("torch._six")
        m.string_classes = (str, bytes)
        sys.modules["torch._six"] = m
    if "omegaconf" not in sys.modules:
        m = types.ModuleType("omegaconf")
        m.OmegaConf = type("OmegaConf", (), {})
        m.DictConfig = dict
        m.ListConfig = list
        sys.modules["omegaconf"] = m
        m2 = types.ModuleType("omegaconf.omegaconf")
        m2.OmegaConf = m.OmegaConf
        sys.modules["omegaconf.omegaconf"] = m2
    for name in ("SimpleITK",):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import nndet  # noqa: F401
    return nndet
