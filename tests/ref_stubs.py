"""TEST INFRASTRUCTURE: permissive stand-ins for the third-party packages the UNMODIFIED reference imports at module level but that are
absent here (pytorch_lightning, SimpleITK, hydra, omegaconf, batchgenerators, mlflow, ...), so that `nndet.ptmodule` -- the registry and
the Lightning module skeleton the plug-in subclasses -- can be imported and executed on CPU (tests/test_registry_cpu.py).  A stub module
hands out classes that accept any construction / attribute access / decorator use; nothing of the reference itself is replaced."""
import sys, types, importlib, importlib.abc, importlib.machinery

class _Meta(type):
    def __getattr__(cls, k):
        if k.startswith("__"): raise AttributeError(k)
        return _make(k)
    def __call__(cls, *a, **k):
        # decorator use: @stub(...) / @stub
        if len(a) == 1 and callable(a[0]) and not k and cls.__dict__.get("_decorator_ok", True) and isinstance(a[0], (types.FunctionType, type)):
            return a[0]
        return super().__call__(*a, **k)
    def __getitem__(cls, k): return cls
    def __or__(cls, o): return cls
    def __ror__(cls, o): return cls

def _make(name):
    return _Meta(name, (), {"__init__": lambda self, *a, **k: None, "__getattr__": lambda self, k: _make(k)(), "__call__": lambda self, *a, **k: (a[0] if a and callable(a[0]) else self)})

class StubModule(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"): raise AttributeError(k)
        v = _make(k); setattr(self, k, v); return v

class Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, roots): self.roots = set(roots)
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
    def create_module(self, spec):
        m = StubModule(spec.name); m.__path__ = []; return m
    def exec_module(self, module): pass

def install(roots):
    sys.meta_path.insert(0, Finder(roots))


ROOTS = ["pytorch_lightning", "loguru", "SimpleITK", "omegaconf", "hydra", "batchgenerators", "mlflow", "nevergrad", "git", "skimage",
         "tqdm.contrib", "matplotlib", "nnunet"]


def import_reference(ref_root="/root/reference"):
    """Import the reference package with the stand-ins installed; returns the `nndet` module."""
    import os
    if not os.path.isdir(os.path.join(ref_root, "nndet")):
        return None
    for k in ROOTS:
        if k in sys.modules and not isinstance(sys.modules[k], StubModule):
            try:
                importlib.import_module(k)          # really installed: keep it
                continue
            except Exception:
                del sys.modules[k]
    install([r for r in ROOTS if r not in sys.modules])
    if "torch._six" not in sys.modules:
        m = types.ModuleType("torch._six"); m.string_classes = (str, bytes); sys.modules["torch._six"] = m
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    return importlib.import_module("nndet")
