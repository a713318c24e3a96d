"""TEST INFRASTRUCTURE ONLY -- builds the REFERENCE's own CUDA NMS (`nndet._C.nms`: nndet/csrc/ops.cpp:13-15,
nndet/csrc/cpu/nms.cpp:18-34, nndet/csrc/cuda/nms.cu:99-221) as the python extension `oracle/_ref/ref_nms*.so` for the on-device
A/B (keep-list equality + timing) against this package's `nndet._C.nms` replacement (SURVEY 8c "Native NMS").

    python oracle/build_ref_nms.py            (in the build container; `/root/reference` must exist)

Nothing of the reference is copied into the repository: `ops.cpp` / `cpu/nms.cpp` / `cuda_helpers.h` are compiled where they lie,
`cuda/nms.cu` is compiled from a scratch copy under /tmp that differs from the original in exactly two tokens -- the
`AT_DISPATCH_FLOATING_TYPES_AND_HALF(dets_sorted.type(), ...)` calls at nms.cu:172,182 become `dets_sorted.scalar_type()`, without
which the file does not compile against torch >= 2.x.  Output only into oracle/_ref/ (git-ignored; it travels to the
GPU box with the snapshot).  The GPU box has no /root/reference: there the prebuilt file is loaded (`load()` below) or the A/B
is skipped.
"""
import glob
import importlib.util
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_CSRC = "/root/reference/nndet/csrc"


def build(verbose: bool = False) -> str:
    import torch
    from torch.utils import cpp_extension as ce
    if not os.path.isdir(REF_CSRC):
        raise FileNotFoundError(REF_CSRC)
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="ref_nms_build_")
    src = open(os.path.join(REF_CSRC, "cuda", "nms.cu")).read()
    assert src.count("dets_sorted.type()") == 2, "reference nms.cu changed: expected exactly the two dispatch sites of nms.cu:172,182"
    patched = os.path.join(tmp, "nms.cu")
    open(patched, "w").write(src.replace("dets_sorted.type()", "dets_sorted.scalar_type()"))
    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}", f"-I{REF_CSRC}", f"-I{os.path.join(REF_CSRC, 'cuda')}"]
    defs = ["-DTORCH_EXTENSION_NAME=ref_nms", "-DTORCH_API_INCLUDE_EXTENSION_H", "-DWITH_CUDA", "-D_GLIBCXX_USE_CXX11_ABI=" + str(int(torch._C._GLIBCXX_USE_CXX11_ABI))]
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    o_cu, o_cpp = os.path.join(tmp, "nms.o"), os.path.join(tmp, "ops.o")
    cmds = [
        [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
         "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__"]
        + defs + inc + ["-c", patched, "-o", o_cu],
        ["g++", "-O2", "-std=c++17", "-fPIC"] + defs + inc + ["-c", os.path.join(REF_CSRC, "ops.cpp"), "-o", o_cpp],
    ]
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    lib = os.path.join(OUT, "ref_nms" + ext)
    libdirs = ce.library_paths("cuda")
    cmds.append(["g++", "-shared", "-o", lib, o_cu, o_cpp] + [f"-L{d}" for d in libdirs] + [f"-Wl,-rpath,{d}" for d in libdirs]
                + ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"])
    for c in cmds:
        r = subprocess.run(c, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(" ".join(c) + "\n" + r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("ref_nms build failed")
    shutil.rmtree(tmp, ignore_errors=True)
    return lib


def load():
    """The built extension module (attribute `nms(dets, scores, iou_threshold)`), or None when it was never built."""
    import torch  # noqa: F401  (libtorch must be loaded first)
    if "ref_nms" in sys.modules:
        return sys.modules["ref_nms"]
    cands = glob.glob(os.path.join(OUT, "ref_nms*.so"))
    if not cands:
        return None
    spec = importlib.util.spec_from_file_location("ref_nms", cands[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["ref_nms"] = mod
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
