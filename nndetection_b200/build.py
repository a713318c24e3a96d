"""In-tree build of libnndet_b200.so (hand-written sm_100a CUDA behind a C ABI).

`python -m nndetection_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU.
The .so lands in nndetection_b200/lib/ (git-ignored, but it travels to the GPU box with the snapshot).
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib")
LIB = os.path.join(OUT, "libnndet_b200.so")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
         "-I", SRC, "-I", os.path.join(os.path.dirname(HERE), "include")]


def _digest(path):
    h = hashlib.sha1()
    for f in sorted(os.listdir(SRC)):
        if f.endswith((".cuh", ".h")) or os.path.join(SRC, f) == path:
            h.update(open(os.path.join(SRC, f), "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src, verbose):
    obj = os.path.join(OUT, os.path.basename(src)[:-3] + ".o")
    stamp = obj + ".sha1"
    dig = _digest(src)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, False
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    open(stamp, "w").write(dig)
    return obj, True


def build(verbose=False, force=False):
    os.makedirs(OUT, exist_ok=True)
    if force:
        for f in os.listdir(OUT):
            os.remove(os.path.join(OUT, f))
    srcs = sorted(os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith(".cu"))
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [o for o, _ in res]
    if any(ch for _, ch in res) or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart", "-lcuda"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
