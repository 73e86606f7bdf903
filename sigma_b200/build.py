"""Build libsigma_b200.so (sm_100a only) in-tree with nvcc.

    python -m sigma_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The shared library links cudart statically and resolves the
one driver entry point it needs (cuTensorMapEncodeTiled) at run time through
cudaGetDriverEntryPoint, so it loads on a machine without libcuda (the CPU test box).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
OBJ = os.path.join(ROOT, "build")
LIB = os.path.join(ROOT, "libsigma_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".inc"))]
    hs.append(os.path.join(os.path.dirname(ROOT), "include", "sigma_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, verbose):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    log = obj + ".log"
    if os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), _headers_mtime()):
        return obj
    cmd = [NVCC, *ARCH, *CFLAGS, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(r.stderr)
    return obj


def build(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link libsigma_b200.so.  Returns its path."""
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    if (not os.path.exists(LIB)) or force or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, *ARCH, "-shared", "-cudart", "static", "-o", LIB, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
