"""In-tree build of libb200cs.so (hand-written sm_100a CUDA kernels + the C-ABI).

    python -m numbacs_b200._build [--force] [--verbose]

nvcc cross-compiles without a GPU.  Each translation unit is compiled in parallel to
``csrc/build/*.o`` and linked into ``numbacs_b200/libb200cs.so`` (git-ignored, travels to the GPU
box with the working tree).  The library links cudart statically and nothing else: no torch, no
third-party code.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libb200cs.so")

SOURCES = ["capi.cu", "flowmap_dispatch.cu", "flowmap_dg.cu", "flowmap_dg_damped.cu", "flowmap_bickley.cu",
           "flowmap_abc.cu", "flowmap_spline.cu", "flowmap_linear.cu", "ftle_kernels.cu", "diag_kernels.cu",
           "tensor_kernels.cu", "ridge_link.cu"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for fn in os.listdir(root):
            if fn.endswith((".cu", ".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, fn)))
    return m


def build(force=False, verbose=False, extra_flags=(), lib=LIB, objdir=OBJ):
    """Compile (if stale) and return the path of libb200cs.so.  `extra_flags` / `lib` / `objdir`
    build an experimental variant next to the product library (tools/build_variant.py)."""
    newest = _deps_mtime()
    if not force and os.path.exists(lib) and os.path.getmtime(lib) >= newest:
        return lib
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(objdir, src[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest:
            return obj
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed on " + src)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return lib


STRICT_LIB = os.path.join(HERE, "libb200cs_strict.so")
STRICT_FLAGS = ["-DB200CS_STRICT=1", "-fmad=false"]


def build_strict(force=False, verbose=False):
    """The parity-calibration build (csrc/dop853.cuh, B200CS_STRICT): reference evaluation order,
    separately rounded operations, CUDA libm.  Test infrastructure: the product never loads it."""
    return build(force=force, verbose=verbose, extra_flags=STRICT_FLAGS, lib=STRICT_LIB,
                 objdir=os.path.join(CSRC, "build_strict"))


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    if "--strict" in sys.argv:
        print(build_strict(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
