"""Builds libdac_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

In-tree so that the built library travels to the GPU box with the repo snapshot; rebuilds only when a
source is newer than the library.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdac_b200.so")
SOURCES = ["api.cu", "conv_igemm.cu", "elementwise.cu", "linattn.cu", "attention.cu", "vit.cu", "imageio.cu", "linattn_qout.cu", "attention_tc.cu", "linattn_kv.cu"]
HEADERS = ["common.h", "ptx.cuh", "tile_common.cuh", "tensormap.h", "conv_kernel.cuh", os.path.join("..", "..", "include", "dac_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--shared", "-cudart", "static",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, debug=False):
    if debug:
        return _build_debug()
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        [os.path.join(CSRC, f) for f in SOURCES] + ["-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libdac_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


def _build_debug():
    """libdac_b200_debug.so: same sources with -DDAC_DEBUG (device printf in the conv kernel)."""
    out = LIB.replace(".so", "_debug.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-DDAC_DEBUG"] + [os.path.join(CSRC, f) for f in SOURCES] + ["-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building the debug library")
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
