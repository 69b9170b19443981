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
SOURCES = ["api.cu", "conv_igemm.cu", "elementwise.cu", "linattn.cu", "attention.cu", "vit.cu", "imageio.cu", "linattn_qout.cu", "attention_tc.cu", "attention_tc2.cu", "attention_vit.cu", "linattn_kv.cu", "linattn_kv2.cu", "linattn_qout2.cu"]
HEADERS = ["common.h", "ptx.cuh", "tile_common.cuh", "tensormap.h", "conv_kernel.cuh", "linattn_kv_common.h", "linattn_qout_common.h", os.path.join("..", "..", "include", "dac_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--shared", "-cudart", "static",
]


STAMP = LIB + ".srchash"      # content hash of the sources the library was built from (git-ignored, travels with the .so)
OBJDIR = os.path.join(HERE, "build")


def _hash(files, extra=""):
    import hashlib
    h = hashlib.sha256((" ".join(NVCC_FLAGS) + extra).encode())
    for f in files:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()


def source_hash():
    """sha256 over every source, header and the compiler flags: staleness is decided by CONTENT, not by mtimes (a repo
    snapshot copied to the GPU box carries arbitrary mtimes)."""
    return _hash(SOURCES + HEADERS)


def _stale():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != source_hash()


def _compile_one(nvcc, src, verbose):
    """One translation unit -> build/<src>.o, skipped when its own content hash (source + all headers) is unchanged."""
    obj = os.path.join(OBJDIR, src + ".o")
    want = _hash([src] + HEADERS)
    stamp = obj + ".srchash"
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == want:
        return obj, ""
    flags = [f for f in NVCC_FLAGS if f != "--shared"]
    cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(want + "\n")
    return obj, r.stderr


def build(force=False, verbose=False, debug=False):
    if debug:
        return _build_debug()
    if not force and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJDIR):
            os.remove(os.path.join(OBJDIR, f))
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        results = list(ex.map(lambda s: _compile_one(nvcc, s, verbose), SOURCES))
    if verbose:
        sys.stderr.write("".join(log for _, log in results))
    tmp = LIB + f".tmp{os.getpid()}"
    r = subprocess.run([nvcc, "--shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]
                       + [o for o, _ in results] + ["-o", tmp], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libdac_b200.so")
    os.replace(tmp, LIB)          # atomic: concurrent ranks never dlopen a half-written library
    with open(STAMP, "w") as fh:
        fh.write(source_hash() + "\n")
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
