"""Builds a hand-made variant of the library with extra -D flags: python tools/build_variant.py OUT.so -DFLAG ...
(load it with DAC_LIB=OUT.so).  Debugging / profiling only."""
import os, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "da-clip_b200"))
import build
out, flags = sys.argv[1], sys.argv[2:]
cmd = ["nvcc"] + build.NVCC_FLAGS + flags + [os.path.join(build.CSRC, f) for f in build.SOURCES] + ["-o", out]
r = subprocess.run(cmd, capture_output=True, text=True)
print(r.returncode, r.stderr[-300:] if r.returncode else "ok")
