#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "groupnorm or layernorm" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_unet_gpu.py -q -x 2>&1 | tail -3
for sw in "FUSE_PRENORM_GN=1" "FUSE_PRENORM_GN=0" "FUSE_PRENORM_GN=1" "FUSE_PRENORM_GN=0"; do
  echo "== $sw"; DAC_SWITCHES=$sw python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_denoiser_step'], d['value'], d['clocks']['sm_mhz'])"
done
