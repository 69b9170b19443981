#!/bin/bash
DAC_CTA2_RES=1 timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_daclip_gpu.py -q -x 2>&1 | tail -2
for v in 0 1 0 1; do echo "== DAC_CTA2_RES=$v"; DAC_CTA2_RES=$v timeout 300 python tools/bench_configs.py 2>&1 | tail -1 | cut -c1-330; done
