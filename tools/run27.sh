#!/bin/bash
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -q -x 2>&1 | tail -2
for v in 1 0 1 0; do echo "== DAC_NO_CTA2_PAIR=$v"; if [ $v = 1 ]; then export DAC_NO_CTA2_PAIR=1; else unset DAC_NO_CTA2_PAIR; fi; timeout 300 python tools/bench_configs.py 2>&1 | tail -1 | cut -c1-330; done
