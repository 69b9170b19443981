#!/bin/bash
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -q -x 2>&1 | tail -2
timeout 300 python tools/cta2_check.py p64 p128 p64res p64skip p64l1 2>&1 | cut -c1-75
DAC_EPI_DEBUG=7 timeout 300 python tools/cta2_check.py p64 p128 2>&1 | cut -c1-50
timeout 300 python tools/bench_configs.py 2>&1 | tail -1 | cut -c1-330
