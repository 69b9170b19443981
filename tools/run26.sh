#!/bin/bash
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -q -x 2>&1 | tail -2
timeout 300 python tools/cta2_check.py l2_128 l3_256 l2_256 l3_512 odd 2>&1 | cut -c1-75
timeout 300 python tools/bench_configs.py 2>&1 | tail -1 | cut -c1-400
