#!/bin/bash
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -q -x 2>&1 | tail -2
timeout 300 python tools/prof_conv.py l0_down l1_up l0_1x1 l0_final_pair l0_qkv 2>&1 | cut -c1-120
timeout 300 python tools/bench_configs.py 2>&1 | tail -1 | cut -c1-330
