#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "device_driven or sde_step" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_driver_gpu.py -q -x 2>&1 | tail -5
timeout 300 python tools/bench_configs.py 2>&1 | tail -8
