#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "in_kernel_prenorm" 2>&1 | tail -3
timeout 200 python tools/kv2_sweep.py
