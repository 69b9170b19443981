#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "layernorm" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_daclip_gpu.py -q -x 2>&1 | tail -2
timeout 300 python tools/prof_encoder_layers.py 256 | grep -E "ln|total"
timeout 300 python tools/bench_configs.py 2>&1 | tail -1 | cut -c200-400
