#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" 2>&1 | tail -8
timeout 600 python -m pytest tests/test_daclip_gpu.py -q -x 2>&1 | tail -3
timeout 300 python tools/prof_encoder_layers.py 256 | grep -E "attn|total"
timeout 300 python tools/bench_configs.py 2>&1 | tail -1 | cut -c200-400
