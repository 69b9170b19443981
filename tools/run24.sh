#!/bin/bash
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x 2>&1 | tail -3
timeout 300 python tools/cta2_check.py vit_qkv vit_fc vit_out vit_proj lin512 geglu l2_256 2>&1 | cut -c1-75
timeout 300 python tools/prof_encoder_layers.py 256 | grep -E "qkv|out|fc|proj|total"
timeout 300 python tools/bench_configs.py 2>&1 | tail -1 | cut -c1-400
