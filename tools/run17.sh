#!/bin/bash
timeout 900 python -m pytest tests/test_daclip_gpu.py tests/test_parity_configs_gpu.py -q -x -s -k "daclip or argmax or encode" 2>&1 | grep -E "parity|passed|failed|Error" | tail -12
echo "== fused zero"; timeout 300 python tools/prof_encoder_layers.py 256 | grep -E "proj|zero|total"
echo "== separate zero"; DAC_FUSE_ZERO=0 timeout 300 python tools/prof_encoder_layers.py 256 | grep -E "proj|zero|total"
timeout 300 python tools/bench_configs.py 2>&1 | tail -1 | cut -c200-400
