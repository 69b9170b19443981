#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "layernorm or groupnorm" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_unet_gpu.py -q -x -s 2>&1 | grep -E "parity|passed|failed|Error|error" | tail -12
