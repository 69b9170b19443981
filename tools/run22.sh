#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "layernorm or groupnorm or transformer or attention_block or spatial" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_unet_gpu.py -q -x 2>&1 | tail -2
for i in 1 2; do python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_denoiser_step'], d['value'], d['clocks']['sm_mhz'])"; done
