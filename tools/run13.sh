#!/bin/bash
# ViT attention: packed / unpacked tcgen05 vs the mma.sync kernel; encoder per-op profile
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_daclip_gpu.py -q -x 2>&1 | tail -3
echo "== packed"; timeout 300 python tools/prof_encoder_layers.py 256 | tee gpurun_out/enc_layers_packed.txt | grep -E "attn|total"
echo "== unpacked"; DAC_ATTN_NO_PACK=1 timeout 300 python tools/prof_encoder_layers.py 256 | grep -E "attn|total"
echo "== mma.sync"; DAC_NO_TC_ATTN=1 timeout 300 python tools/prof_encoder_layers.py 256 | grep -E "attn|total"
timeout 300 python tools/bench_configs.py 2>&1 | tail -1 | cut -c200-400
