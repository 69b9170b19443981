#!/bin/bash
# ncu source-level captures: ViT attention (packed) and the d = 32 attention
python tools/prof_vit_attn.py
python tools/prof_attn.py | head -1
ncu --set full --import-source on --clock-control none -k regex:attn_vit --launch-skip 4 -c 1 -f -o gpurun_out/ncu_vit python tools/prof_vit_attn.py 256,50,12 > gpurun_out/ncu_vit.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:attn_tc2 --launch-skip 4 -c 1 -f -o gpurun_out/ncu_tc2 python tools/prof_attn.py > gpurun_out/ncu_tc2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
