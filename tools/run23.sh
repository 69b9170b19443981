#!/bin/bash
DAC_EPI_DEBUG=7 timeout 300 python tools/cta2_check.py vit_qkv vit_fc geglu 2>&1 | cut -c1-80
ncu --set full --import-source on --clock-control none -k regex:conv_igemm --launch-skip 6 -c 1 -f -o gpurun_out/ncu_cta1 python tools/cta2_check.py vit_qkv > gpurun_out/ncu_cta1.log 2>&1
tail -1 gpurun_out/ncu_cta1.log
