#!/bin/bash
mkdir -p gpurun_out
python tools/prof_conv.py l0_qout_pn > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"linattn_qout2" -s 4 -c 1 -f -o gpurun_out/r2_qout2 python tools/prof_conv.py l0_qout_pn > gpurun_out/ncu_q.log 2>&1
tail -n 2 gpurun_out/ncu_q.log
