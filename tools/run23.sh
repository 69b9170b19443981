#!/bin/bash
ncu --set full --import-source on --clock-control none -k regex:conv_igemm --launch-skip 20 -c 1 -f -o gpurun_out/ncu_p128 python tools/cta2_check.py p128 > gpurun_out/ncu_p128.log 2>&1
tail -1 gpurun_out/ncu_p128.log
