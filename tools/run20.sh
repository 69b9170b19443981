#!/bin/bash
# are the 19-GFLOP layers of levels 2-3 L2 -> SM bound?  standalone times + ncu --set full of both
python tools/prof_conv.py l2_3x3_128 l3_3x3_256 l3_3x3 l2_3x3
ncu --set full --import-source on --clock-control none -k regex:conv_igemm --launch-skip 4 -c 1 -f -o gpurun_out/ncu_l2_128 python tools/prof_conv.py l2_3x3_128 > gpurun_out/ncu_l2_128.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:conv_igemm --launch-skip 4 -c 1 -f -o gpurun_out/ncu_l3_256 python tools/prof_conv.py l3_3x3_256 > gpurun_out/ncu_l3_256.log 2>&1
ls -la gpurun_out/ncu_l2_128.ncu-rep gpurun_out/ncu_l3_256.ncu-rep
