#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x > gpurun_out/r2_kernel_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_kernel_tests.log
tail -4 gpurun_out/r2_kernel_tests.log
timeout 300 python tools/prof_conv.py l0_pair l0_pair_res l0_pair_cat l0_pair_skip l0_pair_plain l0_3x3 l1_3x3 l1_pair l3_3x3 l3_geglu l0_kvtc l0_qout l0_kv l0_toout l0_q l0_final_pair 2>&1 | tee gpurun_out/r2_prof_conv.txt
