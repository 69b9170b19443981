#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "in_kernel_prenorm" 2>&1 | tail -15
timeout 300 python tools/prof_conv.py l0_kvtc l0_kvtc_pn l1_kvtc_pn l0_qout l0_qout_pn l1_qout_pn 2>&1 | cut -c1-60 | tee gpurun_out/r2_prof_pn3.txt
