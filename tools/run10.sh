#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "in_kernel_prenorm" 2>&1 | tail -3
echo turns=2; timeout 100 python tools/prof_conv.py l0_qout_pn l1_qout_pn | cut -c1-50
echo turns=1; DAC_LIB=da-clip_b200/libdac_b200_t1.so timeout 100 python tools/prof_conv.py l0_qout_pn l1_qout_pn | cut -c1-50
echo turns=3; DAC_LIB=da-clip_b200/libdac_b200_t3.so timeout 100 python tools/prof_conv.py l0_qout_pn l1_qout_pn | cut -c1-50
