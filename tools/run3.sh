#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention or attn" > gpurun_out/r2_attn_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_attn_tests.log
tail -15 gpurun_out/r2_attn_tests.log
echo "--- v2"; timeout 120 python tools/prof_attn.py
echo "--- v1"; DAC_ATTN_V1=1 timeout 120 python tools/prof_attn.py
echo "--- v2 dbg1 (no exp)"; DAC_ATTN_DEBUG=1 timeout 120 python tools/prof_attn.py
echo "--- v2 dbg5 (no exp, no P store)"; DAC_ATTN_DEBUG=5 timeout 120 python tools/prof_attn.py
