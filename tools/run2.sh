#!/bin/bash
mkdir -p gpurun_out
python tools/diag_r2.py > gpurun_out/r2_diag.txt 2>&1
tools/_bin/mufu_peak2 > gpurun_out/r2_mufu.txt 2>&1
cat gpurun_out/r2_diag.txt gpurun_out/r2_mufu.txt
