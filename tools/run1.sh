#!/bin/bash
# round-2 GPU call 1: parity suite, bench with layer dump, epilogue diagnostics
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt 2>&1
python -m pytest tests -m gpu -q -x -s > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
python bench.py --steps 2 --warmup 3 --dump-layers gpurun_out/r2_layers_a.txt > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
python tools/diag_r2.py > gpurun_out/r2_diag.txt 2>&1
tail -5 gpurun_out/r2_pytest1.log; cat gpurun_out/r2_bench_a.json | head -c 600; cat gpurun_out/r2_diag.txt
