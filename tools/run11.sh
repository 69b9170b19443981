#!/bin/bash
# full GPU suite + bench with layer dump
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -s > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
grep -E "parity|passed|failed|rc=|Error" gpurun_out/r2_pytest2.log | tail -25
python bench.py --steps 2 --warmup 3 --dump-layers gpurun_out/r2_layers_b.txt > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
head -c 700 gpurun_out/r2_bench_b.json; tail -3 gpurun_out/r2_bench_b.err
