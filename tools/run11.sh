#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -s > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log
grep -E "parity|passed|failed|rc=|Error" gpurun_out/r2_pytest3.log | tail -16
python bench.py --steps 2 --warmup 3 --dump-layers gpurun_out/r2_layers_d.txt > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err
head -c 300 gpurun_out/r2_bench_d.json; tail -n 2 gpurun_out/r2_bench_d.err
