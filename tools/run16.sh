#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -s > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest4.log
grep -E "parity|passed|failed|rc=|Error" gpurun_out/r2_pytest4.log | tail -16
python bench.py --steps 3 --warmup 3 --dump-layers gpurun_out/r2_layers_f.txt > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err
head -c 300 gpurun_out/r2_bench_f.json; tail -n 2 gpurun_out/r2_bench_f.err
python tools/bench_configs.py > gpurun_out/r2_configs_f.json 2> gpurun_out/r2_configs_f.err; tail -c 600 gpurun_out/r2_configs_f.json
