#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --dump-layers gpurun_out/r2_layers_c.txt > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
head -c 400 gpurun_out/r2_bench_c.json; tail -2 gpurun_out/r2_bench_c.err
