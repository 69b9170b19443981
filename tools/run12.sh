#!/bin/bash
# ncu launch list of a bench run in steady state (B=16 256^2) + the same for batch 1; then a plain bench with every leg
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-library-baseline > gpurun_out/plain_b16.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip 16000 -c 264 --csv --log-file gpurun_out/r02_ncu_launches_bench_n1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-library-baseline > gpurun_out/ncu_b16.log 2>&1
tail -n 3 gpurun_out/ncu_b16.log
python bench.py --batch 1 --steps 1 --warmup 1 --no-cpu-baseline --no-library-baseline > gpurun_out/plain_b1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 16000 -c 264 --csv --log-file gpurun_out/r02_ncu_launches_bench_b1.csv python bench.py --batch 1 --steps 1 --warmup 1 --no-cpu-baseline --no-library-baseline > gpurun_out/ncu_b1.log 2>&1
tail -n 3 gpurun_out/ncu_b1.log
python bench.py --steps 3 --warmup 3 --dump-layers gpurun_out/r2_layers_e.txt > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err
head -c 300 gpurun_out/r2_bench_e.json
