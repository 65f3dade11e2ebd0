#!/bin/bash
# Bench line + ncu evidence for profiles/ (round 1).
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err
tail -c 3000 gpurun_out/bench_r01.json
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'slzw_encode|slzw_decode_fast' -s 2 -c 2 -o gpurun_out/r01_bench_kernels -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
