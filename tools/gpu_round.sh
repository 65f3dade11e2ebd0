#!/bin/bash
mkdir -p gpurun_out
python bench.py --streams 8192 --steps 3 --warmup 3 --cpu-sample 256 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
tail -c 900 gpurun_out/bench_small.json; tail -3 gpurun_out/bench_small.err
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
