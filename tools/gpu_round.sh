#!/bin/bash
# One GPU-box visit that regenerates the evidence under profiles/ (run with gpurun from the repo root):
#   parity tests, the bench line, the reference arm, the ncu launch list and the full capture of the
#   two codec kernels of the same bench command.  Copy the results from gpurun_out/ to profiles/ with
#   tools/ncu_summary.py.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest.log
python bench.py > gpurun_out/bench_line.json 2> gpurun_out/bench_line.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_line.err
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'slzw_encode|slzw_decode_fast' -s 2 -c 2 -o gpurun_out/bench_kernels -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
