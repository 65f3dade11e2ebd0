#!/bin/bash
# Bench line + ncu evidence for profiles/ (round 1, final kernels).
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r01e.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','encode_gbs','decode_gbs','gpu_launches')}); print(d['e2e']['value'])
print(d['roofline']); print(d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['clocks'], d['parity'])
PY
python bench.py --impl reference --steps 2 --warmup 1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('reference arm', d['value'], d['cpu_baseline']['cores'])"
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01e_launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'slzw_encode|slzw_decode_fast' -s 2 -c 2 -o gpurun_out/r01e_bench_kernels -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
