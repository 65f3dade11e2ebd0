#!/bin/bash
# Tests under every encoder configuration, then the bench line + ncu evidence for profiles/.
mkdir -p gpurun_out
for c in 0 1 2; do
  SLZW_ENC_CONFIG=$c timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/pytest.log
done
python bench.py > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r01c.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','encode_gbs','decode_gbs','e2e','gpu_launches')})
print(d['roofline']); print(d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01c_launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'slzw_encode' -s 1 -c 1 -o gpurun_out/r01c_bench_encode -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
