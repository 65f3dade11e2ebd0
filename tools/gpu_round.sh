#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
for c in -1 0 1 2 3; do
  echo "== class $c" | tee -a gpurun_out/step.log
  timeout 300 python tools/profile_step.py --streams 16384 --passes 2 --cls $c 2>&1 | grep -v Warning | tee -a gpurun_out/step.log
done
ncu --set full --import-source on --clock-control none -k regex:slzw_decode_fast -c 1 -o gpurun_out/dec_fast_v1 -f python tools/profile_step.py --streams 4096 --passes 1 > gpurun_out/ncu_dec.log 2>&1
tail -3 gpurun_out/ncu_dec.log
