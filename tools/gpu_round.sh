#!/bin/bash
# One GPU-box visit: per-class timing + ncu capture of the encoder (debug aid, not the bench).
mkdir -p gpurun_out
for c in 0 1 2 3; do
  echo "== class $c" | tee -a gpurun_out/step.log
  python tools/profile_step.py --streams 16384 --passes 2 --cls $c 2>&1 | grep -v Warning | tee -a gpurun_out/step.log
done
ncu --set full --import-source on --clock-control none -k regex:slzw_encode -c 1 -o gpurun_out/enc_v6 -f python tools/profile_step.py --streams 4096 --passes 1 --what encode > gpurun_out/ncu_enc.log 2>&1
tail -3 gpurun_out/ncu_enc.log
