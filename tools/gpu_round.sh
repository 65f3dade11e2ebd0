#!/bin/bash
mkdir -p gpurun_out
for c in 0 2; do
  SLZW_ENC_CONFIG=$c timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/pytest.log
done
for c in 0; do
  echo "== SLZW_ENC_CONFIG=$c" | tee -a gpurun_out/step.log
  SLZW_ENC_CONFIG=$c timeout 300 python tools/profile_step.py --streams 16384 --passes 3 --what encode 2>&1 | grep -v Warning | tail -2 | tee -a gpurun_out/step.log
done
echo "== config 5 (fixed), 16384 chunks" | tee -a gpurun_out/step.log
timeout 600 python tools/profile_step.py --config 5 --streams 16384 --passes 2 2>&1 | grep -v Warning | tail -3 | tee -a gpurun_out/step.log
