#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
for c in 0 3 2; do
  echo "== SLZW_ENC_CONFIG=$c" | tee -a gpurun_out/step.log
  SLZW_ENC_CONFIG=$c timeout 300 python tools/profile_step.py --streams 16384 --passes 3 --what encode 2>&1 | grep -v Warning | tee -a gpurun_out/step.log
done
