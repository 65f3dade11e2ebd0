#!/bin/bash
mkdir -p gpurun_out
for c in 0 1 2; do
SLZW_DEC_CONFIG=$c timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1 | tee -a gpurun_out/pytest.log
done
timeout 300 python tools/profile_step.py --streams 65536 --passes 2 2>&1 | grep -v Warning | tail -3 | tee -a gpurun_out/step.log
timeout 600 python tools/profile_step.py --config 4 --streams 4096 --passes 2 2>&1 | grep -v Warning | tail -3 | head -1 | tee -a gpurun_out/step.log
timeout 600 python tools/profile_step.py --config 5 --streams 16384 --passes 2 2>&1 | grep -v Warning | tail -3 | head -1| tee -a gpurun_out/step.log
