#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
echo "== config 4, 512 frames" | tee -a gpurun_out/step.log
timeout 600 python tools/profile_step.py --config 4 --streams 512 --passes 2 2>&1 | grep -v Warning | tee -a gpurun_out/step.log
echo "== config 5, 16384 chunks" | tee -a gpurun_out/step.log
timeout 600 python tools/profile_step.py --config 5 --streams 16384 --passes 2 2>&1 | grep -v Warning | tee -a gpurun_out/step.log
