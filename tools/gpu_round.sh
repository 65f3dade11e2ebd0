#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1 | tee -a gpurun_out/pytest.log
timeout 300 python tools/profile_step.py --streams 65536 --passes 2 2>&1 | grep -v Warning | tail -3 | tee -a gpurun_out/step.log
