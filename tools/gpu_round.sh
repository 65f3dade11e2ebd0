#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee -a gpurun_out/pytest.log
python tools/e2e_probe.py 2>&1 | grep -v Warning | tee gpurun_out/e2e_probe2.log
