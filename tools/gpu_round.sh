#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1 | tee -a gpurun_out/pytest.log
timeout 300 python tools/profile_step.py --streams 65536 --passes 2 2>&1 | grep -v Warning | tail -3 | head -1 | tee -a gpurun_out/step.log
ncu --set full --import-source on --clock-control none -k regex:slzw_decode_fast -c 1 -o gpurun_out/dec_fast_v4 -f python tools/profile_step.py --streams 16384 --passes 1 > gpurun_out/ncu_dec.log 2>&1
tail -1 gpurun_out/ncu_dec.log
