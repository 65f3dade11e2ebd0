#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fuzz.py -x -q 2>&1 | tail -25 | tee gpurun_out/fuzz.log
