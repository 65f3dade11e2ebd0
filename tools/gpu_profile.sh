#!/bin/bash
# ncu evidence for profiles/ (run with gpurun from the repo root, ONE GPU, after the same commands
# have exited 0 without ncu): the launch list of the bench step and the full capture of the two
# codec kernels at the bench configuration (config 3, 65,536 strips).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/r02_profile_plain.json 2> gpurun_out/r02_profile.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_launch.log 2>&1
# encoder: kernel replay (one launch)
ncu --set full --import-source on --clock-control none -k regex:slzw_encode_kernel -s 1 -c 1 -o gpurun_out/r02_encode_65536 -f \
    python tools/profile_step.py --streams 65536 --passes 2 --what encode > gpurun_out/r02_ncu_encode.log 2>&1
tail -2 gpurun_out/r02_ncu_encode.log
# decoder: application replay (kernel replay could not restore the 10 GB of state at this size in round 1)
ncu --set full --import-source on --clock-control none --replay-mode application -k regex:slzw_decode_fast -c 1 -o gpurun_out/r02_decode_65536 -f \
    python tools/profile_step.py --streams 65536 --passes 1 > gpurun_out/r02_ncu_decode.log 2>&1
tail -2 gpurun_out/r02_ncu_decode.log
