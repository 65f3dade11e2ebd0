#!/bin/bash
# One GPU-box visit that regenerates the bench evidence under profiles/ (run with gpurun from the
# repo root, 1 GPU): the contract's line and its reference arm, configs 4 and 5, the single-process
# (slzw_multi_*) line.  tools/gpu_profile.sh takes the ncu captures.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python bench.py > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench_line.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench_line.err
python bench.py --workload config4 --steps 3 --warmup 3 > gpurun_out/r02_config4_1gpu.json 2> gpurun_out/r02_config4.err
python bench.py --workload config5 --endian le --steps 3 --warmup 3 > gpurun_out/r02_config5_le_1gpu.json 2> gpurun_out/r02_config5.err
python bench.py --workload config5 --endian be --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_config5_be_1gpu.json 2>> gpurun_out/r02_config5.err
python bench.py --single-process --gpus 1 --steps 3 --warmup 2 > gpurun_out/r02_single_process_1gpu.json 2> gpurun_out/r02_single_process.err
for f in r02_bench_line r02_bench_reference r02_config4_1gpu r02_config5_le_1gpu r02_config5_be_1gpu r02_single_process_1gpu; do
  echo "== $f"; python - "$f" <<'P'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "encode_gbs", "decode_gbs", "n_gpus")},
          "e2e", (d.get("e2e") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"),
          "parity", d.get("parity"))
except Exception as e:
    print("no line:", e)
P
done
tail -n 3 gpurun_out/*.err
