#!/bin/bash
# Multi-GPU evidence (run with `gpurun --gpus N -- bash tools/gpu_scale.sh N`): the multi-device
# tests, the contract's line at N GPUs (config 3, weak), the strong-scaling lines of configs 3, 4
# and 5, and the single-process (slzw_multi_*) line.
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR tools/host_link_probe.py > gpurun_out/r02_host_link_${N}gpu.json 2> gpurun_out/r02_scale_${N}gpu.err
cat gpurun_out/r02_host_link_${N}gpu.json
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -2 | tee gpurun_out/r02_multi_tests_${N}gpu.log
$TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2>> gpurun_out/r02_scale_${N}gpu.err
$TR bench.py --gpus $N --steps 3 --warmup 3 --scaling strong > gpurun_out/r02_config3_strong_${N}gpu.json 2>> gpurun_out/r02_scale_${N}gpu.err
$TR bench.py --gpus $N --steps 3 --warmup 3 --workload config4 > gpurun_out/r02_config4_${N}gpu.json 2>> gpurun_out/r02_scale_${N}gpu.err
$TR bench.py --gpus $N --steps 3 --warmup 3 --workload config5 --endian le > gpurun_out/r02_config5_le_${N}gpu.json 2>> gpurun_out/r02_scale_${N}gpu.err
$TR bench.py --gpus $N --steps 3 --warmup 3 --workload config5 --endian be --no-e2e > gpurun_out/r02_config5_be_${N}gpu.json 2>> gpurun_out/r02_scale_${N}gpu.err
python bench.py --single-process --gpus $N --steps 3 --warmup 2 > gpurun_out/r02_single_process_${N}gpu.json 2>> gpurun_out/r02_scale_${N}gpu.err
for f in r02_bench_${N}gpu r02_config3_strong_${N}gpu r02_config4_${N}gpu r02_config5_le_${N}gpu r02_config5_be_${N}gpu r02_single_process_${N}gpu; do
  echo "== $f"; python - "$f" <<'P'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
    p = d.get("parity") or {}
    print({k: d.get(k) for k in ("value", "ms_per_step", "encode_gbs", "decode_gbs", "n_gpus", "scaling")},
          "e2e", (d.get("e2e") or {}).get("value"),
          "parity", {k: p.get(k) for k in ("all_ranks_oracle_sample_byte_exact", "all_ranks_round_trip_bytes_equal", "oracle_sample_byte_exact")})
except Exception as e:
    print("no line:", e)
P
done
tail -n 5 gpurun_out/r02_scale_${N}gpu.err
