#!/bin/bash
# A/B of encoder builds on the GPU box: tools/run_variants.sh name1 name2 ... (libs built by tools/build_variants.sh)
export PYTHONUNBUFFERED=1
for v in "$@"; do
  echo "== $v"
  L=$PWD/lzw_b200/csrc/variants/libslzw_$v.so
  SLZW_LIB=$L python tools/enc_variants.py --workload config3 --streams 65536 --configs 0 --reps 5 2>&1 | tail -1 | cut -c1-330
  SLZW_LIB=$L python tools/enc_variants.py --workload config4 --streams 512 --configs 0 2>&1 | tail -1 | cut -c1-330
  SLZW_LIB=$L python tools/enc_variants.py --workload config5 --streams 16384 --configs 0 2>&1 | tail -1 | cut -c1-330
done
