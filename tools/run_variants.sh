export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_stress.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for v in sync0 sync1 sync0 sync1; do
  echo "== $v"; SLZW_LIB=$PWD/lzw_b200/csrc/variants/libslzw_$v.so timeout 300 python tools/enc_variants.py --streams 65536 --configs 2 --reps 5 2>&1 | tail -1 | cut -c1-120
done
