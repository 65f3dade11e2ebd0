export SLZW_ENC_CONFIG=0
SLZW_LIB=$PWD/lzw_b200/csrc/variants/libslzw_redux_u4.so timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -3
unset SLZW_ENC_CONFIG
for v in r1 redux_u4 ballot_u4; do
  echo "== $v"; SLZW_LIB=$PWD/lzw_b200/csrc/variants/libslzw_$v.so timeout 300 python tools/enc_variants.py --streams 65536 --configs 1,0 --reps 5 2>&1 | tail -1
done
