export PYTHONUNBUFFERED=1
for v in tmemfirst smemfirst; do
  echo "== $v"
  for s in 512 1776; do SLZW_LIB=$PWD/lzw_b200/csrc/variants/libslzw_$v.so python tools/enc_variants.py --workload config4 --streams $s --configs 2 2>&1 | tail -1 | cut -c1-110; done
  SLZW_LIB=$PWD/lzw_b200/csrc/variants/libslzw_$v.so python tools/enc_variants.py --workload config3 --streams 1776 --configs 2 2>&1 | tail -1 | cut -c1-110
  SLZW_LIB=$PWD/lzw_b200/csrc/variants/libslzw_$v.so python tools/enc_variants.py --workload config5 --streams 1776 --configs 2 2>&1 | tail -1 | cut -c1-110
done
