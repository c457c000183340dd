#!/bin/bash
# Builds experiment variants of libb2webp.so with different -D flags for one source file
# (VARIANT_SRC, default b2_upsample):
#   tools/variants.sh name "-DB2_BULK_CTAS=5 -DB2_PASS_UNROLL=1" [name2 "flags2" ...]
# Output: go_webp_b200/lib/variants/libb2webp_<name>.so  (run with B2_LIB=<path> tools/quick_bench.py)
set -e
cd "$(dirname "$0")/../go_webp_b200/csrc"
make -s >/dev/null
mkdir -p ../lib/variants build/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  src=${VARIANT_SRC:-b2_upsample}
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v --use_fast_math $flags \
    -c $src.cu -o build/variants/v_$name.o 2> build/variants/v_$name.log
  objs=""
  for o in b2_ctx b2_import b2_import_fast b2_upsample b2_vp8l b2_extra b2_predict b2_convert b2_tools b2_dither b2_anim b2_filters b2_rescale b2_dequant b2_staged b2_palette; do
    if [ $o = $src ]; then objs="$objs build/variants/v_$name.o"; else objs="$objs build/$o.o"; fi
  done
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/variants/libb2webp_$name.so $objs -cudart static
  echo "$name: $(grep -oE 'Used [0-9]+ registers' build/variants/v_$name.log | sort | uniq -c | tr '\n' ' ')"
done
