#!/bin/bash
# Builds experiment variants of libb2webp.so with different -D flags for b2_upsample.cu:
#   tools/variants.sh name "-DB2_BULK_CTAS=5 -DB2_PASS_UNROLL=1" [name2 "flags2" ...]
# Output: go_webp_b200/lib/variants/libb2webp_<name>.so  (run with B2_LIB=<path> tools/quick_bench.py)
set -e
cd "$(dirname "$0")/../go_webp_b200/csrc"
make -s >/dev/null
mkdir -p ../lib/variants build/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v --use_fast_math $flags \
    -c b2_upsample.cu -o build/variants/up_$name.o 2> build/variants/up_$name.log
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/variants/libb2webp_$name.so \
    build/b2_ctx.o build/b2_import.o build/b2_import_fast.o build/variants/up_$name.o build/b2_vp8l.o -cudart static
  echo "$name: $(grep -A2 'upsample_bulk_kernelILi0ELb0' build/variants/up_$name.log | grep -oE 'Used [0-9]+ registers')"
done
