#!/bin/bash
fmt='
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print("%-22s import %.3f  upsample %.3f  chain %.3f" % (d["subset"], d["import"]["frac"], d["upsample"]["frac"], d["chain"]["frac"]))
'
for v in "" i4 i4g33 i8 g25; do
  echo "== variant ${v:-base}"
  if [ -n "$v" ]; then export B2_LIB=go_webp_b200/lib/variants/libb2webp_$v.so; fi
  python tools/mixed_classes.py 130,1080 512,1080 1920,1080 1921,1080 8191,1080 2>&1 | python -c "$fmt"
  B2_MIXED_BREAKDOWN=1 python tools/mixed_bench.py 2>&1 | python -c "$fmt" 
  B2_ARGB_FRAMES=64 B2_OPS=inverse_chain,forward_chain,color_inverse,add_green python tools/quick_bench.py vp8l 2>&1 | grep "chain\|color_inverse\|green" 
done > gpurun_out/flat_variants.txt 2>&1
