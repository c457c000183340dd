#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -q -x -k "upsample or config3 or config4 or mixed or yuv or picture or staged or emit" 2>&1 | tail -5
fmt='
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print("%-22s import %.3f  upsample %.3f  chain %.3f" % (d["subset"], d["import"]["frac"], d["upsample"]["frac"], d["chain"]["frac"]))
'
python tools/mixed_classes.py 130,1080 512,1080 513,1080 1920,1080 1921,1080 8191,1080 2>&1 | python -c "$fmt"
B2_MIXED_BREAKDOWN=1 python tools/mixed_bench.py 2>&1 | python -c "$fmt"
python tools/quick_bench.py upsample up1080 2>&1 | tail -12
