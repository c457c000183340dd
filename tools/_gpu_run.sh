#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x -k "import or config1 or config4 or dither" 2>&1 | tail -3 | cut -c1-300
python tools/mixed_classes.py 128,1080 130,1080 256,1080 1920,1080 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('%-22s import %.3f  upsample %.3f  chain %.3f' % (d['subset'], d['import']['frac'], d['upsample']['frac'], d['chain']['frac']))
"
python tools/quick_bench.py import 2>&1 | grep "sparse  \|opaque"
