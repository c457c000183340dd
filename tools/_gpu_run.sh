#!/bin/bash
echo "== calibrated + refine"; B2_MIXED_CALIBRATED=1 B2_MIXED_REFINE=4 python tools/mixed_bench.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if 'rates_GBps' in d: continue
    if 'refine_round' in d: print('   ', d); continue
    print(d.get('n_gpus'), 'imb', round(d.get('imbalance_max_over_mean',0),3), 'frac', round(d.get('frac_of_measured_peak_per_gpu',0),3))"
