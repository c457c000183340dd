#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02_gpu_tests_f8.log
cat gpurun_out/r02_gpu_tests_f8.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('bench', d['value'], d['roofline']['frac'], d['e2e']['value'], d['other_configs']['configs[4]']['value'])
"
