#!/bin/bash
# round-2 GPU check: new resident tests first, whole gpu suite, smoke, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_resident.py -x -q 2>&1 | tail -25 > gpurun_out/r02_resident_tests.log
cat gpurun_out/r02_resident_tests.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_gpu_tests.log
cat gpurun_out/r02_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
tail -c 600 gpurun_out/r02_bench_a.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_a.json').read().strip().splitlines()[-1])
print('value', d['value'], 'frac', d['roofline']['frac'], d['clocks'])
print('e2e', json.dumps(d['e2e'], indent=0)[:900])
print('sustained', d.get('sustained'))
for k,v in d.get('other_configs',{}).items():
    print(k, json.dumps(v)[:1200])
PY
