timeout 600 python bench.py > gpurun_out/b1.json 2> gpurun_out/b1.err; python -c "import json; d=json.loads(open('gpurun_out/b1.json').read().strip().splitlines()[-1]); print(d['value'], d['clocks'])"
B2_CLOCKS=smi timeout 600 python bench.py > gpurun_out/b2.json 2> gpurun_out/b2.err; python -c "import json; d=json.loads(open('gpurun_out/b2.json').read().strip().splitlines()[-1]); print(d['value'], d['clocks'])"
tail -3 gpurun_out/b2.err
