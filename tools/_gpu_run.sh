#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02_gpu_tests_f6.log
cat gpurun_out/r02_gpu_tests_f6.log
timeout 600 python bench.py > gpurun_out/r02_bench_f6.json 2> gpurun_out/r02_bench_f6.err
tail -c 200 gpurun_out/r02_bench_f6.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
