#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02_gpu_tests_f3.log
cat gpurun_out/r02_gpu_tests_f3.log
timeout 600 python bench.py > gpurun_out/r02_bench_f3.json 2> gpurun_out/r02_bench_f3.err
tail -c 300 gpurun_out/r02_bench_f3.json
