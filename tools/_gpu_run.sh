#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests_final.log 2>&1; tail -2 gpurun_out/r02_gpu_tests_final.log
python bench.py > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err; tail -c 200 gpurun_out/r02_bench_e.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_e_ref.json 2>> gpurun_out/r02_bench_e.err
head -c 300 gpurun_out/r02_bench_e_ref.json
