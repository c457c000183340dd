#!/bin/bash
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench_steps5.csv python bench.py --steps 5 --warmup 3 > gpurun_out/r02_ncu_launches.log 2>&1
tail -c 200 gpurun_out/r02_ncu_launches.log
wc -l gpurun_out/r02_ncu_launches_bench_steps5.csv
ncu --set full --import-source on --clock-control none -k regex:import_fast_kernel -c 1 -s 4 -o gpurun_out/r02c_import_full python bench.py --steps 5 --warmup 3 > gpurun_out/r02c_import_full.log 2>&1
ls -la gpurun_out/r02c_import_full.ncu-rep
