#!/bin/bash
N="--set full --import-source on --clock-control none -c 1"
ncu $N -k regex:upsample_bulk_kernel -s 47 -o gpurun_out/r02c_uprgb python tools/quick_bench.py up1080 > gpurun_out/r02c_prof9.log 2>&1
ncu $N -k regex:upsample_bulk_kernel -s 1 -o gpurun_out/r02c_uprgba python tools/quick_bench.py up1080 > gpurun_out/r02c_prof10.log 2>&1
ls -la gpurun_out/r02c_uprgb.ncu-rep gpurun_out/r02c_uprgba.ncu-rep
