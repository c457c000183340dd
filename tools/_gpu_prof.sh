#!/bin/bash
N="--set full --import-source on --clock-control none -c 1"
B2_FRAMES=128 ncu $N -k regex:alpha_filter_kernel -s 1 -o gpurun_out/r02c_fwdh python tools/quick_bench.py filters > gpurun_out/r02c_prof7.log 2>&1
B2_FRAMES=128 ncu $N -k regex:alpha_unfilter_vertical -s 1 -o gpurun_out/r02c_unfv python tools/quick_bench.py filters > gpurun_out/r02c_prof8.log 2>&1
B2_FRAMES=128 ncu $N -k "regex:upsample_bulk_kernel<3" -s 1 -o gpurun_out/r02c_uprgb python tools/quick_bench.py up1080 > gpurun_out/r02c_prof9.log 2>&1
ls -la gpurun_out/r02c_fwdh.ncu-rep gpurun_out/r02c_unfv.ncu-rep gpurun_out/r02c_uprgb.ncu-rep
