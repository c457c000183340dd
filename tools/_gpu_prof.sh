#!/bin/bash
N="--set full --import-source on --clock-control none -c 1"
B2_FRAMES=64 ncu $N -k regex:alpha_filter_kernel -s 17 -o gpurun_out/r02c_fwdgrad python tools/quick_bench.py filters > gpurun_out/r02c_prof1.log 2>&1
B2_FRAMES=64 ncu $N -k regex:yuv_point_kernel -s 1 -o gpurun_out/r02c_yuv444 python tools/quick_bench.py convert > gpurun_out/r02c_prof2.log 2>&1
B2_FRAMES=64 ncu $N -k regex:yuv_point_kernel -s 9 -o gpurun_out/r02c_yuvpoint python tools/quick_bench.py convert > gpurun_out/r02c_prof3.log 2>&1
B2_FRAMES=64 ncu $N -k regex:alpha_multiply_kernel -s 17 -o gpurun_out/r02c_multrow python tools/quick_bench.py convert > gpurun_out/r02c_prof4.log 2>&1
B2_ARGB_FRAMES=16 ncu $N -k regex:apply_palette_kernel -s 14 -o gpurun_out/r02c_pal16 python tools/quick_bench.py palette > gpurun_out/r02c_prof5.log 2>&1
ls -la gpurun_out/r02c_*.ncu-rep
