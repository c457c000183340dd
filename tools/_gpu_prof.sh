#!/bin/bash
N="--set full --import-source on --clock-control none -c 1"
B2_FRAMES=64 ncu $N -k regex:rescale_expand_fast -s 2 -o gpurun_out/r02c_expand python tools/quick_bench.py rescale > gpurun_out/r02c_prof6.log 2>&1
ls -la gpurun_out/r02c_expand.ncu-rep
