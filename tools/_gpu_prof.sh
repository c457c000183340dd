#!/bin/bash
B2_FRAMES=32 ncu --set full --import-source on --clock-control none -k regex:gradient -c 1 -s 3 -o gpurun_out/r02b_grad_final python tools/quick_bench.py filters > gpurun_out/r02b_grad_ncu.log 2>&1
B2_FRAMES=16 ncu --set full --import-source on --clock-control none -k regex:shrink_fast -c 1 -s 1 -o gpurun_out/r02b_rescale_shrink_final python tools/quick_bench.py rescale > gpurun_out/r02b_rs_ncu.log 2>&1
tools/grad_trace > gpurun_out/r02b_grad_trace.txt
python tools/quick_bench.py filters rescale > gpurun_out/r02b_quick_filters_rescale.txt 2>&1
tail -9 gpurun_out/r02b_quick_filters_rescale.txt
