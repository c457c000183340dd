#!/bin/bash
B2_FRAMES=16 ncu --set full --import-source on --clock-control none -k regex:expand_fast -c 1 -s 1 -o gpurun_out/r02_rescale_expand_fast python tools/quick_bench.py rescale > gpurun_out/r02_rs_ncu.log 2>&1
tail -2 gpurun_out/r02_rs_ncu.log
