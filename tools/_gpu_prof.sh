#!/bin/bash
B2_FRAMES=64 ncu --set full --import-source on --clock-control none -k regex:unfilter_rows -c 1 -s 1 -o gpurun_out/r02b_unfilter_rows python tools/quick_bench.py filters > gpurun_out/r02b_rows_ncu.log 2>&1
tail -2 gpurun_out/r02b_rows_ncu.log
