#!/bin/bash
B2_FRAMES=16 ncu --set full --import-source on --clock-control none -k regex:dq_smooth -c 1 -s 1 -o gpurun_out/r02b_dq_smooth python tools/quick_bench.py dequant > gpurun_out/r02b_dq_ncu.log 2>&1
tail -2 gpurun_out/r02b_dq_ncu.log
