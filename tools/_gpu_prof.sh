#!/bin/bash
B2_FRAMES=64 ncu --set full --import-source on --clock-control none -k regex:cleanup_yuva -c 1 -s 2 -o gpurun_out/r02c_cleanup python tools/quick_bench.py tools > gpurun_out/r02c_cleanup_ncu.log 2>&1
tail -2 gpurun_out/r02c_cleanup_ncu.log
