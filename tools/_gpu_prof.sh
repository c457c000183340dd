#!/bin/bash
B2_ARGB_FRAMES=32 B2_OPS=predictor_forward ncu --set full --import-source on --clock-control none -k regex:predictor_forward -c 1 -s 2 -o gpurun_out/r02c_predfwd python tools/quick_bench.py vp8l > gpurun_out/r02c_predfwd_ncu.log 2>&1
tail -2 gpurun_out/r02c_predfwd_ncu.log
