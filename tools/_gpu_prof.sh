#!/bin/bash
B2_ARGB_FRAMES=64 B2_OPS=predictor_inverse ncu --set full --import-source on --clock-control none -k regex:predictor_inverse -c 1 -s 2 -o gpurun_out/r02_pred_prof python tools/quick_bench.py vp8l > gpurun_out/r02_pred_ncu.log 2>&1
tail -2 gpurun_out/r02_pred_ncu.log
