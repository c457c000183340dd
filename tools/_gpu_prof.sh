#!/bin/bash
B2_FRAMES=32 ncu --set full --import-source on --clock-control none -k regex:gradient -c 1 -s 3 -o gpurun_out/r02_grad_prof python tools/quick_bench.py filters > gpurun_out/r02_grad_ncu.log 2>&1
tail -2 gpurun_out/r02_grad_ncu.log
