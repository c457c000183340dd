#!/bin/bash
B2_MIXED_BREAKDOWN=1 python tools/mixed_bench.py > gpurun_out/r02_mixed_breakdown.txt 2>&1
python tools/mixed_classes.py >> gpurun_out/r02_mixed_breakdown.txt 2>&1
tail -3 gpurun_out/r02_mixed_breakdown.txt | cut -c1-200
