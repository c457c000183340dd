#!/bin/bash
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02_gpu_tests_f2.log
cat gpurun_out/r02_gpu_tests_f2.log
