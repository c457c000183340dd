#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02_gpu_tests_f7.log
cat gpurun_out/r02_gpu_tests_f7.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
