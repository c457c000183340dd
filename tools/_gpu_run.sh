#!/bin/bash
python tools/quick_bench.py rescale 2>&1 | tail -3
for v in e6 e9; do echo $v; B2_LIB=go_webp_b200/lib/variants/libb2webp_$v.so python tools/quick_bench.py rescale 2>&1 | tail -3; done
