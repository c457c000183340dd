#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x -k "filter or rescaled or alpha" 2>&1 | tail -3 | cut -c1-300
for w in 1920 1922 1921; do echo "== $w"; B2_FILT_W=$w python tools/quick_bench.py dequant 2>&1 | grep "random bytes s=100"; done
