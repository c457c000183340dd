#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x -k "convert or rescaled or 4444 or bgra or resident" 2>&1 | tail -2
python tools/quick_bench.py convert 2>&1 | grep -i "ConvertFromBGRA"
