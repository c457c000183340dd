#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x -k "import" 2>&1 | tail -5 | cut -c1-300
python tools/rgb_import_bench.py 2>&1 | tail -3
python tools/quick_bench.py import 2>&1 | tail -5
