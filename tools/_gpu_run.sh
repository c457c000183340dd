timeout 600 python -m pytest tests -m gpu -x -q -k "error_behaviour" 2>&1 | tail -8
