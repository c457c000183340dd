timeout 600 python -m pytest tests -m gpu -x -q -k "rescaler or level_smoothing" 2>&1 | tail -12
