timeout 900 python -m pytest tests -m gpu -x -q -k "maximum_dimension" 2>&1 | tail -12
