timeout 600 python -m pytest tests -m gpu -x -q -k "import_to_argb or import" 2>&1 | tail -12
