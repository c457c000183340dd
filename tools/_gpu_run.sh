timeout 300 python -m pytest tests -m gpu -x -q -k "mixed_pipeline" 2>&1 | grep -E "Error|error|^E|line" | head -20
