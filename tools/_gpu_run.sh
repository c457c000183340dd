timeout 600 python -m pytest tests -m gpu -x -q -k "decoder_golden or upsample" 2>&1 | tail -8
