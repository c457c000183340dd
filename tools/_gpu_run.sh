timeout 600 python -m pytest tests -m gpu -x -q -k "rescaled_decoder or output_conversions or alpha_mult or convert" 2>&1 | tail -8
