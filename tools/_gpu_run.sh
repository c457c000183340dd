#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x -k "alpha_multipl or palette" 2>&1 | tail -3 | cut -c1-300
