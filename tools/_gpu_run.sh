#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -x -k "alpha_filters" 2>&1 | tail -5 | cut -c1-300
