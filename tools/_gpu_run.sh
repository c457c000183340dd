#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -x -k "rescale" 2>&1 | tail -8
