#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x -k "convert_from_bgra or point_sampler" 2>&1 | tail -2
