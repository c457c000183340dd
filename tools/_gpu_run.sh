#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -x -k "unaligned_source_rows or flat_items" 2>&1 | tail -15
