#!/usr/bin/env python
"""Per-kernel rates of the configs[4] pipeline on UNIFORM batches of one size each, to see what a
width class costs on its own (row alignment, partial strips, narrow rows):
  python tools/mixed_classes.py [w,h ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from go_webp_b200 import _abi  # noqa: E402
if os.environ.get("B2_LIB"):
    _abi.LIB_PATH = os.environ["B2_LIB"]
from go_webp_b200.device import Context  # noqa: E402
import tools.mixed_bench as mb  # noqa: E402

DEFAULT = [(128, 1080), (130, 1080), (129, 1080), (512, 1080), (514, 1080), (513, 1080), (1920, 1080), (1922, 1080),
           (1921, 1080), (1921, 1081), (4096, 1080), (4098, 1080), (4097, 1080), (8192, 1080), (8190, 1080), (8191, 1080)]


def main():
    sizes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or DEFAULT
    ctx = Context(0, 1)
    target = int(os.environ.get("B2_CLASS_MPX", "256")) * 1_000_000
    for (w, h) in sizes:
        n = max(1, min(1024, target // (w * h)))
        mb.breakdown(ctx, [(w, h)] * n, "%dx%d x %d" % (w, h, n))


if __name__ == "__main__":
    main()
