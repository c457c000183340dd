#!/usr/bin/env python
"""Short driver for ncu: launches every hot-path kernel 3 times on small
device-resident synthetic batches (inputs larger than L2 where it matters)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from go_webp_b200 import _abi  # noqa: E402
if os.environ.get("B2_LIB"):  # experiment builds (tools/variants.sh)
    _abi.LIB_PATH = os.environ["B2_LIB"]
from go_webp_b200.device import Context  # noqa: E402
from go_webp_b200.pipeline import ArgbPipeline, ImportPipeline, UpsamplePipeline  # noqa: E402

which = sys.argv[1:] or ["import", "upsample", "vp8l"]
ctx = Context(0, 1)
if "import" in which:
    for alpha, mode in (("sparse", 2), ("opaque", 1)):
        p = ImportPipeline(ctx, 64, 1920, 1080, host=False)
        p.fill_synthetic(0xB2000001, 0, mode)
        for _ in range(3):
            p.run_device()
        ctx.sync()
        p.close()
if "upsample" in which:
    p = UpsamplePipeline(ctx, 16, 3840, 2160, _abi.MODE_RGBA, host=False)
    p.fill_synthetic(0xB2000004)
    for _ in range(3):
        p.run_device()
    ctx.sync()
    p.close()
if "vp8l" in which:
    for op, bits in (("inverse_chain", 5), ("add_green", 0), ("color_index_inverse", 0), ("color_index_inverse", 2)):
        p = ArgbPipeline(ctx, 8, 4096, 4096, op, bits)
        p.fill_synthetic(0xB2000003)
        for _ in range(3):
            p.run_device()
        ctx.sync()
        p.close()
print("done, launches:", ctx.launches)
