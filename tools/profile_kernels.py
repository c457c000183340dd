#!/usr/bin/env python
"""Short driver for ncu: launches every hot-path kernel a few times on device-resident
synthetic batches (inputs larger than L2 where it matters).
   ncu --set full -k regex:... python tools/profile_kernels.py [import upsample vp8l predictor tools dither filters]"""
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
if "dither" in which:
    p = ImportPipeline(ctx, 64, 1920, 1080, host=False)
    p.fill_synthetic(0xB2000001, 0, 2)
    for _ in range(2):
        p.run_device(0, 0.5)
    ctx.sync()
    p.close()
if "upsample" in which:
    for (n, w, h) in ((16, 3840, 2160), (64, 1920, 1080)):
        p = UpsamplePipeline(ctx, n, w, h, _abi.MODE_RGBA, host=False)
        p.fill_synthetic(0xB2000004)
        for _ in range(2):
            p.run_device()
        ctx.sync()
        p.close()
if "vp8l" in which:
    for op, bits in (("inverse_chain", 5), ("add_green", 0), ("color_index_inverse", 0), ("color_index_inverse", 2)):
        p = ArgbPipeline(ctx, 8, 4096, 4096, op, bits)
        p.fill_synthetic(0xB2000003)
        for _ in range(2):
            p.run_device()
        ctx.sync()
        p.close()
if "predictor" in which:
    for op in ("predictor_forward", "predictor_inverse"):
        p = ArgbPipeline(ctx, 16, 4096, 4096, op, 5)
        p.fill_synthetic(0xB2000003)
        for _ in range(2):
            p.run_device()
        ctx.sync()
        p.close()
if "tools" in which:
    n = 64
    p = ImportPipeline(ctx, n, 1920, 1080, host=False)
    p.fill_synthetic(0xB2000001, 0, 3)
    p.run_device()
    ctx.sync()
    jobs = (_abi.PictureJob * n)(*[_abi.PictureJob(j.y, j.u, j.v, j.a, None, j.y_stride, j.uv_stride, j.a_stride, 0, j.width, j.height)
                                   for j in p.jobs])
    for _ in range(2):
        ctx._chk(ctx.lib.b2_picture_cleanup_transparent(ctx.h, 0, jobs, n, 0))
    for _ in range(2):
        ctx._chk(ctx.lib.b2_picture_blend_alpha(ctx.h, 0, jobs, n, 0, 0x336699))
    ctx.sync()
    p.close()
print("done, launches:", ctx.launches)
