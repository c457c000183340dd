#!/usr/bin/env python
"""3-byte imports (WebPPictureImportRGB / BGR, picture_csp_enc.c.go step = 3): device-resident rate of
b2_rgba_to_yuv420a on NF 1080p RGB frames (3 B in + 1.5 B out per pixel)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from go_webp_b200 import _abi  # noqa: E402
if os.environ.get("B2_LIB"):
    _abi.LIB_PATH = os.environ["B2_LIB"]
from go_webp_b200.device import Context  # noqa: E402

NF = int(os.environ.get("B2_FRAMES", "128"))
PEAK = 6552.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def main():
    ctx = Context(0, 1)
    w, h = 1920, 1080
    al = lambda v: (v + 255) // 256 * 256  # noqa: E731
    src = ctx.alloc(NF * al(3 * w * h))
    y = ctx.alloc(NF * al(w * h))
    u = ctx.alloc(NF * al(w * h // 4))
    v = ctx.alloc(NF * al(w * h // 4))
    ctx.synth_fill(src.ptr, (3 * w) // 4, h, 3 * w, al(3 * w * h), NF, 0xB2000001, 0, 4)
    ctx.sync()
    jobs = (_abi.ImportJob * NF)(*[_abi.ImportJob(src.ptr + i * al(3 * w * h), y.ptr + i * al(w * h), u.ptr + i * al(w * h // 4), v.ptr + i * al(w * h // 4),
                                                  None, 3 * w, w, w // 2, 0, w, h) for i in range(NF)])
    for order, name in ((_abi.ORDER_RGB, "RGB"), (_abi.ORDER_BGR, "BGR")):
        for _ in range(3):
            ctx._chk(ctx.lib.b2_rgba_to_yuv420a(ctx.h, 0, jobs, NF, order, 0, None))
        ctx.sync(); ctx.event_record(0)
        for _ in range(10):
            ctx._chk(ctx.lib.b2_rgba_to_yuv420a(ctx.h, 0, jobs, NF, order, 0, None))
        ctx.event_record(1); ctx.sync()
        ms = ctx.elapsed_ms(0, 1) / 10
        px = NF * w * h
        print(f"import {NF}x1080p {name} (3 bytes per pixel)   {px / ms / 1e3:10.0f} Mpx/s  {px * 4.5 / ms / 1e6:7.0f} GB/s  {px * 4.5 / ms / 1e6 / PEAK:6.3f} of measured peak")


if __name__ == "__main__":
    main()
