#!/usr/bin/env python
"""Device-resident timing of every hot-path kernel (CUDA events, inputs > L2)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from go_webp_b200 import _abi  # noqa: E402
if os.environ.get("B2_LIB"):  # experiment builds (tools/variants.sh)
    _abi.LIB_PATH = os.environ["B2_LIB"]
from go_webp_b200.device import Context  # noqa: E402
from go_webp_b200.pipeline import ArgbPipeline, ImportPipeline, UpsamplePipeline  # noqa: E402

PEAK = 6552.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
which = sys.argv[1:] or ["import", "upsample", "vp8l"]
NF = int(os.environ.get("B2_FRAMES", "128"))  # 1080p frames of the import runs
ctx = Context(0, 1)


def timed(run, n=20):
    for _ in range(3):
        run()
    ctx.sync()
    ctx.event_record(0)
    for _ in range(n):
        run()
    ctx.event_record(1)
    ctx.sync()
    return ctx.elapsed_ms(0, 1) / n


def report(name, px, bpp, ms):
    print(f"{name:42s} {px / ms / 1e3:10.0f} Mpx/s  {px * bpp / ms / 1e6:7.0f} GB/s  {px * bpp / ms / 1e6 / PEAK:6.3f} of measured peak")


if "import" in which:
    for name, mode in (("opaque", 1), ("sparse", 2), ("raw", 0)):
        p = ImportPipeline(ctx, NF, 1920, 1080, host=False)
        p.fill_synthetic(0xB2000001, 0, mode)
        report(f"import {NF}x1080p alpha={name}", NF * 1920 * 1080, 6.5, timed(p.run_device))
        p.close()
    p = ImportPipeline(ctx, NF, 1920, 1080, host=False)
    p.fill_synthetic(0xB2000001, 0, 2)
    report(f"import {NF}x1080p alpha=sparse DITHERED 0.5", NF * 1920 * 1080, 6.5, timed(lambda: p.run_device(0, 0.5)))
    p.close()
    p = ImportPipeline(ctx, NF, 1920, 1080, import_alpha=False, host=False)
    p.fill_synthetic(0xB2000001, 0, 0)
    report(f"import {NF}x1080p RGBX (A written 0xff)", NF * 1920 * 1080, 6.5, timed(p.run_device))
    p.close()
if "upsample" in which:
    for mode, wa, bpp in ((_abi.MODE_RGBA, False, 5.5), (_abi.MODE_BGRA, True, 6.5), (_abi.MODE_RGB, False, 4.5)):
        p = UpsamplePipeline(ctx, 32, 3840, 2160, mode, with_alpha=wa, host=False)
        p.fill_synthetic(0xB2000004)
        report(f"upsample 32x4K mode={mode} alpha={wa}", 32 * 3840 * 2160, bpp, timed(p.run_device))
        p.close()
    p = UpsamplePipeline(ctx, 128, 1920, 1080, _abi.MODE_RGBA, host=False)
    p.fill_synthetic(0xB2000004)
    report("upsample 128x1080p RGBA", 128 * 1920 * 1080, 5.5, timed(p.run_device))
    p.close()
if "vp8l" in which:
    for op, bits, bpp in (("inverse_chain", 5, 8.004), ("forward_chain", 5, 8.004), ("color_inverse", 5, 8.004),
                          ("add_green", 0, 8.0), ("sub_green", 0, 8.0), ("color_index_inverse", 0, 8.0),
                          ("color_index_inverse", 1, 6.0), ("color_index_inverse", 2, 5.0), ("color_index_inverse", 3, 4.5),
                          ("predictor_forward", 5, 8.004), ("predictor_inverse", 5, 8.004)):
        if os.environ.get("B2_OPS") and op not in os.environ["B2_OPS"].split(","):
            continue
        nf = int(os.environ.get("B2_ARGB_FRAMES", "8"))
        aw, ah = int(os.environ.get("B2_ARGB_W", "4096")), int(os.environ.get("B2_ARGB_H", "4096"))
        p = ArgbPipeline(ctx, nf, aw, ah, op, bits)
        p.fill_synthetic(0xB2000003)
        ms = timed(p.run_device, 5 if "predictor" in op else 20)
        report(f"vp8l {op} bits={bits} {nf}x{aw}x{ah} ({ms * 1e3:.0f} us)", nf * aw * ah, bpp, ms)
        p.close()
if "tools" in which:
    # picture pre-processing on the planes an import just wrote (binary alpha: every 8x8 block is
    # partially transparent, the SmoothenBlock stress case: read A + Y, rewrite Y = 3 B/px; blend:
    # read A, Y, U, V, write all four = 5 B/px)
    import ctypes as C
    p = ImportPipeline(ctx, NF, 1920, 1080, host=False)
    p.fill_synthetic(0xB2000001, 0, 3)
    p.run_device()
    ctx.sync()
    jobs = (_abi.PictureJob * NF)(*[_abi.PictureJob(j.y, j.u, j.v, j.a, None, j.y_stride, j.uv_stride, j.a_stride, 0, j.width, j.height)
                                     for j in p.jobs])
    report(f"cleanup transparent {NF}x1080p YUVA (binary alpha)", NF * 1920 * 1080, 3.0,
           timed(lambda: ctx._chk(ctx.lib.b2_picture_cleanup_transparent(ctx.h, 0, jobs, NF, 0))))
    po = ImportPipeline(ctx, NF, 1920, 1080, host=False)
    po.fill_synthetic(0xB2000001, 0, 1)
    po.run_device()
    ctx.sync()
    ojobs = (_abi.PictureJob * NF)(*[_abi.PictureJob(j.y, j.u, j.v, j.a, None, j.y_stride, j.uv_stride, j.a_stride, 0, j.width, j.height)
                                      for j in po.jobs])
    report(f"cleanup transparent {NF}x1080p YUVA (opaque: reads A only)", NF * 1920 * 1080, 1.0,
           timed(lambda: ctx._chk(ctx.lib.b2_picture_cleanup_transparent(ctx.h, 0, ojobs, NF, 0))))
    po.close()
    report(f"blend alpha {NF}x1080p YUVA", NF * 1920 * 1080, 5.0,
           timed(lambda: ctx._chk(ctx.lib.b2_picture_blend_alpha(ctx.h, 0, jobs, NF, 0, 0x336699))))
    p.close()
    a = ArgbPipeline(ctx, 8, 4096, 4096, "add_green", 0)
    a.fill_synthetic(0xB2000003)
    aj = (_abi.PictureJob * 8)(*[_abi.PictureJob(None, None, None, None, j.src, 0, 0, 0, 4096, 4096, 4096) for j in a.jobs])
    report("cleanup transparent 8x4096^2 ARGB (random alpha: read only)", 8 * 4096 * 4096, 4.0,
           timed(lambda: ctx._chk(ctx.lib.b2_picture_cleanup_transparent(ctx.h, 0, aj, 8, 1))))
    report("blend alpha 8x4096^2 ARGB", 8 * 4096 * 4096, 8.0,
           timed(lambda: ctx._chk(ctx.lib.b2_picture_blend_alpha(ctx.h, 0, aj, 8, 1, 0x336699))))
    report("replace transparent 8x4096^2 ARGB", 8 * 4096 * 4096, 8.0,
           timed(lambda: ctx._chk(ctx.lib.b2_picture_replace_transparent(ctx.h, 0, aj, 8, 0x112233))))
    a.close()
if "filters" in which:
    # alpha-plane filters on NF 1080p planes (read + write = 2 B/px)
    n = NF
    w = int(os.environ.get("B2_FILT_W", "1920"))   # packed rows (stride = width): B2_FILT_W=1922 / 1921 for rows that are not 16-byte aligned
    h, st = 1080, w
    src = ctx.alloc(n * st * h + 16)
    dst = ctx.alloc(n * st * h + 16)
    ctx.synth_fill(src.ptr, (n * st * h + 3) // 4, 1, 4 * ((n * st * h + 3) // 4), 4 * ((n * st * h + 3) // 4), 1, 0xB2000007, 0, 4)
    ctx.sync()
    for f, name in ((1, "horizontal"), (2, "vertical"), (3, "gradient")):
        jobs = (_abi.AlphaFilterJob * n)(*[_abi.AlphaFilterJob(src.ptr + i * st * h, dst.ptr + i * st * h, st, st, w, h, f, 0) for i in range(n)])
        report(f"alpha unfilter {name} {n}x1080p", n * w * h, 2.0, timed(lambda: ctx._chk(ctx.lib.b2_alpha_unfilter(ctx.h, 0, jobs, n)), 5))
        report(f"alpha filter   {name} {n}x1080p", n * w * h, 2.0, timed(lambda: ctx._chk(ctx.lib.b2_alpha_filter(ctx.h, 0, jobs, n)), 5))
    src.free(); dst.free()
if "rescale" in which:
    # rescaler: NF luma planes 3840x2160 -> 1920x1080 (shrink both ways) and 1920x1080 -> 3840x2160 (expand)
    n = min(NF, 64)
    for (w, h, dw, dh) in ((3840, 2160, 1920, 1080), (1920, 1080, 3840, 2160), (1920, 1080, 1280, 720)):
        src = ctx.alloc(n * w * h)
        dst = ctx.alloc(n * dw * dh)
        ctx.synth_fill(src.ptr, w // 4, h, w, w * h, n, 0xB2000008, 0, 4)
        ctx.sync()
        jobs = (_abi.RescaleJob * n)(*[_abi.RescaleJob(src.ptr + i * w * h, dst.ptr + i * dw * dh, w, h, w, dw, dh, dw, 1, 0) for i in range(n)])
        ms = timed(lambda: ctx._chk(ctx.lib.b2_rescale_planes(ctx.h, 0, jobs, n)), 5)
        print(f"rescale {n} x {w}x{h} -> {dw}x{dh}: {n * w * h / ms / 1e3:9.0f} Mpx/s in, {n * (w * h + dw * dh) / ms / 1e6:7.0f} GB/s "
              f"({n * (w * h + dw * dh) / ms / 1e6 / PEAK:.3f} of measured peak)")
        src.free(); dst.free()
if "dequant" in which:
    # alpha level smoothing on NF 1080p planes (3 B/px: the level scan reads the plane, the smoother reads + writes it)
    import numpy as np
    n = NF
    w = int(os.environ.get("B2_FILT_W", "1920"))   # packed rows (stride = width): B2_FILT_W=1922 / 1921 for rows that are not 16-byte aligned
    h, st = 1080, w
    src = ctx.alloc(n * st * h + 16)
    dst = ctx.alloc(n * st * h)
    yy, xx = np.mgrid[0:h, 0:w]
    smooth = ((128 + 100 * np.sin(xx / 90.0) * np.cos(yy / 70.0)).astype(np.int64) // 16 * 16).astype(np.uint8)
    edge = (np.clip((xx - 900) * 2 + 0 * yy, 0, 255) // 16 * 16).astype(np.uint8)   # opaque | 128-column ramp | transparent
    edge[edge >= 240] = 255
    for name in ("random bytes", "smooth, 16 levels apart", "flat with a soft edge"):
        if name == "random bytes":
            ctx.synth_fill(src.ptr, (n * st * h + 3) // 4, 1, 4 * ((n * st * h + 3) // 4), 4 * ((n * st * h + 3) // 4), 1, 0xB2000009, 0, 4)
        else:
            host = smooth if name.startswith("smooth") else edge
            for i in range(n):
                ctx._chk(ctx.lib.b2_h2d_2d_async(ctx.h, 0, src.ptr + i * st * h, st, host.ctypes.data, w, w, h))
        ctx.sync()
        for strength in (100, 50):
            jobs = (_abi.AlphaDequantJob * n)(*[_abi.AlphaDequantJob(src.ptr + i * st * h, dst.ptr + i * st * h, st, st, w, h, strength, 0) for i in range(n)])
            for dither in (1, 0):
                report(f"level smoothing {n}x1080p {name} s={strength} dither={dither}", n * w * h, 3.0,
                       timed(lambda: ctx._chk(ctx.lib.b2_alpha_dequantize_levels(ctx.h, 0, jobs, n, dither)), 5))
    src.free(); dst.free()
if "convert" in which:
    # output conversions (8f-2) on NF 1080p pictures: ARGB words in, packed bytes out; alpha multiplies in place;
    # point-sampled and 4:4:4 YUV -> RGBA; the use_argb import (RGBA bytes -> ARGB words); animation blend
    n = NF
    w, h = int(os.environ.get("B2_CONV_W", "1920")), 1080   # widths that are not multiples of 4: B2_CONV_W=1924 / 1922 (even: the chroma planes stay whole)
    argb = ctx.alloc(n * 4 * w * h)
    out = ctx.alloc(n * 4 * w * h)
    yp = ctx.alloc(n * w * h)
    up = ctx.alloc(n * w * h)
    vp = ctx.alloc(n * w * h)
    ctx.synth_fill(argb.ptr, w, h, 4 * w, 4 * w * h, n, 0xB200000A, 0, 0)
    for pl, tag in ((yp, 1), (up, 2), (vp, 3)):
        ctx.synth_fill(pl.ptr, (w * h) // 4, 1, w * h, w * h, n, 0xB200000A + tag, 0, 4)   # a plane as one row of words: any width
    ctx.sync()
    CJ = _abi.ConvertJob
    for mode, name, bpp in ((_abi.MODE_RGBA, "RGBA", 4), (_abi.MODE_RGB, "RGB", 3), (_abi.MODE_RGB_565, "RGB565", 2), (_abi.MODE_rgbA, "rgbA (premultiplied)", 4)):
        jobs = (CJ * n)(*[CJ(argb.ptr + i * 4 * w * h, None, None, out.ptr + i * 4 * w * h, w, 0, bpp * w, w, h, 0) for i in range(n)])
        report(f"VP8LConvertFromBGRA -> {name} {n}x1080p", n * w * h, 4.0 + bpp,
               timed(lambda: ctx._chk(ctx.lib.b2_vp8l_convert_from_bgra(ctx.h, 0, jobs, n, mode, 0)), 5))
    jobs = (CJ * n)(*[CJ(None, None, None, argb.ptr + i * 4 * w * h, 0, 0, 4 * w, w, h, 0) for i in range(n)])
    report(f"WebPApplyAlphaMultiply (in place) {n}x1080p", n * w * h, 8.0, timed(lambda: ctx._chk(ctx.lib.b2_alpha_multiply(ctx.h, 0, jobs, n, _abi.ALPHA_APPLY, 0)), 5))
    report(f"WebPMultARGBRows (in place) {n}x1080p", n * w * h, 8.0, timed(lambda: ctx._chk(ctx.lib.b2_alpha_multiply(ctx.h, 0, jobs, n, _abi.ALPHA_MULT_ARGB, 0)), 5))
    jobs = (CJ * n)(*[CJ(None, up.ptr + i * w * h, None, yp.ptr + i * w * h, 0, w, w, w, h, 0) for i in range(n)])
    report(f"WebPMultRows (plane, in place) {n}x1080p", n * w * h, 3.0, timed(lambda: ctx._chk(ctx.lib.b2_alpha_multiply(ctx.h, 0, jobs, n, _abi.ALPHA_MULT_ROW, 0)), 5))
    jobs = (CJ * n)(*[CJ(None, up.ptr + i * w * h, None, out.ptr + i * 4 * w * h, 0, w, 4 * w, w, h, 0) for i in range(n)])
    report(f"WebPDispatchAlpha {n}x1080p", n * w * h, 2.0, timed(lambda: ctx._chk(ctx.lib.b2_alpha_multiply(ctx.h, 0, jobs, n, _abi.ALPHA_DISPATCH, 0)), 5))
    jobs = (CJ * n)(*[CJ(yp.ptr + i * w * h, up.ptr + i * w * h, vp.ptr + i * w * h, out.ptr + i * 4 * w * h, w, w, 4 * w, w, h, 0) for i in range(n)])
    report(f"YUV444 -> RGBA {n}x1080p", n * w * h, 7.0, timed(lambda: ctx._chk(ctx.lib.b2_yuv_to_rgb_point(ctx.h, 0, jobs, n, _abi.MODE_RGBA, 1)), 5))
    jobs = (CJ * n)(*[CJ(yp.ptr + i * w * h, up.ptr + i * w * h, vp.ptr + i * w * h, out.ptr + i * 4 * w * h, w, w // 2, 4 * w, w, h, 0) for i in range(n)])
    report(f"point-sampled YUV420 -> RGBA {n}x1080p", n * w * h, 5.5, timed(lambda: ctx._chk(ctx.lib.b2_yuv_to_rgb_point(ctx.h, 0, jobs, n, _abi.MODE_RGBA, 0)), 5))
    jobs = (CJ * n)(*[CJ(argb.ptr + i * 4 * w * h, None, None, out.ptr + i * 4 * w * h, 4 * w, 0, 4 * w, w, h, 0) for i in range(n)])
    report(f"import RGBA -> ARGB (use_argb) {n}x1080p", n * w * h, 8.0, timed(lambda: ctx._chk(ctx.lib.b2_import_to_argb(ctx.h, 0, jobs, n, _abi.ORDER_RGBA, 1)), 5))
    AB = _abi.AnimBlendJob
    jobs = (AB * n)(*[AB(argb.ptr + i * 4 * w * h, out.ptr + i * 4 * w * h, w, w, w, h) for i in range(n)])
    for premult in (0, 1):
        report(f"anim blend rows premult={premult} {n}x1080p", n * w * h, 12.0, timed(lambda: ctx._chk(ctx.lib.b2_anim_blend_rows(ctx.h, 0, jobs, n, premult)), 5))
    for b in (argb, out, yp, up, vp):
        b.free()
if "palette" in which:
    from go_webp_b200.pipeline import PalettePipeline
    for psize, bpp in ((256, 8.0), (16, 4.5), (4, 4.25), (2, 4.125)):
        nf = int(os.environ.get("B2_ARGB_FRAMES", "16"))
        p = PalettePipeline(ctx, nf, 4096, 4096, psize)
        p.fill_synthetic(0xB2000003)
        ms = timed(p.run_device, 10)
        report(f"apply_palette {psize} colours xbits={p.xbits} {nf}x4096^2 ({ms * 1e3:.0f} us)", nf * 4096 * 4096, bpp, ms)
        p.close()
if "up1080" in which:
    for mode, wa, bpp, name in ((_abi.MODE_RGBA, False, 5.5, "RGBA"), (_abi.MODE_BGRA, True, 6.5, "BGRA+A"), (_abi.MODE_RGB, False, 4.5, "RGB")):
        p = UpsamplePipeline(ctx, 256, 1920, 1080, mode, with_alpha=wa, host=False)
        p.fill_synthetic(0xB2000004)
        report(f"upsample 256x1080p {name}", 256 * 1920 * 1080, bpp, timed(p.run_device))
        p.close()
