"""Generate tests/golden/rescaled_decode_libwebp160.npz: the decoder's rescaled output paths
(EmitRescaledRGB / ExportRGB / EmitRescaledAlphaRGB / ExportAlpha and EmitRescaledYUV /
EmitRescaledAlphaYUV, pkg/libwebp/decoder/io_dec.c.go:232-515) pinned on the libwebp 1.6.0
decoder.

    python tests/golden/make_golden_rescaled_decode.py

A lossy(+alpha) file is decoded three ways by the binary: to MODE_YUVA without scaling (the
Y / U / V / A planes the emitters receive), and with options.use_scaling to an RGB-family
mode or to MODE_YUVA (what they must produce).  The oracle's composition -- rescale Y to the
output size, U and V from half size straight to the output size, WebPYUV444Converters, then
the rescaled alpha dispatched into the alpha byte and WebPApplyAlphaMultiply for the
premultiplied modes; for YUVA: luma premultiplied around its rescale as in WebPPictureRescale
-- must reproduce them.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import oracle  # noqa: E402
from oracle import libwebp_ref as ref  # noqa: E402
import make_golden_dequant as gd  # noqa: E402

MODE_RGB, MODE_RGBA, MODE_BGR, MODE_BGRA, MODE_ARGB, MODE_rgbA, MODE_bgrA, MODE_Argb, MODE_YUVA = 0, 1, 2, 3, 4, 7, 8, 9, 12
BPP = {MODE_RGB: 3, MODE_RGBA: 4, MODE_BGR: 3, MODE_BGRA: 4, MODE_ARGB: 4, MODE_rgbA: 4, MODE_bgrA: 4, MODE_Argb: 4}


def decode(webp, mode, scaled=None, bypass_filtering=0):
    """WebPDecode into `mode`; scaled = (width, height) turns options.use_scaling on."""
    L = ref.lib()
    cfg = gd.DecoderConfig()
    assert L.WebPInitDecoderConfigInternal(C.byref(cfg), 0x0209)
    cfg.output.colorspace = mode
    cfg.options.bypass_filtering = bypass_filtering
    if scaled is not None:
        cfg.options.use_scaling, cfg.options.scaled_width, cfg.options.scaled_height = 1, scaled[0], scaled[1]
    L.WebPDecode.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
    st = L.WebPDecode(webp, len(webp), C.byref(cfg))
    assert st == 0, st
    w, h = cfg.output.width, cfg.output.height
    if mode == MODE_YUVA:
        b = cfg.output.u.YUVA
        uvw, uvh = (w + 1) >> 1, (h + 1) >> 1
        out = tuple(ref._plane(p, s, ww, hh) for p, s, ww, hh in ((b.y, b.y_stride, w, h), (b.u, b.u_stride, uvw, uvh),
                                                                  (b.v, b.v_stride, uvw, uvh), (b.a, b.a_stride, w, h)))
    else:
        b = cfg.output.u.RGBA
        out = ref._plane(b.rgba, b.stride, w * BPP[mode], h)
    L.WebPFreeDecBuffer.argtypes = [C.c_void_p]
    L.WebPFreeDecBuffer(C.byref(cfg.output))
    return out


def main():
    rng = np.random.default_rng(0xB2008)
    store, names = {}, []
    n = 0
    for (w, h, dw, dh) in [(64, 48, 32, 24), (40, 30, 90, 70), (33, 17, 20, 40), (17, 33, 16, 33), (50, 20, 50, 19), (24, 24, 61, 7)]:
        yy, xx = np.mgrid[0:h, 0:w]
        rgba = np.zeros((h, w, 4), np.uint8)
        rgba[..., 0] = (xx * 255) // max(1, w - 1)
        rgba[..., 1] = (yy * 255) // max(1, h - 1)
        rgba[..., 2] = rng.integers(0, 256, (h, w))
        rgba[..., 3] = np.clip(255 - ((xx - w / 2) ** 2 + (yy - h / 2) ** 2) * 900 // (w * w + h * h) + rng.integers(-20, 21, (h, w)), 0, 255)
        webp = gd.encode(rgba, 100)
        # a scaled decode skips the loop filter when both dimensions shrink below 3/4 (WebPIoInitFromOptions)
        y, u, v, a = decode(webp, MODE_YUVA, bypass_filtering=int(dw < w * 3 // 4 and dh < h * 3 // 4))
        k = f"c{n}"
        n += 1
        names.append(k)
        for nm, arr in zip("yuva", (y, u, v, a)):
            store[f"{k}/{nm}"] = arr
        store[f"{k}/size"] = np.array([dw, dh], np.int32)
        for mode in (MODE_RGB, MODE_RGBA, MODE_BGR, MODE_BGRA, MODE_ARGB, MODE_rgbA, MODE_bgrA, MODE_Argb):
            exp = decode(webp, mode, (dw, dh))
            got = oracle.emit_rescaled_rgb(y, u, v, a, dw, dh, mode)
            assert (got == exp).all(), (w, h, dw, dh, mode, np.argwhere(got != exp)[:4])
            store[f"{k}/rgb{mode}"] = exp
        exp = decode(webp, MODE_YUVA, (dw, dh))
        got = oracle.picture_rescale_yuva(y, u, v, a, dw, dh)
        for e, g_, nm in zip(exp, got, "yuva"):
            assert (e == g_).all(), (w, h, dw, dh, "yuva", nm)
            store[f"{k}/out_{nm}"] = e
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "rescaled_decode_libwebp160.npz"), **store)
    print(f"{n} files x 9 output modes: oracle == libwebp decoder with use_scaling")


if __name__ == "__main__":
    main()
