"""Generate tests/golden/tools_libwebp160.npz: WebPCleanupTransparentArea and
WebPBlendAlpha (both exported by the libwebp 1.6.0 binary) on YUVA and ARGB pictures,
tests/golden/import_argb_libwebp160.npz: the use_argb branch of WebPPictureImport*,
import_dithered_libwebp160.npz (WebPPictureARGBToYUVADithered) and rescale_libwebp160.npz
(WebPPictureRescale).

    python tests/golden/make_golden_tools.py

Pins oracle/oracle_tools.c (pkg/libwebp/enc/picture_tools_enc.c.go:
110-183, 199-276).  Transparent regions are laid out in 8x8-aligned patches, runs of
patches, partial blocks and ragged right/bottom edges so that the need_reset carry, the
SmoothenBlock averages and the left-over handling are all exercised.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from oracle import libwebp_ref as ref  # noqa: E402

SIZES = [(8, 8), (16, 8), (9, 9), (24, 17), (33, 40), (64, 48), (71, 29), (7, 5), (100, 20)]


def alpha_plane(rng, w, h, kind):
    a = rng.integers(1, 256, (h, w), dtype=np.uint8)
    if kind == "opaque":
        a[:] = 255
    elif kind == "blocks":      # whole 8x8 blocks transparent, in runs
        bw, bh = (w + 7) // 8, (h + 7) // 8
        m = rng.random((bh, bw)) < 0.55
        a[np.kron(m, np.ones((8, 8), bool))[:h, :w]] = 0
    elif kind == "speckle":     # partially transparent blocks
        a[rng.random((h, w)) < 0.4] = 0
    elif kind == "mixed":
        bw, bh = (w + 7) // 8, (h + 7) // 8
        m = rng.random((bh, bw)) < 0.4
        a[np.kron(m, np.ones((8, 8), bool))[:h, :w]] = 0
        a[rng.random((h, w)) < 0.15] = 0
        a[rng.random((h, w)) < 0.3] = 255
    elif kind == "transp":
        a[:] = 0
    return a


def main():
    rng = np.random.default_rng(0xB2004)
    store, names = {}, []
    n = 0
    for (w, h) in SIZES:
        for kind in ("opaque", "blocks", "speckle", "mixed", "transp"):
            uvw, uvh = (w + 1) >> 1, (h + 1) >> 1
            y = rng.integers(0, 256, (h, w), dtype=np.uint8)
            u = rng.integers(0, 256, (uvh, uvw), dtype=np.uint8)
            v = rng.integers(0, 256, (uvh, uvw), dtype=np.uint8)
            a = alpha_plane(rng, w, h, kind)
            argb = (rng.integers(0, 1 << 24, (h, w), dtype=np.uint32)) | (a.astype(np.uint32) << 24)
            bg = int(rng.integers(0, 1 << 24))
            k = f"c{n}"
            names.append(k)
            n += 1
            store[f"{k}/y"], store[f"{k}/u"], store[f"{k}/v"], store[f"{k}/a"], store[f"{k}/argb"] = y, u, v, a, argb
            store[f"{k}/bg"] = np.uint32(bg)
            cy, cu, cv = ref.cleanup_transparent_yuva(y, u, v, a)
            oy, ou, ov = oracle.cleanup_transparent_yuva(y, u, v, a)
            assert (cy == oy).all() and (cu == ou).all() and (cv == ov).all(), (k, "cleanup yuva")
            store[f"{k}/clean_y"], store[f"{k}/clean_u"], store[f"{k}/clean_v"] = cy, cu, cv
            ca = ref.cleanup_transparent_argb(argb)
            assert (ca == oracle.cleanup_transparent_argb(argb)).all(), (k, "cleanup argb")
            store[f"{k}/clean_argb"] = ca
            by, bu, bv, ba = ref.blend_alpha_yuva(y, u, v, a, bg)
            qy, qu, qv, qa = oracle.blend_alpha_yuva(y, u, v, a, bg)
            assert (by == qy).all() and (bu == qu).all() and (bv == qv).all() and (ba == qa).all(), (k, "blend yuva")
            store[f"{k}/blend_y"], store[f"{k}/blend_u"], store[f"{k}/blend_v"], store[f"{k}/blend_a"] = by, bu, bv, ba
            bb = ref.blend_alpha_argb(argb, bg)
            assert (bb == oracle.blend_alpha_argb(argb, bg)).all(), (k, "blend argb")
            store[f"{k}/blend_argb"] = bb
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "tools_libwebp160.npz"), **store)
    print(f"{n} cases: oracle == libwebp binary for cleanup (YUVA, ARGB) and blend (YUVA, ARGB)")


def gen_import_argb():
    """import_argb_libwebp160.npz: WebPPictureImport{RGBA,RGBX,BGRA,BGRX,RGB,BGR} with
    use_argb = 1 (Import(), pkg/libwebp/enc/picture_csp_enc.c.go:138-181)."""
    rng = np.random.default_rng(5)
    store, names = {}, []
    n = 0
    for (w, h) in [(1, 1), (2, 3), (5, 7), (16, 9), (33, 16), (64, 20), (127, 5)]:
        for order in ("RGBA", "RGBX", "BGRA", "BGRX", "RGB", "BGR"):
            step = 3 if order in ("RGB", "BGR") else 4
            src = rng.integers(0, 256, (h, w * step), dtype=np.uint8)
            e = ref.import_to_argb(src, w, h, order)
            assert (e == oracle.import_to_argb(src, w, h, order)).all(), (w, h, order)
            k = f"c{n}"
            n += 1
            names.append(k)
            store[f"{k}/src"], store[f"{k}/argb"], store[f"{k}/order"] = src, e, np.array(order)
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "import_argb_libwebp160.npz"), **store)
    print(f"{n} cases: oracle == libwebp binary for the use_argb import")


def gen_dithered():
    """import_dithered_libwebp160.npz: WebPPictureARGBToYUVADithered (exported by the binary;
    pkg/picture/conversions.go:122-165 with VP8Random, pkg/vp8/random.go).  Translucent
    blocks follow the binary's inverse-alpha table (oracle DIV_TABLE)."""
    rng = np.random.default_rng(11)
    store, names = {}, []
    n = 0
    for (w, h) in [(1, 1), (2, 2), (5, 7), (16, 9), (33, 16), (64, 33), (127, 20), (300, 11)]:
        for kind in ("opaque", "sparse", "rand"):
            for d in (0.05, 0.5, 1.0):
                argb = rng.integers(0, 1 << 32, (h, w), dtype=np.uint32)
                if kind == "opaque":
                    argb |= 0xff000000
                elif kind == "sparse":
                    argb[rng.random((h, w)) < 0.7] |= 0xff000000
                y, u, v, a, ha = ref.argb_to_yuva_dithered(argb, d)
                oy, ou, ov, oa, oha = oracle.import_yuva_dithered(argb.view(np.uint8).reshape(h, 4 * w), w, h, d, "BGRA",
                                                                  divide_mode=oracle.DIV_TABLE)
                assert ha == oha and (y == oy).all() and (u == ou).all() and (v == ov).all(), (w, h, kind, d)
                assert not ha or (a == oa).all()
                k = f"c{n}"
                n += 1
                names.append(k)
                store[f"{k}/argb"], store[f"{k}/d"], store[f"{k}/y"], store[f"{k}/u"], store[f"{k}/v"] = argb, np.float32(d), y, u, v
                store[f"{k}/ha"] = np.bool_(ha)
                if ha:
                    store[f"{k}/a"] = a
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "import_dithered_libwebp160.npz"), **store)
    print(f"{n} cases: oracle == libwebp binary for the dithered import")


def gen_rescale():
    """rescale_libwebp160.npz: WebPPictureRescale (exported) on YUV420 pictures without
    alpha: RescalePlane on the Y, U and V planes (pkg/libwebp/enc/picture_rescale_enc.c.go:
    125-136, 149-200), every combination of horizontal / vertical expansion and shrinking."""
    rng = np.random.default_rng(3)
    store, names = {}, []
    n = 0
    for (w, h, dw, dh) in [(64, 48, 32, 24), (64, 48, 100, 70), (33, 17, 20, 40), (50, 50, 50, 50), (7, 9, 3, 2), (3, 2, 9, 11),
                           (100, 3, 37, 3), (1, 1, 5, 5), (20, 30, 1, 1), (120, 80, 119, 81), (16, 16, 17, 15), (300, 40, 77, 111),
                           (2, 64, 2, 64), (1, 40, 2, 40)]:
        y = rng.integers(0, 256, (h, w), dtype=np.uint8)
        u = rng.integers(0, 256, ((h + 1) // 2, (w + 1) // 2), dtype=np.uint8)
        v = rng.integers(0, 256, ((h + 1) // 2, (w + 1) // 2), dtype=np.uint8)
        ey, eu, ev = ref.picture_rescale_yuv(y, u, v, dw, dh)
        for src, exp in ((y, ey), (u, eu), (v, ev)):
            assert (oracle.rescale_plane(src, exp.shape[1], exp.shape[0]) == exp).all(), (w, h, dw, dh)
            k = f"c{n}"
            n += 1
            names.append(k)
            store[f"{k}/src"], store[f"{k}/dst"] = src, exp
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "rescale_libwebp160.npz"), **store)
    print(f"{n} planes: oracle == libwebp binary for the rescaler")


def gen_rescale_alpha():
    """rescale_alpha_libwebp160.npz: WebPPictureRescale on YUV420A and ARGB pictures: the
    premultiplying branches around RescalePlane (picture_rescale_enc.c.go:138-147, 174-212:
    WebPMultRows / WebPMultARGBRows before, their inverses after)."""
    rng = np.random.default_rng(5)
    store, names = {}, []
    n = 0
    for (w, h, dw, dh) in [(64, 48, 32, 24), (40, 30, 90, 70), (33, 17, 20, 40), (7, 9, 3, 2), (3, 2, 9, 11), (1, 1, 4, 4),
                           (120, 80, 119, 81), (50, 20, 50, 20)]:
        for alpha_kind in ("random", "binary", "smooth"):
            yy, xx = np.mgrid[0:h, 0:w]
            a = {"random": rng.integers(0, 256, (h, w)), "binary": rng.integers(0, 2, (h, w)) * 255,
                 "smooth": (xx * 255) // max(1, w - 1) + 0 * yy}[alpha_kind].astype(np.uint8)
            y = rng.integers(0, 256, (h, w), dtype=np.uint8)
            u = rng.integers(0, 256, ((h + 1) // 2, (w + 1) // 2), dtype=np.uint8)
            v = rng.integers(0, 256, ((h + 1) // 2, (w + 1) // 2), dtype=np.uint8)
            exp = ref.picture_rescale_yuva(y, u, v, a, dw, dh)
            got = oracle.picture_rescale_yuva(y, u, v, a, dw, dh)
            for e, g_ in zip(exp, got):
                assert (e == g_).all(), ("yuva", w, h, dw, dh, alpha_kind)
            argb = rng.integers(0, 2**32, (h, w), dtype=np.uint32)
            argb = (argb & 0x00ffffff) | (a.astype(np.uint32) << 24)
            eargb = ref.picture_rescale_argb(argb, dw, dh)
            assert (oracle.picture_rescale_argb(argb, dw, dh) == eargb).all(), ("argb", w, h, dw, dh, alpha_kind)
            if w * h > 3200 and alpha_kind != "random":
                continue   # checked above, not stored (fixture size)
            k = f"c{n}"
            n += 1
            names.append(k)
            for nm, arr in zip("yuva", (y, u, v, a)):
                store[f"{k}/{nm}"] = arr
            for nm, arr in zip("yuva", exp):
                store[f"{k}/out_{nm}"] = arr
            store[f"{k}/argb"], store[f"{k}/out_argb"] = argb, eargb
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "rescale_alpha_libwebp160.npz"), **store)
    print(f"{n} pictures (YUVA + ARGB each): oracle == libwebp binary for WebPPictureRescale with alpha")


if __name__ == "__main__":
    main()
    gen_import_argb()
    gen_dithered()
    gen_rescale()
    gen_rescale_alpha()
