"""Generate tests/golden/*.npz from the upstream libwebp 1.6.0 binary.

Run once in the build container (needs Pillow's bundled libwebp; the GPU box
only reads the committed .npz files):

    python tests/golden/make_golden.py

What is pinned:
  import_*.npz    WebPPictureImport{RGBA,BGRA,RGBX,BGRX,RGB,BGR}, use_argb=0
                  -> Y,U,V,(A), has_alpha.  Translucent 2x2 blocks follow the
                  binary's USE_INVERSE_ALPHA_TABLE divide (oracle DIV_TABLE).
  upsample_*.npz  WebPPictureYUVAToARGB -> u32 0xAARRGGBB.
  decode_*.npz    real lossy .webp files (Pillow, q=80): WebPDecode to
                  MODE_YUVA gives the planes, WebPDecode to every RGB mode gives
                  the bytes the fancy upsampler + per-pixel writers + alpha
                  dispatch must reproduce (pins pkg/libwebp/dsp/
                  upsampling.c.go:107-114 writers that the Go tree leaves
                  undefined, and io_dec.c.go:65-115,175-195).
"""
from __future__ import annotations

import ctypes as C
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import libwebp_ref as ref  # noqa: E402

SIZES = [(1, 1), (1, 2), (2, 1), (2, 2), (3, 3), (4, 4), (5, 7), (6, 8),
         (7, 10), (16, 33), (33, 16), (64, 64), (65, 31), (127, 129)]
ALPHA_KINDS = ["opaque", "rand", "sparse", "binary", "transp"]


def make_rgba(rng, w, h, kind):
    rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    if kind == "opaque":
        rgba[..., 3] = 255
    elif kind == "sparse":
        rgba[..., 3][rng.random((h, w)) < 0.7] = 255
    elif kind == "binary":
        rgba[..., 3] = np.where(rng.random((h, w)) < 0.5, 0, 255)
    elif kind == "transp":
        rgba[..., 3] = 0
    return rgba


def gen_import():
    rng = np.random.default_rng(0xB200)
    out = {}
    n = 0
    for (w, h) in SIZES:
        for kind in ALPHA_KINDS:
            rgba = make_rgba(rng, w, h, kind)
            y, u, v, a, ha = ref.import_yuva(rgba, w, h, "RGBA")
            k = f"c{n}"
            out[k + "_src"] = rgba
            out[k + "_y"], out[k + "_u"], out[k + "_v"] = y, u, v
            out[k + "_a"] = a if a is not None else np.zeros((0, 0), np.uint8)
            out[k + "_meta"] = np.array([w, h, int(ha)], np.int32)
            n += 1
    out["n"] = np.array([n], np.int32)
    # channel orders / steps / padded stride
    for order in ["BGRA", "RGBX", "BGRX", "RGB", "BGR"]:
        w, h = 37, 23
        step = 3 if len(order) == 3 else 4
        pad = 12
        src = rng.integers(0, 256, (h, w * step + pad), dtype=np.uint8)
        y, u, v, a, ha = ref.import_yuva(src, w, h, order, stride=w * step + pad)
        out[f"o_{order}_src"] = src
        out[f"o_{order}_y"], out[f"o_{order}_u"], out[f"o_{order}_v"] = y, u, v
        out[f"o_{order}_a"] = a if a is not None else np.zeros((0, 0), np.uint8)
        out[f"o_{order}_meta"] = np.array([w, h, int(ha), step, w * step + pad],
                                          np.int32)
    np.savez_compressed(os.path.join(HERE, "import_libwebp160.npz"), **out)
    return n


def gen_upsample():
    rng = np.random.default_rng(0xB201)
    out = {}
    n = 0
    for (w, h) in SIZES + [(258, 6), (6, 258)]:
        uw, uh = (w + 1) // 2, (h + 1) // 2
        for with_a in (False, True):
            y = rng.integers(0, 256, (h, w), dtype=np.uint8)
            u = rng.integers(0, 256, (uh, uw), dtype=np.uint8)
            v = rng.integers(0, 256, (uh, uw), dtype=np.uint8)
            a = rng.integers(0, 256, (h, w), dtype=np.uint8) if with_a else None
            argb = ref.yuva_to_argb(y, u, v, a)
            k = f"c{n}"
            out[k + "_y"], out[k + "_u"], out[k + "_v"] = y, u, v
            out[k + "_a"] = a if a is not None else np.zeros((0, 0), np.uint8)
            out[k + "_argb"] = argb
            n += 1
    out["n"] = np.array([n], np.int32)
    np.savez_compressed(os.path.join(HERE, "upsample_libwebp160.npz"), **out)
    return n


# ---- advanced decode API (decode.h of libwebp 1.6.0) ------------------------
class BitstreamFeatures(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("has_alpha", C.c_int),
                ("has_animation", C.c_int), ("format", C.c_int),
                ("pad", C.c_uint32 * 5)]


class RGBABuffer(C.Structure):
    _fields_ = [("rgba", C.c_void_p), ("stride", C.c_int), ("size", C.c_size_t)]


class YUVABuffer(C.Structure):
    _fields_ = [("y", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p),
                ("a", C.c_void_p), ("y_stride", C.c_int), ("u_stride", C.c_int),
                ("v_stride", C.c_int), ("a_stride", C.c_int),
                ("y_size", C.c_size_t), ("u_size", C.c_size_t),
                ("v_size", C.c_size_t), ("a_size", C.c_size_t)]


class BufUnion(C.Union):
    _fields_ = [("RGBA", RGBABuffer), ("YUVA", YUVABuffer)]


class DecBuffer(C.Structure):
    _fields_ = [("colorspace", C.c_int), ("width", C.c_int), ("height", C.c_int),
                ("is_external_memory", C.c_int), ("u", BufUnion),
                ("pad", C.c_uint32 * 4), ("private_memory", C.c_void_p)]


class DecoderOptions(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "bypass_filtering", "no_fancy_upsampling", "use_cropping", "crop_left",
        "crop_top", "crop_width", "crop_height", "use_scaling", "scaled_width",
        "scaled_height", "use_threads", "dithering_strength", "flip",
        "alpha_dithering_strength")] + [("pad", C.c_uint32 * 5)]


class DecoderConfig(C.Structure):
    _fields_ = [("input", BitstreamFeatures), ("output", DecBuffer),
                ("options", DecoderOptions)]


MODE_YUVA = 12
BPP = {0: 3, 1: 4, 2: 3, 3: 4, 4: 4, 5: 2, 6: 2}


def decode(data: bytes, mode: int, flip=0):
    L = ref.lib()
    cfg = DecoderConfig()
    assert L.WebPInitDecoderConfigInternal(C.byref(cfg), 0x0210)
    cfg.output.colorspace = mode
    cfg.options.flip = flip
    st = L.WebPDecode(data, C.c_size_t(len(data)), C.byref(cfg))
    assert st == 0, st
    w, h = cfg.output.width, cfg.output.height
    if mode == MODE_YUVA:
        b = cfg.output.u.YUVA
        uw, uh = (w + 1) // 2, (h + 1) // 2

        def pl(ptr, stride, ww, hh):
            if not ptr:
                return None
            buf = (C.c_uint8 * (stride * hh)).from_address(ptr)
            return np.frombuffer(buf, np.uint8).reshape(hh, stride)[:, :ww].copy()
        res = (pl(b.y, b.y_stride, w, h), pl(b.u, b.u_stride, uw, uh),
               pl(b.v, b.v_stride, uw, uh), pl(b.a, b.a_stride, w, h))
    else:
        b = cfg.output.u.RGBA
        bpp = BPP[mode]
        buf = (C.c_uint8 * (b.stride * h)).from_address(b.rgba)
        res = np.frombuffer(buf, np.uint8).reshape(h, b.stride)[:, :w * bpp].copy()
    L.WebPFreeDecBuffer(C.byref(cfg.output))
    return res


def gen_decode():
    from PIL import Image
    rng = np.random.default_rng(0xB202)
    out = {}
    n = 0
    for (w, h, alpha) in [(17, 23, False), (32, 48, False), (33, 47, True),
                          (64, 64, True), (1, 1, False), (2, 3, True),
                          (120, 36, False)]:
        # smooth-ish content so the lossy codec produces varied chroma
        base = rng.integers(0, 256, (h // 4 + 2, w // 4 + 2, 4), dtype=np.uint8)
        img = np.kron(base, np.ones((4, 4, 1), np.uint8))[:h, :w]
        img = (img.astype(np.int32) + rng.integers(-20, 20, img.shape)).clip(0, 255).astype(np.uint8)
        if not alpha:
            pil = Image.fromarray(img[..., :3], "RGB")
        else:
            pil = Image.fromarray(img, "RGBA")
        bio = io.BytesIO()
        pil.save(bio, "WEBP", quality=80, method=4)
        data = bio.getvalue()
        y, u, v, a = decode(data, MODE_YUVA)
        k = f"c{n}"
        out[k + "_y"], out[k + "_u"], out[k + "_v"] = y, u, v
        out[k + "_a"] = a if a is not None else np.zeros((0, 0), np.uint8)
        for mode in range(7):
            out[k + f"_m{mode}"] = decode(data, mode)
        n += 1
    out["n"] = np.array([n], np.int32)
    np.savez_compressed(os.path.join(HERE, "decode_libwebp160.npz"), **out)
    return n


if __name__ == "__main__":
    print("import cases  :", gen_import())
    print("upsample cases:", gen_upsample())
    print("decode cases  :", gen_decode())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
