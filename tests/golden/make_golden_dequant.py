"""Generate tests/golden/alpha_dequant_libwebp160.npz: alpha level smoothing
(WebPDequantizeLevels, pkg/libwebp/utils/quant_levels_dec_utils.c.go:259-286, called from
pkg/libwebp/decoder/alpha_dec.go:198-203) pinned on the libwebp 1.6.0 binary.

    python tests/golden/make_golden_dequant.py

The binary encodes RGBA pictures with alpha_quality < 100, which quantises the alpha levels
and marks the ALPH chunk "pre-processed"; WebPDecode with alpha_dithering_strength = 0 gives
the quantised plane (the smoother's input), with strength S the smoothed plane.  The oracle
must turn the former into the latter for every S.
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
import make_golden_alpha as ga  # noqa: E402


class Features(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("has_alpha", C.c_int), ("has_animation", C.c_int),
                ("format", C.c_int), ("pad", C.c_uint32 * 5)]


class RGBABuffer(C.Structure):
    _fields_ = [("rgba", C.c_void_p), ("stride", C.c_int), ("size", C.c_size_t)]


class YUVABuffer(C.Structure):
    _fields_ = [("y", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("a", C.c_void_p),
                ("y_stride", C.c_int), ("u_stride", C.c_int), ("v_stride", C.c_int), ("a_stride", C.c_int),
                ("y_size", C.c_size_t), ("u_size", C.c_size_t), ("v_size", C.c_size_t), ("a_size", C.c_size_t)]


class BufUnion(C.Union):
    _fields_ = [("RGBA", RGBABuffer), ("YUVA", YUVABuffer)]


class DecBuffer(C.Structure):
    _fields_ = [("colorspace", C.c_int), ("width", C.c_int), ("height", C.c_int), ("is_external_memory", C.c_int),
                ("u", BufUnion), ("pad", C.c_uint32 * 4), ("private_memory", C.c_void_p)]


class DecoderOptions(C.Structure):
    _fields_ = [("bypass_filtering", C.c_int), ("no_fancy_upsampling", C.c_int), ("use_cropping", C.c_int),
                ("crop_left", C.c_int), ("crop_top", C.c_int), ("crop_width", C.c_int), ("crop_height", C.c_int),
                ("use_scaling", C.c_int), ("scaled_width", C.c_int), ("scaled_height", C.c_int), ("use_threads", C.c_int),
                ("dithering_strength", C.c_int), ("flip", C.c_int), ("alpha_dithering_strength", C.c_int),
                ("pad", C.c_uint32 * 5)]


class DecoderConfig(C.Structure):
    _fields_ = [("input", Features), ("output", DecBuffer), ("options", DecoderOptions), ("tail", C.c_uint32 * 8)]


def encode(rgba, alpha_quality, alpha_filtering=1):
    L = ref.lib()
    h, w, _ = rgba.shape
    cfg = ga.WebPConfig()
    assert L.WebPConfigInitInternal(C.byref(cfg), 0, C.c_float(90.0), 0x0210)
    cfg.alpha_filtering, cfg.alpha_quality, cfg.alpha_compression, cfg.method = alpha_filtering, alpha_quality, 1, 4
    assert L.WebPValidateConfig(C.byref(cfg))
    p = ref.WebPPicture()
    assert L.WebPPictureInitInternal(C.byref(p), 0x0210)
    p.width, p.height, p.use_argb = w, h, 0
    rgba = np.ascontiguousarray(rgba)
    assert L.WebPPictureImportRGBA(C.byref(p), rgba.ctypes.data_as(C.c_void_p), 4 * w)
    wr = ga.WebPMemoryWriter()
    L.WebPMemoryWriterInit(C.byref(wr))
    p.writer = C.cast(L.WebPMemoryWrite, C.c_void_p).value
    p.custom_ptr = C.addressof(wr)
    assert L.WebPEncode(C.byref(cfg), C.byref(p)), p.error_code
    data = C.string_at(wr.mem, wr.size)
    L.WebPMemoryWriterClear(C.byref(wr))
    L.WebPPictureFree(C.byref(p))
    return data


def decode_alpha(webp, strength):
    """Alpha plane of WebPDecode(MODE_RGBA) with options.alpha_dithering_strength = strength."""
    L = ref.lib()
    cfg = DecoderConfig()
    assert L.WebPInitDecoderConfigInternal(C.byref(cfg), 0x0209)
    cfg.output.colorspace = 1   # MODE_RGBA (not premultiplied)
    cfg.options.alpha_dithering_strength = strength
    L.WebPDecode.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
    st = L.WebPDecode(webp, len(webp), C.byref(cfg))
    assert st == 0, st
    w, h, stride = cfg.output.width, cfg.output.height, cfg.output.u.RGBA.stride
    raw = np.frombuffer(C.string_at(cfg.output.u.RGBA.rgba, stride * h), np.uint8).reshape(h, stride)
    a = raw[:, :4 * w].reshape(h, w, 4)[..., 3].copy()
    L.WebPFreeDecBuffer.argtypes = [C.c_void_p]
    L.WebPFreeDecBuffer(C.byref(cfg.output))
    return a


def main():
    rng = np.random.default_rng(0xB2007)
    store, names = {}, []
    n = changed = 0
    for (w, h) in [(5, 7), (16, 16), (33, 20), (64, 48), (97, 31), (130, 75)]:
        yy, xx = np.mgrid[0:h, 0:w]
        pats = {
            "hramp": (xx * 255) // max(1, w - 1) + 0 * yy,
            "diag": ((xx + yy) * 255) // max(1, w + h - 2),
            "smooth": (128 + 120 * np.sin(xx / 11.0) * np.cos(yy / 9.0)).astype(np.int64),
            "blob": np.clip(255 - ((xx - w / 2) ** 2 + (yy - h / 2) ** 2) * 600 // (w * w + h * h), 0, 255),
            "noise": rng.integers(0, 256, (h, w)),
            "two": np.where(xx > w // 2, 255, 0),
        }
        for name, a in pats.items():
            rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
            rgba[..., 3] = np.asarray(a, np.int64).astype(np.uint8)
            for aq in (10, 40, 80):
                webp = encode(rgba, aq)
                alph = ga.riff_chunk(webp, b"ALPH")
                if alph is None:
                    continue
                assert (alph[0] >> 4) & 3 == 1, "ALPH chunk not marked pre-processed"
                quantised = decode_alpha(webp, 0)
                for strength in (20, 50, 75, 100):
                    expected = decode_alpha(webp, strength)
                    got = oracle.dequantize_levels(quantised, strength)
                    assert (got == expected).all(), (w, h, name, aq, strength)
                    changed += int((expected != quantised).any())
                    if strength in (50, 100) and (aq != 40 or w <= 64):
                        k = f"c{n}"
                        n += 1
                        names.append(k)
                        store[f"{k}/strength"], store[f"{k}/in"], store[f"{k}/out"] = np.int32(strength), quantised, expected
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "alpha_dequant_libwebp160.npz"), **store)
    print(f"{n} cases stored; {changed} (file, strength) pairs where smoothing changed the plane")


if __name__ == "__main__":
    main()
