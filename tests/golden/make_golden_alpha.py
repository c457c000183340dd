"""Generate tests/golden/alpha_unfilter_libwebp160.npz: the alpha-plane unfilters
(HorizontalUnfilter_C / VerticalUnfilter_C / GradientUnfilter_C, pkg/libwebp/dsp/filters.go:
130-171, driven row by row as ALPHDecode / AlphaApplyFilter do, pkg/vp8/vp8l_dec.c.go:
802-814) cut out of lossy+alpha files written by the libwebp 1.6.0 binary.

    python tests/golden/make_golden_alpha.py

The binary encodes RGBA pictures with alpha_filtering = 2 (every filter is tried, the
smallest wins) and alpha_quality = 100 (no level quantisation).  The ALPH chunk's header
byte gives the filter; its payload is a headerless VP8L stream (or raw bytes) whose green
channel is the FILTERED alpha plane (vp8l_bitstream.py + the oracle's inverse transforms);
WebPDecodeRGBA of the same file gives the plane the unfilter must produce.
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
import vp8l_bitstream as vb  # noqa: E402
from make_golden_vp8l import apply_inverse  # noqa: E402


class WebPConfig(C.Structure):
    _fields_ = [("lossless", C.c_int), ("quality", C.c_float), ("method", C.c_int), ("image_hint", C.c_int),
                ("target_size", C.c_int), ("target_PSNR", C.c_float), ("segments", C.c_int), ("sns_strength", C.c_int),
                ("filter_strength", C.c_int), ("filter_sharpness", C.c_int), ("filter_type", C.c_int), ("autofilter", C.c_int),
                ("alpha_compression", C.c_int), ("alpha_filtering", C.c_int), ("alpha_quality", C.c_int), ("pass_", C.c_int),
                ("show_compressed", C.c_int), ("preprocessing", C.c_int), ("partitions", C.c_int), ("partition_limit", C.c_int),
                ("emulate_jpeg_size", C.c_int), ("thread_level", C.c_int), ("low_memory", C.c_int), ("near_lossless", C.c_int),
                ("exact", C.c_int), ("use_delta_palette", C.c_int), ("use_sharp_yuv", C.c_int), ("qmin", C.c_int), ("qmax", C.c_int),
                ("tail", C.c_uint32 * 8)]


class WebPMemoryWriter(C.Structure):
    _fields_ = [("mem", C.c_void_p), ("size", C.c_size_t), ("max_size", C.c_size_t), ("pad", C.c_uint32 * 1)]


def encode_lossy_alpha(rgba, alpha_filtering=2, method=4):
    L = ref.lib()
    h, w, _ = rgba.shape
    cfg = WebPConfig()
    assert L.WebPConfigInitInternal(C.byref(cfg), 0, C.c_float(90.0), 0x0210)
    cfg.alpha_filtering, cfg.alpha_quality, cfg.alpha_compression, cfg.method = alpha_filtering, 100, 1, method
    assert L.WebPValidateConfig(C.byref(cfg))
    p = ref.WebPPicture()
    assert L.WebPPictureInitInternal(C.byref(p), 0x0210)
    p.width, p.height, p.use_argb = w, h, 0
    rgba = np.ascontiguousarray(rgba)
    assert L.WebPPictureImportRGBA(C.byref(p), rgba.ctypes.data_as(C.c_void_p), 4 * w)
    wr = WebPMemoryWriter()
    L.WebPMemoryWriterInit(C.byref(wr))
    p.writer = C.cast(L.WebPMemoryWrite, C.c_void_p).value
    p.custom_ptr = C.addressof(wr)
    ok = L.WebPEncode(C.byref(cfg), C.byref(p))
    assert ok, p.error_code
    data = C.string_at(wr.mem, wr.size)
    L.WebPMemoryWriterClear(C.byref(wr))
    L.WebPPictureFree(C.byref(p))
    return data


def decode_rgba(webp):
    L = ref.lib()
    w, h = C.c_int(), C.c_int()
    L.WebPDecodeRGBA.restype = C.c_void_p
    L.WebPDecodeRGBA.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    p = L.WebPDecodeRGBA(webp, len(webp), C.byref(w), C.byref(h))
    assert p
    a = np.frombuffer(C.string_at(p, 4 * w.value * h.value), np.uint8).copy().reshape(h.value, w.value, 4)
    L.WebPFree.argtypes = [C.c_void_p]
    L.WebPFree(p)
    return a


def riff_chunk(webp, tag):
    p = 12
    while p + 8 <= len(webp):
        t, size = webp[p:p + 4], int.from_bytes(webp[p + 4:p + 8], "little")
        if t == tag:
            return webp[p + 8:p + 8 + size]
        p += 8 + size + (size & 1)
    return None


def filtered_plane(alph, w, h):
    """(filter, filtered alpha plane) from an ALPH chunk payload (pkg/libwebp/dec alpha header:
    bits 0-1 compression, 2-3 filter, 4-5 pre-processing)."""
    hdr = alph[0]
    comp, filt, pre = hdr & 3, (hdr >> 2) & 3, (hdr >> 4) & 3
    assert pre == 0
    if comp == 0:
        return filt, np.frombuffer(alph[1:1 + w * h], np.uint8).reshape(h, w).copy()
    br = vb.BitReader(alph[1:])
    transforms, pix, fx = vb.decode_image_stream(br, w, h, True)
    img = apply_inverse({"transforms": transforms, "residual": pix.reshape(h, fx)})
    return filt, ((img >> 8) & 0xff).astype(np.uint8)


def main():
    rng = np.random.default_rng(0xB2006)
    store, names, seen = {}, [], {}
    n = 0
    for (w, h) in [(16, 16), (33, 20), (64, 48), (97, 31), (130, 40)]:
        yy, xx = np.mgrid[0:h, 0:w]
        pats = {
            "hramp": (xx * 5 + rng.integers(0, 2, (h, w))) % 256,
            "vramp": (yy * 7 + rng.integers(0, 2, (h, w))) % 256,
            "vbands": ((yy // 3) * 37) % 256 + 0 * xx,
            "diag": ((xx * 3 + yy * 2) + rng.integers(0, 2, (h, w))) % 256,
            "bilinear": np.clip(xx * 2 + yy * 3 + rng.integers(-1, 2, (h, w)), 0, 255),
            "smooth": (128 + 100 * np.sin(xx / 9.0) * np.cos(yy / 7.0)).astype(np.int64),
            "noise": rng.integers(0, 256, (h, w)),
        }
        for name, a in pats.items():
            rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
            rgba[..., 3] = np.asarray(a, np.int64).astype(np.uint8)
            for method in (2, 4):
                webp = encode_lossy_alpha(rgba, 2, method)
                alph = riff_chunk(webp, b"ALPH")
                if alph is None:
                    continue
                filt, fin = filtered_plane(alph, w, h)
                expected = decode_rgba(webp)[..., 3]
                got = oracle.alpha_unfilter(fin, filt)
                assert (got == expected).all(), (w, h, name, method, filt)
                assert (oracle.alpha_filter(expected, filt) == fin).all(), (w, h, name, "forward", filt)
                seen[filt] = seen.get(filt, 0) + 1
                if seen[filt] > 14:
                    continue
                k = f"c{n}"
                n += 1
                names.append(k)
                store[f"{k}/filter"], store[f"{k}/in"], store[f"{k}/out"] = np.int32(filt), fin, expected
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "alpha_unfilter_libwebp160.npz"), **store)
    print(f"{n} cases stored; files per filter (none, horizontal, vertical, gradient): {[seen.get(f, 0) for f in range(4)]}")


if __name__ == "__main__":
    main()
