"""Generate tests/golden/anim_libwebp160.npz: canvas blending of the animation decoder
(BlendPixelRowNonPremult / BlendPixelRowPremult, pkg/libwebp/demux/anim_decode.go:195-270)
from the libwebp 1.6.0 binaries bundled with Pillow (libwebp, libwebpmux, libwebpdemux).

    python tests/golden/make_golden_anim.py

Two frames are encoded losslessly one by one, assembled into an animation with WebPMux
(second frame: full canvas, blend = WEBP_MUX_BLEND, dispose = NONE) and decoded with
WebPAnimDecoder in MODE_RGBA (non-premultiplied blend) and MODE_rgbA (premultiplied blend).
Stored per case: src = the second frame as the decoder sees it before blending (a one-frame
animation of the same bitstream), dst = the canvas after frame 1, expected = the canvas
after frame 2.  The script asserts the oracle reproduces expected from (src, dst).
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import sys

import numpy as np
import PIL

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from oracle import libwebp_ref as ref  # noqa: E402


class WebPData(C.Structure):
    _fields_ = [("bytes", C.c_void_p), ("size", C.c_size_t)]


class WebPMuxFrameInfo(C.Structure):
    _fields_ = [("bitstream", WebPData), ("x_offset", C.c_int), ("y_offset", C.c_int), ("duration", C.c_int),
                ("id", C.c_int), ("dispose_method", C.c_int), ("blend_method", C.c_int), ("pad", C.c_uint32 * 1)]


class WebPMuxAnimParams(C.Structure):
    _fields_ = [("bgcolor", C.c_uint32), ("loop_count", C.c_int)]


class WebPAnimDecoderOptions(C.Structure):
    _fields_ = [("color_mode", C.c_int), ("use_threads", C.c_int), ("padding", C.c_uint32 * 7)]


def _libs():
    ref.lib()
    root = os.path.join(os.path.dirname(os.path.dirname(PIL.__file__)), "pillow.libs")
    mux = C.CDLL(glob.glob(os.path.join(root, "libwebpmux-*.so*"))[0])
    demux = C.CDLL(glob.glob(os.path.join(root, "libwebpdemux-*.so*"))[0])
    return mux, demux


def encode_lossless(rgba):
    L = ref.lib()
    h, w, _ = rgba.shape
    out = C.c_void_p()
    L.WebPEncodeLosslessRGBA.restype = C.c_size_t
    L.WebPEncodeLosslessRGBA.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    rgba = np.ascontiguousarray(rgba)
    n = L.WebPEncodeLosslessRGBA(rgba.ctypes.data, w, h, 4 * w, C.byref(out))
    assert n
    data = C.string_at(out.value, n)
    L.WebPFree.argtypes = [C.c_void_p]
    L.WebPFree(out)
    return data


def assemble(mux, frames, w, h):
    mux.WebPNewInternal.restype = C.c_void_p
    m = None
    for ver in (0x0109, 0x0108, 0x0107):
        m = mux.WebPNewInternal(ver)
        if m:
            break
    assert m, "WebPMuxNew"
    m = C.c_void_p(m)
    keep = []
    for bits in frames:
        buf = C.create_string_buffer(bits, len(bits))
        keep.append(buf)
        info = WebPMuxFrameInfo()
        info.bitstream = WebPData(C.cast(buf, C.c_void_p), len(bits))
        info.duration, info.id, info.dispose_method, info.blend_method = 100, 3, 0, 0   # ANMF, DISPOSE_NONE, BLEND
        assert mux.WebPMuxPushFrame(m, C.byref(info), 1) == 1
    assert mux.WebPMuxSetCanvasSize(m, w, h) == 1
    params = WebPMuxAnimParams(0, 0)
    assert mux.WebPMuxSetAnimationParams(m, C.byref(params)) == 1
    out = WebPData()
    assert mux.WebPMuxAssemble(m, C.byref(out)) == 1
    data = C.string_at(out.bytes, out.size)
    mux.WebPMuxDelete(m)
    return data


def decode_canvases(demux, webp, mode, w, h):
    opts = WebPAnimDecoderOptions()
    dec = None
    buf = C.create_string_buffer(webp, len(webp))
    wd = WebPData(C.cast(buf, C.c_void_p), len(webp))
    demux.WebPAnimDecoderNewInternal.restype = C.c_void_p
    for ver in (0x0107, 0x0106, 0x0105):
        if not demux.WebPAnimDecoderOptionsInitInternal(C.byref(opts), ver):
            continue
        opts.color_mode, opts.use_threads = mode, 0
        dec = demux.WebPAnimDecoderNewInternal(C.byref(wd), C.byref(opts), ver)
        if dec:
            break
    assert dec, "WebPAnimDecoderNew"
    dec = C.c_void_p(dec)
    out = []
    p, ts = C.c_void_p(), C.c_int()
    while demux.WebPAnimDecoderHasMoreFrames(dec):
        assert demux.WebPAnimDecoderGetNext(dec, C.byref(p), C.byref(ts))
        out.append(np.frombuffer(C.string_at(p.value, 4 * w * h), np.uint8).copy().view(np.uint32).reshape(h, w))
    demux.WebPAnimDecoderDelete(dec)
    return out


def main():
    mux, demux = _libs()
    rng = np.random.default_rng(0xB2005)
    store, names = {}, []
    n = 0
    for (w, h) in [(1, 1), (4, 3), (17, 9), (33, 20), (64, 31)]:
        for kind in ("rand", "sparse", "opaque_over", "transp_over"):
            f1 = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
            f2 = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
            if kind == "sparse":
                f2[..., 3][rng.random((h, w)) < 0.4] = 255
                f2[..., 3][rng.random((h, w)) < 0.3] = 0
                f1[..., 3][rng.random((h, w)) < 0.5] = 255
            elif kind == "opaque_over":
                f1[..., 3] = 255
            elif kind == "transp_over":
                f1[..., 3][rng.random((h, w)) < 0.5] = 0
            b1, b2 = encode_lossless(f1), encode_lossless(f2)
            for mode, premult in ((1, 0), (7, 1)):    # MODE_RGBA, MODE_rgbA
                c = decode_canvases(demux, assemble(mux, [b1, b2], w, h), mode, w, h)
                s = decode_canvases(demux, assemble(mux, [b2], w, h), mode, w, h)
                assert len(c) == 2 and len(s) == 1
                src, dst, exp = s[0], c[0], c[1]
                got = oracle.anim_blend_rows(src, dst, premult)
                assert (got == exp).all(), (w, h, kind, mode, np.argwhere(got != exp)[:3])
                k = f"c{n}"
                n += 1
                names.append(k)
                store[f"{k}/src"], store[f"{k}/dst"], store[f"{k}/exp"], store[f"{k}/premult"] = src, dst, exp, np.int32(premult)
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "anim_libwebp160.npz"), **store)
    print(f"{n} cases: oracle == WebPAnimDecoder for the non-premultiplied and premultiplied canvas blend")


if __name__ == "__main__":
    main()
