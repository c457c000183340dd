"""Second oracle: the upstream libwebp 1.6.0 binary bundled with Pillow.

TEST INFRASTRUCTURE ONLY.  go-webp is a transliteration of libwebp 1.6.0
(pkg/vp8/constants.go:18-24); the Go tree itself does not compile and no Go
toolchain is in the image, so the closest runnable "reference itself" is this
shared object.  It pins parts 1 and 2 of the hot path:

  import   : WebPPictureImportRGBA/BGRA/RGBX/BGRX/RGB/BGR with use_argb=0
             (pkg/libwebp/enc/picture_csp_enc.c.go:189-225)
  upsample : WebPPictureYUVAToARGB (picture_csp_enc.c.go:62-118)

Caveat: the binary is built with USE_INVERSE_ALPHA_TABLE, so on
translucent 2x2 blocks it matches the oracle's DIV_TABLE mode, not the exact
division the Go reference states (pkg/libwebp/dsp/yuv.go:231).

tests/golden/make_golden.py uses this module to write the committed fixtures;
tests then check the C oracle against the fixtures without needing Pillow.
"""
from __future__ import annotations

import ctypes as C
import glob
import os

import numpy as np

_L = None


class WebPPicture(C.Structure):
    # field order documented by pkg/picture/picture.go:33-105 (sizeof == 256)
    _fields_ = [
        ("use_argb", C.c_int), ("colorspace", C.c_int),
        ("width", C.c_int), ("height", C.c_int),
        ("y", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p),
        ("y_stride", C.c_int), ("uv_stride", C.c_int),
        ("a", C.c_void_p), ("a_stride", C.c_int), ("pad1", C.c_uint32 * 2),
        ("argb", C.c_void_p), ("argb_stride", C.c_int),
        ("pad2", C.c_uint32 * 3),
        ("writer", C.c_void_p), ("custom_ptr", C.c_void_p),
        ("extra_info_type", C.c_int), ("extra_info", C.c_void_p),
        ("stats", C.c_void_p), ("error_code", C.c_int),
        ("progress_hook", C.c_void_p), ("user_data", C.c_void_p),
        ("pad3", C.c_uint32 * 3), ("pad4", C.c_void_p), ("pad5", C.c_void_p),
        ("pad6", C.c_uint32 * 8),
        ("memory_", C.c_void_p), ("memory_argb_", C.c_void_p),
        ("pad7", C.c_void_p * 2),
    ]


def available() -> bool:
    try:
        lib()
        return True
    except OSError:
        return False


def lib():
    global _L
    if _L is None:
        import PIL  # noqa: F401  (locates pillow.libs)
        root = os.path.join(os.path.dirname(os.path.dirname(PIL.__file__)),
                            "pillow.libs")
        sharp = glob.glob(os.path.join(root, "libsharpyuv-*.so*"))
        webp = [p for p in glob.glob(os.path.join(root, "libwebp-*.so*"))]
        if not sharp or not webp:
            raise OSError("bundled libwebp not found under " + root)
        C.CDLL(sharp[0], mode=C.RTLD_GLOBAL)
        L = C.CDLL(webp[0])
        assert C.sizeof(WebPPicture) == 256
        L.WebPGetEncoderVersion.restype = C.c_int
        if L.WebPGetEncoderVersion() != 0x010600:
            raise OSError("bundled libwebp is not 1.6.0")
        _L = L
    return _L


def _plane(ptr, stride, w, h):
    buf = (C.c_uint8 * (stride * h)).from_address(ptr)
    return np.frombuffer(buf, np.uint8).reshape(h, stride)[:, :w].copy()


_IMPORTERS = {"RGBA": ("WebPPictureImportRGBA", 4), "BGRA": ("WebPPictureImportBGRA", 4),
              "RGBX": ("WebPPictureImportRGBX", 4), "BGRX": ("WebPPictureImportBGRX", 4),
              "RGB": ("WebPPictureImportRGB", 3), "BGR": ("WebPPictureImportBGR", 3)}


def import_yuva(src: np.ndarray, width: int, height: int, order="RGBA",
                stride=None):
    """Returns (y, u, v, a_or_None, has_alpha) from the libwebp binary."""
    L = lib()
    fn, step = _IMPORTERS[order]
    src = np.ascontiguousarray(src, np.uint8).reshape(-1)
    if stride is None:
        stride = step * width
    p = WebPPicture()
    assert L.WebPPictureInitInternal(C.byref(p), 0x0210)
    p.width, p.height, p.use_argb = width, height, 0
    ok = getattr(L, fn)(C.byref(p), src.ctypes.data_as(C.c_void_p), stride)
    if not ok:
        raise ValueError("libwebp import failed")
    uvw, uvh = (width + 1) >> 1, (height + 1) >> 1
    y = _plane(p.y, p.y_stride, width, height)
    u = _plane(p.u, p.uv_stride, uvw, uvh)
    v = _plane(p.v, p.uv_stride, uvw, uvh)
    has_alpha = bool(p.colorspace & 4)
    a = _plane(p.a, p.a_stride, width, height) if has_alpha else None
    L.WebPPictureFree(C.byref(p))
    return y, u, v, a, has_alpha


def yuva_to_argb(y, u, v, a=None):
    """WebPPictureYUVAToARGB -> uint32 (h, w) array of 0xAARRGGBB."""
    L = lib()
    h, w = y.shape
    p = WebPPicture()
    assert L.WebPPictureInitInternal(C.byref(p), 0x0210)
    p.width, p.height, p.use_argb = w, h, 0
    p.colorspace = 4 if a is not None else 0
    assert L.WebPPictureAlloc(C.byref(p))

    def put(ptr, stride, arr):
        hh, ww = arr.shape
        buf = (C.c_uint8 * (stride * hh)).from_address(ptr)
        np.frombuffer(buf, np.uint8).reshape(hh, stride)[:, :ww] = arr

    put(p.y, p.y_stride, y)
    put(p.u, p.uv_stride, u)
    put(p.v, p.uv_stride, v)
    if a is not None:
        put(p.a, p.a_stride, a)
    assert L.WebPPictureYUVAToARGB(C.byref(p))
    buf = (C.c_uint32 * (p.argb_stride * h)).from_address(p.argb)
    out = np.frombuffer(buf, np.uint32).reshape(h, p.argb_stride)[:, :w].copy()
    L.WebPPictureFree(C.byref(p))
    return out


def _picture_yuva(y, u, v, a):
    L = lib()
    h, w = y.shape
    p = WebPPicture()
    assert L.WebPPictureInitInternal(C.byref(p), 0x0210)
    p.width, p.height, p.use_argb = w, h, 0
    p.colorspace = 4
    assert L.WebPPictureAlloc(C.byref(p))
    for ptr, stride, arr in ((p.y, p.y_stride, y), (p.u, p.uv_stride, u), (p.v, p.uv_stride, v), (p.a, p.a_stride, a)):
        hh, ww = arr.shape
        buf = (C.c_uint8 * (stride * hh)).from_address(ptr)
        np.frombuffer(buf, np.uint8).reshape(hh, stride)[:, :ww] = arr
    return p


def _read_yuva(p):
    w, h = p.width, p.height
    uvw, uvh = (w + 1) >> 1, (h + 1) >> 1
    return (_plane(p.y, p.y_stride, w, h), _plane(p.u, p.uv_stride, uvw, uvh),
            _plane(p.v, p.uv_stride, uvw, uvh), _plane(p.a, p.a_stride, w, h))


def _picture_argb(argb):
    L = lib()
    h, w = argb.shape
    p = WebPPicture()
    assert L.WebPPictureInitInternal(C.byref(p), 0x0210)
    p.width, p.height, p.use_argb = w, h, 1
    assert L.WebPPictureAlloc(C.byref(p))
    buf = (C.c_uint32 * (p.argb_stride * h)).from_address(p.argb)
    np.frombuffer(buf, np.uint32).reshape(h, p.argb_stride)[:, :w] = argb
    return p


def _read_argb(p):
    buf = (C.c_uint32 * (p.argb_stride * p.height)).from_address(p.argb)
    return np.frombuffer(buf, np.uint32).reshape(p.height, p.argb_stride)[:, :p.width].copy()


def cleanup_transparent_yuva(y, u, v, a):
    """WebPCleanupTransparentArea (exported by the binary) on a YUVA picture -> (y, u, v)."""
    L = lib()
    p = _picture_yuva(y, u, v, a)
    L.WebPCleanupTransparentArea(C.byref(p))
    out = _read_yuva(p)[:3]
    L.WebPPictureFree(C.byref(p))
    return out


def cleanup_transparent_argb(argb):
    L = lib()
    p = _picture_argb(argb)
    L.WebPCleanupTransparentArea(C.byref(p))
    out = _read_argb(p)
    L.WebPPictureFree(C.byref(p))
    return out


def blend_alpha_yuva(y, u, v, a, background_rgb):
    L = lib()
    p = _picture_yuva(y, u, v, a)
    L.WebPBlendAlpha(C.byref(p), C.c_uint32(background_rgb))
    out = _read_yuva(p)
    L.WebPPictureFree(C.byref(p))
    return out


def blend_alpha_argb(argb, background_rgb):
    L = lib()
    p = _picture_argb(argb)
    L.WebPBlendAlpha(C.byref(p), C.c_uint32(background_rgb))
    out = _read_argb(p)
    L.WebPPictureFree(C.byref(p))
    return out


def import_to_argb(src: np.ndarray, width: int, height: int, order="RGBA", stride=None):
    """WebPPictureImport* with use_argb = 1 -> uint32 (h, w) 0xAARRGGBB from the binary."""
    L = lib()
    fn, step = _IMPORTERS[order]
    src = np.ascontiguousarray(src, np.uint8).reshape(-1)
    if stride is None:
        stride = step * width
    p = WebPPicture()
    assert L.WebPPictureInitInternal(C.byref(p), 0x0210)
    p.width, p.height, p.use_argb = width, height, 1
    if not getattr(L, fn)(C.byref(p), src.ctypes.data_as(C.c_void_p), stride):
        raise ValueError("libwebp import failed")
    out = _read_argb(p)
    L.WebPPictureFree(C.byref(p))
    return out


def argb_to_yuva_dithered(argb: np.ndarray, dithering: float):
    """WebPPictureARGBToYUVADithered (exported) on an ARGB picture -> (y, u, v, a_or_None,
    has_alpha).  Translucent blocks follow the binary's inverse-alpha table (DIV_TABLE)."""
    L = lib()
    p = _picture_argb(np.ascontiguousarray(argb, np.uint32))
    L.WebPPictureARGBToYUVADithered.argtypes = [C.c_void_p, C.c_int, C.c_float]
    if not L.WebPPictureARGBToYUVADithered(C.byref(p), 0, C.c_float(dithering)):   # WEBP_YUV420
        raise ValueError("WebPPictureARGBToYUVADithered failed")
    w, h = p.width, p.height
    uvw, uvh = (w + 1) >> 1, (h + 1) >> 1
    has_alpha = bool(p.colorspace & 4)
    out = (_plane(p.y, p.y_stride, w, h), _plane(p.u, p.uv_stride, uvw, uvh), _plane(p.v, p.uv_stride, uvw, uvh),
           _plane(p.a, p.a_stride, w, h) if has_alpha else None, has_alpha)
    L.WebPPictureFree(C.byref(p))
    return out


def picture_rescale_yuv(y, u, v, width, height):
    """WebPPictureRescale (exported) on a YUV420 picture without alpha -> (y, u, v)."""
    L = lib()
    h, w = y.shape
    p = WebPPicture()
    assert L.WebPPictureInitInternal(C.byref(p), 0x0210)
    p.width, p.height, p.use_argb, p.colorspace = w, h, 0, 0
    assert L.WebPPictureAlloc(C.byref(p))
    for ptr, stride, arr in ((p.y, p.y_stride, y), (p.u, p.uv_stride, u), (p.v, p.uv_stride, v)):
        hh, ww = arr.shape
        buf = (C.c_uint8 * (stride * hh)).from_address(ptr)
        np.frombuffer(buf, np.uint8).reshape(hh, stride)[:, :ww] = arr
    if not L.WebPPictureRescale(C.byref(p), width, height):
        raise ValueError("WebPPictureRescale failed")
    uvw, uvh = (p.width + 1) >> 1, (p.height + 1) >> 1
    out = (_plane(p.y, p.y_stride, p.width, p.height), _plane(p.u, p.uv_stride, uvw, uvh), _plane(p.v, p.uv_stride, uvw, uvh))
    L.WebPPictureFree(C.byref(p))
    return out


def picture_rescale_yuva(y, u, v, a, width, height):
    """WebPPictureRescale on a YUV420A picture (alpha rescaled first, luma premultiplied
    around its rescale, picture_rescale_enc.c.go:174-194) -> (y, u, v, a)."""
    L = lib()
    p = _picture_yuva(y, u, v, a)
    if not L.WebPPictureRescale(C.byref(p), width, height):
        raise ValueError("WebPPictureRescale failed")
    out = _read_yuva(p)
    L.WebPPictureFree(C.byref(p))
    return out


def picture_rescale_argb(argb, width, height):
    """WebPPictureRescale on an ARGB picture (premultiply, 4-channel rescale, un-premultiply,
    picture_rescale_enc.c.go:196-212) -> argb."""
    L = lib()
    p = _picture_argb(argb)
    if not L.WebPPictureRescale(C.byref(p), width, height):
        raise ValueError("WebPPictureRescale failed")
    out = _read_argb(p)
    L.WebPPictureFree(C.byref(p))
    return out
