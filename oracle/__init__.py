"""ctypes binding of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(go_webp_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: dict[str, C.CDLL] = {}

MODE_RGB, MODE_RGBA, MODE_BGR, MODE_BGRA, MODE_ARGB = 0, 1, 2, 3, 4
MODE_RGBA_4444, MODE_RGB_565 = 5, 6
MODE_rgbA, MODE_bgrA, MODE_Argb, MODE_rgbA_4444 = 7, 8, 9, 10
DIV_EXACT, DIV_TABLE = 0, 1
PREDICTOR, CROSS_COLOR, SUBTRACT_GREEN, COLOR_INDEXING = 0, 1, 2, 3

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)


def build(force: bool = False) -> None:
    """Compile liboracle.so / liboracle_scalar.so with the committed Makefile."""
    newest = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("oracle.c", "oracle_post.c", "oracle_tools.c"))
    need = force or not all(
        os.path.exists(os.path.join(_HERE, n))
        and os.path.getmtime(os.path.join(_HERE, n)) >= newest
        for n in ("liboracle.so", "liboracle_scalar.so"))
    if need:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "all"])


def lib(scalar: bool = False) -> C.CDLL:
    name = "liboracle_scalar.so" if scalar else "liboracle.so"
    if name not in _LIBS:
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_init()
        L.orc_LinearToGamma.argtypes = [C.c_uint32, C.c_int]
        L.orc_synth_word.restype = C.c_uint32
        L.orc_synth_word.argtypes = [C.c_uint32] * 5
        L.orc_predict.restype = C.c_uint32
        L.orc_predict.argtypes = [C.c_int] + [C.c_uint32] * 4
        L.orc_ColorTransformDelta.argtypes = [C.c_int8, C.c_int8]
        L.orc_vp8l_batch_mt.argtypes = [
            C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
            C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int]
        _LIBS[name] = L
    return _LIBS[name]


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def gamma_tables():
    g2l = np.zeros(256, np.uint16)
    l2g = np.zeros(33, np.int32)
    lib().orc_gamma_tables(_p(g2l), _p(l2g))
    return g2l, l2g


# byte offsets (r, g, b, a) inside a pixel for the import entry points of
# pkg/libwebp/enc/picture_csp_enc.c.go:189-225 and pkg/picture/conversions.go:22-26
ORDERS = {
    "RGBA": (0, 1, 2, 3, 4), "BGRA": (2, 1, 0, 3, 4),
    "RGBX": (0, 1, 2, -1, 4), "BGRX": (2, 1, 0, -1, 4),
    "RGB": (0, 1, 2, -1, 3), "BGR": (2, 1, 0, -1, 3),
    "ARGB32": (2, 1, 0, 3, 4),  # u32 0xAARRGGBB little-endian == B,G,R,A bytes
}


def import_yuva(src: np.ndarray, width: int, height: int, order: str = "RGBA",
                stride: int | None = None, divide_mode: int = DIV_EXACT,
                want_a: bool = True, scalar: bool = False):
    """src: uint8 buffer of interleaved pixels.  Returns (y,u,v,a,has_alpha)."""
    r, g, b, a_off, step = ORDERS[order]
    src = np.ascontiguousarray(src, np.uint8).reshape(-1)
    if stride is None:
        stride = step * width
    uvw, uvh = (width + 1) >> 1, (height + 1) >> 1
    y = np.zeros((height, width), np.uint8)
    u = np.zeros((uvh, uvw), np.uint8)
    v = np.zeros((uvh, uvw), np.uint8)
    a = np.zeros((height, width), np.uint8) if want_a else None
    base = src
    if stride < 0:  # pointer addresses the last row (buffer_dec.c.go:161-184)
        off = (-stride) * (height - 1)
        ptr = C.c_void_p(src.ctypes.data + off)
    else:
        ptr = _p(base)
    ha = lib(scalar).orc_import_yuva(
        ptr, r, g, b, a_off, step, stride, width, height, _p(y), width, _p(u),
        _p(v), uvw, _p(a), width, divide_mode)
    if ha < 0:
        raise ValueError("orc_import_yuva: bad arguments")
    return y, u, v, a, bool(ha)


def _swap16(out, mode, swap16):
    """16-bit modes in the go-webp text's byte order (WEBP_SWAP_16BIT_CSP: ba / gb first) instead of the
    stock libwebp build's (rg first), which is what the C functions here and the golden binary produce."""
    if swap16 and mode in (5, 6, 10):
        h = out.shape[0]
        return np.ascontiguousarray(out.reshape(h, -1, 2)[:, :, ::-1]).reshape(h, -1)
    return out


def yuva_to_rgba(y, u, v, a=None, mode: int = MODE_RGBA, swap16: bool = False):
    h, w = y.shape
    y = np.ascontiguousarray(y, np.uint8)
    u = np.ascontiguousarray(u, np.uint8)
    v = np.ascontiguousarray(v, np.uint8)
    if a is not None:
        a = np.ascontiguousarray(a, np.uint8)
    bpp = lib().orc_mode_bpp(mode)
    dst = np.zeros((h, w * bpp), np.uint8)
    ok = lib().orc_yuva_to_rgba(mode, _p(y), w, _p(u), _p(v), u.shape[1],
                                _p(a), w, w, h, _p(dst), w * bpp)
    if not ok:
        raise ValueError("orc_yuva_to_rgba: bad arguments")
    return _swap16(dst, mode, swap16)


def emit_fancy_stream(y, u, v, splits, mode: int = MODE_RGBA, swap16: bool = False):
    h, w = y.shape
    bpp = lib().orc_mode_bpp(mode)
    dst = np.zeros((h, w * bpp), np.uint8)
    sp = np.asarray(splits, np.int32)
    n = lib().orc_emit_fancy_rgb_stream(
        mode, _p(np.ascontiguousarray(y)), w, _p(np.ascontiguousarray(u)),
        _p(np.ascontiguousarray(v)), u.shape[1], w, h, _p(sp), len(sp), _p(dst),
        w * bpp)
    return _swap16(dst, mode, swap16), n


def upsample_line_pair(mode, top_y, bottom_y, top_u, top_v, cur_u, cur_v):
    n = len(top_y)
    bpp = lib().orc_mode_bpp(mode)
    top = np.zeros(n * bpp, np.uint8)
    bot = np.zeros(n * bpp, np.uint8) if bottom_y is not None else None
    lib().orc_upsample_line_pair(mode, _p(top_y), _p(bottom_y), _p(top_u),
                                 _p(top_v), _p(cur_u), _p(cur_v), _p(top),
                                 _p(bot), n)
    return top, bot


def add_green(src):
    src = np.ascontiguousarray(src, np.uint32)
    dst = np.empty_like(src)
    lib().orc_add_green(_p(src), src.size, _p(dst))
    return dst


def sub_green(src):
    dst = np.array(src, np.uint32, copy=True)
    lib().orc_sub_green(_p(dst), dst.size)
    return dst


def color_space_inverse(src, codes, bits):
    src = np.ascontiguousarray(src, np.uint32)
    codes = np.ascontiguousarray(codes, np.uint32)
    h, w = src.shape
    dst = np.empty_like(src)
    lib().orc_color_space_inverse(bits, w, _p(codes), 0, h, _p(src), _p(dst))
    return dst


def color_space_forward(src, codes, bits):
    dst = np.array(src, np.uint32, copy=True)
    codes = np.ascontiguousarray(codes, np.uint32)
    h, w = dst.shape
    lib().orc_color_space_forward(bits, w, h, _p(codes), _p(dst))
    return dst


def color_index_inverse(src, color_map, bits, width):
    src = np.ascontiguousarray(src, np.uint32)
    color_map = np.ascontiguousarray(color_map, np.uint32)
    assert color_map.size >= (1 << (8 >> bits))
    h = src.shape[0]
    dst = np.empty((h, width), np.uint32)
    lib().orc_color_index_inverse(bits, width, _p(color_map), 0, h, _p(src),
                                  _p(dst))
    return dst


def color_index_inverse_alpha(src, color_map, bits, width):
    src = np.ascontiguousarray(src, np.uint8)
    color_map = np.ascontiguousarray(color_map, np.uint32)
    h = src.shape[0]
    dst = np.empty((h, width), np.uint8)
    lib().orc_color_index_inverse_alpha(bits, width, _p(color_map), 0, h,
                                        _p(src), _p(dst))
    return dst


def expand_color_map(data, bits):
    data = np.ascontiguousarray(data, np.uint32)
    out = np.zeros(1 << (8 >> bits), np.uint32)
    lib().orc_expand_color_map(data.size, bits, _p(data), _p(out))
    return out


def bundle_color_map(idx, xbits):
    idx = np.ascontiguousarray(idx, np.uint8)
    h, w = idx.shape
    ow = (w + (1 << xbits) - 1) >> xbits
    dst = np.zeros((h, ow), np.uint32)
    for r in range(h):
        lib().orc_bundle_color_map(_p(idx[r]), w, xbits, _p(dst[r]))
    return dst


def apply_palette(src, palette, xbits):
    """ApplyPalette (pkg/vp8/vp8l_encoder.go:1130-1212): ARGB [h, w] -> packed index rows
    [h, ceil(w / 2^xbits)]; also returns the strategy (0 greedy, 1..3 hash, 4 sorted)."""
    src = np.ascontiguousarray(src, np.uint32)
    pal = np.ascontiguousarray(palette, np.uint32).reshape(-1)
    h, w = src.shape
    ow = (w + (1 << xbits) - 1) >> xbits
    dst = np.zeros((h, ow), np.uint32)
    st = lib().orc_apply_palette(_p(src), w, _p(dst), ow, _p(pal), pal.size, w, h, xbits)
    if st < 0:
        raise MemoryError("orc_apply_palette")
    return dst, st


def palette_xbits(palette_size: int) -> int:
    """MapImageFromPalette's rule (vp8l_encoder.go:1229-1233)."""
    if palette_size <= 4:
        return 3 if palette_size <= 2 else 2
    return 1 if palette_size <= 16 else 0


def predictor_inverse(src, modes, bits):
    src = np.ascontiguousarray(src, np.uint32)
    modes = np.ascontiguousarray(modes, np.uint32)
    h, w = src.shape
    dst = np.empty_like(src)
    lib().orc_predictor_inverse(bits, w, h, _p(modes), _p(src), _p(dst))
    return dst


def predictor_forward(src, modes, bits):
    src = np.ascontiguousarray(src, np.uint32)
    modes = np.ascontiguousarray(modes, np.uint32)
    h, w = src.shape
    dst = np.empty_like(src)
    lib().orc_predictor_forward(bits, w, h, _p(modes), _p(src), _p(dst))
    return dst


def argb_to_y(row):
    row = np.ascontiguousarray(row, np.uint32)
    y = np.zeros(row.size, np.uint8)
    lib().orc_ConvertARGBToY(_p(row), _p(y), row.size)
    return y


def argb_to_uv(row, u=None, v=None, do_store=True):
    row = np.ascontiguousarray(row, np.uint32)
    n = (row.size + 1) >> 1
    u = np.zeros(n, np.uint8) if u is None else np.array(u, np.uint8)
    v = np.zeros(n, np.uint8) if v is None else np.array(v, np.uint8)
    lib().orc_ConvertARGBToUV(_p(row), _p(u), _p(v), row.size, int(do_store))
    return u, v


def synth_u32(seed, img, plane, words_per_row, rows, alpha_mode=0):
    out = np.empty((rows, words_per_row), np.uint32)
    lib().orc_synth_fill_u32(C.c_uint32(seed), C.c_uint32(img),
                             C.c_uint32(plane), _p(out), words_per_row, rows,
                             words_per_row * 4, alpha_mode)
    return out


# ---- output conversions (oracle_post.c) ------------------------------
def convert_from_bgra(src, mode, swap16=0):
    """VP8LConvertFromBGRA (lossless.go:616-666) over a [h, w] uint32 image -> uint8 [h, w*bpp]."""
    src = np.ascontiguousarray(src, np.uint32)
    h, w = src.shape
    bpp = lib().orc_mode_bpp(mode)
    dst = np.zeros((h, w * bpp), np.uint8)
    for r in range(h):
        lib().orc_convert_from_bgra(_p(src[r]), w, mode, swap16, _p(dst[r]))
    return dst


def apply_alpha_multiply(rgba, alpha_first):
    out = np.array(rgba, np.uint8)
    h, wb = out.shape
    lib().orc_apply_alpha_multiply(_p(out), int(alpha_first), wb // 4, h, wb)
    return out


def apply_alpha_multiply_4444(rgba4444, rg_byte_pos=1):
    out = np.array(rgba4444, np.uint8)
    h, wb = out.shape
    lib().orc_apply_alpha_multiply_4444(_p(out), wb // 2, h, wb, rg_byte_pos)
    return out


def mult_argb_rows(argb, inverse):
    out = np.array(argb, np.uint32)
    for r in range(out.shape[0]):
        lib().orc_mult_argb_row(_p(out[r]), out.shape[1], int(inverse))
    return out


def mult_rows(plane, alpha, inverse):
    out = np.array(plane, np.uint8)
    alpha = np.ascontiguousarray(alpha, np.uint8)
    for r in range(out.shape[0]):
        lib().orc_mult_row(_p(out[r]), _p(alpha[r]), out.shape[1], int(inverse))
    return out


def sampler_process_plane(y, u, v, mode, swap16: bool = False):
    """WebPSamplerProcessPlane (dsp/yuv.go:47-58) = EmitSampledRGB: point-sampled chroma."""
    y, u, v = (np.ascontiguousarray(x, np.uint8) for x in (y, u, v))
    h, w = y.shape
    bpp = lib().orc_mode_bpp(mode)
    dst = np.zeros((h, w * bpp), np.uint8)
    lib().orc_sampler_process_plane(mode, _p(y), w, _p(u), _p(v), u.shape[1], _p(dst), w * bpp, w, h)
    return _swap16(dst, mode, swap16)


def yuv444_to_rgb(y, u, v, mode, swap16: bool = False):
    y, u, v = (np.ascontiguousarray(x, np.uint8) for x in (y, u, v))
    h, w = y.shape
    bpp = lib().orc_mode_bpp(mode)
    dst = np.zeros((h, w * bpp), np.uint8)
    for r in range(h):
        lib().orc_yuv444_row(mode, _p(y[r]), _p(u[r]), _p(v[r]), _p(dst[r]), w)
    return _swap16(dst, mode, swap16)


# ---- picture pre-processing (oracle_tools.c) -------------------------
def cleanup_transparent_yuva(y, u, v, a):
    """WebPCleanupTransparentArea on Y,U,V,A planes; returns new (y, u, v)."""
    y, u, v = (np.array(p, np.uint8, order="C") for p in (y, u, v))
    a = np.ascontiguousarray(a, np.uint8)
    h, w = y.shape
    lib().orc_cleanup_transparent_yuva(_p(y), _p(u), _p(v), _p(a), y.shape[1], u.shape[1], a.shape[1], w, h)
    return y, u, v


def cleanup_transparent_argb(argb):
    argb = np.array(argb, np.uint32, order="C")
    h, w = argb.shape
    lib().orc_cleanup_transparent_argb(_p(argb), w, w, h)
    return argb


def replace_transparent(argb, color):
    argb = np.array(argb, np.uint32, order="C")
    h, w = argb.shape
    lib().orc_replace_transparent(_p(argb), w, w, h, C.c_uint32(color))
    return argb


def blend_alpha_yuva(y, u, v, a, background_rgb):
    """WebPBlendAlpha on Y,U,V,A planes; returns new (y, u, v, a)."""
    y, u, v, a = (np.array(p, np.uint8, order="C") for p in (y, u, v, a))
    h, w = y.shape
    lib().orc_blend_alpha_yuva(_p(y), _p(u), _p(v), _p(a), y.shape[1], u.shape[1], a.shape[1], w, h,
                               C.c_uint32(background_rgb))
    return y, u, v, a


def blend_alpha_argb(argb, background_rgb):
    argb = np.array(argb, np.uint32, order="C")
    h, w = argb.shape
    lib().orc_blend_alpha_argb(_p(argb), w, w, h, C.c_uint32(background_rgb))
    return argb


_ARGB_ORDERS = {"RGBA": (4, 0, 1), "RGBX": (4, 0, 0), "BGRA": (4, 1, 1), "BGRX": (4, 1, 0), "RGB": (3, 0, 0), "BGR": (3, 1, 0)}


def import_to_argb(src, width, height, order="RGBA", stride=None):
    """Import() with use_argb = 1 -> uint32 (h, w) 0xAARRGGBB."""
    step, swap_rb, import_alpha = _ARGB_ORDERS[order]
    src = np.ascontiguousarray(src, np.uint8).reshape(-1)
    if stride is None:
        stride = step * width
    dst = np.empty((height, width), np.uint32)
    base = src.ctypes.data + (abs(stride) * (height - 1) if stride < 0 else 0)
    ok = lib().orc_import_to_argb(C.c_void_p(base), stride, step, swap_rb, import_alpha, width, height, _p(dst), width)
    assert ok
    return dst


def import_yuva_dithered(src: np.ndarray, width: int, height: int, dithering: float, order: str = "RGBA",
                         stride: int | None = None, divide_mode: int = DIV_EXACT):
    """The dithered import (conversions.go:122-165).  Returns (y,u,v,a,has_alpha)."""
    r, g, b, a_off, step = ORDERS[order]
    src = np.ascontiguousarray(src, np.uint8).reshape(-1)
    if stride is None:
        stride = step * width
    uvw, uvh = (width + 1) >> 1, (height + 1) >> 1
    y = np.zeros((height, width), np.uint8)
    u = np.zeros((uvh, uvw), np.uint8)
    v = np.zeros((uvh, uvw), np.uint8)
    a = np.zeros((height, width), np.uint8)
    ptr = C.c_void_p(src.ctypes.data + ((-stride) * (height - 1) if stride < 0 else 0))
    ha = lib().orc_import_yuva_dithered(ptr, r, g, b, a_off, step, stride, width, height, _p(y), width, _p(u), _p(v),
                                        uvw, _p(a), width, divide_mode, C.c_float(dithering))
    if ha < 0:
        raise ValueError("orc_import_yuva_dithered: bad arguments")
    return y, u, v, a, bool(ha)


def vp8_random(dithering: float, num_bits: int, n: int):
    """First n values of VP8RandomBits(rg, num_bits) after VP8InitRandom(dithering)."""
    out = np.zeros(n, np.int32)
    lib().orc_vp8_random(C.c_float(dithering), num_bits, n, _p(out))
    return out


# ---- animation (oracle_tools.c) ---------------------------------------
def anim_blend_rows(src, dst, premult):
    """BlendPixelRowNonPremult / Premult of every row of src over dst -> new src."""
    src = np.array(src, np.uint32, order="C")
    dst = np.ascontiguousarray(dst, np.uint32)
    h, w = src.shape
    lib().orc_anim_blend_rows(_p(src), w, _p(dst), w, w, h, int(premult))
    return src


def compare_pixels(src, src_step, dst, dst_step, length, lossless=True, max_allowed_diff=0):
    src = np.ascontiguousarray(src, np.uint32)
    dst = np.ascontiguousarray(dst, np.uint32)
    if lossless:
        return int(lib().orc_compare_pixels_lossless(_p(src), src_step, _p(dst), dst_step, length))
    return int(lib().orc_compare_pixels_lossy(_p(src), src_step, _p(dst), dst_step, length, max_allowed_diff))


# ---- alpha-plane filters (oracle_tools.c) -----------------------------
def alpha_unfilter(plane, filt):
    plane = np.ascontiguousarray(plane, np.uint8)
    h, w = plane.shape
    out = np.empty_like(plane)
    lib().orc_alpha_unfilter(int(filt), _p(plane), w, _p(out), w, w, h)
    return out


def alpha_filter(plane, filt):
    plane = np.ascontiguousarray(plane, np.uint8)
    h, w = plane.shape
    out = np.empty_like(plane)
    lib().orc_alpha_filter(int(filt), _p(plane), w, _p(out), w, w, h)
    return out


# ---- rescaler (oracle_tools.c) ----------------------------------------
def rescale_plane(plane, dst_width, dst_height, num_channels=1):
    """RescalePlane: plane uint8 [h, w * num_channels] -> [dst_height, dst_width * num_channels]."""
    plane = np.ascontiguousarray(plane, np.uint8)
    h, wc = plane.shape
    w = wc // num_channels
    out = np.zeros((dst_height, dst_width * num_channels), np.uint8)
    ok = lib().orc_rescale_plane(_p(plane), w, h, wc, _p(out), dst_width, dst_height, dst_width * num_channels, num_channels)
    assert ok
    return out


def picture_rescale_yuva(y, u, v, a, width, height):
    """WebPPictureRescale, YUV420 branch with alpha (picture_rescale_enc.c.go:174-194):
    alpha first, luma premultiplied by the source alpha, rescaled, divided by the new alpha."""
    hw, hh = (width + 1) >> 1, (height + 1) >> 1
    a2 = rescale_plane(a, width, height)
    y2 = rescale_plane(mult_rows(y, a, 0), width, height)
    return mult_rows(y2, a2, 1), rescale_plane(u, hw, hh), rescale_plane(v, hw, hh), a2


def picture_rescale_argb(argb, width, height):
    """WebPPictureRescale, ARGB branch (picture_rescale_enc.c.go:196-212)."""
    pre = np.ascontiguousarray(mult_argb_rows(argb, 0))
    h, w = pre.shape
    out = rescale_plane(pre.view(np.uint8).reshape(h, 4 * w), width, height, 4)
    return mult_argb_rows(np.ascontiguousarray(out).view(np.uint32).reshape(height, width), 1)


def emit_rescaled_rgb(y, u, v, a, width, height, mode):
    """The decoder's rescaled RGB output over a whole picture (io_dec.c.go:361-515): Y to the
    output size, U / V from half size to the output size (InitRGBRescaler :476-478),
    WebPYUV444Converters per row (ExportRGB :361-381), the rescaled alpha dispatched into the
    alpha byte and WebPApplyAlphaMultiply for premultiplied modes (ExportAlpha :405-430).
    a = None or a mode without alpha: the converter's 0xff stays."""
    ys, us, vs = rescale_plane(y, width, height), rescale_plane(u, width, height), rescale_plane(v, width, height)
    out = yuv444_to_rgb(ys, us, vs, mode)
    alpha_modes = {1: 3, 3: 3, 4: 0, 7: 3, 8: 3, 9: 0}     # mode -> alpha byte (RGBA BGRA ARGB rgbA bgrA Argb)
    if a is not None and mode in alpha_modes:
        a_s = rescale_plane(a, width, height)
        out = out.copy()
        out[:, alpha_modes[mode]::4] = a_s
        if mode in (7, 8, 9) and (a_s != 255).any():
            out = apply_alpha_multiply(out, int(alpha_modes[mode] == 0))
    return out


def dequantize_levels(plane, strength, ordered_dither=False):
    """WebPDequantizeLevels on an alpha plane -> new plane.  ordered_dither: the Go file's
    DFIX = 4 branch; False is the stock libwebp build (DFIX = 0)."""
    plane = np.array(plane, np.uint8, order="C")
    h, w = plane.shape
    ok = lib().orc_dequantize_levels(_p(plane), w, h, w, int(strength), int(bool(ordered_dither)))
    assert ok
    return plane
