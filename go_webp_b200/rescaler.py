"""Host-side mirror of the rescaler: RescalePlane (pkg/libwebp/enc/
picture_rescale_enc.c.go:125-136) over WebPRescalerImport / WebPRescalerExport
(pkg/libwebp/utils/rescaler_utils.c.go:119-158), WebPRescalerGetScaledDimensions (:88-113)
and WebPPictureRescale (:149-220: YUV420 with or without alpha, ARGB), plus one added
batched entry point.  The work happens in b2_rescale.cu through the C ABI (no CPU fallback)."""
from __future__ import annotations

import numpy as np

from . import _abi
from .device import Context
from .picture import (ENC_ERROR_BAD_DIMENSION, ENC_ERROR_OUT_OF_MEMORY, Picture, _al, _arena, _STATUS_TO_ENC,
                      default_context)
from ._abi import B2Error


def WebPRescalerGetScaledDimensions(src_width: int, src_height: int, scaled_width: int, scaled_height: int):
    """rescaler_utils.c.go:88-113: returns (ok, width, height); a 0 dimension follows the ratio."""
    width, height = scaled_width, scaled_height
    max_size = (2**31 - 1) // 2
    if width == 0 and src_height > 0:
        width = (src_width * height + src_height - 1) // src_height
    if height == 0 and src_width > 0:
        height = (src_height * width + src_width - 1) // src_width
    if width <= 0 or height <= 0 or width > max_size or height > max_size:
        return 0, scaled_width, scaled_height
    return 1, width, height


def RescalePlanesBatch(planes, sizes, num_channels: int = 1, ctx: Context | None = None):
    """Added batched entry point: planes[i] (uint8 [h, w * num_channels]) rescaled to
    sizes[i] = (dst_width, dst_height).  Returns the list of outputs."""
    ctx = ctx or default_context()
    lib, hnd = ctx.lib, ctx.h
    n = len(planes)
    if n == 0:
        return []
    metas, need = [], 0
    for p, (dw, dh) in zip(planes, sizes):
        h, wc = p.shape
        ss, ds = _al(wc, 16), _al(dw * num_channels, 16)
        metas.append((wc // num_channels, h, ss, dw, dh, ds))
        need += _al(ss * h) + _al(ds * dh)
    ar = _arena(ctx)
    ar.reset(need)
    jobs, ptrs = [], []
    for (w, h, ss, dw, dh, ds), p in zip(metas, planes):
        ps, pd = ar.take(ss * h), ar.take(ds * dh)
        p = np.ascontiguousarray(p, np.uint8)
        ctx._chk(lib.b2_h2d_2d_async(hnd, 0, ps, ss, p.ctypes.data, w * num_channels, w * num_channels, h))
        jobs.append(_abi.RescaleJob(ps, pd, w, h, ss, dw, dh, ds, num_channels, 0))
        ptrs.append(pd)
    arr = (_abi.RescaleJob * n)(*jobs)
    ctx._chk(lib.b2_rescale_planes(hnd, 0, arr, n))
    outs = []
    for (w, h, ss, dw, dh, ds), pd in zip(metas, ptrs):
        o = np.empty((dh, dw * num_channels), np.uint8)
        ctx._chk(lib.b2_d2h_2d_async(hnd, 0, o.ctypes.data, dw * num_channels, pd, ds, dw * num_channels, dh))
        outs.append(o)
    ctx.sync()
    return outs


def RescalePlane(src: np.ndarray, dst_width: int, dst_height: int, num_channels: int = 1) -> np.ndarray:
    """picture_rescale_enc.c.go:125-136."""
    return RescalePlanesBatch([src], [(dst_width, dst_height)], num_channels)[0]


def WebPPictureRescale(pic: Picture, width: int, height: int) -> int:
    """picture_rescale_enc.c.go:149-220.  YUV420: alpha (if any) is rescaled first, luma is
    premultiplied by the source alpha (AlphaMultiplyY :143-147), Y / U / V are rescaled, luma
    is divided by the new alpha.  ARGB: premultiply (AlphaMultiplyARGB :138-141), 4-channel
    rescale, un-premultiply."""
    from . import dsp
    if pic is None:
        return 0
    prev_w, prev_h = pic.Width, pic.Height
    ok, width, height = WebPRescalerGetScaledDimensions(prev_w, prev_h, width, height)
    if not ok:
        return pic.SetEncodingError(ENC_ERROR_BAD_DIMENSION)
    halve = lambda x: (x + 1) >> 1  # noqa: E731
    try:
        if not pic.UseARGB:
            crop = lambda p, w, h: np.ascontiguousarray(np.asarray(p)[:h, :w])  # noqa: E731
            y, u, v = crop(pic.Y, prev_w, prev_h), crop(pic.U, halve(prev_w), halve(prev_h)), crop(pic.V, halve(prev_w), halve(prev_h))
            a2 = None
            if pic.A is not None:
                a = crop(pic.A, prev_w, prev_h)
                y = dsp.WebPMultRows(y, a, prev_w, prev_h, 0)
                y2, a2 = RescalePlanesBatch([y, a], [(width, height)] * 2)
                y2 = dsp.WebPMultRows(y2, a2, width, height, 1)
            else:
                y2 = RescalePlane(y, width, height)
            u2, v2 = RescalePlanesBatch([u, v], [(halve(width), halve(height))] * 2)
            pic.Y, pic.U, pic.V, pic.A = y2, u2, v2, a2
            pic.YStride, pic.UVStride, pic.AStride = width, halve(width), (width if a2 is not None else 0)
        else:
            argb = np.ascontiguousarray(np.asarray(pic.ARGB, np.uint32)[:prev_h, :prev_w])
            pre = np.ascontiguousarray(dsp.WebPMultARGBRows(argb, prev_w, prev_h, 0))
            out = RescalePlane(pre.view(np.uint8).reshape(prev_h, 4 * prev_w), width, height, 4)
            pic.ARGB = dsp.WebPMultARGBRows(np.ascontiguousarray(out).view(np.uint32).reshape(height, width), width, height, 1)
            pic.ARGBStride = width
    except B2Error as e:
        return pic.SetEncodingError(_STATUS_TO_ENC.get(e.status, ENC_ERROR_OUT_OF_MEMORY))
    pic.Width, pic.Height = width, height
    return 1


_ALPHA_BYTE = {_abi.MODE_RGBA: 3, _abi.MODE_BGRA: 3, _abi.MODE_ARGB: 0, _abi.MODE_rgbA: 3, _abi.MODE_bgrA: 3, _abi.MODE_Argb: 0}
_PREMULTIPLIED = (_abi.MODE_rgbA, _abi.MODE_bgrA, _abi.MODE_Argb)


def EmitRescaledRGB(y: np.ndarray, u: np.ndarray, v: np.ndarray, a: np.ndarray | None, scaled_width: int, scaled_height: int,
                    mode: int) -> np.ndarray:
    """The decoder's rescaled RGB output over a whole picture (io_dec.c.go:361-515; the row
    batches of EmitRescaledRGB feed the same streaming rescalers, so the result is the same):
    InitRGBRescaler's three rescalers (Y to the output size, U and V from half size straight
    to the output size, :476-478), WebPYUV444Converters[mode] per exported row (ExportRGB
    :361-381), then EmitRescaledAlphaRGB / ExportAlpha (:405-430): the rescaled alpha through
    WebPDispatchAlpha and, for premultiplied modes, WebPApplyAlphaMultiply.  Returns
    uint8 [scaled_height, scaled_width * bpp]."""
    from . import dsp
    ys = RescalePlane(y, scaled_width, scaled_height)
    us, vs = RescalePlanesBatch([u, v], [(scaled_width, scaled_height)] * 2)
    out = dsp.WebPYUV444ToRGB(ys, us, vs, mode)
    if a is not None and mode in _ALPHA_BYTE:
        alpha_first = int(_ALPHA_BYTE[mode] == 0)
        a_s = RescalePlane(a, scaled_width, scaled_height)
        out = dsp.WebPDispatchAlpha(a_s, scaled_width, scaled_height, out, 4 * scaled_width, alpha_first)
        if mode in _PREMULTIPLIED:
            out = dsp.WebPApplyAlphaMultiply(out, alpha_first, scaled_width, scaled_height, 4 * scaled_width)
    return out


def EmitRescaledYUV(y: np.ndarray, u: np.ndarray, v: np.ndarray, a: np.ndarray | None, scaled_width: int, scaled_height: int):
    """EmitRescaledYUV + EmitRescaledAlphaYUV over a whole picture (io_dec.c.go:244-281) for a
    MODE_YUVA output: luma premultiplied by the source alpha before its rescale, divided by
    the rescaled alpha after; an absent alpha plane is delivered opaque (:272-276).
    Returns (y, u, v, a)."""
    from . import dsp
    h, w = y.shape
    halve = lambda x: (x + 1) >> 1  # noqa: E731
    if a is not None:
        y = dsp.WebPMultRows(y, a, w, h, 0)
        ys, a_s = RescalePlanesBatch([y, a], [(scaled_width, scaled_height)] * 2)
        ys = dsp.WebPMultRows(ys, a_s, scaled_width, scaled_height, 1)
    else:
        ys = RescalePlane(y, scaled_width, scaled_height)
        a_s = np.full((scaled_height, scaled_width), 255, np.uint8)
    us, vs = RescalePlanesBatch([u, v], [(halve(scaled_width), halve(scaled_height))] * 2)
    return ys, us, vs, a_s
