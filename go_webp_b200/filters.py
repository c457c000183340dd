"""Host-side mirror of the alpha-plane filters: the plane-level forward
filters `WebPFilters[...]` (HorizontalFilter_C / VerticalFilter_C / GradientFilter_C,
pkg/libwebp/dsp/filters.go:114-124) and the unfilter loop of the alpha decoder
(`WebPUnfilters[...]` row after row, pkg/vp8/vp8l_dec.c.go:802-814), plus one batched entry
point each.  The work happens in b2_filters.cu through the C ABI (no CPU fallback)."""
from __future__ import annotations

import numpy as np

from . import _abi
from .device import Context
from .picture import _al, _arena, default_context

WEBP_FILTER_NONE, WEBP_FILTER_HORIZONTAL, WEBP_FILTER_VERTICAL, WEBP_FILTER_GRADIENT = 0, 1, 2, 3


def _run(planes, filters, unfilter: bool, ctx: Context | None, row_align: int = 16, ptr_offset: int = 0):
    """row_align: alignment of the device row stride (1 = tightly packed rows); ptr_offset: bytes added to
    both plane pointers -- the two knobs that take the kernels off their word-aligned paths."""
    ctx = ctx or default_context()
    lib, hnd = ctx.lib, ctx.h
    n = len(planes)
    if n == 0:
        return []
    metas, need = [], 0
    for p in planes:
        h, w = p.shape
        st = _al(w, row_align)
        metas.append((w, h, st))
        need += 2 * _al(st * h + ptr_offset)
    ar = _arena(ctx)
    ar.reset(need)
    jobs, outs_ptr = [], []
    for (w, h, st), p, f in zip(metas, planes, filters):
        ps, pd = ar.take(st * h + ptr_offset) + ptr_offset, ar.take(st * h + ptr_offset) + ptr_offset
        p = np.ascontiguousarray(p, np.uint8)
        ctx._chk(lib.b2_h2d_2d_async(hnd, 0, ps, st, p.ctypes.data, w, w, h))
        jobs.append(_abi.AlphaFilterJob(ps, pd, st, st, w, h, int(f), 0))
        outs_ptr.append(pd)
    arr = (_abi.AlphaFilterJob * n)(*jobs)
    ctx._chk((lib.b2_alpha_unfilter if unfilter else lib.b2_alpha_filter)(hnd, 0, arr, n))
    outs = []
    for (w, h, st), pd in zip(metas, outs_ptr):
        o = np.empty((h, w), np.uint8)
        ctx._chk(lib.b2_d2h_2d_async(hnd, 0, o.ctypes.data, w, pd, st, w, h))
        outs.append(o)
    ctx.sync()
    return outs


def DequantizeLevelsBatch(planes, strengths, ordered_dither: bool = True, ctx: Context | None = None,
                          row_align: int = 16):
    """Added batched entry point: WebPDequantizeLevels on every plane (uint8 [h, w]) with
    strengths[i] in [0, 100] -> new planes.  ordered_dither=True is the Go file as written
    (DFIX = 4, quant_levels_dec_utils.c.go:36-45); False is the stock libwebp build.
    row_align: alignment of the device row stride (1 = tightly packed rows)."""
    ctx = ctx or default_context()
    lib, hnd = ctx.lib, ctx.h
    n = len(planes)
    if n == 0:
        return []
    metas, need = [], 0
    for p in planes:
        h, w = p.shape
        st = _al(w, row_align)
        metas.append((w, h, st))
        need += 2 * _al(st * h)
    ar = _arena(ctx)
    ar.reset(need)
    jobs, outs_ptr = [], []
    for (w, h, st), p, strength in zip(metas, planes, strengths):
        ps, pd = ar.take(st * h), ar.take(st * h)
        p = np.ascontiguousarray(p, np.uint8)
        ctx._chk(lib.b2_h2d_2d_async(hnd, 0, ps, st, p.ctypes.data, w, w, h))
        jobs.append(_abi.AlphaDequantJob(ps, pd, st, st, w, h, int(strength), 0))
        outs_ptr.append(pd)
    arr = (_abi.AlphaDequantJob * n)(*jobs)
    ctx._chk(lib.b2_alpha_dequantize_levels(hnd, 0, arr, n, int(bool(ordered_dither))))
    outs = []
    for (w, h, st), pd in zip(metas, outs_ptr):
        o = np.empty((h, w), np.uint8)
        ctx._chk(lib.b2_d2h_2d_async(hnd, 0, o.ctypes.data, w, pd, st, w, h))
        outs.append(o)
    ctx.sync()
    return outs


def WebPDequantizeLevels(data: np.ndarray, width: int, height: int, stride: int, strength: int,
                         ordered_dither: bool = True) -> bool:
    """quant_levels_dec_utils.c.go:259-286: smooths the alpha plane `data` (flat uint8, row
    stride `stride`) in place; False for a strength outside [0, 100] or an empty plane."""
    if strength < 0 or strength > 100 or data is None or width <= 0 or height <= 0:
        return False
    flat = np.asarray(data).reshape(-1)
    plane = np.stack([flat[y * stride:y * stride + width] for y in range(height)])
    out = DequantizeLevelsBatch([plane], [strength], ordered_dither)[0]
    for y in range(height):
        flat[y * stride:y * stride + width] = out[y]
    return True


def AlphaUnfilterBatch(planes, filters, ctx: Context | None = None, row_align: int = 16, ptr_offset: int = 0):
    """Added batched entry point: planes[i] (uint8 [h, w], the filtered alpha plane the ALPH
    chunk decodes to) unfiltered with WebPUnfilters[filters[i]] exactly as the row loop of
    the alpha decoder does (prev = the previous output row, nil for the first)."""
    return _run(planes, filters, True, ctx, row_align, ptr_offset)


def AlphaFilterBatch(planes, filters, ctx: Context | None = None, row_align: int = 16, ptr_offset: int = 0):
    """Added batched entry point: WebPFilters[filters[i]] on whole planes (filter 0 copies)."""
    return _run(planes, filters, False, ctx, row_align, ptr_offset)


def HorizontalFilter_C(data: np.ndarray, width: int, height: int, stride: int, filtered_data: np.ndarray) -> None:
    """filters.go:114-116 (data / filtered_data: flat uint8 with row stride `stride`)."""
    _plane_filter(WEBP_FILTER_HORIZONTAL, data, width, height, stride, filtered_data)


def VerticalFilter_C(data, width, height, stride, filtered_data) -> None:
    """filters.go:118-120."""
    _plane_filter(WEBP_FILTER_VERTICAL, data, width, height, stride, filtered_data)


def GradientFilter_C(data, width, height, stride, filtered_data) -> None:
    """filters.go:122-124."""
    _plane_filter(WEBP_FILTER_GRADIENT, data, width, height, stride, filtered_data)


def _plane_filter(f, data, width, height, stride, filtered_data):
    src = np.asarray(data, np.uint8).reshape(-1)[:stride * (height - 1) + width]
    plane = np.stack([src[y * stride:y * stride + width] for y in range(height)])
    out = AlphaFilterBatch([plane], [f])[0]
    dst = np.asarray(filtered_data).reshape(-1)
    for y in range(height):
        dst[y * stride:y * stride + width] = out[y]
