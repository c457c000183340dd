"""Host-side mirror of the animation pixel functions:
BlendPixelRowNonPremult / BlendPixelRowPremult (pkg/libwebp/demux/anim_decode.go:234-270)
and ComparePixelsLossless / ComparePixelsLossy (pkg/libwebp/mux/anim_encode.go:318-367),
same names and argument meaning, plus one batched entry point each.  The work happens in
b2_anim.cu through the C ABI (no CPU fallback)."""
from __future__ import annotations

import numpy as np

from . import _abi
from .device import Context
from .picture import _al, _arena, default_context


def BlendRowsBatch(srcs, dsts, premultiplied: bool, ctx: Context | None = None):
    """Added batched entry point: srcs[i] (uint32 [h, w], the frame as decoded into the
    canvas) blended over dsts[i] (the previous canvas rows).  Returns the blended arrays."""
    ctx = ctx or default_context()
    lib, hnd = ctx.lib, ctx.h
    n = len(srcs)
    if n == 0:
        return []
    metas, need = [], 0
    for s, d in zip(srcs, dsts):
        h, w = s.shape
        assert d.shape == s.shape
        st = _al(w, 4)
        metas.append((w, h, st))
        need += 2 * _al(4 * st * h)
    ar = _arena(ctx)
    ar.reset(need)
    jobs, ptrs = [], []
    for (w, h, st), s, d in zip(metas, srcs, dsts):
        ps, pd = ar.take(4 * st * h), ar.take(4 * st * h)
        s = np.ascontiguousarray(s, np.uint32)
        d = np.ascontiguousarray(d, np.uint32)
        ctx._chk(lib.b2_h2d_2d_async(hnd, 0, ps, 4 * st, s.ctypes.data, 4 * w, 4 * w, h))
        ctx._chk(lib.b2_h2d_2d_async(hnd, 0, pd, 4 * st, d.ctypes.data, 4 * w, 4 * w, h))
        jobs.append(_abi.AnimBlendJob(ps, pd, st, st, w, h))
        ptrs.append(ps)
    arr = (_abi.AnimBlendJob * n)(*jobs)
    ctx._chk(lib.b2_anim_blend_rows(hnd, 0, arr, n, int(bool(premultiplied))))
    outs = []
    for (w, h, st), ps in zip(metas, ptrs):
        o = np.empty((h, w), np.uint32)
        ctx._chk(lib.b2_d2h_2d_async(hnd, 0, o.ctypes.data, 4 * w, ps, 4 * st, 4 * w, h))
        outs.append(o)
    ctx.sync()
    return outs


def BlendPixelRowNonPremult(src: np.ndarray, dst: np.ndarray, num_pixels: int) -> None:
    """anim_decode.go:234-243: src[:num_pixels] blended in place over dst."""
    src[:num_pixels] = BlendRowsBatch([src[:num_pixels].reshape(1, -1)], [np.asarray(dst[:num_pixels]).reshape(1, -1)], False)[0][0]


def BlendPixelRowPremult(src: np.ndarray, dst: np.ndarray, num_pixels: int) -> None:
    """anim_decode.go:259-270."""
    src[:num_pixels] = BlendRowsBatch([src[:num_pixels].reshape(1, -1)], [np.asarray(dst[:num_pixels]).reshape(1, -1)], True)[0][0]


def ComparePixelsBatch(src_img: np.ndarray, dst_img: np.ndarray, runs, lossless: bool, max_allowed_diff: int = 0,
                       ctx: Context | None = None):
    """Added batched entry point: runs = [(src_offset, src_step, dst_offset, dst_step, length)]
    in pixels over two flat uint32 images (rows and columns of the frame rectangles the
    encoder scans, anim_encode.go:369-480).  Returns a list of 0 / 1."""
    ctx = ctx or default_context()
    lib, hnd = ctx.lib, ctx.h
    if not runs:
        return []
    a = np.ascontiguousarray(src_img, np.uint32).reshape(-1)
    b = np.ascontiguousarray(dst_img, np.uint32).reshape(-1)
    ar = _arena(ctx)
    ar.reset(_al(a.nbytes) + _al(b.nbytes) + _al(4 * len(runs)))
    pa, pb, pr = ar.take(a.nbytes), ar.take(b.nbytes), ar.take(4 * len(runs))
    ctx._chk(lib.b2_h2d_async(hnd, 0, pa, a.ctypes.data, a.nbytes))
    ctx._chk(lib.b2_h2d_async(hnd, 0, pb, b.ctypes.data, b.nbytes))
    jobs = []
    for (so, ss, do, ds, ln) in runs:
        assert ln > 0 and 0 <= so and so + (ln - 1) * ss < a.size and 0 <= do and do + (ln - 1) * ds < b.size
        jobs.append(_abi.AnimCompareJob(pa + 4 * so, pb + 4 * do, ss, ds, ln, 0))
    arr = (_abi.AnimCompareJob * len(jobs))(*jobs)
    ctx._chk(lib.b2_anim_compare_pixels(hnd, 0, arr, len(jobs), int(bool(lossless)), int(max_allowed_diff), pr))
    out = np.empty(len(jobs), np.int32)
    ctx._chk(lib.b2_d2h_async(hnd, 0, out.ctypes.data, pr, out.nbytes))
    ctx.sync()
    return out.tolist()


def ComparePixelsLossless(src, src_step, dst, dst_step, length, max_allowed_diff=0) -> int:
    """anim_encode.go:318-332 (max_allowed_diff unused, as there)."""
    return ComparePixelsBatch(src, dst, [(0, src_step, 0, dst_step, length)], True)[0]


def ComparePixelsLossy(src, src_step, dst, dst_step, length, max_allowed_diff) -> int:
    """anim_encode.go:355-367."""
    return ComparePixelsBatch(src, dst, [(0, src_step, 0, dst_step, length)], False, max_allowed_diff)[0]
