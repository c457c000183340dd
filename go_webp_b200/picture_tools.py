"""Host-side mirror of pkg/libwebp/enc/picture_tools_enc.c.go: the picture
pre-processing the encoder runs after the import.  Same names, argument meaning and
"nothing to do" rules as the reference; the work happens in b2_tools.cu through the C
ABI (no CPU fallback), on the pictures' planes staged to the device and copied back in
place.  One added batched entry point per function (`...Batch`)."""
from __future__ import annotations

import numpy as np

from . import _abi
from ._abi import B2Error
from .device import Context
from .picture import (ENC_ERROR_OUT_OF_MEMORY, WEBP_CSP_ALPHA_BIT, Picture, _al, _arena,
                      _STATUS_TO_ENC, default_context)


def _launch(ctx, op, arr, n, use_argb, arg):
    lib, hnd = ctx.lib, ctx.h
    if op == "cleanup":
        ctx._chk(lib.b2_picture_cleanup_transparent(hnd, 0, arr, n, int(use_argb)))
    elif op == "blend":
        ctx._chk(lib.b2_picture_blend_alpha(hnd, 0, arr, n, int(use_argb), arg & 0xffffffff))
    else:
        ctx._chk(lib.b2_picture_replace_transparent(hnd, 0, arr, n, arg & 0xffffffff))


def _run_resident(pics, op: str, arg: int, ctx: Context, download: bool):
    """Pictures whose planes already live on the device (ImportRGBABatch, YUVAToARGBBatch):
    the operation runs in place on the device planes; nothing is uploaded, and the result
    comes back (one copy per picture) only when download is set."""
    lib, hnd = ctx.lib, ctx.h
    for use_argb in (False, True):
        sel = [p for p in pics if bool(p.UseARGB) == use_argb]
        if not sel:
            continue
        jobs = (_abi.PictureJob * len(sel))()
        for j, pic in zip(jobs, sel):
            if use_argb:
                m = pic.memory_argb_
                j.argb, j.argb_stride, j.width, j.height = m.dev("argb"), m.stride, m.w, m.h
            else:
                m = pic.memory_
                j.y, j.u, j.v, j.a = m.dev("y"), m.dev("u"), m.dev("v"), m.dev("a")
                j.y_stride, j.uv_stride, j.a_stride, j.width, j.height = m.ys, m.uvs, m.ys, m.w, m.h
        _launch(ctx, op, jobs, len(sel), use_argb, arg)
        for k, pic in enumerate(sel):
            m = pic.memory_argb_ if use_argb else pic.memory_
            m.host_valid = False
            if download:
                ctx._chk(lib.b2_d2h_async(hnd, 0, m.block.pin.ptr + m.off,
                                          m.block.dev.ptr + m.off, m.nbytes))
                m.host_valid = True
    ctx.sync()


def _run(pics, op: str, arg: int = 0, ctx: Context | None = None, download: bool = True):
    ctx = ctx or default_context()
    lib, hnd = ctx.lib, ctx.h
    res = [p for p in pics if p.resident("argb" if p.UseARGB else "yuva")
           and (p.memory_argb_ if p.UseARGB else p.memory_).block.ctx is ctx]
    if res:
        _run_resident(res, op, arg, ctx, download)
        pics = [p for p in pics if not any(p is r for r in res)]
        if not pics:
            return
    metas, need = [], 0
    for pic in pics:
        w, h = pic.Width, pic.Height
        if pic.UseARGB:
            st = _al(w, 4)
            sz = dict(argb=_al(4 * st * h))
            metas.append((True, w, h, st, 0, 0, sz))
        else:
            uvw, uvh = (w + 1) >> 1, (h + 1) >> 1
            ys, uvs = _al(w, 16), _al(uvw, 16)
            sz = dict(y=_al(ys * h), u=_al(uvs * uvh), v=_al(uvs * uvh), a=_al(ys * h))
            metas.append((False, w, h, ys, uvs, (uvw, uvh), sz))
        need += sum(sz.values())
    ar = _arena(ctx)
    ar.reset(need)
    for use_argb in (False, True):
        idx = [i for i, m in enumerate(metas) if m[0] == use_argb]
        if not idx:
            continue
        jobs, ptrs = [], []
        for i in idx:
            _, w, h, s0, s1, uvdim, sz = metas[i]
            p = {k: ar.take(v) for k, v in sz.items()}
            ptrs.append(p)
            pic = pics[i]
            if use_argb:
                A = np.ascontiguousarray(pic.ARGB, np.uint32)
                ctx._chk(lib.b2_h2d_2d_async(hnd, 0, p["argb"], 4 * s0, A.ctypes.data, 4 * A.shape[1], 4 * w, h))
                jobs.append(_abi.PictureJob(None, None, None, None, p["argb"], 0, 0, 0, s0, w, h))
            else:
                uvw, uvh = uvdim
                for key, arr, stride, ww, hh in (("y", pic.Y, s0, w, h), ("u", pic.U, s1, uvw, uvh),
                                                 ("v", pic.V, s1, uvw, uvh), ("a", pic.A, s0, w, h)):
                    arr = np.ascontiguousarray(arr, np.uint8)
                    ctx._chk(lib.b2_h2d_2d_async(hnd, 0, p[key], stride, arr.ctypes.data, arr.shape[1], ww, hh))
                jobs.append(_abi.PictureJob(p["y"], p["u"], p["v"], p["a"], None, s0, s1, s0, 0, w, h))
        arr = (_abi.PictureJob * len(jobs))(*jobs)
        _launch(ctx, op, arr, len(jobs), use_argb, arg)
        for i, p in zip(idx, ptrs):
            _, w, h, s0, s1, uvdim, sz = metas[i]
            pic = pics[i]
            if use_argb:
                out = np.empty((h, w), np.uint32)
                ctx._chk(lib.b2_d2h_2d_async(hnd, 0, out.ctypes.data, 4 * w, p["argb"], 4 * s0, 4 * w, h))
                pic.ARGB, pic.ARGBStride = out, w
            else:
                uvw, uvh = uvdim
                outs = {}
                for key, stride, ww, hh in (("y", s0, w, h), ("u", s1, uvw, uvh), ("v", s1, uvw, uvh), ("a", s0, w, h)):
                    o = np.empty((hh, ww), np.uint8)
                    ctx._chk(lib.b2_d2h_2d_async(hnd, 0, o.ctypes.data, ww, p[key], stride, ww, hh))
                    outs[key] = o
                pic.Y, pic.U, pic.V, pic.A = outs["y"], outs["u"], outs["v"], outs["a"]
                pic.YStride, pic.UVStride, pic.AStride = w, uvw, w
    ctx.sync()


def _guarded(pics, op, arg, ctx, download=True):
    try:
        _run(pics, op, arg, ctx, download)
    except B2Error as e:
        for pic in pics:
            pic.SetEncodingError(_STATUS_TO_ENC.get(e.status, ENC_ERROR_OUT_OF_MEMORY))


def WebPCleanupTransparentAreaBatch(pics, ctx: Context | None = None, download: bool = True) -> None:
    """Added batched entry point.  picture_tools_enc.c.go:110-183: pictures without the
    planes the function needs are left alone (:142-144).  Device-resident pictures are
    processed where they lie (webp_encoding.go:49-72: import -> cleanup without a round trip);
    download=False leaves the result on the device for the next call."""
    todo = [p for p in pics if p is not None and
            (p._ARGB is not None if p.UseARGB else all(x is not None for x in (p._Y, p._U, p._V, p._A)))]
    if todo:
        _guarded(todo, "cleanup", 0, ctx, download)


def WebPCleanupTransparentArea(pic: Picture) -> None:
    """picture_tools_enc.c.go:110."""
    if pic is not None:
        WebPCleanupTransparentAreaBatch([pic])


def WebPBlendAlphaBatch(pics, background_rgb: int, ctx: Context | None = None, download: bool = True) -> None:
    """Added batched entry point.  picture_tools_enc.c.go:199-276: a YUV picture without
    alpha (colour space bit or plane) has nothing to blend (:218-219); afterwards every
    alpha value is 0xff."""
    todo = []
    for p in pics:
        if p is None:
            continue
        if p.UseARGB:
            if p._ARGB is not None:
                todo.append(p)
        elif (p.ColorSpace & WEBP_CSP_ALPHA_BIT) and p._A is not None:
            todo.append(p)
    if todo:
        _guarded(todo, "blend", background_rgb, ctx, download)


def WebPBlendAlpha(pic: Picture, background_rgb: int) -> None:
    """picture_tools_enc.c.go:199."""
    if pic is not None:
        WebPBlendAlphaBatch([pic], background_rgb)


def WebPReplaceTransparentPixelsBatch(pics, color: int, ctx: Context | None = None, download: bool = True) -> None:
    """Added batched entry point.  picture_tools_enc.c.go:95-108: ARGB pictures only."""
    todo = [p for p in pics if p is not None and p.UseARGB and p._ARGB is not None]
    if todo:
        _guarded(todo, "replace", color, ctx, download)


def WebPReplaceTransparentPixels(pic: Picture, color: int) -> None:
    """picture_tools_enc.c.go:95."""
    if pic is not None:
        WebPReplaceTransparentPixelsBatch([pic], color)
