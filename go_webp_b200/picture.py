"""Host-side mirror of go-webp's picture API for the pixel hot path.

Names, argument meaning and error behaviour follow the reference:
  Picture, WebPPictureAllocYUVA/AllocARGB      pkg/picture/picture.go:33-105,
                                               pkg/picture/allocations.go:32-122
  WebPPictureImportRGBA/RGBX/BGRA/BGRX/RGB/BGR pkg/libwebp/enc/picture_csp_enc.c.go:123-225
  PictureARGBToYUVA / WebPPictureARGBToYUVA    pkg/picture/conversions.go:12-48
  WebPPictureYUVAToARGB                        pkg/libwebp/enc/picture_csp_enc.c.go:62-118
plus the added batched entry points ImportRGBABatch and YUVAToRGBABatch
(SURVEY.md 8b).  Every conversion dispatches to the CUDA library through the C
ABI; there is no CPU path here (the Go version of this file is go/pkg/picture,
see INTEGRATION.md).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _abi
from ._abi import B2Error
from .device import Context

# pkg/color/colorspace/colorspace.go:4-16
WEBP_YUV420, WEBP_YUV420A, WEBP_CSP_UV_MASK, WEBP_CSP_ALPHA_BIT = 0, 4, 3, 4
WEBP_MAX_DIMENSION = 16383  # pkg/libwebp/webp/encode.h.go:95

# pkg/picture/errors.go:8-21 (sentinels; ENC_OK is None like Go's nil)
ENC_OK = None
ENC_ERROR_OUT_OF_MEMORY = "memory error allocating objects"
ENC_ERROR_nil_PARAMETER = "a pointer parameter is nil"
ENC_ERROR_INVALID_CONFIGURATION = "configuration is invalid"
ENC_ERROR_BAD_DIMENSION = "picture has invalid width/height"

_STATUS_TO_ENC = {
    _abi.B2_ERR_OOM: ENC_ERROR_OUT_OF_MEMORY,
    _abi.B2_ERR_BAD_DIM: ENC_ERROR_BAD_DIMENSION,
    _abi.B2_ERR_NULL: ENC_ERROR_nil_PARAMETER,
    _abi.B2_ERR_INVALID_CONFIG: ENC_ERROR_INVALID_CONFIGURATION,
    _abi.B2_ERR_CUDA: ENC_ERROR_OUT_OF_MEMORY,
}

_CTX: dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    if device not in _CTX:
        _CTX[device] = Context(device, n_streams=4)
    return _CTX[device]


@dataclass
class Picture:
    """pkg/picture/picture.go:33-105 with the fields the code actually uses
    (SURVEY Appendix B-10): planes as numpy arrays, strides in bytes/pixels."""
    Width: int = 0
    Height: int = 0
    UseARGB: bool = False
    ColorSpace: int = WEBP_YUV420
    Y: np.ndarray | None = None
    U: np.ndarray | None = None
    V: np.ndarray | None = None
    A: np.ndarray | None = None
    YStride: int = 0
    UVStride: int = 0
    AStride: int = 0
    ARGB: np.ndarray | None = None
    ARGBStride: int = 0
    ErrorCode: str | None = None
    UserData: object = field(default=None, repr=False)

    def SetEncodingError(self, err):  # picture.go:220-229: the oldest error wins
        if self.ErrorCode is None:
            self.ErrorCode = err
        return 0


def WebPValidatePicture(pic: Picture):
    if pic is None:
        return ENC_ERROR_nil_PARAMETER
    if pic.Width <= 0 or pic.Height <= 0 or pic.Width > WEBP_MAX_DIMENSION or pic.Height > WEBP_MAX_DIMENSION:
        return ENC_ERROR_BAD_DIMENSION
    return None


def WebPPictureAllocYUVA(pic: Picture) -> int:
    """allocations.go:32-93: y_stride = width, uv_stride = ceil(w/2), A iff the
    colourspace has the alpha bit."""
    err = WebPValidatePicture(pic)
    if err:
        return pic.SetEncodingError(err)
    w, h = pic.Width, pic.Height
    uvw, uvh = (w + 1) >> 1, (h + 1) >> 1
    has_alpha = bool(pic.ColorSpace & WEBP_CSP_ALPHA_BIT)
    pic.Y = np.empty((h, w), np.uint8)
    pic.U = np.empty((uvh, uvw), np.uint8)
    pic.V = np.empty((uvh, uvw), np.uint8)
    pic.A = np.empty((h, w), np.uint8) if has_alpha else None
    pic.YStride, pic.UVStride, pic.AStride = w, uvw, (w if has_alpha else 0)
    return 1


def WebPPictureAllocARGB(pic: Picture) -> int:
    err = WebPValidatePicture(pic)
    if err:
        return pic.SetEncodingError(err)
    pic.ARGB = np.empty((pic.Height, pic.Width), np.uint32)
    pic.ARGBStride = pic.Width
    return 1


class _Arena:
    """Grow-only device arena reused across batched calls of one context."""

    def __init__(self, ctx: Context):
        self.ctx, self.buf, self.cap, self.off = ctx, None, 0, 0

    def reset(self, need: int):
        if need > self.cap:
            if self.buf is not None:
                self.ctx.sync()
                self.buf.free()
            self.cap = int(need * 1.25) + (1 << 20)
            self.buf = self.ctx.alloc(self.cap)
        self.off = 0

    def take(self, nbytes: int) -> int:
        p = self.buf.ptr + self.off
        self.off += (nbytes + 255) & ~255
        assert self.off <= self.cap
        return p


_ARENAS: dict[int, _Arena] = {}


def _arena(ctx: Context) -> _Arena:
    a = _ARENAS.get(id(ctx))
    if a is None:
        a = _ARENAS[id(ctx)] = _Arena(ctx)
    return a


def _al(n: int, a: int = 256) -> int:
    return (n + a - 1) & ~(a - 1)


_ORDERS = {  # name -> (b2 order, step, import_alpha)   picture_csp_enc.c.go:189-225
    "RGBA": (_abi.ORDER_RGBA, 4, 1), "RGBX": (_abi.ORDER_RGBA, 4, 0),
    "BGRA": (_abi.ORDER_BGRA, 4, 1), "BGRX": (_abi.ORDER_BGRA, 4, 0),
    "RGB": (_abi.ORDER_RGB, 3, 0), "BGR": (_abi.ORDER_BGR, 3, 0),
}


def _import_argb_batch(pics, srcs, strides, b2order, step, import_alpha, ctx) -> int:
    """Import() into ARGB pictures: WebPPictureAlloc + WebPPackRGB / row copy /
    VP8LConvertBGRAToRGBA (picture_csp_enc.c.go:138-181) as one b2_import_to_argb launch."""
    lib, hnd = ctx.lib, ctx.h
    metas, need = [], 0
    for pic, stride in zip(pics, strides):
        w, h = pic.Width, pic.Height
        ds = _al(4 * w, 16)
        sz = dict(src=_al(abs(stride) * h), dst=_al(ds * h))
        metas.append((w, h, ds, sz))
        need += sum(sz.values())
    ar = _arena(ctx)
    ar.reset(need)
    jobs, dptr = [], []
    try:
        for (w, h, ds, sz), stride, src in zip(metas, strides, srcs):
            p = {k: ar.take(v) for k, v in sz.items()}
            dptr.append(p)
            flat = np.ascontiguousarray(src).reshape(-1)
            ctx._chk(lib.b2_h2d_async(hnd, 0, p["src"], flat.ctypes.data, abs(stride) * (h - 1) + step * w))
            src0 = p["src"] + (abs(stride) * (h - 1) if stride < 0 else 0)
            jobs.append(_abi.ConvertJob(src0, None, None, p["dst"], stride, 0, ds, w, h, 0))
        arr = (_abi.ConvertJob * len(jobs))(*jobs)
        ctx._chk(lib.b2_import_to_argb(hnd, 0, arr, len(jobs), b2order, int(import_alpha)))
        for pic, (w, h, ds, sz), p in zip(pics, metas, dptr):
            WebPPictureAllocARGB(pic)
            ctx._chk(lib.b2_d2h_2d_async(hnd, 0, pic.ARGB.ctypes.data, 4 * w, p["dst"], ds, 4 * w, h))
        ctx.sync()
    except B2Error as e:
        for pic in pics:
            pic.SetEncodingError(_STATUS_TO_ENC.get(e.status, ENC_ERROR_OUT_OF_MEMORY))
        return 0
    return 1


def ImportRGBABatch(pics, srcs, strides, order: str = "RGBA", ctx: Context | None = None,
                    divide_table: bool = False, chunk: int = 16, dithering: float = 0.0) -> int:
    """Added batched entry point (one per part): imports srcs[i]
    (interleaved bytes, row stride strides[i]) into pics[i] as YUV420(A).

    Per image this is enc.Import(...) -> picture.ImportYUVAFromRGBA(...) with
    dithering=0 and use_iterative_conversion=0 (picture_csp_enc.c.go:123-137,
    conversions.go:50-121).  The batch is cut into chunks that are pipelined
    over the context's streams: H2D, kernel, D2H of different chunks overlap.
    Returns 1 on success, 0 after recording the error on the offending picture.
    """
    ctx = ctx or default_context()
    b2order, step, import_alpha = _ORDERS[order]
    n = len(pics)
    if n == 0:
        return 1
    for pic, src, stride in zip(pics, srcs, strides):
        if pic is None or src is None:
            return 0
        err = WebPValidatePicture(pic)
        if err:
            return pic.SetEncodingError(err)
        if abs(stride) < (4 if import_alpha else 3) * pic.Width or abs(stride) < step * pic.Width:
            return 0  # picture_csp_enc.c.go:132
    argb_idx = [i for i, pic in enumerate(pics) if pic.UseARGB]
    if argb_idx:  # use_argb branch of Import() (picture_csp_enc.c.go:138-181)
        if not _import_argb_batch([pics[i] for i in argb_idx], [srcs[i] for i in argb_idx],
                                  [strides[i] for i in argb_idx], b2order, step, import_alpha, ctx):
            return 0
        rest = [i for i in range(n) if not pics[i].UseARGB]
        if not rest:
            return 1
        return ImportRGBABatch([pics[i] for i in rest], [srcs[i] for i in rest], [strides[i] for i in rest],
                               order, ctx, divide_table, chunk, dithering)
    # device layout per image: src rows (caller stride kept), then Y | U | V | A
    metas = []
    need = 0
    for pic, stride in zip(pics, strides):
        w, h = pic.Width, pic.Height
        uvw, uvh = (w + 1) >> 1, (h + 1) >> 1
        ys, uvs = _al(w, 16), _al(uvw, 16)
        sz = dict(src=_al(abs(stride) * h), y=_al(ys * h), u=_al(uvs * uvh), v=_al(uvs * uvh), a=_al(ys * h))
        metas.append((w, h, uvw, uvh, ys, uvs, sz))
        need += sum(sz.values())
    need += _al(4 * n)
    ar = _arena(ctx)
    ar.reset(need)
    flags_ptr = ar.take(4 * n)
    jobs, dptr = [], []
    for (w, h, uvw, uvh, ys, uvs, sz), stride in zip(metas, strides):
        p = {k: ar.take(v) for k, v in sz.items()}
        dptr.append(p)
        src0 = p["src"] + (abs(stride) * (h - 1) if stride < 0 else 0)
        jobs.append(_abi.ImportJob(src0, p["y"], p["u"], p["v"], p["a"], stride, ys, uvs, ys, w, h))
    flags = (_abi.IMPORT_ALPHA if import_alpha else 0) | (_abi.IMPORT_DIV_TABLE if divide_table else 0)
    lib, hnd = ctx.lib, ctx.h
    ns = ctx.n_streams
    flag_host = ctx.pinned(4 * n)
    # pass 1: allocate host planes (alpha kept only if the image turns out non-opaque)
    tmp_a = []
    for pic in pics:
        pic.ColorSpace = WEBP_YUV420A
        pic.UseARGB = False
        WebPPictureAllocYUVA(pic)
        tmp_a.append(pic.A)
    try:
        for c0 in range(0, n, chunk):
            c1 = min(n, c0 + chunk)
            s = (c0 // chunk) % ns
            for i in range(c0, c1):
                w, h, uvw, uvh, ys, uvs, sz = metas[i]
                src = np.ascontiguousarray(srcs[i]).reshape(-1)
                ctx._chk(lib.b2_h2d_async(hnd, s, dptr[i]["src"], src.ctypes.data, abs(strides[i]) * (h - 1) + step * w))
            arr = (_abi.ImportJob * (c1 - c0))(*jobs[c0:c1])
            if dithering > 0.0:  # conversions.go:122-165; one VP8Random per image, whatever the chunking
                ctx._chk(lib.b2_rgba_to_yuv420a_dithered(hnd, s, arr, c1 - c0, b2order, flags, flags_ptr + 4 * c0,
                                                         float(dithering)))
            else:
                ctx._chk(lib.b2_rgba_to_yuv420a(hnd, s, arr, c1 - c0, b2order, flags, flags_ptr + 4 * c0))
            for i in range(c0, c1):
                w, h, uvw, uvh, ys, uvs, sz = metas[i]
                pic = pics[i]
                ctx._chk(lib.b2_d2h_2d_async(hnd, s, pic.Y.ctypes.data, w, dptr[i]["y"], ys, w, h))
                ctx._chk(lib.b2_d2h_2d_async(hnd, s, pic.U.ctypes.data, uvw, dptr[i]["u"], uvs, uvw, uvh))
                ctx._chk(lib.b2_d2h_2d_async(hnd, s, pic.V.ctypes.data, uvw, dptr[i]["v"], uvs, uvw, uvh))
                ctx._chk(lib.b2_d2h_2d_async(hnd, s, tmp_a[i].ctypes.data, w, dptr[i]["a"], ys, w, h))
            ctx._chk(lib.b2_d2h_async(hnd, s, flag_host.ptr + 4 * c0, flags_ptr + 4 * c0, 4 * (c1 - c0)))
        ctx.sync()
    except B2Error as e:
        for pic in pics:
            pic.SetEncodingError(_STATUS_TO_ENC.get(e.status, ENC_ERROR_OUT_OF_MEMORY))
        return 0
    has_alpha = flag_host.array.view(np.int32).copy()
    flag_host.free()
    for i, pic in enumerate(pics):
        if has_alpha[i]:  # conversions.go:56-58
            pic.ColorSpace, pic.A, pic.AStride = WEBP_YUV420A, tmp_a[i], pic.Width
        else:
            pic.ColorSpace, pic.A, pic.AStride = WEBP_YUV420, None, 0
    return 1


def _import_one(pic, buf, stride, order):
    if pic is None or buf is None:
        return 0
    return ImportRGBABatch([pic], [buf], [stride], order)


# picture_csp_enc.c.go:189-225
def WebPPictureImportRGB(pic, rgb, rgb_stride): return _import_one(pic, rgb, rgb_stride, "RGB")
def WebPPictureImportRGBA(pic, rgba, rgba_stride): return _import_one(pic, rgba, rgba_stride, "RGBA")
def WebPPictureImportRGBX(pic, rgbx, rgbx_stride): return _import_one(pic, rgbx, rgbx_stride, "RGBX")
def WebPPictureImportBGR(pic, bgr, bgr_stride): return _import_one(pic, bgr, bgr_stride, "BGR")
def WebPPictureImportBGRA(pic, bgra, bgra_stride): return _import_one(pic, bgra, bgra_stride, "BGRA")
def WebPPictureImportBGRX(pic, bgrx, bgrx_stride): return _import_one(pic, bgrx, bgrx_stride, "BGRX")


def PictureARGBToYUVA(pic: Picture, csp: int, dithering: float = 0.0, use_iterative_conversion: int = 0,
                      divide_table: bool = False):
    """conversions.go:12-30: ARGB (uint32 0xAARRGGBB) -> YUV420(A).  Returns an
    error sentinel or None, like the Go function."""
    if pic is None:
        return None
    if pic.ARGB is None:
        pic.SetEncodingError(ENC_ERROR_nil_PARAMETER)
        return ENC_ERROR_nil_PARAMETER
    if (csp & WEBP_CSP_UV_MASK) != WEBP_YUV420:
        pic.SetEncodingError(ENC_ERROR_INVALID_CONFIGURATION)
        return ENC_ERROR_INVALID_CONFIGURATION
    if use_iterative_conversion:
        raise NotImplementedError("sharp-YUV import is not on the hot path (SURVEY 5; the sharpyuv package is absent)")
    argb = np.ascontiguousarray(pic.ARGB, np.uint32)
    stride = 4 * (pic.ARGBStride or pic.Width)
    pic.UseARGB = False  # ImportYUVAFromRGBA sets use_argb = 0 (conversions.go:59)
    ok = ImportRGBABatch([pic], [argb.view(np.uint8)], [stride], "BGRA", divide_table=divide_table,
                         dithering=max(float(dithering), 0.0))
    return None if ok else (pic.ErrorCode or ENC_ERROR_OUT_OF_MEMORY)


def WebPPictureARGBToYUVA(pic: Picture, csp: int = WEBP_YUV420):
    return PictureARGBToYUVA(pic, csp, 0.0, 0)


def WebPPictureARGBToYUVADithered(pic: Picture, csp: int, dithering: float):
    """conversions.go:36-38."""
    return PictureARGBToYUVA(pic, csp, dithering, 0)


def YUVAToRGBABatch(pics, mode: int = _abi.MODE_BGRA, dsts=None, ctx: Context | None = None,
                    chunk: int = 16):
    """Added batched entry point for part 2: fancy-upsamples pics[i].Y/U/V(/A)
    into dsts[i] (uint8 [h, w*bpp]; allocated when None) in WEBP_CSP_MODE 'mode'.
    Per image this is the schedule of WebPPictureYUVAToARGB
    (picture_csp_enc.c.go:62-118) with WebPUpsamplers[mode].  Returns the list of
    outputs, or None after recording the error."""
    ctx = ctx or default_context()
    n = len(pics)
    bpp = _abi.MODE_BPP[mode]
    if dsts is None:
        dsts = [None] * n
    out = []
    metas = []
    need = 0
    for i, pic in enumerate(pics):
        if pic is None:
            return None
        if pic.Y is None or pic.U is None or pic.V is None:
            pic.SetEncodingError(ENC_ERROR_nil_PARAMETER)
            return None
        if (pic.ColorSpace & WEBP_CSP_ALPHA_BIT) and pic.A is None:
            pic.SetEncodingError(ENC_ERROR_nil_PARAMETER)
            return None
        if (pic.ColorSpace & WEBP_CSP_UV_MASK) != WEBP_YUV420:
            pic.SetEncodingError(ENC_ERROR_INVALID_CONFIGURATION)
            return None
        w, h = pic.Width, pic.Height
        uvw, uvh = (w + 1) >> 1, (h + 1) >> 1
        ys, uvs, ds = _al(w, 16), _al(uvw, 16), _al(w * bpp, 16)
        has_a = bool(pic.ColorSpace & WEBP_CSP_ALPHA_BIT) and (bpp == 4 or mode in (_abi.MODE_RGBA_4444, _abi.MODE_rgbA_4444))
        sz = dict(y=_al(ys * h), u=_al(uvs * uvh), v=_al(uvs * uvh), a=_al(ys * h) if has_a else 0, d=_al(ds * h))
        metas.append((w, h, uvw, uvh, ys, uvs, ds, has_a, sz))
        need += sum(sz.values())
        out.append(dsts[i] if dsts[i] is not None else np.empty((h, w * bpp), np.uint8))
    ar = _arena(ctx)
    ar.reset(need)
    lib, hnd, ns = ctx.lib, ctx.h, ctx.n_streams
    jobs, dptr = [], []
    for (w, h, uvw, uvh, ys, uvs, ds, has_a, sz) in metas:
        p = {k: (ar.take(v) if v else None) for k, v in sz.items()}
        dptr.append(p)
        jobs.append(_abi.UpsampleJob(p["y"], p["u"], p["v"], p["a"], p["d"], ys, uvs, ys, ds, w, h))
    try:
        for c0 in range(0, n, chunk):
            c1 = min(n, c0 + chunk)
            s = (c0 // chunk) % ns
            for i in range(c0, c1):
                w, h, uvw, uvh, ys, uvs, ds, has_a, sz = metas[i]
                pic = pics[i]
                Y, U, V = (np.ascontiguousarray(x) for x in (pic.Y, pic.U, pic.V))
                ctx._chk(lib.b2_h2d_2d_async(hnd, s, dptr[i]["y"], ys, Y.ctypes.data, Y.shape[1], w, h))
                ctx._chk(lib.b2_h2d_2d_async(hnd, s, dptr[i]["u"], uvs, U.ctypes.data, U.shape[1], uvw, uvh))
                ctx._chk(lib.b2_h2d_2d_async(hnd, s, dptr[i]["v"], uvs, V.ctypes.data, V.shape[1], uvw, uvh))
                if has_a:
                    A = np.ascontiguousarray(pic.A)
                    ctx._chk(lib.b2_h2d_2d_async(hnd, s, dptr[i]["a"], ys, A.ctypes.data, A.shape[1], w, h))
            arr = (_abi.UpsampleJob * (c1 - c0))(*jobs[c0:c1])
            ctx._chk(lib.b2_yuv420a_to_rgba_fancy(hnd, s, arr, c1 - c0, mode))
            for i in range(c0, c1):
                w, h, uvw, uvh, ys, uvs, ds, has_a, sz = metas[i]
                ctx._chk(lib.b2_d2h_2d_async(hnd, s, out[i].ctypes.data, w * bpp, dptr[i]["d"], ds, w * bpp, h))
        ctx.sync()
    except B2Error as e:
        for pic in pics:
            pic.SetEncodingError(_STATUS_TO_ENC.get(e.status, ENC_ERROR_OUT_OF_MEMORY))
        return None
    return out


def EmitFancyRGB(pic: Picture, mode: int, mb_y: int, mb_h: int, dst: np.ndarray, ctx: Context | None = None) -> int:
    """EmitFancyRGB (pkg/libwebp/decoder/io_dec.c.go:65-115) for the macroblock-row batch
    [mb_y, mb_y + mb_h) of a decoded picture: writes the output rows that batch completes
    into dst (uint8 [h, w*bpp]) and returns their number.  First batch: row 0 is
    mirrored; every batch leaves its last row for the next call (it needs the chroma
    row below), the last batch also emits the final row.  The reference carries the
    unfinished row in tmp_y/u/v; here the planes are whole, so the carried row is simply
    read again, and EmitAlphaRGB (io_dec.c.go:175-195) is fused when pic has alpha."""
    ctx = ctx or default_context()
    w, h = pic.Width, pic.Height
    bpp = _abi.MODE_BPP[mode]
    y_end = min(mb_y + mb_h, h)
    row_begin = max(mb_y - 1, 0)
    row_end = h if y_end >= h else y_end - 1
    if row_end <= row_begin:
        return 0
    uvw, uvh = (w + 1) >> 1, (h + 1) >> 1
    ys, uvs, ds = _al(w, 16), _al(uvw, 16), _al(w * bpp, 16)
    has_a = bool(pic.ColorSpace & WEBP_CSP_ALPHA_BIT) and (bpp == 4 or mode in (_abi.MODE_RGBA_4444, _abi.MODE_rgbA_4444)) \
        and pic.A is not None
    sz = dict(y=_al(ys * h), u=_al(uvs * uvh), v=_al(uvs * uvh), a=_al(ys * h) if has_a else 0, d=_al(ds * h))
    ar = _arena(ctx)
    ar.reset(sum(sz.values()))
    p = {k: (ar.take(v) if v else None) for k, v in sz.items()}
    lib, hnd = ctx.lib, ctx.h
    for key, plane, stride, pw, ph in (("y", pic.Y, ys, w, h), ("u", pic.U, uvs, uvw, uvh), ("v", pic.V, uvs, uvw, uvh)) + \
            ((("a", pic.A, ys, w, h),) if has_a else ()):
        arr = np.ascontiguousarray(plane)
        ctx._chk(lib.b2_h2d_2d_async(hnd, 0, p[key], stride, arr.ctypes.data, arr.shape[1], pw, ph))
    job = _abi.UpsampleJob(p["y"], p["u"], p["v"], p["a"], p["d"], ys, uvs, ys, ds, w, h, row_begin, row_end)
    ctx._chk(lib.b2_yuv420a_to_rgba_fancy(hnd, 0, (_abi.UpsampleJob * 1)(job), 1, mode))
    rows = dst[row_begin:row_end]
    ctx._chk(lib.b2_d2h_2d_async(hnd, 0, rows.ctypes.data, dst.strides[0], p["d"] + row_begin * ds, ds,
                                 w * bpp, row_end - row_begin))
    ctx.sync()
    return row_end - row_begin


def WebPPictureYUVAToARGB(pic: Picture) -> int:
    """picture_csp_enc.c.go:62-118: fills pic.ARGB (uint32 0xAARRGGBB, i.e. the
    MODE_BGRA byte order on little-endian, :85) and sets UseARGB."""
    if pic is None:
        return 0
    if pic.Y is None or pic.U is None or pic.V is None:
        return pic.SetEncodingError(ENC_ERROR_nil_PARAMETER)
    if (pic.ColorSpace & WEBP_CSP_ALPHA_BIT) and pic.A is None:
        return pic.SetEncodingError(ENC_ERROR_nil_PARAMETER)
    if (pic.ColorSpace & WEBP_CSP_UV_MASK) != WEBP_YUV420:
        return pic.SetEncodingError(ENC_ERROR_INVALID_CONFIGURATION)
    if not WebPPictureAllocARGB(pic):
        return 0
    res = YUVAToRGBABatch([pic], _abi.MODE_BGRA, [pic.ARGB.view(np.uint8).reshape(pic.Height, 4 * pic.Width)])
    if res is None:
        return 0
    pic.UseARGB = True
    return 1
