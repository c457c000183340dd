"""Host-side mirror of go-webp's dsp function tables for the pixel hot path.

The reference dispatches through package-level function variables
(pkg/libwebp/dsp: WebPUpsamplers[] upsampling.c.go:31,275-289; VP8L function
types lossless.h.go:33-45, bound at lossless_enc.go:488-519; VP8LInverseTransform
lossless.go:511-547).  The functions below keep those names and argument
meanings, take HOST numpy arrays and run on the GPU through the C ABI.  The
decoder-side symbols the Go tree uses but never binds (VP8LAddGreenToBlueAndRed,
VP8LTransformColorInverse, VP8LMapColor32b) are supplied
here.  No CPU implementation lives in this module.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _abi
from .device import Context
from .picture import _al, _arena, default_context

# pkg/vp8/enums.go:33-40
PREDICTOR_TRANSFORM, CROSS_COLOR_TRANSFORM, SUBTRACT_GREEN_TRANSFORM, COLOR_INDEXING_TRANSFORM = 0, 1, 2, 3


def VP8LSubSampleSize(size: int, sampling_bits: int) -> int:  # lossless_common.go:34-36
    return (size + (1 << sampling_bits) - 1) >> sampling_bits


@dataclass
class VP8LTransform:  # pkg/vp8/models.go:126-132
    vtype: int
    bits: int
    xsize: int
    ysize: int
    data: np.ndarray | None = None


@dataclass
class VP8LImage:
    """One image of the batched entry point: 'pixels' is what DecodeImageData
    produced (uint32, row length = width after colour-index subsampling)."""
    pixels: np.ndarray
    width: int
    height: int


def _run(op: str, items, ctx: Context | None = None, chunk: int = 4):
    """items: list of (src u32/u8 array, out_shape, data array or None, width, height, bits).
    Returns the list of uint32 outputs.  One b2_vp8l_batch_staged call: chunks are pipelined
    over the streams (H2D, kernel and D2H of different chunks overlap)."""
    ctx = ctx or default_context()
    n = len(items)
    if n == 0:
        return []
    metas, need = [], 0
    for (src, oshape, data, w, h, bits) in items:
        src = np.ascontiguousarray(src)
        data = None if data is None else np.ascontiguousarray(data, np.uint32)
        ob = int(np.prod(oshape)) * 4
        sz = (_al(src.nbytes), _al(ob), _al(data.nbytes) if data is not None else 0)
        metas.append((src, data, ob, sz))
        need += sum(sz)
    ar = _arena(ctx)
    ar.reset(need)
    jobs = (_abi.ArgbJob * n)()
    hosts = (_abi.ArgbHost * n)()
    outs = []
    for i, ((src, data, ob, sz), (_, oshape, _, w, h, bits)) in enumerate(zip(metas, items)):
        j, hh = jobs[i], hosts[i]
        j.src, j.dst = ar.take(sz[0]), ar.take(sz[1])
        j.data = ar.take(sz[2]) if sz[2] else None
        j.width, j.height, j.bits = w, h, bits
        o = np.empty(oshape, np.uint32)
        outs.append(o)
        hh.src, hh.src_bytes = src.ctypes.data, src.nbytes
        hh.dst, hh.dst_bytes = o.ctypes.data, ob
        if data is not None:
            hh.data, hh.data_bytes = data.ctypes.data, data.nbytes
    ctx._chk(ctx.lib.b2_vp8l_batch_staged(ctx.h, _abi.VP8L_OPS[op], jobs, hosts, n, chunk))
    return outs


def _flat(op, src, data=None, width=None, height=None, bits=0, ctx=None):
    src = np.ascontiguousarray(src, np.uint32)
    if width is None:
        height, width = (src.shape if src.ndim == 2 else (1, src.size))
    return _run(op, [(src, src.shape, data, width, height, bits)], ctx)[0]


# ---- row-level functions with the reference's names -----------------------------
def VP8LAddGreenToBlueAndRed(src: np.ndarray, num_pixels: int | None = None) -> np.ndarray:
    """VP8LProcessDecBlueAndRedFunc (lossless.h.go:35); lossless.go:337-347."""
    src = np.ascontiguousarray(src, np.uint32).reshape(-1)
    n = src.size if num_pixels is None else num_pixels
    return _flat("add_green", src[:n], None, n, 1)


def VP8LSubtractGreenFromBlueAndRed(argb_data: np.ndarray, num_pixels: int | None = None) -> np.ndarray:
    """VP8LProcessEncBlueAndRedFunc (lossless.h.go:44); lossless_enc.go:207-216.
    Returns the transformed copy (the Go function works in place)."""
    src = np.ascontiguousarray(argb_data, np.uint32).reshape(-1)
    n = src.size if num_pixels is None else num_pixels
    return _flat("sub_green", src[:n], None, n, 1)


def MultipliersToColorCode(green_to_red: int, green_to_blue: int, red_to_blue: int) -> int:
    """pkg/libwebp/enc/predictor_enc.go:729-731"""
    return 0xff000000 | ((red_to_blue & 0xff) << 16) | ((green_to_blue & 0xff) << 8) | (green_to_red & 0xff)


def VP8LTransformColorInverse(m_code: int, src: np.ndarray, num_pixels: int | None = None) -> np.ndarray:
    """VP8LTransformColorInverseFunc (lossless.h.go:36); lossless.go:359-374.
    m_code: the multipliers as a colour code (ColorCodeToMultipliers, :353-357)."""
    src = np.ascontiguousarray(src, np.uint32).reshape(-1)
    n = src.size if num_pixels is None else num_pixels
    # one tile spanning the whole segment (bits = 9 covers up to 512 px per tile row)
    out = np.empty(n, np.uint32)
    for x0 in range(0, n, 512):
        seg = src[x0:min(n, x0 + 512)]
        out[x0:x0 + seg.size] = _flat("color_inverse", seg, np.array([m_code], np.uint32), seg.size, 1, 9)
    return out


def VP8LTransformColor(m_code: int, data: np.ndarray, num_pixels: int | None = None) -> np.ndarray:
    """VP8LTransformColorFunc (lossless.h.go:45); lossless_enc.go:224-239."""
    src = np.ascontiguousarray(data, np.uint32).reshape(-1)
    n = src.size if num_pixels is None else num_pixels
    out = np.empty(n, np.uint32)
    for x0 in range(0, n, 512):
        seg = src[x0:min(n, x0 + 512)]
        out[x0:x0 + seg.size] = _flat("color_forward", seg, np.array([m_code], np.uint32), seg.size, 1, 9)
    return out


def VP8LMapColor32b(src: np.ndarray, color_map: np.ndarray, y_start: int, y_end: int, width: int) -> np.ndarray:
    """VP8LMapARGBFunc (lossless.h.go:40); MapARGB_C lossless.go:415-426."""
    src = np.ascontiguousarray(src, np.uint32).reshape(y_end - y_start, width)
    pal = np.zeros(256, np.uint32)
    cm = np.ascontiguousarray(color_map, np.uint32).reshape(-1)
    pal[:cm.size] = cm
    return _run("color_index_inverse", [(src, src.shape, pal, width, y_end - y_start, 0)])[0]


def _simple(fn_name: str, job_type, uploads, make_job, out_nbytes: int, out_dtype, out_shape, ctx=None):
    """Upload 'uploads' (arrays), run one job, download one output."""
    ctx = ctx or default_context()
    ar = _arena(ctx)
    ar.reset(sum(_al(a.nbytes) for a in uploads) + _al(out_nbytes))
    ptrs = []
    for a in uploads:
        pa = ar.take(_al(a.nbytes))
        ctx._chk(ctx.lib.b2_h2d_async(ctx.h, 0, pa, a.ctypes.data, a.nbytes))
        ptrs.append(pa)
    po = ar.take(_al(out_nbytes))
    job = make_job(ptrs, po)
    ctx._chk(getattr(ctx.lib, fn_name)(ctx.h, 0, (job_type * 1)(job), 1))
    out = np.empty(out_shape, out_dtype)
    ctx._chk(ctx.lib.b2_d2h_async(ctx.h, 0, out.ctypes.data, po, out_nbytes))
    ctx.sync()
    return out


def VP8LColorIndexInverseTransformAlpha(transform: VP8LTransform, y_start: int, y_end: int,
                                        src: np.ndarray) -> np.ndarray:
    """lossless.go:461-505 (MapAlpha_C for bits == 0): 8-bit source rows of
    VP8LSubSampleSize(xsize, bits) bytes -> 8-bit rows of xsize bytes holding the
    green byte of the palette entry (VP8GetAlphaValue, lossless_common.go:31)."""
    rows, w, bits = y_end - y_start, transform.xsize, transform.bits
    sw = VP8LSubSampleSize(w, bits)
    srcb = np.ascontiguousarray(src, np.uint8).reshape(rows, sw)
    pal = np.zeros(1 << (8 >> bits), np.uint32)
    cm = np.ascontiguousarray(transform.data, np.uint32).reshape(-1)
    pal[:min(cm.size, pal.size)] = cm[:pal.size]
    return _simple("b2_vp8l_color_index_inverse_alpha", _abi.AlphaIndexJob, [srcb, pal],
                   lambda p, po: _abi.AlphaIndexJob(p[0], po, p[1], w, rows, bits, 0),
                   rows * w, np.uint8, (rows, w))


def VP8LMapColor8b(src: np.ndarray, color_map: np.ndarray, y_start: int, y_end: int, width: int) -> np.ndarray:
    """VP8LMapAlphaFunc (lossless.h.go:41); MapAlpha_C lossless.go:461-472."""
    t = VP8LTransform(COLOR_INDEXING_TRANSFORM, 0, width, y_end, color_map)
    return VP8LColorIndexInverseTransformAlpha(t, y_start, y_end, src)


def ExpandColorMap(num_colors: int, transform: VP8LTransform) -> np.ndarray:
    """pkg/vp8/vp8l_dec.c.go:1196-1219: un-delta the palette read from the bitstream
    and zero-pad it to 1 << (8 >> bits) entries; returns the new colour map (the
    reference stores it back into transform.data)."""
    bits = transform.bits
    final = 1 << (8 >> bits)
    data = np.ascontiguousarray(transform.data, np.uint32).reshape(-1)[:num_colors]
    return _simple("b2_vp8l_expand_color_map", _abi.ExpandJob, [data],
                   lambda p, po: _abi.ExpandJob(p[0], po, num_colors, bits), final * 4, np.uint32, (final,))


def ConvertToYUVA(argb: np.ndarray, with_alpha: bool = False, y_pos: int = 0, u=None, v=None):
    """ConvertToYUVA (pkg/vp8/vp8l_dec.c.go:543-562) over rows y_pos .. y_pos + h - 1:
    every row through WebPConvertARGBToY and WebPConvertARGBToUV (pkg/libwebp/dsp/
    yuv.go:74-120; rows at even positions store U/V, rows at odd positions average
    into them), alpha = byte 3 when asked.  argb: [h, w] uint32 0xAARRGGBB.  When
    y_pos is odd, u / v must hold the chroma rows stored by the previous call (first
    row = chroma row y_pos >> 1).  Returns (y, u, v, a-or-None)."""
    argb = np.ascontiguousarray(argb, np.uint32)
    h, w = argb.shape
    uw, uh = (w + 1) // 2, (h + (y_pos & 1) + 1) // 2
    ctx = default_context()
    ar = _arena(ctx)
    sizes = [_al(argb.nbytes), _al(w * h), _al(uw * uh), _al(uw * uh), _al(w * h)]
    ar.reset(sum(sizes))
    ps, py, pu, pv, pa = (ar.take(n) for n in sizes)
    ctx._chk(ctx.lib.b2_h2d_async(ctx.h, 0, ps, argb.ctypes.data, argb.nbytes))
    if y_pos & 1:
        for plane, ptr in ((u, pu), (v, pv)):
            prev = np.ascontiguousarray(plane, np.uint8).reshape(-1)[:uw]
            ctx._chk(ctx.lib.b2_h2d_async(ctx.h, 0, ptr, prev.ctypes.data, prev.nbytes))
    job = _abi.PlainJob(ps, py, pu, pv, pa if with_alpha else None, w, w, uw, w, w, h, y_pos, 0)
    ctx._chk(ctx.lib.b2_argb_to_yuva_plain(ctx.h, 0, (_abi.PlainJob * 1)(job), 1))
    outs = [np.empty((h, w), np.uint8), np.empty((uh, uw), np.uint8), np.empty((uh, uw), np.uint8)]
    for o, p_ in zip(outs, (py, pu, pv)):
        ctx._chk(ctx.lib.b2_d2h_async(ctx.h, 0, o.ctypes.data, p_, o.nbytes))
    a = None
    if with_alpha:
        a = np.empty((h, w), np.uint8)
        ctx._chk(ctx.lib.b2_d2h_async(ctx.h, 0, a.ctypes.data, pa, a.nbytes))
    ctx.sync()
    return outs[0], outs[1], outs[2], a


def WebPConvertARGBToY(argb: np.ndarray, width: int) -> np.ndarray:
    """pkg/libwebp/dsp/yuv.go:74-80 for one row."""
    row = np.ascontiguousarray(argb, np.uint32).reshape(-1)[:width].reshape(1, width)
    return ConvertToYUVA(row)[0][0]


def WebPConvertARGBToUV(argb: np.ndarray, u: np.ndarray, v: np.ndarray, src_width: int, do_store: bool):
    """pkg/libwebp/dsp/yuv.go:82-120 for one row: returns the new (u, v) rows; with
    do_store false the previous values are averaged in, (old + new + 1) >> 1."""
    row = np.ascontiguousarray(argb, np.uint32).reshape(-1)[:src_width].reshape(1, src_width)
    _, nu, nv, _ = ConvertToYUVA(row, y_pos=0 if do_store else 1, u=u, v=v)
    return nu[0], nv[0]


# ---- output conversions after parts 2 and 3 ---------------------------
def _convert(fn_name, uploads, make_job, out_ptr_index, out_shape, out_dtype, extra, ctx=None):
    """uploads: list of (array, role); the output is uploads[out_ptr_index] (in place) or,
    when out_ptr_index is None, a fresh buffer of out_shape/out_dtype."""
    ctx = ctx or default_context()
    out_nbytes = int(np.prod(out_shape)) * np.dtype(out_dtype).itemsize
    ar = _arena(ctx)
    ar.reset(sum(_al(a.nbytes) for a in uploads) + _al(out_nbytes))
    ptrs = []
    for a in uploads:
        pa = ar.take(_al(a.nbytes))
        ctx._chk(ctx.lib.b2_h2d_async(ctx.h, 0, pa, a.ctypes.data, a.nbytes))
        ptrs.append(pa)
    po = ptrs[out_ptr_index] if out_ptr_index is not None else ar.take(_al(out_nbytes))
    job = make_job(ptrs, po)
    ctx._chk(getattr(ctx.lib, fn_name)(ctx.h, 0, (_abi.ConvertJob * 1)(job), 1, *extra))
    out = np.empty(out_shape, out_dtype)
    ctx._chk(ctx.lib.b2_d2h_async(ctx.h, 0, out.ctypes.data, po, out_nbytes))
    ctx.sync()
    return out


def VP8LConvertFromBGRA(in_data: np.ndarray, num_pixels: int | None, out_colorspace: int,
                        upstream_16bit_order: bool = False) -> np.ndarray:
    """lossless.go:616-666.  in_data: uint32 pixels ([h, w] or flat); returns the bytes of
    'out_colorspace' ([h, w*bpp]).  Premultiplied modes include WebPApplyAlphaMultiply.
    16-bit modes: the context's order (Context.swap_16bit_csp, default the Go text's);
    upstream_16bit_order forces the stock libwebp order for this call."""
    src = np.ascontiguousarray(in_data, np.uint32)
    if src.ndim == 1:
        src = src[:num_pixels].reshape(1, -1)
    h, w = src.shape
    bpp = _abi.MODE_BPP[out_colorspace]
    return _convert("b2_vp8l_convert_from_bgra", [src],
                    lambda p, po: _abi.ConvertJob(p[0], None, None, po, w, 0, w * bpp, w, h, 0),
                    None, (h, w * bpp), np.uint8,
                    (out_colorspace, _abi.CONVERT_UPSTREAM_16BIT if upstream_16bit_order else 0))


def WebPApplyAlphaMultiply(rgba: np.ndarray, alpha_first: int, w: int, h: int, stride: int) -> np.ndarray:
    """ApplyAlphaMultiply_C (alpha_processing.go:110-126); returns the premultiplied copy."""
    buf = np.ascontiguousarray(rgba, np.uint8).reshape(h, stride)
    return _convert("b2_alpha_multiply", [buf], lambda p, po: _abi.ConvertJob(None, None, None, po, 0, 0, stride, w, h, 0),
                    0, (h, stride), np.uint8, (_abi.ALPHA_APPLY, int(alpha_first)))


def WebPApplyAlphaMultiply4444(rgba4444: np.ndarray, w: int, h: int, stride: int, rg_byte_pos: int | None = None) -> np.ndarray:
    """ApplyAlphaMultiply4444_C (alpha_processing.go:141-163).  rg_byte_pos defaults to the context's 16-bit
    order: 1 with the Go text's (the value it binds), 0 with the stock libwebp order."""
    if rg_byte_pos is None:
        rg_byte_pos = 1 if default_context().swap_16bit_csp else 0
    buf = np.ascontiguousarray(rgba4444, np.uint8).reshape(h, stride)
    return _convert("b2_alpha_multiply", [buf], lambda p, po: _abi.ConvertJob(None, None, None, po, 0, 0, stride, w, h, 0),
                    0, (h, stride), np.uint8, (_abi.ALPHA_APPLY_4444, int(rg_byte_pos)))


def WebPMultARGBRows(ptr: np.ndarray, width: int, num_rows: int, inverse: int) -> np.ndarray:
    """WebPMultARGBRow_C over rows (alpha_processing.go:33-52,73-79); ptr: [num_rows, width] uint32."""
    buf = np.ascontiguousarray(ptr, np.uint32).reshape(num_rows, width)
    return _convert("b2_alpha_multiply", [buf], lambda p, po: _abi.ConvertJob(None, None, None, po, 0, 0, 4 * width, width, num_rows, 0),
                    0, (num_rows, width), np.uint32, (_abi.ALPHA_MULT_ARGB, int(inverse)))


def WebPMultRows(ptr: np.ndarray, alpha: np.ndarray, width: int, num_rows: int, inverse: int) -> np.ndarray:
    """WebPMultRow_C over rows (alpha_processing.go:54-67,81-89)."""
    buf = np.ascontiguousarray(ptr, np.uint8).reshape(num_rows, width)
    al = np.ascontiguousarray(alpha, np.uint8).reshape(num_rows, width)
    return _convert("b2_alpha_multiply", [buf, al], lambda p, po: _abi.ConvertJob(None, p[1], None, po, 0, width, width, width, num_rows, 0),
                    0, (num_rows, width), np.uint8, (_abi.ALPHA_MULT_ROW, int(inverse)))


def WebPDispatchAlpha(alpha: np.ndarray, width: int, height: int, dst: np.ndarray, dst_stride: int, alpha_first: int = 0) -> np.ndarray:
    """DispatchAlpha_C (alpha_processing.go:165-180) as ExportAlpha / EmitAlphaRGB call it: the
    alpha plane written into the alpha byte of the 4-byte pixels of dst; returns the new dst."""
    buf = np.ascontiguousarray(dst, np.uint8).reshape(height, dst_stride)
    al = np.ascontiguousarray(alpha, np.uint8).reshape(height, width)
    return _convert("b2_alpha_multiply", [buf, al], lambda p, po: _abi.ConvertJob(None, p[1], None, po, 0, width, dst_stride, width, height, 0),
                    0, (height, dst_stride), np.uint8, (_abi.ALPHA_DISPATCH, int(alpha_first)))


def WebPSamplerProcessPlane(y: np.ndarray, u: np.ndarray, v: np.ndarray, mode: int) -> np.ndarray:
    """dsp/yuv.go:47-58 with WebPSamplers[mode] (== EmitSampledRGB, io_dec.c.go:53-59):
    point-sampled chroma, y [h, w], u/v [ceil(h/2), ceil(w/2)] -> bytes [h, w*bpp]."""
    y, u, v = (np.ascontiguousarray(x, np.uint8) for x in (y, u, v))
    h, w = y.shape
    bpp = _abi.MODE_BPP[mode]
    return _convert("b2_yuv_to_rgb_point", [y, u, v],
                    lambda p, po: _abi.ConvertJob(p[0], p[1], p[2], po, w, u.shape[1], w * bpp, w, h, 0),
                    None, (h, w * bpp), np.uint8, (mode, 0))


def WebPYUV444ToRGB(y: np.ndarray, u: np.ndarray, v: np.ndarray, mode: int) -> np.ndarray:
    """WebPYUV444Converters[mode] over rows (upsampling.c.go:191-245): full-resolution chroma."""
    y, u, v = (np.ascontiguousarray(x, np.uint8) for x in (y, u, v))
    h, w = y.shape
    bpp = _abi.MODE_BPP[mode]
    return _convert("b2_yuv_to_rgb_point", [y, u, v],
                    lambda p, po: _abi.ConvertJob(p[0], p[1], p[2], po, w, w, w * bpp, w, h, 0),
                    None, (h, w * bpp), np.uint8, (mode, 1))


def VP8LPredictorForward(argb: np.ndarray, modes: np.ndarray, bits: int) -> np.ndarray:
    """Residuals of a whole image for GIVEN predictor modes: PredictorSub0..13_C
    (lossless_enc.go:339-557) driven as predictor_enc.go:666-712 does.  argb: [h, w]
    uint32; modes: the mode image (mode in bits 8..11)."""
    src = np.ascontiguousarray(argb, np.uint32)
    h, w = src.shape
    return _run("predictor_forward", [(src, (h, w), modes, w, h, bits)])[0]


def VP8LBundleColorMap(row: np.ndarray, width: int, xbits: int) -> np.ndarray:
    """lossless_enc.go:293-312 for one row (or a [h, w] block of rows)."""
    idx = np.ascontiguousarray(row, np.uint8)
    if idx.ndim == 1:
        idx = idx.reshape(1, -1)
    h = idx.shape[0]
    ow = VP8LSubSampleSize(width, xbits)
    return _run("bundle_color_map", [(idx, (h, ow), None, width, h, xbits)])[0]


def PaletteXBits(palette_size: int) -> int:
    """MapImageFromPalette's packing rule (pkg/vp8/vp8l_encoder.go:1229-1233)."""
    if palette_size <= 4:
        return 3 if palette_size <= 2 else 2
    return 1 if palette_size <= 16 else 0


def ApplyPaletteBatch(items, ctx: Context | None = None):
    """Added batched form of ApplyPalette: items = [(src [h, w] uint32, palette, xbits), ...];
    returns the packed index images ([h, ceil(w / 2^xbits)] uint32).  One launch per xbits
    value present; lookup and VP8LBundleColorMap are one pass on the device."""
    ctx = ctx or default_context()
    n = len(items)
    if n == 0:
        return []
    metas, need = [], 0
    for (src, palette, xbits) in items:
        src = np.ascontiguousarray(src, np.uint32)
        pal = np.ascontiguousarray(palette, np.uint32).reshape(-1)
        h, w = src.shape
        ow = VP8LSubSampleSize(w, xbits)
        metas.append((src, pal, h, w, ow))
        need += _al(src.nbytes) + _al(4 * ow * h) + _al(pal.nbytes)
    ar = _arena(ctx)
    ar.reset(need)
    jobs = (_abi.PaletteJob * n)()
    outs, dptr = [], []
    for i, ((src, pal, h, w, ow), (_, _, xbits)) in enumerate(zip(metas, items)):
        j = jobs[i]
        j.src, j.dst, j.palette = ar.take(src.nbytes), ar.take(4 * ow * h), ar.take(pal.nbytes)
        j.src_stride, j.dst_stride, j.width, j.height, j.palette_size, j.xbits = w, ow, w, h, pal.size, xbits
        ctx._chk(ctx.lib.b2_h2d_async(ctx.h, 0, j.src, src.ctypes.data, src.nbytes))
        ctx._chk(ctx.lib.b2_h2d_async(ctx.h, 0, j.palette, pal.ctypes.data, pal.nbytes))
        outs.append(np.empty((h, ow), np.uint32))
        dptr.append(j.dst)
    ctx._chk(ctx.lib.b2_vp8l_apply_palette(ctx.h, 0, jobs, n))
    for o, p in zip(outs, dptr):
        ctx._chk(ctx.lib.b2_d2h_async(ctx.h, 0, o.ctypes.data, p, o.nbytes))
    ctx.sync()
    return outs


def ApplyPalette(src: np.ndarray, src_stride: int, palette: np.ndarray, palette_size: int, width: int, height: int,
                 xbits: int) -> np.ndarray:
    """ApplyPalette (pkg/vp8/vp8l_encoder.go:1149-1212): src = ARGB rows src_stride pixels apart;
    returns dst ([height, VP8LSubSampleSize(width, xbits)] uint32, the packed palette indices)."""
    rows = np.ascontiguousarray(src, np.uint32).reshape(-1)[:src_stride * (height - 1) + width]
    img = np.lib.stride_tricks.as_strided(rows, (height, width), (4 * src_stride, 4))
    pal = np.ascontiguousarray(palette, np.uint32).reshape(-1)[:palette_size]
    return ApplyPaletteBatch([(img, pal, xbits)])[0]


def VP8LInverseTransform(transform: VP8LTransform, row_start: int, row_end: int, in_rows: np.ndarray) -> np.ndarray:
    """lossless.go:511-547 for rows [row_start,row_end) of one transform; 'in_rows'
    holds exactly those rows.  Out-of-place (the in-place colour-index trick of
    :531-541 is a host-memory optimisation the device path does not need)."""
    assert row_start < row_end <= transform.ysize
    rows = row_end - row_start
    w = transform.xsize
    if transform.vtype == SUBTRACT_GREEN_TRANSFORM:
        src = np.ascontiguousarray(in_rows, np.uint32).reshape(rows, w)
        return _flat("add_green", src, None, w, rows)
    if transform.vtype == CROSS_COLOR_TRANSFORM:
        src = np.ascontiguousarray(in_rows, np.uint32).reshape(rows, w)
        tpr = VP8LSubSampleSize(w, transform.bits)
        # tile rows covering [row_start,row_end): hand the kernel a window whose
        # row 0 is aligned to a tile boundary
        t0 = row_start >> transform.bits
        pad = row_start - (t0 << transform.bits)
        data = np.ascontiguousarray(transform.data, np.uint32).reshape(-1, tpr)[t0:]
        if pad:
            src = np.concatenate([np.zeros((pad, w), np.uint32), src])
        out = _flat("color_inverse", src, data, w, rows + pad, transform.bits)
        return out[pad:]
    if transform.vtype == COLOR_INDEXING_TRANSFORM:
        sw = VP8LSubSampleSize(w, transform.bits)
        src = np.ascontiguousarray(in_rows, np.uint32).reshape(rows, sw)
        return _run("color_index_inverse", [(src, (rows, w), transform.data, w, rows, transform.bits)])[0]
    if transform.vtype == PREDICTOR_TRANSFORM:
        if row_start != 0:
            raise NotImplementedError("predictor inverse needs the rows above: pass the image from row 0")
        src = np.ascontiguousarray(in_rows, np.uint32).reshape(rows, w)
        return _run("predictor_inverse", [(src, (rows, w), transform.data, w, rows, transform.bits)])[0]
    raise ValueError("unknown transform type")


def _plan_chain(c):
    """ApplyInverseTransforms order (vp8l_dec.c.go:710-729): last read, first applied.  Steps that have a
    fused kernel are merged: [cross-colour] -> predictor -> [add-green] is ONE pass of the predictor
    wavefront (the order libwebp's encoder always writes), cross-colour -> add-green without a predictor is
    the flat fused kernel.  Returns [(op, transforms)], op 'predictor_chain' taking (cross or None,
    predictor, subtract-green or None)."""
    todo = [t for t in reversed(list(c))]
    steps, k = [], 0
    while k < len(todo):
        t = todo[k]
        nxt = todo[k + 1] if k + 1 < len(todo) else None
        nx2 = todo[k + 2] if k + 2 < len(todo) else None
        if t.vtype == CROSS_COLOR_TRANSFORM and nxt is not None and nxt.vtype == PREDICTOR_TRANSFORM:
            green = nx2 if (nx2 is not None and nx2.vtype == SUBTRACT_GREEN_TRANSFORM) else None
            steps.append(("predictor_chain", (t, nxt, green)))
            k += 3 if green is not None else 2
            continue
        if t.vtype == PREDICTOR_TRANSFORM and nxt is not None and nxt.vtype == SUBTRACT_GREEN_TRANSFORM:
            steps.append(("predictor_chain", (None, t, nxt)))
            k += 2
            continue
        if t.vtype == CROSS_COLOR_TRANSFORM and nxt is not None and nxt.vtype == SUBTRACT_GREEN_TRANSFORM:
            steps.append(("inverse_chain", (t,)))
            k += 2
            continue
        op = {CROSS_COLOR_TRANSFORM: "color_inverse", SUBTRACT_GREEN_TRANSFORM: "add_green",
              COLOR_INDEXING_TRANSFORM: "color_index_inverse", PREDICTOR_TRANSFORM: "predictor_inverse"}.get(t.vtype)
        if op is None:
            raise ValueError("unknown transform type")
        steps.append((op, (t,)))
        k += 1
    return steps


def VP8LInverseTransformBatch(imgs, chain, ctx: Context | None = None, chunk: int = 4, out_mode: int = _abi.MODE_BGRA):
    """Added batched entry point for part 3.  imgs: list[VP8LImage];
    chain: list (per image) of the transforms in READ order (ReadTransform,
    vp8l_dec.c.go:1221-1269); they are applied last-read-first exactly like
    ApplyInverseTransforms (vp8l_dec.c.go:710-729).  Fused kernels are used where the chain
    allows (see _plan_chain): libwebp's own chain -- cross-colour, predictor, subtract-green --
    is ONE kernel pass.  The pixels cross PCIe twice, whatever the length of the chain: decoded
    pixels up with the first step, final pixels down with the last; everything in between stays
    on the device (two ping-pong buffers per image).
    out_mode: MODE_BGRA returns uint32 0xAARRGGBB pixels; MODE_RGBA / MODE_ARGB the 4-byte orders of
    EmitRows -> VP8LConvertFromBGRA (vp8l_dec.c.go:771), fused into the last step when that step is the
    predictor chain (else one more device pass)."""
    ctx = ctx or default_context()
    n = len(imgs)
    cur = [np.ascontiguousarray(im.pixels, np.uint32) for im in imgs]
    plans = [_plan_chain(c) for c in chain]
    if n == 0:
        return cur
    if out_mode not in (_abi.MODE_BGRA, _abi.MODE_RGBA, _abi.MODE_ARGB):
        raise ValueError("out_mode: a 4-byte, not premultiplied mode (VP8LConvertFromBGRA does the others)")
    sizes, need = [], 0
    for im, src, plan in zip(imgs, cur, plans):
        full = max([4 * ts[-1].xsize * ts[-1].ysize if op != "predictor_chain" else 4 * ts[1].xsize * ts[1].ysize
                    for op, ts in plan] + [src.nbytes])
        dsz = [[_al(np.asarray(t.data).size * 4) if (t is not None and t.data is not None) else 0 for t in ts] for _, ts in plan]
        sizes.append((_al(full), dsz))
        need += 2 * _al(full) + sum(sum(d) for d in dsz)
    ar = _arena(ctx)
    ar.reset(need)
    bufs, dptr, hdata = [], [], []
    for (full, dsz), plan in zip(sizes, plans):
        bufs.append([ar.take(full), ar.take(full)])
        dptr.append([[ar.take(z) if z else None for z in d] for d in dsz])
        hdata.append([[np.ascontiguousarray(t.data, np.uint32) if (t is not None and t.data is not None) else None for t in ts]
                      for _, ts in plan])
    outs = list(cur)
    converted = [False] * n
    depth = max((len(p) for p in plans), default=0)
    for k in range(depth):
        groups: dict[str, list] = {}
        for i, plan in enumerate(plans):
            if k < len(plan):
                groups.setdefault(plan[k][0], []).append(i)
        for op, members in groups.items():
            m = len(members)
            if op == "predictor_chain":
                # the byte order rides on the last step only; images that stop later convert at the end
                for mode_now in sorted({(out_mode if k == len(plans[i]) - 1 else _abi.MODE_BGRA) for i in members}):
                    sel = [i for i in members if (out_mode if k == len(plans[i]) - 1 else _abi.MODE_BGRA) == mode_now]
                    jobs = (_abi.ChainJob * len(sel))()
                    hosts = (_abi.ChainHost * len(sel))()
                    for q, i in enumerate(sel):
                        cc, pr, gr = plans[i][k][1]
                        j, hh = jobs[q], hosts[q]
                        j.src, j.dst = bufs[i][k & 1], bufs[i][(k + 1) & 1]
                        j.width, j.height, j.predictor_bits = pr.xsize, pr.ysize, pr.bits
                        j.predictor_data = dptr[i][k][1]
                        hh.predictor_data, hh.predictor_bytes = hdata[i][k][1].ctypes.data, hdata[i][k][1].nbytes
                        if cc is not None:
                            j.color_data, j.color_bits = dptr[i][k][0], cc.bits
                            hh.color_data, hh.color_bytes = hdata[i][k][0].ctypes.data, hdata[i][k][0].nbytes
                        j.add_green = int(gr is not None)
                        if k == 0:
                            hh.src = cur[i].ctypes.data
                        if k == len(plans[i]) - 1:
                            outs[i] = np.empty((pr.ysize, pr.xsize), np.uint32)
                            hh.dst = outs[i].ctypes.data
                            converted[i] = True
                    ctx._chk(ctx.lib.b2_vp8l_chain_batch_staged(ctx.h, jobs, hosts, len(sel), mode_now, chunk))
                continue
            jobs = (_abi.ArgbJob * m)()
            hosts = (_abi.ArgbHost * m)()
            for q, i in enumerate(members):
                t = plans[i][k][1][0]
                j, hh = jobs[q], hosts[q]
                j.src, j.dst, j.data = bufs[i][k & 1], bufs[i][(k + 1) & 1], dptr[i][k][0]
                j.width, j.height, j.bits = t.xsize, t.ysize, t.bits
                if k == 0:
                    hh.src, hh.src_bytes = cur[i].ctypes.data, cur[i].nbytes
                if hdata[i][k][0] is not None:
                    hh.data, hh.data_bytes = hdata[i][k][0].ctypes.data, hdata[i][k][0].nbytes
                if k == len(plans[i]) - 1:
                    outs[i] = np.empty((t.ysize, t.xsize), np.uint32)
                    hh.dst, hh.dst_bytes = outs[i].ctypes.data, outs[i].nbytes
            ctx._chk(ctx.lib.b2_vp8l_batch_staged(ctx.h, _abi.VP8L_OPS[op], jobs, hosts, m, chunk))
    if out_mode != _abi.MODE_BGRA:
        for i in range(n):
            if not converted[i]:
                o = np.ascontiguousarray(outs[i], np.uint32)
                outs[i] = VP8LConvertFromBGRA(o, None, out_mode).view(np.uint32).reshape(o.shape)
    return outs


# ---- fancy upsampler line-pair table ----------------------------------------------
def _line_pair(mode):
    def fn(top_y, bottom_y, top_u, top_v, cur_u, cur_v, length):
        """WebPUpsampleLinePairFunc (dsp.h.go:144; upsampling.c.go:43-103): returns
        (top_dst, bottom_dst or None).  A pair is a 2-row image whose chroma rows
        are (top, cur); a single row (bottom_y None) is the mirrored-row case."""
        from .picture import Picture, YUVAToRGBABatch, WEBP_YUV420
        uw = (length + 1) >> 1
        bpp = _abi.MODE_BPP[mode]
        if bottom_y is None:
            # first/last-row call sites pass top == cur (picture_csp_enc.c.go:88,103)
            # -> a 1-row image with that chroma row.  Different top/cur with a nil
            # bottom row never occurs in the reference.
            assert np.array_equal(top_u[:uw], cur_u[:uw]) and np.array_equal(top_v[:uw], cur_v[:uw])
            pic = Picture(Width=length, Height=1, ColorSpace=WEBP_YUV420,
                          Y=np.asarray(top_y[:length], np.uint8).reshape(1, -1),
                          U=np.asarray(top_u[:uw], np.uint8).reshape(1, -1),
                          V=np.asarray(top_v[:uw], np.uint8).reshape(1, -1))
            out = YUVAToRGBABatch([pic], mode)[0]
            return out[0], None
        # rows (1,2) of a 3-row image: row 0 is a dummy, rows 1 and 2 are the pair
        # between chroma rows 0 (top) and 1 (cur)
        y = np.stack([np.zeros(length, np.uint8), np.asarray(top_y[:length], np.uint8),
                      np.asarray(bottom_y[:length], np.uint8)])
        u = np.stack([np.asarray(top_u[:uw], np.uint8), np.asarray(cur_u[:uw], np.uint8)])
        v = np.stack([np.asarray(top_v[:uw], np.uint8), np.asarray(cur_v[:uw], np.uint8)])
        pic = Picture(Width=length, Height=3, ColorSpace=WEBP_YUV420, Y=y, U=u, V=v)
        out = YUVAToRGBABatch([pic], mode)[0]
        return out[1][:length * bpp], out[2][:length * bpp]
    return fn


# WebPUpsamplers[MODE_LAST] (upsampling.c.go:275-289)
WebPUpsamplers = {m: _line_pair(m) for m in range(11)}
