"""Batched, pinned-host-staged entry points for uniform batches.

This is the shape north_star gives pkg/picture: device-resident Y/U/V/A (or
RGBA) planes with pinned-host staging.  A pipeline owns pinned host buffers for
a batch of n equally sized images plus the matching device buffers; run() moves
the batch host -> device, converts it and brings the results back, cut into
chunks that are spread over the context's streams so that H2D, the kernel and
D2H of different chunks overlap (PCIe is full duplex).  run_device() runs only
the kernel on the resident buffers.

Reference semantics per image: ImportPipeline == enc.Import ->
picture.ImportYUVAFromRGBA (pkg/libwebp/enc/picture_csp_enc.c.go:123-137,
pkg/picture/conversions.go:50-121); UpsamplePipeline == WebPPictureYUVAToARGB /
EmitFancyRGB (picture_csp_enc.c.go:62-118, pkg/libwebp/decoder/io_dec.c.go:65-115).
"""
from __future__ import annotations

import numpy as np

from . import _abi
from .device import Context


def _al(n: int, a: int = 256) -> int:
    return (n + a - 1) & ~(a - 1)


class ImportPipeline:
    def __init__(self, ctx: Context, n: int, width: int, height: int, order: str = "RGBA",
                 import_alpha: bool = True, divide_table: bool = False, chunk: int = 8,
                 host: bool = True):
        self.ctx, self.n, self.w, self.h = ctx, n, width, height
        self.order = {"RGBA": _abi.ORDER_RGBA, "BGRA": _abi.ORDER_BGRA}[order]
        self.flags = (_abi.IMPORT_ALPHA if import_alpha else 0) | (_abi.IMPORT_DIV_TABLE if divide_table else 0)
        self.chunk = max(1, min(chunk, n))
        self.uvw, self.uvh = (width + 1) >> 1, (height + 1) >> 1
        self.ys, self.uvs = _al(width, 16), _al(self.uvw, 16)
        self.src_img = _al(4 * width * height)
        self.y_img, self.uv_img = _al(self.ys * height), _al(self.uvs * self.uvh)
        # device: all src | all Y | all U | all V | all A | flags  (plane-major, so a
        # chunk of images is one contiguous range per plane -> one copy per plane)
        self.d_src = ctx.alloc(self.src_img * n)
        self.d_y = ctx.alloc(self.y_img * n)
        self.d_u = ctx.alloc(self.uv_img * n)
        self.d_v = ctx.alloc(self.uv_img * n)
        self.d_a = ctx.alloc(self.y_img * n)
        self.d_flags = ctx.alloc(_al(4 * n))
        self.jobs = [_abi.ImportJob(self.d_src.ptr + i * self.src_img, self.d_y.ptr + i * self.y_img,
                                    self.d_u.ptr + i * self.uv_img, self.d_v.ptr + i * self.uv_img,
                                    self.d_a.ptr + i * self.y_img, 4 * width, self.ys, self.uvs, self.ys,
                                    width, height) for i in range(n)]
        self._job_arr = (_abi.ImportJob * n)(*self.jobs)
        self.host = host
        if host:
            self.p_src = ctx.pinned(self.src_img * n)
            self.p_y = ctx.pinned(self.y_img * n)
            self.p_u = ctx.pinned(self.uv_img * n)
            self.p_v = ctx.pinned(self.uv_img * n)
            self.p_a = ctx.pinned(self.y_img * n)
            self.p_flags = ctx.pinned(_al(4 * n))

    # pinned host views ----------------------------------------------------------
    def src_view(self, i: int) -> np.ndarray:
        o = i * self.src_img
        return self.p_src.array[o:o + 4 * self.w * self.h].reshape(self.h, self.w, 4)

    def planes(self, i: int):
        """(Y, U, V, A or None, has_alpha) of image i after run(); A is dropped when the
        image is opaque, as conversions.go:56-58 / allocations.go:52-58 do."""
        y = self.p_y.array[i * self.y_img:i * self.y_img + self.ys * self.h].reshape(self.h, self.ys)[:, :self.w]
        u = self.p_u.array[i * self.uv_img:i * self.uv_img + self.uvs * self.uvh].reshape(self.uvh, self.uvs)[:, :self.uvw]
        v = self.p_v.array[i * self.uv_img:i * self.uv_img + self.uvs * self.uvh].reshape(self.uvh, self.uvs)[:, :self.uvw]
        ha = bool(self.p_flags.array.view(np.int32)[i])
        a = self.p_a.array[i * self.y_img:i * self.y_img + self.ys * self.h].reshape(self.h, self.ys)[:, :self.w] if ha else None
        return y, u, v, a, ha

    @property
    def h2d_bytes(self) -> int:
        return 4 * self.w * self.h * self.n

    @property
    def d2h_bytes(self) -> int:
        return (2 * self.ys * self.h + 2 * self.uvs * self.uvh + 4) * self.n

    # execution --------------------------------------------------------------------
    def run_device(self, stream: int = 0, dithering: float = 0.0):
        """One kernel launch over the resident batch (inputs already in HBM); with
        dithering > 0 the dithered import (skip-ahead state kernel + import kernel)."""
        c = self.ctx
        if dithering > 0.0:
            c._chk(c.lib.b2_rgba_to_yuv420a_dithered(c.h, stream, self._job_arr, self.n, self.order, self.flags,
                                                     self.d_flags.ptr, float(dithering)))
            return
        c._chk(c.lib.b2_rgba_to_yuv420a(c.h, stream, self._job_arr, self.n, self.order, self.flags,
                                        self.d_flags.ptr))

    def run(self):
        """Host -> device -> kernel -> host for the whole batch, chunk-pipelined."""
        c, lib, hnd = self.ctx, self.ctx.lib, self.ctx.h
        ns = c.n_streams
        for k, c0 in enumerate(range(0, self.n, self.chunk)):
            c1 = min(self.n, c0 + self.chunk)
            m = c1 - c0
            s = k % ns
            c._chk(lib.b2_h2d_async(hnd, s, self.d_src.ptr + c0 * self.src_img,
                                    self.p_src.ptr + c0 * self.src_img, m * self.src_img))
            arr = (_abi.ImportJob * m).from_address(
                _abi.C.addressof(self._job_arr) + c0 * _abi.C.sizeof(_abi.ImportJob))
            c._chk(lib.b2_rgba_to_yuv420a(hnd, s, arr, m, self.order, self.flags, self.d_flags.ptr + 4 * c0))
            c._chk(lib.b2_d2h_async(hnd, s, self.p_y.ptr + c0 * self.y_img, self.d_y.ptr + c0 * self.y_img, m * self.y_img))
            c._chk(lib.b2_d2h_async(hnd, s, self.p_u.ptr + c0 * self.uv_img, self.d_u.ptr + c0 * self.uv_img, m * self.uv_img))
            c._chk(lib.b2_d2h_async(hnd, s, self.p_v.ptr + c0 * self.uv_img, self.d_v.ptr + c0 * self.uv_img, m * self.uv_img))
            c._chk(lib.b2_d2h_async(hnd, s, self.p_a.ptr + c0 * self.y_img, self.d_a.ptr + c0 * self.y_img, m * self.y_img))
            c._chk(lib.b2_d2h_async(hnd, s, self.p_flags.ptr + 4 * c0, self.d_flags.ptr + 4 * c0, 4 * m))
        c.sync()

    def run_copies_only(self):
        """The H2D and D2H copies of run() without the kernel: what the box's PCIe / host path
        gives this traffic mix (bench.py reports it as the end-to-end ceiling)."""
        c, lib, hnd = self.ctx, self.ctx.lib, self.ctx.h
        ns = c.n_streams
        for k, c0 in enumerate(range(0, self.n, self.chunk)):
            m = min(self.n, c0 + self.chunk) - c0
            s = k % ns
            c._chk(lib.b2_h2d_async(hnd, s, self.d_src.ptr + c0 * self.src_img,
                                    self.p_src.ptr + c0 * self.src_img, m * self.src_img))
            for dp, pp, sz in ((self.d_y, self.p_y, self.y_img), (self.d_u, self.p_u, self.uv_img),
                               (self.d_v, self.p_v, self.uv_img), (self.d_a, self.p_a, self.y_img)):
                c._chk(lib.b2_d2h_async(hnd, s, pp.ptr + c0 * sz, dp.ptr + c0 * sz, m * sz))
        c.sync()

    def fill_synthetic(self, seed: int, first_img: int = 0, alpha_mode: int = _abi.SYNTH_OPAQUE,
                       to_host: bool = False):
        """Device-side generator (identical to the oracle's orc_synth_fill_u32)."""
        c = self.ctx
        c.synth_fill(self.d_src.ptr, self.w, self.h, 4 * self.w, self.src_img, self.n, seed, first_img, 0, alpha_mode)
        c.sync()
        if to_host and self.host:
            c._chk(c.lib.b2_d2h_async(c.h, 0, self.p_src.ptr, self.d_src.ptr, self.src_img * self.n))
            c.sync()

    def device_planes(self, i: int):
        """Copies image i's planes back from the device (after run_device())."""
        c = self.ctx
        y = c.d2h(self.d_y.ptr + i * self.y_img, self.ys * self.h).reshape(self.h, self.ys)[:, :self.w]
        u = c.d2h(self.d_u.ptr + i * self.uv_img, self.uvs * self.uvh).reshape(self.uvh, self.uvs)[:, :self.uvw]
        v = c.d2h(self.d_v.ptr + i * self.uv_img, self.uvs * self.uvh).reshape(self.uvh, self.uvs)[:, :self.uvw]
        a = c.d2h(self.d_a.ptr + i * self.y_img, self.ys * self.h).reshape(self.h, self.ys)[:, :self.w]
        ha = bool(c.d2h(self.d_flags.ptr + 4 * i, 4).view(np.int32)[0])
        return y, u, v, a, ha

    def close(self):
        for b in ("d_src", "d_y", "d_u", "d_v", "d_a", "d_flags", "p_src", "p_y", "p_u", "p_v", "p_a", "p_flags"):
            if hasattr(self, b):
                getattr(self, b).free()


class UpsamplePipeline:
    def __init__(self, ctx: Context, n: int, width: int, height: int, mode: int = _abi.MODE_RGBA,
                 with_alpha: bool = False, chunk: int = 8, host: bool = True):
        self.ctx, self.n, self.w, self.h, self.mode = ctx, n, width, height, mode
        self.bpp = _abi.MODE_BPP[mode]
        self.chunk = max(1, min(chunk, n))
        self.uvw, self.uvh = (width + 1) >> 1, (height + 1) >> 1
        self.ys, self.uvs, self.ds = _al(width, 16), _al(self.uvw, 16), _al(width * self.bpp, 16)
        self.y_img, self.uv_img, self.d_img = _al(self.ys * height), _al(self.uvs * self.uvh), _al(self.ds * height)
        self.with_alpha = with_alpha
        self.d_y = ctx.alloc(self.y_img * n)
        self.d_u = ctx.alloc(self.uv_img * n)
        self.d_v = ctx.alloc(self.uv_img * n)
        self.d_a = ctx.alloc(self.y_img * n) if with_alpha else None
        self.d_dst = ctx.alloc(self.d_img * n)
        self.jobs = [_abi.UpsampleJob(self.d_y.ptr + i * self.y_img, self.d_u.ptr + i * self.uv_img,
                                      self.d_v.ptr + i * self.uv_img,
                                      self.d_a.ptr + i * self.y_img if with_alpha else None,
                                      self.d_dst.ptr + i * self.d_img, self.ys, self.uvs, self.ys, self.ds,
                                      width, height) for i in range(n)]
        self._job_arr = (_abi.UpsampleJob * n)(*self.jobs)
        self.host = host
        if host:
            self.p_y = ctx.pinned(self.y_img * n)
            self.p_u = ctx.pinned(self.uv_img * n)
            self.p_v = ctx.pinned(self.uv_img * n)
            self.p_a = ctx.pinned(self.y_img * n) if with_alpha else None
            self.p_dst = ctx.pinned(self.d_img * n)

    @property
    def h2d_bytes(self) -> int:
        return (self.ys * self.h * (2 if self.with_alpha else 1) + 2 * self.uvs * self.uvh) * self.n

    @property
    def d2h_bytes(self) -> int:
        return self.ds * self.h * self.n

    def run_device(self, stream: int = 0):
        c = self.ctx
        c._chk(c.lib.b2_yuv420a_to_rgba_fancy(c.h, stream, self._job_arr, self.n, self.mode))

    def run(self):
        c, lib, hnd = self.ctx, self.ctx.lib, self.ctx.h
        ns = c.n_streams
        for k, c0 in enumerate(range(0, self.n, self.chunk)):
            c1 = min(self.n, c0 + self.chunk)
            m = c1 - c0
            s = k % ns
            c._chk(lib.b2_h2d_async(hnd, s, self.d_y.ptr + c0 * self.y_img, self.p_y.ptr + c0 * self.y_img, m * self.y_img))
            c._chk(lib.b2_h2d_async(hnd, s, self.d_u.ptr + c0 * self.uv_img, self.p_u.ptr + c0 * self.uv_img, m * self.uv_img))
            c._chk(lib.b2_h2d_async(hnd, s, self.d_v.ptr + c0 * self.uv_img, self.p_v.ptr + c0 * self.uv_img, m * self.uv_img))
            if self.with_alpha:
                c._chk(lib.b2_h2d_async(hnd, s, self.d_a.ptr + c0 * self.y_img, self.p_a.ptr + c0 * self.y_img, m * self.y_img))
            arr = (_abi.UpsampleJob * m).from_address(
                _abi.C.addressof(self._job_arr) + c0 * _abi.C.sizeof(_abi.UpsampleJob))
            c._chk(lib.b2_yuv420a_to_rgba_fancy(hnd, s, arr, m, self.mode))
            c._chk(lib.b2_d2h_async(hnd, s, self.p_dst.ptr + c0 * self.d_img, self.d_dst.ptr + c0 * self.d_img, m * self.d_img))
        c.sync()

    def fill_synthetic(self, seed: int, first_img: int = 0):
        c = self.ctx
        c.synth_fill(self.d_y.ptr, self.ys // 4, self.h, self.ys, self.y_img, self.n, seed, first_img, 1)
        c.synth_fill(self.d_u.ptr, self.uvs // 4, self.uvh, self.uvs, self.uv_img, self.n, seed, first_img, 2)
        c.synth_fill(self.d_v.ptr, self.uvs // 4, self.uvh, self.uvs, self.uv_img, self.n, seed, first_img, 3)
        if self.with_alpha:
            c.synth_fill(self.d_a.ptr, self.ys // 4, self.h, self.ys, self.y_img, self.n, seed, first_img, 4)
        c.sync()

    def device_image(self, i: int):
        c = self.ctx
        y = c.d2h(self.d_y.ptr + i * self.y_img, self.ys * self.h).reshape(self.h, self.ys)[:, :self.w]
        u = c.d2h(self.d_u.ptr + i * self.uv_img, self.uvs * self.uvh).reshape(self.uvh, self.uvs)[:, :self.uvw]
        v = c.d2h(self.d_v.ptr + i * self.uv_img, self.uvs * self.uvh).reshape(self.uvh, self.uvs)[:, :self.uvw]
        a = None
        if self.with_alpha:
            a = c.d2h(self.d_a.ptr + i * self.y_img, self.ys * self.h).reshape(self.h, self.ys)[:, :self.w]
        d = c.d2h(self.d_dst.ptr + i * self.d_img, self.ds * self.h).reshape(self.h, self.ds)[:, :self.w * self.bpp]
        return y, u, v, a, d

    def close(self):
        for b in ("d_y", "d_u", "d_v", "d_a", "d_dst", "p_y", "p_u", "p_v", "p_a", "p_dst"):
            x = getattr(self, b, None)
            if x is not None:
                x.free()


class ArgbPipeline:
    """Device-resident batch of n equal ARGB images for the VP8L transforms."""

    def __init__(self, ctx: Context, n: int, width: int, height: int, op: str, bits: int,
                 host: bool = False, host_images: int = 4):
        self.ctx, self.n, self.w, self.h, self.op, self.bits = ctx, n, width, height, op, bits
        self.host, self.host_images = host, min(host_images, n)
        self.index = op == "color_index_inverse"
        sub = lambda v, b: (v + (1 << b) - 1) >> b  # noqa: E731
        self.src_w = sub(width, bits) if self.index else width
        self.src_img = _al(4 * self.src_w * height)
        self.dst_img = _al(4 * width * height)
        if self.index:
            self.data_words = 1 << (8 >> bits)
        elif op in ("add_green", "sub_green"):
            self.data_words = 0
        else:
            self.data_words = sub(width, bits) * sub(height, bits)
        self.data_img = _al(4 * max(self.data_words, 1))
        self.d_src = ctx.alloc(self.src_img * n)
        self.d_dst = ctx.alloc(self.dst_img * n)
        self.d_data = ctx.alloc(self.data_img * n)
        self.jobs = [_abi.ArgbJob(self.d_src.ptr + i * self.src_img, self.d_dst.ptr + i * self.dst_img,
                                  self.d_data.ptr + i * self.data_img if self.data_words else None,
                                  width, height, bits, 0) for i in range(n)]
        self._job_arr = (_abi.ArgbJob * n)(*self.jobs)
        self._fn = getattr(ctx.lib, "b2_vp8l_" + op)
        if self.host:   # pinned staging for the first host_images images (end-to-end runs)
            m = self.host_images
            self.p_src = ctx.pinned(self.src_img * m)
            self.p_dst = ctx.pinned(self.dst_img * m)
            self.p_data = ctx.pinned(self.data_img * m)
            self._host_arr = (_abi.ArgbHost * m)(*[
                _abi.ArgbHost(self.p_src.ptr + i * self.src_img, self.p_dst.ptr + i * self.dst_img,
                              self.p_data.ptr + i * self.data_img if self.data_words else None,
                              4 * self.src_w * height, 4 * width * height, 4 * self.data_words) for i in range(m)])

    @property
    def h2d_bytes(self) -> int:
        return self.host_images * (4 * self.src_w * self.h + 4 * self.data_words)

    @property
    def d2h_bytes(self) -> int:
        return self.host_images * 4 * self.w * self.h

    def run(self, chunk: int = 1):
        """pinned host -> device -> kernel -> pinned host for the staged images (b2_vp8l_batch_staged)."""
        c = self.ctx
        c._chk(c.lib.b2_vp8l_batch_staged(c.h, _abi.VP8L_OPS[self.op], self._job_arr, self._host_arr, self.host_images, chunk))

    def fill_synthetic(self, seed: int, first_img: int = 0):
        c = self.ctx
        c.synth_fill(self.d_src.ptr, self.src_w, self.h, 4 * self.src_w, self.src_img, self.n, seed, first_img, 5)
        if self.data_words:
            c.synth_fill(self.d_data.ptr, self.data_words, 1, 4 * self.data_words, self.data_img, self.n, seed, first_img, 6)
        c.sync()
        if self.host:
            m = self.host_images
            c._chk(c.lib.b2_d2h_async(c.h, 0, self.p_src.ptr, self.d_src.ptr, self.src_img * m))
            if self.data_words:
                c._chk(c.lib.b2_d2h_async(c.h, 0, self.p_data.ptr, self.d_data.ptr, self.data_img * m))
            c.sync()

    def run_device(self, stream: int = 0):
        c = self.ctx
        c._chk(self._fn(c.h, stream, self._job_arr, self.n))

    def device_image(self, i: int):
        c = self.ctx
        src = c.d2h(self.d_src.ptr + i * self.src_img, 4 * self.src_w * self.h).view(np.uint32).reshape(self.h, self.src_w)
        data = c.d2h(self.d_data.ptr + i * self.data_img, 4 * self.data_words).view(np.uint32) if self.data_words else None
        dst = c.d2h(self.d_dst.ptr + i * self.dst_img, 4 * self.w * self.h).view(np.uint32).reshape(self.h, self.w)
        return src, data, dst

    def close(self):
        for b in ("d_src", "d_dst", "d_data", "p_src", "p_dst", "p_data"):
            if hasattr(self, b):
                getattr(self, b).free()


class ChainPipeline:
    """Device-resident batch for the whole VP8L inverse chain of libwebp's encoder (cross-colour inverse ->
    predictor inverse -> add-green, 32x32 tiles): run_device() = the fused single pass
    (b2_vp8l_predictor_inverse_chain), run_unfused() = the same result as three launches."""

    def __init__(self, ctx: Context, n: int, width: int, height: int, bits: int = 5):
        self.ctx, self.n, self.w, self.h, self.bits = ctx, n, width, height, bits
        sub = lambda v, b: (v + (1 << b) - 1) >> b  # noqa: E731
        self.words = sub(width, bits) * sub(height, bits)
        self.img, self.dat = _al(4 * width * height), _al(4 * self.words)
        self.d_src, self.d_dst, self.d_tmp = ctx.alloc(self.img * n), ctx.alloc(self.img * n), ctx.alloc(self.img * n)
        self.d_modes, self.d_codes = ctx.alloc(self.dat * n), ctx.alloc(self.dat * n)
        A = _abi
        self._chain = (A.ChainJob * n)(*[A.ChainJob(self.d_src.ptr + i * self.img, self.d_dst.ptr + i * self.img,
                                                    self.d_modes.ptr + i * self.dat, self.d_codes.ptr + i * self.dat,
                                                    width, height, bits, bits, 1, 0) for i in range(n)])
        mk = lambda s, d, t: (A.ArgbJob * n)(*[A.ArgbJob(s.ptr + i * self.img, d.ptr + i * self.img,  # noqa: E731
                                                        t.ptr + i * self.dat if t else None, width, height, bits, 0) for i in range(n)])
        self._cc = mk(self.d_src, self.d_tmp, self.d_codes)
        self._pred = mk(self.d_tmp, self.d_dst, self.d_modes)
        self._green = mk(self.d_dst, self.d_dst, None)

    def fill_synthetic(self, seed: int, first_img: int = 0):
        c = self.ctx
        c.synth_fill(self.d_src.ptr, self.w, self.h, 4 * self.w, self.img, self.n, seed, first_img, 5)
        c.synth_fill(self.d_modes.ptr, self.words, 1, 4 * self.words, self.dat, self.n, seed, first_img, 6)
        c.synth_fill(self.d_codes.ptr, self.words, 1, 4 * self.words, self.dat, self.n, seed, first_img, 7)
        c.sync()

    def run_device(self, stream: int = 0, out_mode: int = _abi.MODE_BGRA):
        c = self.ctx
        c._chk(c.lib.b2_vp8l_predictor_inverse_chain(c.h, stream, self._chain, self.n, out_mode))

    def run_unfused(self, stream: int = 0):
        c = self.ctx
        c._chk(c.lib.b2_vp8l_color_inverse(c.h, stream, self._cc, self.n))
        c._chk(c.lib.b2_vp8l_predictor_inverse(c.h, stream, self._pred, self.n))
        c._chk(c.lib.b2_vp8l_add_green(c.h, stream, self._green, self.n))

    def device_image(self, i: int):
        c = self.ctx
        f = lambda b, nb: c.d2h(b.ptr + i * (self.img if nb == 4 * self.w * self.h else self.dat), nb).view(np.uint32)  # noqa: E731
        return (f(self.d_src, 4 * self.w * self.h).reshape(self.h, self.w), f(self.d_modes, 4 * self.words),
                f(self.d_codes, 4 * self.words), f(self.d_dst, 4 * self.w * self.h).reshape(self.h, self.w))

    def close(self):
        for b in ("d_src", "d_dst", "d_tmp", "d_modes", "d_codes"):
            getattr(self, b).free()


class PalettePipeline:
    """Device-resident batch of n equal ARGB images whose pixels are all colours of a per-image palette,
    for ApplyPalette (colour -> packed index).  The images are made on the device: synthetic packed
    indices through the colour-index inverse, so run_device() must give those indices back."""

    def __init__(self, ctx: Context, n: int, width: int, height: int, palette_size: int):
        self.ctx, self.n, self.w, self.h, self.psize = ctx, n, width, height, palette_size
        self.xbits = (3 if palette_size <= 2 else 2) if palette_size <= 4 else (1 if palette_size <= 16 else 0)
        self.ow = (width + (1 << self.xbits) - 1) >> self.xbits
        self.pal_words = 1 << (8 >> self.xbits)
        self.img, self.pk, self.pal_img = _al(4 * width * height), _al(4 * self.ow * height), _al(4 * self.pal_words)
        self.d_img = ctx.alloc(self.img * n)
        self.d_idx = ctx.alloc(self.pk * n)     # synthetic packed indices (the expected answer's green bytes)
        self.d_out = ctx.alloc(self.pk * n)
        self.d_pal = ctx.alloc(self.pal_img * n)
        self._jobs = (_abi.PaletteJob * n)(*[
            _abi.PaletteJob(self.d_img.ptr + i * self.img, self.d_out.ptr + i * self.pk, self.d_pal.ptr + i * self.pal_img,
                            width, self.ow, width, height, palette_size, self.xbits) for i in range(n)])
        self._inv = (_abi.ArgbJob * n)(*[
            _abi.ArgbJob(self.d_idx.ptr + i * self.pk, self.d_img.ptr + i * self.img, self.d_pal.ptr + i * self.pal_img,
                         width, height, self.xbits, 0) for i in range(n)])

    def fill_synthetic(self, seed: int, palettes: np.ndarray | None = None):
        """palettes: [n, palette_size] uint32 of distinct colours (host); indices uniform below palette_size
        when it is a power of two (the synthetic words are masked on the host for the other sizes)."""
        c = self.ctx
        c.synth_fill(self.d_idx.ptr, self.ow, self.h, 4 * self.ow, self.pk, self.n, seed, 0, 5)
        if palettes is None:
            rng = np.random.default_rng(seed)
            palettes = np.stack([rng.choice(1 << 24, self.psize, replace=False).astype(np.uint32) | 0xff000000 for _ in range(self.n)])
        full = np.zeros((self.n, self.pal_img // 4), np.uint32)
        full[:, :self.psize] = palettes
        # indices >= palette_size would hit the zero padding: fold them onto entry index % psize by repeating the palette
        for k in range(self.psize, self.pal_words):
            full[:, k] = full[:, k % self.psize]
        c.h2d(self.d_pal.ptr, full)
        c._chk(c.lib.b2_vp8l_color_index_inverse(c.h, 0, self._inv, self.n))
        c.sync()
        self.palettes = palettes

    def run_device(self, stream: int = 0):
        c = self.ctx
        c._chk(c.lib.b2_vp8l_apply_palette(c.h, stream, self._jobs, self.n))

    def device_image(self, i: int):
        c = self.ctx
        img = c.d2h(self.d_img.ptr + i * self.img, 4 * self.w * self.h).view(np.uint32).reshape(self.h, self.w)
        out = c.d2h(self.d_out.ptr + i * self.pk, 4 * self.ow * self.h).view(np.uint32).reshape(self.h, self.ow)
        return img, out

    def close(self):
        for b in ("d_img", "d_idx", "d_out", "d_pal"):
            getattr(self, b).free()


def mixed_sizes(n: int = 256, seed: int = 0xB2000005, lo: int = 64, hi: int = 8192):
    """The configs[4] batch: w and h independently log-uniform in
    [lo, hi] from the counter-based generator, parities forced so the four (odd/even w
    x odd/even h) classes are equally represented, corner cases always included."""
    def lowbias32(x):
        x &= 0xFFFFFFFF
        x ^= x >> 16; x = (x * 0x7FEB352D) & 0xFFFFFFFF
        x ^= x >> 15; x = (x * 0x846CA68B) & 0xFFFFFFFF
        x ^= x >> 16
        return x
    fixed = [(64, 64), (8192, 8192), (65, 8191), (8191, 65)]
    out = list(fixed[:max(0, min(len(fixed), n))])
    ratio = hi / lo
    for i in range(len(out), n):
        uw = lowbias32(seed ^ (2 * i * 0x9E3779B1)) / 2.0**32
        uh = lowbias32(seed ^ ((2 * i + 1) * 0x9E3779B1)) / 2.0**32
        w = int(lo * ratio**uw)
        h = int(lo * ratio**uh)
        w = min(hi, max(lo, (w & ~1) | (i & 1)))
        h = min(hi, max(lo, (h & ~1) | ((i >> 1) & 1)))
        out.append((w, h))
    return out


class MixedPipeline:
    """Device-resident batch of images of DIFFERENT sizes through the whole pixel path
    (configs[4]): RGBA -> YUV420A import, fancy upsample back to BGRA, and the VP8L
    inverse chain (cross-colour inverse + add-green, bits = 5) on the result -- three
    launches per pass, each over the whole batch as one flat work list."""

    def __init__(self, ctx: Context, sizes, mode: int = _abi.MODE_BGRA, bits: int = 5):
        self.ctx, self.sizes, self.mode, self.bits = ctx, list(sizes), mode, bits
        n = self.n = len(self.sizes)
        sub = lambda v, b: (v + (1 << b) - 1) >> b  # noqa: E731
        off = {k: [0] for k in ("src", "y", "uv", "dst", "data")}
        self.geo = []
        for (w, h) in self.sizes:
            uvw, uvh = (w + 1) >> 1, (h + 1) >> 1
            ys, uvs, ds = _al(w, 16), _al(uvw, 16), _al(4 * w, 16)
            self.geo.append((w, h, uvw, uvh, ys, uvs, ds))
            off["src"].append(off["src"][-1] + _al(4 * w * h))
            off["y"].append(off["y"][-1] + _al(ys * h))
            off["uv"].append(off["uv"][-1] + _al(uvs * uvh))
            off["dst"].append(off["dst"][-1] + _al(ds * h))
            off["data"].append(off["data"][-1] + _al(4 * sub(w, bits) * sub(h, bits)))
        self.off = off
        self.d_src = ctx.alloc(off["src"][-1])
        self.d_y = ctx.alloc(off["y"][-1])
        self.d_a = ctx.alloc(off["y"][-1])
        self.d_u = ctx.alloc(off["uv"][-1])
        self.d_v = ctx.alloc(off["uv"][-1])
        self.d_dst = ctx.alloc(off["dst"][-1])
        self.d_out = ctx.alloc(off["src"][-1])
        self.d_data = ctx.alloc(off["data"][-1])
        self.d_flags = ctx.alloc(_al(4 * n))
        imp, ups, arg = [], [], []
        for i, (w, h, uvw, uvh, ys, uvs, ds) in enumerate(self.geo):
            y, a = self.d_y.ptr + off["y"][i], self.d_a.ptr + off["y"][i]
            u, v = self.d_u.ptr + off["uv"][i], self.d_v.ptr + off["uv"][i]
            imp.append(_abi.ImportJob(self.d_src.ptr + off["src"][i], y, u, v, a, 4 * w, ys, uvs, ys, w, h))
            ups.append(_abi.UpsampleJob(y, u, v, a, self.d_dst.ptr + off["dst"][i], ys, uvs, ys, ds, w, h, 0, h))
            # the ARGB chain runs on the tightly packed source image (same pixel count)
            arg.append(_abi.ArgbJob(self.d_src.ptr + off["src"][i], self.d_out.ptr + off["src"][i],
                                    self.d_data.ptr + off["data"][i], w, h, bits, 0))
        self._imp = (_abi.ImportJob * n)(*imp)
        self._ups = (_abi.UpsampleJob * n)(*ups)
        self._arg = (_abi.ArgbJob * n)(*arg)
        self.pixels = sum(w * h for (w, h) in self.sizes)

    def algorithmic_bytes(self) -> int:
        """import 4+1+1 (+UV), upsample 1+1 (+UV) + 4, chain 4+4 per pixel (tile codes negligible)."""
        t = 0
        for (w, h, uvw, uvh, ys, uvs, ds) in self.geo:
            t += (6 * w * h + 2 * uvw * uvh) + (6 * w * h + 2 * uvw * uvh) + 8 * w * h
        return t

    def fill_synthetic(self, seed: int):
        c = self.ctx
        sub = lambda v, b: (v + (1 << b) - 1) >> b  # noqa: E731
        for i, (w, h, *_rest) in enumerate(self.geo):
            c.synth_fill(self.d_src.ptr + self.off["src"][i], w, h, 4 * w, _al(4 * w * h), 1, seed, i, 0, _abi.SYNTH_SPARSE)
            words = sub(w, self.bits) * sub(h, self.bits)
            c.synth_fill(self.d_data.ptr + self.off["data"][i], words, 1, 4 * words, _al(4 * words), 1, seed, i, 6)
        c.sync()

    def run_device(self, stream: int = 0):
        c = self.ctx
        c._chk(c.lib.b2_rgba_to_yuv420a(c.h, stream, self._imp, self.n, _abi.ORDER_RGBA, _abi.IMPORT_ALPHA, self.d_flags.ptr))
        c._chk(c.lib.b2_yuv420a_to_rgba_fancy(c.h, stream, self._ups, self.n, self.mode))
        c._chk(c.lib.b2_vp8l_inverse_chain(c.h, stream, self._arg, self.n))

    def close(self):
        for b in ("d_src", "d_y", "d_a", "d_u", "d_v", "d_dst", "d_out", "d_data", "d_flags"):
            getattr(self, b).free()


def calibrate_mixed_rates(ctx: Context, sizes, seed: int = 0xB2000005) -> dict:
    """One warm-up pass for the time-based partition of a mixed-size batch (configs[4]): the images of
    every class (partition.size_class: row alignment x width bin) run as one batch through the full
    pipeline; rate = algorithmic bytes / (time - the fixed cost of three launches with next to nothing
    to do), in bytes per millisecond.  Returns class -> rate."""
    from . import partition

    def time_pass(sub):
        p = MixedPipeline(ctx, sub)
        p.fill_synthetic(seed)
        for _ in range(2):
            p.run_device()
        ctx.sync()
        ctx.event_record(8)
        for _ in range(3):
            p.run_device()
        ctx.event_record(9)
        ctx.sync()
        ms, by = ctx.elapsed_ms(8, 9) / 3, p.algorithmic_bytes()
        p.close()
        return ms, by

    t0, _ = time_pass([(64, 64)])
    groups = [partition.size_class(w, h) for (w, h) in sizes]
    rates = {}
    for g in sorted(set(groups)):
        ms, by = time_pass([s for s, gg in zip(sizes, groups) if gg == g])
        rates[g] = by / max(ms - t0, 0.05 * ms)
    return rates
