"""Context / buffer plumbing over the C ABI (host side of pkg/cuda in the Go shim).

One Context per GPU and per host thread, exactly as include/b2webp.h states.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from ._abi import B2Error


class DevBuf:
    """A device allocation owned by the library (b2_dev_alloc)."""

    def __init__(self, ctx: "Context", nbytes: int):
        self.ctx, self.nbytes = ctx, int(nbytes)
        p = C.c_void_p()
        ctx._chk(ctx.lib.b2_dev_alloc(ctx.h, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def free(self):
        if self.ptr:
            self.ctx._chk(self.ctx.lib.b2_dev_free(self.ctx.h, self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedBuf:
    """Pinned host staging memory (b2_pinned_alloc) viewed as a numpy uint8 array."""

    def __init__(self, ctx: "Context", nbytes: int):
        self.ctx, self.nbytes = ctx, int(nbytes)
        p = C.c_void_p()
        ctx._chk(ctx.lib.b2_pinned_alloc(ctx.h, self.nbytes, C.byref(p)))
        self.ptr = p.value
        self.array = np.frombuffer((C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr),
                                   np.uint8, count=self.nbytes)

    def free(self):
        if self.ptr:
            self.array = None
            self.ctx._chk(self.ctx.lib.b2_pinned_free(self.ctx.h, self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    def __init__(self, device: int = 0, n_streams: int = 4):
        self.lib = _abi.load()
        h = C.c_void_p()
        st = self.lib.b2_ctx_create(device, n_streams, C.byref(h))
        if st != 0:
            raise B2Error(st, self.lib.b2_strerror(st).decode())
        self.h = h
        self.n_streams = n_streams
        self.device = device

    # -- plumbing -------------------------------------------------------------
    def _chk(self, st: int):
        if st != 0:
            msg = self.lib.b2_strerror(st).decode()
            if st == _abi.B2_ERR_CUDA:
                msg += ": " + self.lib.b2_last_error(self.h).decode()
            raise B2Error(st, msg)

    def close(self):
        if self.h:
            self.lib.b2_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def swap_16bit_csp(self) -> bool:
        """Byte order of the 16-bit output modes, library-wide: True (default) = the go-webp text
        (WEBP_SWAP_16BIT_CSP: ba / gb byte first), False = the stock libwebp build (rg first)."""
        return bool(self.lib.b2_ctx_swap_16bit_csp(self.h))

    @swap_16bit_csp.setter
    def swap_16bit_csp(self, swap: bool):
        self._chk(self.lib.b2_ctx_set_swap_16bit_csp(self.h, int(bool(swap))))

    @property
    def sm_count(self) -> int:
        return self.lib.b2_ctx_sm_count(self.h)

    @property
    def launches(self) -> int:
        return int(self.lib.b2_launch_count(self.h))

    def copy_stats(self) -> dict:
        """Copies issued through this context so far (calls and bytes per direction)."""
        out = (C.c_uint64 * 4)()
        self._chk(self.lib.b2_copy_stats(self.h, out))
        return {"h2d_calls": int(out[0]), "h2d_bytes": int(out[1]), "d2h_calls": int(out[2]), "d2h_bytes": int(out[3])}

    def alloc(self, nbytes: int) -> DevBuf:
        return DevBuf(self, nbytes)

    def pinned(self, nbytes: int) -> PinnedBuf:
        return PinnedBuf(self, nbytes)

    def sync(self, stream: int = -1):
        self._chk(self.lib.b2_sync(self.h, stream))

    def h2d(self, dst_ptr: int, arr: np.ndarray, stream: int = 0):
        arr = np.ascontiguousarray(arr)
        self._chk(self.lib.b2_h2d_async(self.h, stream, dst_ptr, arr.ctypes.data, arr.nbytes))
        self.sync(stream)  # pageable source: keep it alive until the copy is done

    def d2h(self, src_ptr: int, nbytes: int, stream: int = 0) -> np.ndarray:
        out = np.empty(nbytes, np.uint8)
        self._chk(self.lib.b2_d2h_async(self.h, stream, out.ctypes.data, src_ptr, nbytes))
        self.sync(stream)
        return out

    def upload(self, arr: np.ndarray, stream: int = 0) -> DevBuf:
        arr = np.ascontiguousarray(arr)
        buf = self.alloc(max(arr.nbytes, 16))
        if arr.nbytes:
            self.h2d(buf.ptr, arr, stream)
        return buf

    def memset(self, ptr: int, value: int, nbytes: int, stream: int = 0):
        self._chk(self.lib.b2_memset_async(self.h, stream, ptr, value, nbytes))

    def event_record(self, slot: int, stream: int = 0):
        self._chk(self.lib.b2_event_record(self.h, stream, slot))

    def elapsed_ms(self, start_slot: int, stop_slot: int) -> float:
        ms = C.c_float()
        self._chk(self.lib.b2_event_elapsed_ms(self.h, start_slot, stop_slot, C.byref(ms)))
        return float(ms.value)

    # -- kernels (device pointers) ----------------------------------------------
    def rgba_to_yuv420a(self, jobs, order=_abi.ORDER_RGBA, flags=_abi.IMPORT_ALPHA,
                        has_alpha_ptr=None, stream: int = 0):
        arr = (_abi.ImportJob * len(jobs))(*jobs)
        self._chk(self.lib.b2_rgba_to_yuv420a(self.h, stream, arr, len(jobs), order, flags,
                                              has_alpha_ptr))

    def yuv420a_to_rgba_fancy(self, jobs, mode=_abi.MODE_RGBA, stream: int = 0):
        arr = (_abi.UpsampleJob * len(jobs))(*jobs)
        self._chk(self.lib.b2_yuv420a_to_rgba_fancy(self.h, stream, arr, len(jobs), mode))

    def vp8l(self, op: str, jobs, stream: int = 0):
        fn = getattr(self.lib, "b2_vp8l_" + op)
        arr = (_abi.ArgbJob * len(jobs))(*jobs)
        self._chk(fn(self.h, stream, arr, len(jobs)))

    def synth_fill(self, ptr, words_per_row, rows, stride_bytes, image_bytes, n_images,
                   seed, first_img, plane, alpha_mode=_abi.SYNTH_RAW, stream: int = 0):
        self._chk(self.lib.b2_synth_fill_batch_u32(
            self.h, stream, ptr, words_per_row, rows, stride_bytes, image_bytes, n_images,
            seed & 0xffffffff, first_img, plane, alpha_mode))


def device_count() -> int:
    n = C.c_int()
    st = _abi.load().b2_device_count(C.byref(n))
    return n.value if st == 0 else 0
