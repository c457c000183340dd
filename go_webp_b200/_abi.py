"""ctypes declarations of include/b2webp.h (the C ABI of libb2webp.so).

There is no CPU fallback: loading fails loudly when the CUDA library has not
been built (python -c 'import __graft_entry__ as g; g.build()').
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libb2webp.so")

B2_OK, B2_ERR_OOM, B2_ERR_BAD_DIM, B2_ERR_NULL, B2_ERR_INVALID_CONFIG, B2_ERR_CUDA = 0, -1, -2, -3, -4, -5
ORDER_RGBA, ORDER_BGRA, ORDER_RGB, ORDER_BGR = 0, 1, 2, 3
IMPORT_ALPHA, IMPORT_DIV_TABLE = 1, 2
(MODE_RGB, MODE_RGBA, MODE_BGR, MODE_BGRA, MODE_ARGB, MODE_RGBA_4444, MODE_RGB_565,
 MODE_rgbA, MODE_bgrA, MODE_Argb, MODE_rgbA_4444) = range(11)
MODE_BPP = {0: 3, 1: 4, 2: 3, 3: 4, 4: 4, 5: 2, 6: 2, 7: 4, 8: 4, 9: 4, 10: 2}
SYNTH_RAW, SYNTH_OPAQUE, SYNTH_SPARSE, SYNTH_BINARY = 0, 1, 2, 3


class ImportJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("y", C.c_void_p), ("u", C.c_void_p),
                ("v", C.c_void_p), ("a", C.c_void_p), ("src_stride", C.c_int32),
                ("y_stride", C.c_int32), ("uv_stride", C.c_int32),
                ("a_stride", C.c_int32), ("width", C.c_int32), ("height", C.c_int32)]


class UpsampleJob(C.Structure):
    _fields_ = [("y", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p),
                ("a", C.c_void_p), ("dst", C.c_void_p), ("y_stride", C.c_int32),
                ("uv_stride", C.c_int32), ("a_stride", C.c_int32),
                ("dst_stride", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("row_begin", C.c_int32), ("row_end", C.c_int32)]


class ArgbJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("data", C.c_void_p),
                ("width", C.c_int32), ("height", C.c_int32), ("bits", C.c_int32),
                ("reserved", C.c_int32)]


class PlainJob(C.Structure):
    _fields_ = [("argb", C.c_void_p), ("y", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p),
                ("a", C.c_void_p), ("argb_stride", C.c_int32), ("y_stride", C.c_int32),
                ("uv_stride", C.c_int32), ("a_stride", C.c_int32), ("width", C.c_int32),
                ("height", C.c_int32), ("y_pos", C.c_int32), ("reserved", C.c_int32)]


class AlphaIndexJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("palette", C.c_void_p),
                ("width", C.c_int32), ("height", C.c_int32), ("bits", C.c_int32),
                ("reserved", C.c_int32)]


class ExpandJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("num_colors", C.c_int32),
                ("bits", C.c_int32)]


class ConvertJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("dst", C.c_void_p),
                ("src_stride", C.c_int32), ("uv_stride", C.c_int32), ("dst_stride", C.c_int32),
                ("width", C.c_int32), ("height", C.c_int32), ("reserved", C.c_int32)]


class PictureJob(C.Structure):
    _fields_ = [("y", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("a", C.c_void_p),
                ("argb", C.c_void_p), ("y_stride", C.c_int32), ("uv_stride", C.c_int32),
                ("a_stride", C.c_int32), ("argb_stride", C.c_int32), ("width", C.c_int32),
                ("height", C.c_int32)]


class AnimBlendJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("src_stride", C.c_int32), ("dst_stride", C.c_int32),
                ("width", C.c_int32), ("height", C.c_int32)]


class AnimCompareJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("src_step", C.c_int32), ("dst_step", C.c_int32),
                ("length", C.c_int32), ("reserved", C.c_int32)]


class AlphaFilterJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("src_stride", C.c_int32), ("dst_stride", C.c_int32),
                ("width", C.c_int32), ("height", C.c_int32), ("filter", C.c_int32), ("reserved", C.c_int32)]


class AlphaDequantJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("src_stride", C.c_int32), ("dst_stride", C.c_int32),
                ("width", C.c_int32), ("height", C.c_int32), ("strength", C.c_int32), ("reserved", C.c_int32)]


class RescaleJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("src_width", C.c_int32), ("src_height", C.c_int32),
                ("src_stride", C.c_int32), ("dst_width", C.c_int32), ("dst_height", C.c_int32), ("dst_stride", C.c_int32),
                ("num_channels", C.c_int32), ("reserved", C.c_int32)]


class PaletteJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("palette", C.c_void_p), ("src_stride", C.c_int32),
                ("dst_stride", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("palette_size", C.c_int32),
                ("xbits", C.c_int32)]


class ChainJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("predictor_data", C.c_void_p), ("color_data", C.c_void_p),
                ("width", C.c_int32), ("height", C.c_int32), ("predictor_bits", C.c_int32), ("color_bits", C.c_int32),
                ("add_green", C.c_int32), ("reserved", C.c_int32)]


class ChainHost(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("predictor_data", C.c_void_p), ("color_data", C.c_void_p),
                ("predictor_bytes", C.c_size_t), ("color_bytes", C.c_size_t)]


class ImportHost(C.Structure):
    _fields_ = [("src", C.c_void_p), ("y", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("a", C.c_void_p),
                ("y_stride", C.c_int32), ("uv_stride", C.c_int32), ("a_stride", C.c_int32), ("reserved", C.c_int32)]


class UpsampleHost(C.Structure):
    _fields_ = [("y", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("a", C.c_void_p), ("dst", C.c_void_p),
                ("y_stride", C.c_int32), ("uv_stride", C.c_int32), ("a_stride", C.c_int32), ("dst_stride", C.c_int32)]


class ArgbHost(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("data", C.c_void_p),
                ("src_bytes", C.c_size_t), ("dst_bytes", C.c_size_t), ("data_bytes", C.c_size_t)]


VP8L_OPS = {"add_green": 0, "sub_green": 1, "color_inverse": 2, "color_forward": 3, "inverse_chain": 4,
            "forward_chain": 5, "color_index_inverse": 6, "bundle_color_map": 7, "predictor_inverse": 8,
            "predictor_forward": 9}

CONVERT_UPSTREAM_16BIT = 1
ALPHA_APPLY, ALPHA_APPLY_4444, ALPHA_MULT_ARGB, ALPHA_MULT_ROW, ALPHA_DISPATCH = 0, 1, 2, 3, 4

# every symbol include/b2webp.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
_ctx = C.c_void_p
SYMBOLS = {
    "b2_abi_version": (C.c_int, []),
    "b2_strerror": (C.c_char_p, [C.c_int]),
    "b2_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "b2_ctx_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(_ctx)]),
    "b2_ctx_destroy": (None, [_ctx]),
    "b2_ctx_set_swap_16bit_csp": (C.c_int, [_ctx, C.c_int]),
    "b2_ctx_swap_16bit_csp": (C.c_int, [_ctx]),
    "b2_ctx_device": (C.c_int, [_ctx]),
    "b2_ctx_sm_count": (C.c_int, [_ctx]),
    "b2_last_error": (C.c_char_p, [_ctx]),
    "b2_launch_count": (C.c_uint64, [_ctx]),
    "b2_pinned_alloc": (C.c_int, [_ctx, C.c_size_t, C.POINTER(_vp)]),
    "b2_pinned_free": (C.c_int, [_ctx, _vp]),
    "b2_host_register": (C.c_int, [_ctx, _vp, C.c_size_t]),
    "b2_host_unregister": (C.c_int, [_ctx, _vp]),
    "b2_dev_alloc": (C.c_int, [_ctx, C.c_size_t, C.POINTER(_vp)]),
    "b2_dev_free": (C.c_int, [_ctx, _vp]),
    "b2_h2d_async": (C.c_int, [_ctx, C.c_int, _vp, _vp, C.c_size_t]),
    "b2_d2h_async": (C.c_int, [_ctx, C.c_int, _vp, _vp, C.c_size_t]),
    "b2_h2d_2d_async": (C.c_int, [_ctx, C.c_int, _vp, C.c_size_t, _vp, C.c_size_t, C.c_size_t, C.c_size_t]),
    "b2_d2h_2d_async": (C.c_int, [_ctx, C.c_int, _vp, C.c_size_t, _vp, C.c_size_t, C.c_size_t, C.c_size_t]),
    "b2_memset_async": (C.c_int, [_ctx, C.c_int, _vp, C.c_int, C.c_size_t]),
    "b2_sync": (C.c_int, [_ctx, C.c_int]),
    "b2_event_record": (C.c_int, [_ctx, C.c_int, C.c_int]),
    "b2_event_elapsed_ms": (C.c_int, [_ctx, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "b2_event_sync": (C.c_int, [_ctx, C.c_int]),
    "b2_copy_stats": (C.c_int, [_ctx, C.POINTER(C.c_uint64)]),
    "b2_import_batch_staged": (C.c_int, [_ctx, C.POINTER(ImportJob), C.POINTER(ImportHost), C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_float, _vp]),
    "b2_upsample_batch_staged": (C.c_int, [_ctx, C.POINTER(UpsampleJob), C.POINTER(UpsampleHost), C.c_int, C.c_int, C.c_int]),
    "b2_vp8l_batch_staged": (C.c_int, [_ctx, C.c_int, C.POINTER(ArgbJob), C.POINTER(ArgbHost), C.c_int, C.c_int]),
    "b2_rgba_to_yuv420a": (C.c_int, [_ctx, C.c_int, C.POINTER(ImportJob), C.c_int, C.c_int, C.c_int, _vp]),
    "b2_rgba_to_yuv420a_dithered": (C.c_int, [_ctx, C.c_int, C.POINTER(ImportJob), C.c_int, C.c_int, C.c_int, _vp, C.c_float]),
    "b2_yuv420a_to_rgba_fancy": (C.c_int, [_ctx, C.c_int, C.POINTER(UpsampleJob), C.c_int, C.c_int]),
    "b2_vp8l_add_green": (C.c_int, [_ctx, C.c_int, C.POINTER(ArgbJob), C.c_int]),
    "b2_vp8l_sub_green": (C.c_int, [_ctx, C.c_int, C.POINTER(ArgbJob), C.c_int]),
    "b2_vp8l_color_inverse": (C.c_int, [_ctx, C.c_int, C.POINTER(ArgbJob), C.c_int]),
    "b2_vp8l_color_forward": (C.c_int, [_ctx, C.c_int, C.POINTER(ArgbJob), C.c_int]),
    "b2_vp8l_inverse_chain": (C.c_int, [_ctx, C.c_int, C.POINTER(ArgbJob), C.c_int]),
    "b2_vp8l_forward_chain": (C.c_int, [_ctx, C.c_int, C.POINTER(ArgbJob), C.c_int]),
    "b2_vp8l_color_index_inverse": (C.c_int, [_ctx, C.c_int, C.POINTER(ArgbJob), C.c_int]),
    "b2_vp8l_bundle_color_map": (C.c_int, [_ctx, C.c_int, C.POINTER(ArgbJob), C.c_int]),
    "b2_vp8l_apply_palette": (C.c_int, [_ctx, C.c_int, C.POINTER(PaletteJob), C.c_int]),
    "b2_vp8l_predictor_inverse": (C.c_int, [_ctx, C.c_int, C.POINTER(ArgbJob), C.c_int]),
    "b2_vp8l_predictor_inverse_chain": (C.c_int, [_ctx, C.c_int, C.POINTER(ChainJob), C.c_int, C.c_int]),
    "b2_vp8l_chain_batch_staged": (C.c_int, [_ctx, C.POINTER(ChainJob), C.POINTER(ChainHost), C.c_int, C.c_int, C.c_int]),
    "b2_vp8l_predictor_forward": (C.c_int, [_ctx, C.c_int, C.POINTER(ArgbJob), C.c_int]),
    "b2_vp8l_convert_from_bgra": (C.c_int, [_ctx, C.c_int, C.POINTER(ConvertJob), C.c_int, C.c_int, C.c_int]),
    "b2_alpha_multiply": (C.c_int, [_ctx, C.c_int, C.POINTER(ConvertJob), C.c_int, C.c_int, C.c_int]),
    "b2_yuv_to_rgb_point": (C.c_int, [_ctx, C.c_int, C.POINTER(ConvertJob), C.c_int, C.c_int, C.c_int]),
    "b2_argb_to_yuva_plain": (C.c_int, [_ctx, C.c_int, C.POINTER(PlainJob), C.c_int]),
    "b2_import_to_argb": (C.c_int, [_ctx, C.c_int, C.POINTER(ConvertJob), C.c_int, C.c_int, C.c_int]),
    "b2_anim_blend_rows": (C.c_int, [_ctx, C.c_int, C.POINTER(AnimBlendJob), C.c_int, C.c_int]),
    "b2_anim_compare_pixels": (C.c_int, [_ctx, C.c_int, C.POINTER(AnimCompareJob), C.c_int, C.c_int, C.c_int, _vp]),
    "b2_alpha_unfilter": (C.c_int, [_ctx, C.c_int, C.POINTER(AlphaFilterJob), C.c_int]),
    "b2_alpha_filter": (C.c_int, [_ctx, C.c_int, C.POINTER(AlphaFilterJob), C.c_int]),
    "b2_rescale_planes": (C.c_int, [_ctx, C.c_int, C.POINTER(RescaleJob), C.c_int]),
    "b2_alpha_dequantize_levels": (C.c_int, [_ctx, C.c_int, C.POINTER(AlphaDequantJob), C.c_int, C.c_int]),
    "b2_picture_cleanup_transparent": (C.c_int, [_ctx, C.c_int, C.POINTER(PictureJob), C.c_int, C.c_int]),
    "b2_picture_blend_alpha": (C.c_int, [_ctx, C.c_int, C.POINTER(PictureJob), C.c_int, C.c_int, C.c_uint32]),
    "b2_picture_replace_transparent": (C.c_int, [_ctx, C.c_int, C.POINTER(PictureJob), C.c_int, C.c_uint32]),
    "b2_vp8l_color_index_inverse_alpha": (C.c_int, [_ctx, C.c_int, C.POINTER(AlphaIndexJob), C.c_int]),
    "b2_vp8l_expand_color_map": (C.c_int, [_ctx, C.c_int, C.POINTER(ExpandJob), C.c_int]),
    "b2_synth_fill_batch_u32": (C.c_int, [_ctx, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_size_t,
                                          C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]),
}

_LIB = None


class B2Error(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"b2webp status {status}: {msg}")
        self.status = status


def load() -> C.CDLL:
    """Loads libb2webp.so; raises (never falls back) when it is missing."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA library must be built "
                "(__graft_entry__.build()); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        if lib.b2_abi_version() != 0x0105:
            raise ImportError("libb2webp.so ABI version mismatch; rebuild")
        _LIB = lib
    return _LIB
