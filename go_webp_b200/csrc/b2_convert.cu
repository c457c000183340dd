// b2_convert.cu -- the output conversions that follow parts 2 and 3 in the decoder
// (SURVEY.md 8f-2):
//   VP8LConvertFromBGRA and its rows          pkg/libwebp/dsp/lossless.go:549-666
//   ApplyAlphaMultiply_C / ..4444_C           pkg/libwebp/dsp/alpha_processing.go:100-163
//   WebPMultARGBRow_C / WebPMultRow_C         pkg/libwebp/dsp/alpha_processing.go:17-89
//   point samplers (row_func, WebPSamplerProcessPlane == EmitSampledRGB)
//                                             pkg/libwebp/dsp/yuv.go:19-72, decoder/io_dec.c.go:53-59
//   YUV444 converters                         pkg/libwebp/dsp/upsampling.c.go:191-245
// All per-pixel and HBM-bound; one thread handles 4 pixels of a row.
//
// 16-bit modes follow the Go text (ba / gb byte first, rg_byte_pos 1: upstream built
// with WEBP_SWAP_16BIT_CSP); B2_CONVERT_UPSTREAM_16BIT flips it.
// The YUV writers keep the upstream default order, like b2_upsample.cu.
#include <stdlib.h>

#include "b2_common.cuh"
#include "b2_upsample_core.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct ConvJobDev {
  const void* src;     // uint32 pixels, or the Y plane
  const uint8_t* u;    // chroma planes (samplers) or alpha plane (mult rows)
  const uint8_t* v;
  uint8_t* dst;
  int32_t src_stride;  // pixels for uint32 sources, bytes for planes
  int32_t uv_stride;
  int32_t dst_stride;  // bytes
  int32_t width, height;
  uint32_t chunks_x;   // kItemPx-pixel chunks per row
};

enum {  // canonical store modes
  ST_RGB = 0, ST_RGBA = 1, ST_BGR = 2, ST_BGRA = 3, ST_ARGB = 4, ST_4444 = 5, ST_565 = 6
};
__host__ __device__ inline int canonical(int mode, bool* premult) {
  *premult = (mode == B2_MODE_rgbA || mode == B2_MODE_bgrA || mode == B2_MODE_Argb || mode == B2_MODE_rgbA_4444);
  switch (mode) {
    case B2_MODE_RGB: return ST_RGB;
    case B2_MODE_BGR: return ST_BGR;
    case B2_MODE_RGBA: case B2_MODE_rgbA: return ST_RGBA;
    case B2_MODE_BGRA: case B2_MODE_bgrA: return ST_BGRA;
    case B2_MODE_ARGB: case B2_MODE_Argb: return ST_ARGB;
    case B2_MODE_RGBA_4444: case B2_MODE_rgbA_4444: return ST_4444;
    case B2_MODE_RGB_565: return ST_565;
    default: return -1;
  }
}
__host__ __device__ inline int bpp_of(int st) { return st == ST_RGB || st == ST_BGR ? 3 : (st >= ST_4444 ? 2 : 4); }

// PREMULTIPLY(x, MULTIPLIER1(a)), alpha_processing.go:103-108
__device__ __forceinline__ uint32_t premul(uint32_t x, uint32_t a) { return (x * (a * 32897u)) >> 23; }
// ApplyAlphaMultiply4444_C on one pixel (alpha_processing.go:128-156)
__device__ __forceinline__ void premul4444(uint32_t* rg, uint32_t* ba) {
  const uint32_t a = *ba & 0x0f;
  const uint32_t mult = a * 0x1111u;
  const uint32_t hi_rg = (*rg & 0xf0) | (*rg >> 4), lo_rg = ((*rg & 0x0f) | (*rg << 4)) & 0xff;
  const uint32_t hi_ba = (*ba & 0xf0) | (*ba >> 4);
  const uint32_t r = (hi_rg * mult) >> 16, g = (lo_rg * mult) >> 16, b = (hi_ba * mult) >> 16;
  *rg = (r & 0xf0) | ((g >> 4) & 0x0f);
  *ba = (b & 0xf0) | a;
}
// One pixel (a, r, g, b bytes) into 'st' byte order at d.
__device__ __forceinline__ void store_px(int st, bool premult, bool upstream16, uint32_t a, uint32_t r, uint32_t g,
                                         uint32_t b, uint8_t* d) {
  if (st == ST_4444) {
    uint32_t rg = (r & 0xf0) | (g >> 4), ba = (b & 0xf0) | (a >> 4);
    if (premult) premul4444(&rg, &ba);
    d[upstream16 ? 0 : 1] = (uint8_t)rg;
    d[upstream16 ? 1 : 0] = (uint8_t)ba;
    return;
  }
  if (st == ST_565) {
    const uint32_t rg = (r & 0xf8) | (g >> 5), gb = ((g << 3) & 0xe0) | (b >> 3);
    d[upstream16 ? 0 : 1] = (uint8_t)rg;
    d[upstream16 ? 1 : 0] = (uint8_t)gb;
    return;
  }
  if (premult && a != 0xff) { r = premul(r, a); g = premul(g, a); b = premul(b, a); }
  switch (st) {
    case ST_RGB: d[0] = r; d[1] = g; d[2] = b; break;
    case ST_BGR: d[0] = b; d[1] = g; d[2] = r; break;
    case ST_RGBA: d[0] = r; d[1] = g; d[2] = b; d[3] = a; break;
    case ST_BGRA: d[0] = b; d[1] = g; d[2] = r; d[3] = a; break;
    default: d[0] = a; d[1] = r; d[2] = g; d[3] = b; break;  // ST_ARGB
  }
}
__device__ __forceinline__ uint32_t pack_px(int st, bool premult, uint32_t a, uint32_t r, uint32_t g, uint32_t b) {
  if (premult && a != 0xff) { r = premul(r, a); g = premul(g, a); b = premul(b, a); }
  if (st == ST_RGBA) return r | (g << 8) | (b << 16) | (a << 24);
  if (st == ST_BGRA) return b | (g << 8) | (r << 16) | (a << 24);
  return a | (r << 8) | (g << 16) | (b << 24);  // ST_ARGB
}
// Four consecutive pixels into 'st' byte order at d: whole words when d allows it (16, 12 or 8
// bytes per lane: one 128-bit, three 32-bit or one 64-bit store), else byte by byte.
__device__ __forceinline__ void store_px4(int st, bool premult, bool upstream16, const uint32_t* a, const uint32_t* r,
                                          const uint32_t* g, const uint32_t* b, uint8_t* d) {
  if ((((uintptr_t)d) & 3) != 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) store_px(st, premult, upstream16, a[k], r[k], g[k], b[k], d + k * bpp_of(st));
    return;
  }
  if (st == ST_4444 || st == ST_565) {
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t hi, lo;   // hi = the RG byte, lo = the BA / GB byte
      if (st == ST_4444) {
        hi = (r[k] & 0xf0) | (g[k] >> 4); lo = (b[k] & 0xf0) | (a[k] >> 4);
        if (premult) premul4444(&hi, &lo);
      } else {
        hi = (r[k] & 0xf8) | (g[k] >> 5); lo = ((g[k] << 3) & 0xe0) | (b[k] >> 3);
      }
      v[k] = upstream16 ? ((hi & 0xff) | ((lo & 0xff) << 8)) : ((lo & 0xff) | ((hi & 0xff) << 8));
    }
    uint32_t* o = (uint32_t*)d;
    o[0] = v[0] | (v[1] << 16);
    o[1] = v[2] | (v[3] << 16);
    return;
  }
  if (st == ST_RGB || st == ST_BGR) {
    uint32_t c0[4], c2[4];   // first and third byte of a pixel
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      c0[k] = (st == ST_RGB) ? r[k] : b[k];
      c2[k] = (st == ST_RGB) ? b[k] : r[k];
    }
    uint32_t* o = (uint32_t*)d;
    o[0] = c0[0] | (g[0] << 8) | (c2[0] << 16) | (c0[1] << 24);
    o[1] = g[1] | (c2[1] << 8) | (c0[2] << 16) | (g[2] << 24);
    o[2] = c2[2] | (c0[3] << 8) | (g[3] << 16) | (c2[3] << 24);
    return;
  }
  uint4 q;
  q.x = pack_px(st, premult, a[0], r[0], g[0], b[0]);
  q.y = pack_px(st, premult, a[1], r[1], g[1], b[1]);
  q.z = pack_px(st, premult, a[2], r[2], g[2], b[2]);
  q.w = pack_px(st, premult, a[3], r[3], g[3], b[3]);
  if ((((uintptr_t)d) & 15) == 0) {
    b2_stg_stream_v4(d, q);
  } else {
    uint32_t* o = (uint32_t*)d;
    o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
  }
}

// Item walk of the ARGB converters: a warp takes 128 pixels of one row, a lane 4 of them.
#define B2_CONV_ITEM_PROLOGUE                                                                        \
  const int lane = threadIdx.x & 31;                                                                 \
  const uint32_t warp = threadIdx.x >> 5;                                                            \
  for (uint32_t item = blockIdx.x * kWarps + warp; item < total_items; item += gridDim.x * kWarps) { \
    uint32_t local;                                                                                  \
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, item, &local);                       \
    const ConvJobDev j = jobs[img];                                                                  \
    const uint32_t y = local / j.chunks_x;                                                           \
    const int x0 = (int)(local - y * j.chunks_x) * 128 + lane * 4;                                   \
    if (x0 >= j.width) continue;                                                                     \
    const int nv = min(4, j.width - x0);

// Item walk of the in-place multiplies and the YUV converters: a warp takes kItemPx pixels of one row, 128 at a time, a lane 4 of them
// (the item's image lookup and job descriptor are paid once per kItemPx pixels).
constexpr int kItemPx = 512;   // the ARGB converter, the in-place multiplies and the YUV converters; the use_argb import keeps 128
#define B2_CONV_ITEM_PROLOGUE_WIDE(ITEM_PX)                                                                        \
  const int lane = threadIdx.x & 31;                                                                 \
  const uint32_t warp = threadIdx.x >> 5;                                                            \
  for (uint32_t item = blockIdx.x * kWarps + warp; item < total_items; item += gridDim.x * kWarps) { \
    uint32_t local;                                                                                  \
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, item, &local);                       \
    const ConvJobDev j = jobs[img];                                                                  \
    const uint32_t y = local / j.chunks_x;                                                           \
    const int xc = (int)(local - y * j.chunks_x) * (ITEM_PX);                                        \
    const int xe = min(xc + (ITEM_PX), j.width);                                                     \
    for (int x0 = xc + lane * 4; x0 < xe; x0 += 128) {                                               \
    const int nv = min(4, j.width - x0);

// ---- VP8LConvertFromBGRA ----------------------------------------------------------
// A warp takes 512 pixels of one row: the four 128-bit loads of a lane are all requested before the first is
// converted (with one load per lane in flight the converters sat at 0.6 of the copy bandwidth: 64 warps x 512 bytes
// per SM is not enough to cover the memory latency), and the image lookup is paid once per 512 pixels.
__global__ void __launch_bounds__(kThreads, 4)
convert_from_bgra_kernel(const ConvJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix, uint32_t n_jobs,
                         uint32_t uniform, uint32_t total_items, int st, int premult, int upstream16) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5;
  for (uint32_t item = blockIdx.x * kWarps + warp; item < total_items; item += gridDim.x * kWarps) {
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, item, &local);
    const ConvJobDev j = jobs[img];
    const uint32_t y = local / j.chunks_x;
    const int xc = (int)(local - y * j.chunks_x) * kItemPx;
    const uint32_t* srow = (const uint32_t*)j.src + (ptrdiff_t)y * j.src_stride;
    uint32_t p[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int x0 = xc + u * 128 + lane * 4, nv = min(4, j.width - x0);
      const uint32_t* s = srow + x0;
      if (nv == 4 && (((uintptr_t)s) & 15) == 0) {
        const uint4 q = b2_ldg_stream_v4(s);
        p[u][0] = q.x; p[u][1] = q.y; p[u][2] = q.z; p[u][3] = q.w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) p[u][k] = (k < nv) ? __ldg(s + k) : 0u;
      }
    }
    const int bpp = bpp_of(st);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int x0 = xc + u * 128 + lane * 4, nv = min(4, j.width - x0);
      if (nv <= 0) continue;
      uint8_t* d = j.dst + (ptrdiff_t)y * j.dst_stride + x0 * bpp;
      if (bpp == 4 && nv == 4 && (((uintptr_t)d) & 15) == 0) {
        uint4 o;
        o.x = pack_px(st, premult, p[u][0] >> 24, (p[u][0] >> 16) & 0xff, (p[u][0] >> 8) & 0xff, p[u][0] & 0xff);
        o.y = pack_px(st, premult, p[u][1] >> 24, (p[u][1] >> 16) & 0xff, (p[u][1] >> 8) & 0xff, p[u][1] & 0xff);
        o.z = pack_px(st, premult, p[u][2] >> 24, (p[u][2] >> 16) & 0xff, (p[u][2] >> 8) & 0xff, p[u][2] & 0xff);
        o.w = pack_px(st, premult, p[u][3] >> 24, (p[u][3] >> 16) & 0xff, (p[u][3] >> 8) & 0xff, p[u][3] & 0xff);
        b2_stg_stream_v4(d, o);
      } else if (nv == 4) {
        uint32_t a4[4], r4[4], g4[4], b4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { a4[k] = p[u][k] >> 24; r4[k] = (p[u][k] >> 16) & 0xff; g4[k] = (p[u][k] >> 8) & 0xff; b4[k] = p[u][k] & 0xff; }
        store_px4(st, premult, upstream16, a4, r4, g4, b4, d);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < nv) store_px(st, premult, upstream16, p[u][k] >> 24, (p[u][k] >> 16) & 0xff, (p[u][k] >> 8) & 0xff, p[u][k] & 0xff, d + k * bpp);
      }
    }
  }
}

// ---- Import() into an ARGB picture (picture_csp_enc.c.go:138-181) -----------------------
// WebPPackRGB (step 3 or 4, alpha forced to 0xff), the memcpy of BGRA input and
// VP8LConvertBGRAToRGBA of RGBA input all give out = A << 24 | R << 16 | G << 8 | B.
// src_stride is in BYTES here (the caller's rgb_stride, possibly negative).
__global__ void __launch_bounds__(kThreads, 8)
import_argb_kernel(const ConvJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix, uint32_t n_jobs,
                   uint32_t uniform, uint32_t total_items, int order, int import_alpha) {
  const int step = (order == B2_ORDER_RGB || order == B2_ORDER_BGR) ? 3 : 4;
  const bool swap_rb = (order == B2_ORDER_BGRA || order == B2_ORDER_BGR);   // b,g,r(,a) in memory
  B2_CONV_ITEM_PROLOGUE
    const uint8_t* s = (const uint8_t*)j.src + (ptrdiff_t)y * j.src_stride + (ptrdiff_t)x0 * step;
    uint32_t* d = (uint32_t*)(j.dst + (ptrdiff_t)y * j.dst_stride) + x0;
    uint32_t o[4];
    if (step == 4 && nv == 4 && (((uintptr_t)s) & 15) == 0) {
      const uint4 q = b2_ldg_stream_v4(s);
      const uint32_t p[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t v = swap_rb ? p[k] : __byte_perm(p[k], 0, 0x3012);   // R,G,B,A bytes -> B,G,R,A
        o[k] = import_alpha ? v : (v | 0xff000000u);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        o[k] = 0u;
        if (k < nv) {
          const uint8_t* q = s + k * step;
          const uint32_t c0 = q[0], c1 = q[1], c2 = q[2];
          const uint32_t a = (step == 4 && import_alpha) ? q[3] : 0xffu;
          o[k] = (a << 24) | ((swap_rb ? c2 : c0) << 16) | (c1 << 8) | (swap_rb ? c0 : c2);
        }
      }
    }
    if (nv == 4 && (((uintptr_t)d) & 15) == 0) {
      b2_stg_stream_v4(d, make_uint4(o[0], o[1], o[2], o[3]));
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) if (k < nv) d[k] = o[k];
    }
  }
}

// ---- alpha multiplies (in place) ----------------------------------------------------
enum { AM_APPLY = 0, AM_APPLY_4444 = 1, AM_MULT_ARGB = 2, AM_MULT_ROW = 3, AM_DISPATCH = 4 };
#define MFIX 24
// Mult / GetScale, alpha_processing.go:17-31
__device__ __forceinline__ uint32_t mult_scale(uint32_t a, bool inverse) {
  return inverse ? (255u << MFIX) / a : a * ((1u << MFIX) / 255u);
}
__device__ __forceinline__ uint32_t mult_fix(uint32_t x, uint32_t scale) { return (x * scale + (1u << (MFIX - 1))) >> MFIX; }

__global__ void __launch_bounds__(kThreads, 8)
alpha_multiply_kernel(const ConvJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix, uint32_t n_jobs,
                      uint32_t uniform, uint32_t total_items, int kind, int flag) {
  B2_CONV_ITEM_PROLOGUE_WIDE(kItemPx)
    if (kind == AM_APPLY) {  // ApplyAlphaMultiply_C: 4 bytes per pixel, flag = alpha_first
      uint8_t* d = j.dst + (ptrdiff_t)y * j.dst_stride + x0 * 4;
      if (nv == 4 && (((uintptr_t)d) & 15) == 0) {   // a pixel per word
        uint4 q = *(const uint4*)d;
        uint32_t px[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t a = flag ? (px[k] & 0xff) : (px[k] >> 24);
          if (a != 0xff) {
            const int sh = flag ? 8 : 0;
            const uint32_t c0 = premul((px[k] >> sh) & 0xff, a), c1 = premul((px[k] >> (sh + 8)) & 0xff, a), c2 = premul((px[k] >> (sh + 16)) & 0xff, a);
            px[k] = flag ? (a | (c0 << 8) | (c1 << 16) | (c2 << 24)) : (c0 | (c1 << 8) | (c2 << 16) | (a << 24));
          }
        }
        *(uint4*)d = make_uint4(px[0], px[1], px[2], px[3]);
        continue;
      }
      for (int k = 0; k < nv; ++k) {
        uint8_t* px = d + 4 * k;
        const uint32_t a = px[flag ? 0 : 3];
        if (a != 0xff) {
          uint8_t* rgb = px + (flag ? 1 : 0);
          rgb[0] = (uint8_t)premul(rgb[0], a); rgb[1] = (uint8_t)premul(rgb[1], a); rgb[2] = (uint8_t)premul(rgb[2], a);
        }
      }
    } else if (kind == AM_APPLY_4444) {  // flag = rg_byte_pos
      uint8_t* d = j.dst + (ptrdiff_t)y * j.dst_stride + x0 * 2;
      for (int k = 0; k < nv; ++k) {
        uint32_t rg = d[2 * k + flag], ba = d[2 * k + (flag ^ 1)];
        premul4444(&rg, &ba);
        d[2 * k + flag] = (uint8_t)rg; d[2 * k + (flag ^ 1)] = (uint8_t)ba;
      }
    } else if (kind == AM_MULT_ARGB) {  // WebPMultARGBRow_C, flag = inverse
      uint32_t* d = (uint32_t*)(j.dst + (ptrdiff_t)y * j.dst_stride) + x0;
      if (nv == 4 && (((uintptr_t)d) & 15) == 0) {   // one 128-bit access each way
        const uint4 q = *(const uint4*)d;
        uint32_t px[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t argb = px[k];
          if (argb < 0xff000000u) {
            if (argb <= 0x00ffffffu) {
              px[k] = 0;
            } else {
              const uint32_t scale = mult_scale(argb >> 24, flag != 0);
              px[k] = (argb & 0xff000000u) | mult_fix(argb & 0xff, scale) | (mult_fix((argb >> 8) & 0xff, scale) << 8) |
                      (mult_fix((argb >> 16) & 0xff, scale) << 16);
            }
          }
        }
        *(uint4*)d = make_uint4(px[0], px[1], px[2], px[3]);
        continue;
      }
      for (int k = 0; k < nv; ++k) {
        const uint32_t argb = d[k];
        if (argb < 0xff000000u) {
          if (argb <= 0x00ffffffu) {
            d[k] = 0;
          } else {
            const uint32_t scale = mult_scale(argb >> 24, flag != 0);
            d[k] = (argb & 0xff000000u) | mult_fix(argb & 0xff, scale) | (mult_fix((argb >> 8) & 0xff, scale) << 8) |
                   (mult_fix((argb >> 16) & 0xff, scale) << 16);
          }
        }
      }
    } else if (kind == AM_DISPATCH) {  // DispatchAlpha_C (alpha_processing.go:165-180): alpha plane u -> alpha byte, flag = alpha_first
      uint8_t* d = j.dst + (ptrdiff_t)y * j.dst_stride + x0 * 4 + (flag ? 0 : 3);
      const uint8_t* al = j.u + (ptrdiff_t)y * j.uv_stride + x0;
      uint8_t* dw = j.dst + (ptrdiff_t)y * j.dst_stride + x0 * 4;
      if (nv == 4 && (((uintptr_t)dw) & 15) == 0 && (((uintptr_t)al) & 3) == 0) {   // the lane owns these 4 pixels: whole words
        const uint32_t aw = __ldg((const uint32_t*)al);
        uint4 q = *(const uint4*)dw;
        uint32_t px[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t a = (aw >> (8 * k)) & 0xff;
          px[k] = flag ? ((px[k] & 0xffffff00u) | a) : ((px[k] & 0x00ffffffu) | (a << 24));
        }
        *(uint4*)dw = make_uint4(px[0], px[1], px[2], px[3]);
        continue;
      }
      for (int k = 0; k < nv; ++k) d[4 * k] = al[k];
    } else {  // WebPMultRow_C: plane dst, alpha plane u, flag = inverse
      if (!flag && x0 == xc + lane * 4) {
        // Premultiplying direction, whole item on aligned rows: 16 samples per lane, the alpha and the plane load in
        // flight together, and no special cases -- with scale = a * 0x010101 the general formula already gives
        // v for a == 255 ((v * 0xffffff + 2^23) >> 24 == v) and 0 for a == 0, so a sample is two extracts, two
        // multiply-adds and a share of a byte permute (the general code below spends 28 instructions per sample
        // on the two tests, both directions of GetScale and the byte inserts: 0.35 of the peak, issue-bound).
        uint8_t* d16 = j.dst + (ptrdiff_t)y * j.dst_stride + xc + lane * 16;
        const uint8_t* a16 = j.u + (ptrdiff_t)y * j.uv_stride + xc + lane * 16;
        if ((((uintptr_t)d16 | (uintptr_t)a16) & 15) == 0) {   // the same for every lane
          const int left = j.width - (xc + lane * 16);
          if (left < 16) {   // the row ends inside (or before) this lane's 16 samples
            for (int k = 0; k < left; ++k) d16[k] = (uint8_t)mult_fix(d16[k], mult_scale(a16[k], false));
            break;
          }
          const uint4 aq = b2_ldg_stream_v4((const uint32_t*)a16);
          const uint4 dq = *(const uint4*)d16;
          const uint32_t aw4[4] = {aq.x, aq.y, aq.z, aq.w}, dw4[4] = {dq.x, dq.y, dq.z, dq.w};
          uint32_t o4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t t[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t a = (aw4[q] >> (8 * k)) & 0xff, v = (dw4[q] >> (8 * k)) & 0xff;
              t[k] = (a * v) * 0x010101u + (1u << (MFIX - 1));   // the result is byte 3
            }
            o4[q] = __byte_perm(__byte_perm(t[0], t[1], 0x0073), __byte_perm(t[2], t[3], 0x0073), 0x5410);
          }
          *(uint4*)d16 = make_uint4(o4[0], o4[1], o4[2], o4[3]);
          break;
        }
      }
      uint8_t* d = j.dst + (ptrdiff_t)y * j.dst_stride + x0;
      const uint8_t* al = j.u + (ptrdiff_t)y * j.uv_stride + x0;
      if (nv == 4 && (((uintptr_t)d | (uintptr_t)al) & 3) == 0) {
        const uint32_t aw = __ldg((const uint32_t*)al);
        if (aw != 0xffffffffu) {
          const uint32_t dw = *(const uint32_t*)d;
          uint32_t o = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t a = (aw >> (8 * k)) & 0xff, v = (dw >> (8 * k)) & 0xff;
            const uint32_t r = (a == 255) ? v : (a == 0) ? 0u : (mult_fix(v, mult_scale(a, flag != 0)) & 0xff);
            o |= r << (8 * k);
          }
          *(uint32_t*)d = o;
        }
        continue;
      }
      for (int k = 0; k < nv; ++k) {
        const uint32_t a = al[k];
        if (a != 255) d[k] = (a == 0) ? 0 : (uint8_t)mult_fix(d[k], mult_scale(a, flag != 0));
      }
    }
    }
  }
}

// ---- point-sampled / 4:4:4 YUV -> RGB -----------------------------------------------
// VP8YUVToR/G/B (pkg/color/yuv/conversion.go:28-48)
__device__ __forceinline__ void yuv_rgb(uint32_t y, uint32_t u, uint32_t v, uint32_t* r, uint32_t* g, uint32_t* b) {
  const int yy = (int)(y * 19077u) >> 8;
  const int r6 = (yy + ((int)(v * 26149u) >> 8) - 14234) >> 6;
  const int g6 = (yy - ((int)(u * 6419u) >> 8) - ((int)(v * 13320u) >> 8) + 8708) >> 6;
  const int b6 = (yy + ((int)(u * 33050u) >> 8) - 17685) >> 6;
  *r = (uint32_t)min(max(r6, 0), 255); *g = (uint32_t)min(max(g6, 0), 255); *b = (uint32_t)min(max(b6, 0), 255);
}
template <bool YUV444>
__global__ void __launch_bounds__(kThreads, 4)
yuv_point_kernel(const ConvJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix, uint32_t n_jobs,
                 uint32_t uniform, uint32_t total_items, int st, bool upstream16) {
  B2_CONV_ITEM_PROLOGUE_WIDE(kItemPx)
    const uint8_t* yr = (const uint8_t*)j.src + (ptrdiff_t)y * j.src_stride + x0;
    // WebPSamplerProcessPlane advances u/v after every odd row (yuv.go:52-55): chroma row y >> 1
    const ptrdiff_t crow = (ptrdiff_t)(YUV444 ? y : (y >> 1)) * j.uv_stride;
    const uint8_t* ur = j.u + crow;
    const uint8_t* vr = j.v + crow;
    const int bpp = bpp_of(st);
    uint8_t* d = j.dst + (ptrdiff_t)y * j.dst_stride + x0 * bpp;
    const int ng = (xe - xc) >> 7;   // whole 128-pixel groups of the item (the last item of a row may have fewer than 4)
    if (x0 == xc + lane * 4 && ((xe - xc) & 127) == 0 &&
        ((((uintptr_t)yr) & 3) | (((uintptr_t)(ur + (YUV444 ? x0 : (x0 >> 1))) | (uintptr_t)(vr + (YUV444 ? x0 : (x0 >> 1)))) & (YUV444 ? 3 : 1))) == 0) {
      // an item of whole 128-pixel groups on aligned rows (warp-uniform): the twelve loads of the lane's four groups first
      uint32_t yw4[4], uw4[4], vw4[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        yw4[g] = uw4[g] = vw4[g] = 0u;
        if (g >= ng) continue;
        yw4[g] = __ldg((const uint32_t*)(yr + 128 * g));
        if (YUV444) {
          uw4[g] = __ldg((const uint32_t*)(ur + x0 + 128 * g));
          vw4[g] = __ldg((const uint32_t*)(vr + x0 + 128 * g));
        } else {
          uw4[g] = __byte_perm((uint32_t)__ldg((const uint16_t*)(ur + (x0 >> 1) + 64 * g)), 0, 0x1100);
          vw4[g] = __byte_perm((uint32_t)__ldg((const uint16_t*)(vr + (x0 >> 1) + 64 * g)), 0, 0x1100);
        }
      }
      const B2YuvConsts kc = b2_yuv_consts();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (g >= ng) break;
        uint32_t a4[4] = {0xff, 0xff, 0xff, 0xff}, r4[4], g4[4], b4[4];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const uint32_t sel = p ? 0x4342u : 0x4140u;
          const B2Rgb14 c = b2_yuv_pair(b2_luma14(__byte_perm(yw4[g], 0u, sel), kc), __byte_perm(uw4[g], 0u, sel), __byte_perm(vw4[g], 0u, sel), kc);
          r4[2 * p] = (c.r >> 6) & 0xffu; r4[2 * p + 1] = c.r >> 22;
          g4[2 * p] = (c.g >> 6) & 0xffu; g4[2 * p + 1] = c.g >> 22;
          b4[2 * p] = (c.b >> 6) & 0xffu; b4[2 * p + 1] = c.b >> 22;
        }
        store_px4(st, false, upstream16, a4, r4, g4, b4, d + 128 * g * bpp);
      }
      break;   // the item is done
    }
    if (nv == 4) {   // whole words in (x0 is a multiple of 4: its chroma pair starts at x0 / 2), whole words out
      const uint32_t yw = ((((uintptr_t)yr) & 3) == 0) ? __ldg((const uint32_t*)yr)
                                                       : (uint32_t)yr[0] | ((uint32_t)yr[1] << 8) | ((uint32_t)yr[2] << 16) | ((uint32_t)yr[3] << 24);
      uint32_t uw, vw;   // the 4 chroma samples of the 4 pixels, one per byte
      if (YUV444) {
        const uint8_t* up = ur + x0;
        const uint8_t* vp = vr + x0;
        uw = ((((uintptr_t)up) & 3) == 0) ? __ldg((const uint32_t*)up) : (uint32_t)up[0] | ((uint32_t)up[1] << 8) | ((uint32_t)up[2] << 16) | ((uint32_t)up[3] << 24);
        vw = ((((uintptr_t)vp) & 3) == 0) ? __ldg((const uint32_t*)vp) : (uint32_t)vp[0] | ((uint32_t)vp[1] << 8) | ((uint32_t)vp[2] << 16) | ((uint32_t)vp[3] << 24);
      } else {
        const uint8_t* up = ur + (x0 >> 1);
        const uint8_t* vp = vr + (x0 >> 1);
        const uint32_t u2 = ((((uintptr_t)up) & 1) == 0) ? (uint32_t)__ldg((const uint16_t*)up) : (uint32_t)up[0] | ((uint32_t)up[1] << 8);
        const uint32_t v2 = ((((uintptr_t)vp) & 1) == 0) ? (uint32_t)__ldg((const uint16_t*)vp) : (uint32_t)vp[0] | ((uint32_t)vp[1] << 8);
        uw = __byte_perm(u2, 0, 0x1100);   // row_func: two pixels share a sample (yuv.go:24-30)
        vw = __byte_perm(v2, 0, 0x1100);
      }
      // two pixels per 32-bit register through the fancy upsampler's packed core (b2_upsample_core.cuh: the same
      // MultHi / Clip8 arithmetic on 16-bit lanes, checked over all 2^24 (y, u, v) on the host): half the multiplies
      // and clamps of the scalar form; uw / vw already hold one chroma byte per pixel in both samplings
      uint32_t a4[4] = {0xff, 0xff, 0xff, 0xff}, r4[4], g4[4], b4[4];
      const B2YuvConsts kc = b2_yuv_consts();
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const uint32_t sel = p ? 0x4342u : 0x4140u;   // bytes (2p, 2p + 1) into the two 16-bit lanes
        const B2Rgb14 c = b2_yuv_pair(b2_luma14(__byte_perm(yw, 0u, sel), kc), __byte_perm(uw, 0u, sel), __byte_perm(vw, 0u, sel), kc);
        r4[2 * p] = (c.r >> 6) & 0xffu; r4[2 * p + 1] = c.r >> 22;
        g4[2 * p] = (c.g >> 6) & 0xffu; g4[2 * p + 1] = c.g >> 22;
        b4[2 * p] = (c.b >> 6) & 0xffu; b4[2 * p + 1] = c.b >> 22;
      }
      store_px4(st, false, upstream16, a4, r4, g4, b4, d);
      continue;
    }
    for (int k = 0; k < nv; ++k) {
      const int cx = YUV444 ? (x0 + k) : ((x0 + k) >> 1);  // row_func: two pixels share a sample (yuv.go:24-30)
      uint32_t r, g, b;
      yuv_rgb(yr[k], ur[cx], vr[cx], &r, &g, &b);
      store_px(st, false, upstream16, 0xff, r, g, b, d + k * bpp);
    }
    }
  }
}

// flat_src / flat_uv: how to tell that a job is a per-pixel map between tightly packed buffers -- flat_src = what one
// pixel counts in src_stride (0: the kernel has no src; -1: src_stride is in pixels; k > 0: k bytes), flat_uv = 1 when the
// u (and v) planes are read at the pixel's own position.  Such a job whose width is NOT a multiple of 4 is handed to the
// kernel as ONE row of width * height pixels: its rows start at every byte alignment, the single row is as aligned as
// the buffer, so the 128-bit accesses apply again (1922-wide pictures ran these kernels at 0.23-0.60 of the peak
// against 0.73-0.88 for 1920-wide ones).  FLAT_NONE for kernels whose rows mean something (point-sampled 4:2:0).
enum { FLAT_NONE = -2 };
int prepare_conv(b2_ctx* ctx, int stream, const b2_convert_job* jobs, int n_jobs, bool need_uv, int min_dst_bpp, int item_px,
                 const ConvJobDev** dj, const uint32_t** dp, uint32_t* uniform, uint32_t* total,
                 int flat_src = FLAT_NONE, int flat_uv = 0) {
  const size_t jb = (sizeof(ConvJobDev) * (size_t)n_jobs + 15) & ~(size_t)15;
  const size_t pb = sizeof(uint32_t) * ((size_t)n_jobs + 1);
  if (jb + pb > B2_DESC_RING_BYTES) return B2_ERR_INVALID_CONFIG;
  uint8_t* host = (uint8_t*)malloc(jb + pb);
  if (host == NULL) return B2_ERR_OOM;
  ConvJobDev* d = (ConvJobDev*)host;
  uint32_t* pre = (uint32_t*)(host + jb);
  uint64_t tot = 0;
  uint32_t uni = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const b2_convert_job& j = jobs[i];
    int st = B2_OK;
    if (j.dst == NULL || (need_uv && j.u == NULL)) st = B2_ERR_NULL;
    else if (j.width <= 0 || j.height <= 0 || j.width > B2_MAX_DIMENSION || j.height > B2_MAX_DIMENSION) st = B2_ERR_BAD_DIM;
    else if ((j.dst_stride < 0 ? -(int64_t)j.dst_stride : j.dst_stride) < (int64_t)min_dst_bpp * j.width) st = B2_ERR_BAD_DIM;
    if (st != B2_OK) { free(host); return st; }
    d[i].src = j.src; d[i].u = j.u; d[i].v = j.v; d[i].dst = j.dst;
    d[i].src_stride = j.src_stride; d[i].uv_stride = j.uv_stride; d[i].dst_stride = j.dst_stride;
    d[i].width = j.width; d[i].height = j.height;
    const int64_t n_px = (int64_t)j.width * j.height;
    if (flat_src != FLAT_NONE && (j.width & 3) != 0 && j.height > 1 && n_px <= (1 << 28) &&
        j.dst_stride == min_dst_bpp * j.width &&
        (flat_src == 0 || j.src == NULL || j.src_stride == (flat_src < 0 ? j.width : flat_src * j.width)) &&
        (!flat_uv || j.u == NULL || j.uv_stride == j.width)) {
      d[i].width = (int32_t)n_px; d[i].height = 1;
      d[i].src_stride = (int32_t)(flat_src < 0 ? n_px : (flat_src > 0 ? flat_src * n_px : 0));
      d[i].uv_stride = (int32_t)n_px;
      d[i].dst_stride = (int32_t)(min_dst_bpp * n_px);
    }
    d[i].chunks_x = (uint32_t)((d[i].width + item_px - 1) / item_px);
    const uint32_t items = d[i].chunks_x * (uint32_t)d[i].height;
    if (i == 0) uni = items; else if (items != uni) uni = 0;
    pre[i] = (uint32_t)tot;
    tot += items;
    if (tot > 0x7fffffffull) { free(host); return B2_ERR_BAD_DIM; }
  }
  pre[n_jobs] = (uint32_t)tot;
  void* dev = NULL;
  const int st = b2_upload_desc(ctx, stream, host, jb + pb, &dev);
  free(host);
  if (st != B2_OK) return st;
  *dj = (const ConvJobDev*)dev; *dp = (const uint32_t*)((uint8_t*)dev + jb); *uniform = uni; *total = (uint32_t)tot;
  return B2_OK;
}

}  // namespace

#define B2_CONV_ENTRY_CHECKS          \
  B2_CHECK_CTX(ctx, stream);          \
  if (n_jobs < 0) return B2_ERR_INVALID_CONFIG; \
  if (n_jobs == 0) return B2_OK;      \
  if (jobs == NULL) return B2_ERR_NULL;

extern "C" int b2_vp8l_convert_from_bgra(b2_ctx* ctx, int stream, const b2_convert_job* jobs, int n_jobs, int mode, int flags) {
  B2_CONV_ENTRY_CHECKS
  bool premult;
  const int st = canonical(mode, &premult);
  if (st < 0) return B2_ERR_INVALID_CONFIG;
  for (int i = 0; i < n_jobs; ++i) if (jobs[i].src == NULL) return B2_ERR_NULL;
  const ConvJobDev* dj; const uint32_t* dp; uint32_t uni, tot;
  const int rc = prepare_conv(ctx, stream, jobs, n_jobs, false, bpp_of(st), kItemPx, &dj, &dp, &uni, &tot, -1, 0);
  if (rc != B2_OK) return rc;
  convert_from_bgra_kernel<<<b2_grid_for(ctx, 8, (tot + kWarps - 1) / kWarps), kThreads, 0, ctx->streams[stream]>>>(
      dj, dp, (uint32_t)n_jobs, uni, tot, st, premult ? 1 : 0, ((flags & B2_CONVERT_UPSTREAM_16BIT) || !ctx->swap16) ? 1 : 0);
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}

extern "C" int b2_alpha_multiply(b2_ctx* ctx, int stream, const b2_convert_job* jobs, int n_jobs, int kind, int flag) {
  B2_CONV_ENTRY_CHECKS
  if (kind < 0 || kind > 4) return B2_ERR_INVALID_CONFIG;
  const int bpp = (kind == B2_ALPHA_APPLY || kind == B2_ALPHA_MULT_ARGB || kind == B2_ALPHA_DISPATCH) ? 4 : (kind == B2_ALPHA_APPLY_4444 ? 2 : 1);
  const ConvJobDev* dj; const uint32_t* dp; uint32_t uni, tot;
  const int rc = prepare_conv(ctx, stream, jobs, n_jobs, kind == B2_ALPHA_MULT_ROW || kind == B2_ALPHA_DISPATCH, bpp, kItemPx, &dj, &dp, &uni, &tot,
                              0, (kind == B2_ALPHA_MULT_ROW || kind == B2_ALPHA_DISPATCH) ? 1 : 0);
  if (rc != B2_OK) return rc;
  alpha_multiply_kernel<<<b2_grid_for(ctx, 8, (tot + kWarps - 1) / kWarps), kThreads, 0, ctx->streams[stream]>>>(
      dj, dp, (uint32_t)n_jobs, uni, tot, kind, flag ? 1 : 0);
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}

extern "C" int b2_yuv_to_rgb_point(b2_ctx* ctx, int stream, const b2_convert_job* jobs, int n_jobs, int mode, int yuv444) {
  B2_CONV_ENTRY_CHECKS
  bool premult;
  const int st = canonical(mode, &premult);  // premultiplied modes alias the plain samplers (yuv.go:67-71)
  if (st < 0) return B2_ERR_INVALID_CONFIG;
  for (int i = 0; i < n_jobs; ++i) if (jobs[i].src == NULL || jobs[i].v == NULL) return B2_ERR_NULL;
  const ConvJobDev* dj; const uint32_t* dp; uint32_t uni, tot;
  // 4:4:4 is a per-pixel map (flat when the planes are packed); point-sampled 4:2:0 shares a chroma row between two rows
  const int rc = prepare_conv(ctx, stream, jobs, n_jobs, true, bpp_of(st), kItemPx, &dj, &dp, &uni, &tot, yuv444 ? 1 : FLAT_NONE, yuv444 ? 1 : 0);
  if (rc != B2_OK) return rc;
  const int grid = b2_grid_for(ctx, 8, (tot + kWarps - 1) / kWarps);
  const bool up16 = !ctx->swap16;
  if (yuv444) yuv_point_kernel<true><<<grid, kThreads, 0, ctx->streams[stream]>>>(dj, dp, (uint32_t)n_jobs, uni, tot, st, up16);
  else yuv_point_kernel<false><<<grid, kThreads, 0, ctx->streams[stream]>>>(dj, dp, (uint32_t)n_jobs, uni, tot, st, up16);
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}

extern "C" int b2_import_to_argb(b2_ctx* ctx, int stream, const b2_convert_job* jobs, int n_jobs, int order, int import_alpha) {
  B2_CONV_ENTRY_CHECKS
  if (order < B2_ORDER_RGBA || order > B2_ORDER_BGR) return B2_ERR_INVALID_CONFIG;
  const int step = (order == B2_ORDER_RGB || order == B2_ORDER_BGR) ? 3 : 4;
  if (import_alpha && step != 4) return B2_ERR_INVALID_CONFIG;   // assert(step == 4), picture_csp_enc.c.go:146
  for (int i = 0; i < n_jobs; ++i) {
    if (jobs[i].src == NULL) return B2_ERR_NULL;
    const int64_t st = jobs[i].src_stride < 0 ? -(int64_t)jobs[i].src_stride : jobs[i].src_stride;
    if (jobs[i].width > 0 && st < (int64_t)(import_alpha ? 4 : 3) * jobs[i].width) return B2_ERR_BAD_DIM;   // :132
    if (jobs[i].width > 0 && st < (int64_t)step * jobs[i].width) return B2_ERR_BAD_DIM;
  }
  const ConvJobDev* dj; const uint32_t* dp; uint32_t uni, tot;
  const int rc = prepare_conv(ctx, stream, jobs, n_jobs, false, 4, 128, &dj, &dp, &uni, &tot, step, 0);
  if (rc != B2_OK) return rc;
  import_argb_kernel<<<b2_grid_for(ctx, 8, (tot + kWarps - 1) / kWarps), kThreads, 0, ctx->streams[stream]>>>(
      dj, dp, (uint32_t)n_jobs, uni, tot, order, import_alpha ? 1 : 0);
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}
