// b2_tools.cu -- picture pre-processing that follows the import in the encoder
// (SURVEY.md 8f-3):
//   WebPCleanupTransparentArea   pkg/libwebp/enc/picture_tools_enc.c.go:110-183 (+ helpers :31-91)
//   WebPReplaceTransparentPixels pkg/libwebp/enc/picture_tools_enc.c.go:95-108
//   WebPBlendAlpha               pkg/libwebp/enc/picture_tools_enc.c.go:199-276
// The encoder runs the first one right after the import when !config.exact
// (pkg/libwebp/enc/webp_encoding.go:49-72), on the planes the import just wrote.
//
// All three work in place.  The only serial piece of the reference is the value a run of
// fully transparent 8x8 blocks is flattened to: the top-left sample of the run's FIRST
// block, carried along the block row by need_reset.  Here a warp owns a block row, a lane
// a block; the run start of every block comes from a ballot of the "transparent" flags
// (highest clear bit below the lane, plus a carry between 32-block chunks), and the run
// value is re-read from that block -- it is never changed by the flattening itself
// (the block is flattened to its own top-left value).
#include <stdlib.h>

#include "b2_common.cuh"

#ifndef B2_CLEANUP_CTAS
#define B2_CLEANUP_CTAS 3
#endif
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct PicJobDev {
  uint8_t* y;
  uint8_t* u;
  uint8_t* v;
  uint8_t* a;
  uint32_t* argb;
  int32_t y_stride, uv_stride, a_stride, argb_stride;
  int32_t width, height;
  uint32_t units_x;   // per-kernel: segments per row (flat kernels)
  uint32_t pad;
};

// run start (block index within the row) of a transparent block at position `lane` of a
// chunk whose first block is `base`; tmask = ballot of the transparent flags; carry = run
// start inherited from the previous chunk, or -1
__device__ __forceinline__ int run_start(uint32_t tmask, int lane, int base, int carry) {
  const uint32_t below_clear = ~tmask & ((1u << lane) - 1u);     // non-transparent blocks left of me
  if (below_clear != 0u) return base + (32 - __clz(below_clear));  // the block after the nearest one
  return (carry >= 0) ? carry : base;
}
__device__ __forceinline__ int carry_out(uint32_t tmask, int base, int carry, int nblocks) {
  // state after the last block of the chunk (nblocks = blocks of the chunk that exist)
  const uint32_t valid = (nblocks >= 32) ? 0xffffffffu : ((1u << nblocks) - 1u);
  const uint32_t clear = ~tmask & valid;
  if (nblocks < 32 || !((tmask >> 31) & 1u)) return -1;           // row ends, or last block not transparent
  if (clear != 0u) return base + (32 - __clz(clear));
  return (carry >= 0) ? carry : base;
}

// ---- WebPCleanupTransparentArea, YUVA form ------------------------------------------
// unit = one band of 8 rows (the last band may be shorter: SmoothenBlock only, :173-181)
__global__ void __launch_bounds__(kThreads, B2_CLEANUP_CTAS)
cleanup_yuva_kernel(const PicJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix, uint32_t n_jobs,
                    uint32_t uniform, uint32_t total) {
  const int lane = threadIdx.x & 31;
  for (uint32_t unit = blockIdx.x * kWarps + (threadIdx.x >> 5); unit < total; unit += gridDim.x * kWarps) {
    uint32_t band;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, unit, &band);
    const PicJobDev* __restrict__ j = jobs + img;
    const int width = j->width, height = j->height;
    const int y0 = (int)band * 8;
    const int bh = min(8, height - y0);
    const bool full_band = (bh == 8);
    const int ys = j->y_stride, as = j->a_stride, uvs = j->uv_stride;
    uint8_t* yb = j->y + (ptrdiff_t)y0 * ys;
    const uint8_t* ab = j->a + (ptrdiff_t)y0 * as;
    uint8_t* ub = j->u + (ptrdiff_t)(y0 >> 1) * uvs;
    uint8_t* vb = j->v + (ptrdiff_t)(y0 >> 1) * uvs;
    // 64-bit accesses when every row of Y and A starts 8-byte aligned, 32-bit U/V likewise
    const bool al8 = (((uintptr_t)j->y | (uintptr_t)j->a | (uintptr_t)(uint32_t)ys | (uintptr_t)(uint32_t)as) & 7) == 0;
    const bool al4 = (((uintptr_t)j->u | (uintptr_t)j->v | (uintptr_t)(uint32_t)uvs) & 3) == 0;
    const int nblocks = (width + 7) >> 3;   // the last one may be narrower
    int carry = -1;
    for (int base = 0; base < nblocks; base += 32) {
      const int bx = base + lane;
      const int x0 = bx * 8;
      const int bw = min(8, width - x0);     // <= 0: no block for this lane
      uint32_t a_lo[8], a_hi[8];
      int count = 0, sum = 0;
      const int n_blk = max(bw, 0) * bh;
      if (full_band && al8 && (base + 32) * 8 <= width) {
        // Whole chunk of whole 8x8 blocks (warp-uniform): straight-line code, so the eight alpha loads of
        // a block are in flight together, and so are its eight luma loads (the general code below
        // decides row by row and waits for every load on its own: 16 round trips per chunk).  The
        // luma stays in registers between the sum and the rewrite.
        const uint8_t* ap = ab + x0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const uint2 av = *reinterpret_cast<const uint2*>(ap + (ptrdiff_t)r * as);
          a_lo[r] = av.x; a_hi[r] = av.y;
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) count += (__popc(__vcmpne4(a_lo[r], 0u)) + __popc(__vcmpne4(a_hi[r], 0u))) >> 3;
        const bool partial = count > 0 && count < 64;
        if (__any_sync(0xffffffffu, partial)) {
          uint8_t* yp = yb + x0;
          uint32_t y_lo[8], y_hi[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            y_lo[r] = y_hi[r] = 0u;
            if (partial) {
              const uint2 yv = *reinterpret_cast<const uint2*>(yp + (ptrdiff_t)r * ys);
              y_lo[r] = yv.x; y_hi[r] = yv.y;
            }
          }
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const uint32_t m_lo = __vcmpne4(a_lo[r], 0u), m_hi = __vcmpne4(a_hi[r], 0u);
            sum += (int)__dp4a(y_lo[r] & m_lo, 0x01010101u, 0u) + (int)__dp4a(y_hi[r] & m_hi, 0x01010101u, 0u);
          }
          if (partial) {
            const uint32_t avg4 = ((uint32_t)(sum / count) & 0xffu) * 0x01010101u;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const uint32_t m_lo = __vcmpne4(a_lo[r], 0u), m_hi = __vcmpne4(a_hi[r], 0u);
              *reinterpret_cast<uint2*>(yp + (ptrdiff_t)r * ys) =
                  make_uint2((y_lo[r] & m_lo) | (avg4 & ~m_lo), (y_hi[r] & m_hi) | (avg4 & ~m_hi));
            }
          }
        }
      } else {
      auto load_luma = [&](int r, uint32_t* lo, uint32_t* hi) {
        const uint8_t* yp = yb + (ptrdiff_t)r * ys + x0;
        *lo = *hi = 0u;
        if (bw == 8 && al8) {
          const uint2 yv = *reinterpret_cast<const uint2*>(yp);
          *lo = yv.x; *hi = yv.y;
        } else {
          for (int k = 0; k < bw; ++k) {
            const uint32_t yv = yp[k];
            if (k < 4) *lo |= yv << (8 * k); else *hi |= yv << (8 * (k - 4));
          }
        }
      };
      if (bw > 0) {
        // alpha first: opaque and fully transparent blocks (the usual case) never touch luma
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          a_lo[r] = a_hi[r] = 0u;
          if (r < bh) {
            const uint8_t* ap = ab + (ptrdiff_t)r * as + x0;
            if (bw == 8 && al8) {
              const uint2 av = *reinterpret_cast<const uint2*>(ap);
              a_lo[r] = av.x; a_hi[r] = av.y;
            } else {
              for (int k = 0; k < bw; ++k) {
                const uint32_t av = ap[k];
                if (k < 4) a_lo[r] |= av << (8 * k); else a_hi[r] |= av << (8 * (k - 4));
              }
            }
            // pixels with alpha != 0 (SmoothenBlock :66-76); bytes past bw are zero and are
            // not part of n_blk
            count += (__popc(__vcmpne4(a_lo[r], 0u)) + __popc(__vcmpne4(a_hi[r], 0u))) >> 3;
          }
        }
        if (count > 0 && count < n_blk) {   // partially transparent: sum the luma of the visible pixels
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            if (r < bh) {
              uint32_t ylo, yhi;
              load_luma(r, &ylo, &yhi);
              const uint32_t m_lo = __vcmpne4(a_lo[r], 0u), m_hi = __vcmpne4(a_hi[r], 0u);
              sum += (int)__dp4a(ylo & m_lo, 0x01010101u, 0u) + (int)__dp4a(yhi & m_hi, 0x01010101u, 0u);
            }
          }
        }
      }
      if (bw > 0 && count > 0 && count < n_blk) {   // partially transparent: transparent pixels take the average luma
        const uint32_t avg = (uint32_t)(sum / count) & 0xffu;
        const uint32_t avg4 = avg * 0x01010101u;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          if (r < bh) {
            const uint32_t m_lo = __vcmpne4(a_lo[r], 0u), m_hi = __vcmpne4(a_hi[r], 0u);
            uint32_t ylo, yhi;
            load_luma(r, &ylo, &yhi);   // second read of the block's luma: an L1 hit
            const uint32_t n_lo = (ylo & m_lo) | (avg4 & ~m_lo), n_hi = (yhi & m_hi) | (avg4 & ~m_hi);
            uint8_t* yp = yb + (ptrdiff_t)r * ys + x0;
            if (bw == 8 && al8) {
              *reinterpret_cast<uint2*>(yp) = make_uint2(n_lo, n_hi);
            } else {
              for (int k = 0; k < bw; ++k) yp[k] = (uint8_t)((k < 4 ? n_lo >> (8 * k) : n_hi >> (8 * (k - 4))));
            }
          }
        }
      }
      }
      // fully transparent 8x8 blocks of a full band are flattened to their run's first values
      const bool transp = full_band && bw == 8 && count == 0;
      const uint32_t tmask = __ballot_sync(0xffffffffu, transp);
      if (transp) {
        const int sx = run_start(tmask, lane, base, carry) * 8;
        const uint32_t yv = yb[sx] * 0x01010101u;
        const uint32_t uv = ub[sx >> 1] * 0x01010101u, vv = vb[sx >> 1] * 0x01010101u;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          uint8_t* yp = yb + (ptrdiff_t)r * ys + x0;
          if (al8) *reinterpret_cast<uint2*>(yp) = make_uint2(yv, yv);
          else for (int k = 0; k < 8; ++k) yp[k] = (uint8_t)yv;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          uint8_t* up = ub + (ptrdiff_t)r * uvs + (x0 >> 1);
          uint8_t* vp = vb + (ptrdiff_t)r * uvs + (x0 >> 1);
          if (al4) { *reinterpret_cast<uint32_t*>(up) = uv; *reinterpret_cast<uint32_t*>(vp) = vv; }
          else for (int k = 0; k < 4; ++k) { up[k] = (uint8_t)uv; vp[k] = (uint8_t)vv; }
        }
      }
      carry = carry_out(tmask, base, carry, min(32, nblocks - base));
      // the run value of the next chunk is read from blocks this chunk may have flattened:
      // the value there is unchanged, but make the order explicit
      __syncwarp();
    }
  }
}

// ---- WebPCleanupTransparentArea, ARGB form -------------------------------------------
// unit = one row of whole 8x8 blocks (left-overs on the right / bottom are ignored, :114)
__global__ void __launch_bounds__(kThreads, 4)
cleanup_argb_kernel(const PicJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix, uint32_t n_jobs,
                    uint32_t uniform, uint32_t total) {
  const int lane = threadIdx.x & 31;
  for (uint32_t unit = blockIdx.x * kWarps + (threadIdx.x >> 5); unit < total; unit += gridDim.x * kWarps) {
    uint32_t band;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, unit, &band);
    const PicJobDev* __restrict__ j = jobs + img;
    const int stride = j->argb_stride;
    uint32_t* rb = j->argb + (ptrdiff_t)band * 8 * stride;
    const bool al16 = (((uintptr_t)j->argb & 15) == 0) && ((stride & 3) == 0);
    const int nblocks = j->width >> 3;
    int carry = -1;
    for (int base = 0; base < nblocks; base += 32) {
      const int bx = base + lane;
      const bool have = bx < nblocks;
      uint32_t acc = 0u;
      if (have) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const uint32_t* p = rb + (ptrdiff_t)r * stride + bx * 8;
          if (al16) {
            const uint4 q0 = *reinterpret_cast<const uint4*>(p), q1 = *reinterpret_cast<const uint4*>(p + 4);
            acc |= q0.x | q0.y | q0.z | q0.w | q1.x | q1.y | q1.z | q1.w;
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) acc |= p[k];
          }
        }
      }
      const bool transp = have && (acc & 0xff000000u) == 0u;   // IsTransparentARGBArea :31-42
      const uint32_t tmask = __ballot_sync(0xffffffffu, transp);
      if (transp) {
        const uint32_t v = rb[run_start(tmask, lane, base, carry) * 8];
        const uint4 q = make_uint4(v, v, v, v);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          uint32_t* p = rb + (ptrdiff_t)r * stride + bx * 8;
          if (al16) { *reinterpret_cast<uint4*>(p) = q; *reinterpret_cast<uint4*>(p + 4) = q; }
          else {
#pragma unroll
            for (int k = 0; k < 8; ++k) p[k] = v;
          }
        }
      }
      carry = carry_out(tmask, base, carry, min(32, nblocks - base));
      __syncwarp();
    }
  }
}

// ---- WebPBlendAlpha ---------------------------------------------------------------------
// picture_tools_enc.c.go:191-194
__device__ __forceinline__ uint32_t blend8(uint32_t v0, uint32_t v1, uint32_t alpha) {
  return ((v0 * (255u - alpha) + v1 * alpha) * 0x101u + 256u) >> 16;
}
__device__ __forceinline__ uint32_t blend10(uint32_t v0, uint32_t v1, uint32_t alpha) {
  return ((v0 * (1020u - alpha) + v1 * alpha) * 0x101u + 1024u) >> 18;
}
// YUVA form (:208-252): unit = (row pair, 256-pixel segment); a lane owns 4 chroma samples
// and their 2x2 luma / alpha samples (64-bit accesses on Y and A, 32-bit on U and V when
// the rows are aligned; per sample otherwise and on the ragged right edge).  Luma of both
// rows, chroma once (the reference blends chroma on even rows with the alpha of rows y
// and y+1), then alpha <- 0xff.
__device__ __forceinline__ void blend_sample(const PicJobDev* __restrict__ j, int rp, int cx, bool has_row1,
                                             uint32_t Y0, uint32_t U0, uint32_t V0) {
  const int width = j->width;
  const int x = 2 * cx;
  if (x >= width) return;
  const bool two = x + 1 < width;
  uint8_t* yp0 = j->y + (ptrdiff_t)(2 * rp) * j->y_stride + x;
  uint8_t* ap0 = j->a + (ptrdiff_t)(2 * rp) * j->a_stride + x;
  uint8_t* yp1 = yp0 + j->y_stride;
  uint8_t* ap1 = ap0 + j->a_stride;
  const uint32_t a00 = ap0[0], a01 = two ? ap0[1] : 0u;
  const uint32_t a10 = has_row1 ? ap1[0] : a00, a11 = has_row1 ? (two ? ap1[1] : 0u) : a01;   // a_ptr2 = a_ptr on the last row (:228)
  if (a00 < 0xffu) yp0[0] = (uint8_t)blend8(Y0, yp0[0], a00);
  if (two && a01 < 0xffu) yp0[1] = (uint8_t)blend8(Y0, yp0[1], a01);
  if (has_row1) {
    if (a10 < 0xffu) yp1[0] = (uint8_t)blend8(Y0, yp1[0], a10);
    if (two && a11 < 0xffu) yp1[1] = (uint8_t)blend8(Y0, yp1[1], a11);
  }
  const uint32_t asum = two ? (a00 + a01 + a10 + a11) : 2u * (a00 + a10);   // rightmost pixel of an odd width (:238-242)
  uint8_t* up = j->u + (ptrdiff_t)rp * j->uv_stride + cx;
  uint8_t* vp = j->v + (ptrdiff_t)rp * j->uv_stride + cx;
  *up = (uint8_t)blend10(U0, *up, asum);
  *vp = (uint8_t)blend10(V0, *vp, asum);
  ap0[0] = 0xff;
  if (two) ap0[1] = 0xff;
  if (has_row1) { ap1[0] = 0xff; if (two) ap1[1] = 0xff; }
}
// 4 luma samples of one word against their alphas
__device__ __forceinline__ uint32_t blend_luma4(uint32_t y4, uint32_t a4, uint32_t Y0) {
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t a = (a4 >> (8 * k)) & 0xff, y = (y4 >> (8 * k)) & 0xff;
    r |= ((a < 0xffu) ? blend8(Y0, y, a) : y) << (8 * k);
  }
  return r;
}
__global__ void __launch_bounds__(kThreads, 8)
blend_yuva_kernel(const PicJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix, uint32_t n_jobs,
                  uint32_t uniform, uint32_t total, uint32_t Y0, uint32_t U0, uint32_t V0) {
  const int lane = threadIdx.x & 31;
  for (uint32_t unit = blockIdx.x * kWarps + (threadIdx.x >> 5); unit < total; unit += gridDim.x * kWarps) {
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, unit, &local);
    const PicJobDev* __restrict__ j = jobs + img;
    const int width = j->width, height = j->height;
    const int rp = (int)(local / j->units_x), seg = (int)(local % j->units_x);
    const bool has_row1 = 2 * rp + 1 < height;
    const int cx0 = seg * 128 + lane * 4;     // first of this lane's 4 chroma columns
    const int x = 2 * cx0;
    if (x >= width) continue;
    const bool al = ((((uintptr_t)j->y | (uintptr_t)j->a | (uintptr_t)(uint32_t)j->y_stride | (uintptr_t)(uint32_t)j->a_stride) & 7) == 0) &&
                    ((((uintptr_t)j->u | (uintptr_t)j->v | (uintptr_t)(uint32_t)j->uv_stride) & 3) == 0);
    if (al && x + 8 <= width) {
      uint8_t* yp0 = j->y + (ptrdiff_t)(2 * rp) * j->y_stride + x;
      uint8_t* ap0 = j->a + (ptrdiff_t)(2 * rp) * j->a_stride + x;
      uint8_t* yp1 = has_row1 ? yp0 + j->y_stride : yp0;
      uint8_t* ap1 = has_row1 ? ap0 + j->a_stride : ap0;   // a_ptr2 = a_ptr on the last row (:228)
      uint8_t* up = j->u + (ptrdiff_t)rp * j->uv_stride + cx0;
      uint8_t* vp = j->v + (ptrdiff_t)rp * j->uv_stride + cx0;
      // every load of the unit before its first store (the planes may alias as far as the compiler knows,
      // so a store in between would split the loads into three dependent round trips)
      const uint2 a0 = *reinterpret_cast<const uint2*>(ap0), a1 = *reinterpret_cast<const uint2*>(ap1);
      const uint2 y0v = *reinterpret_cast<const uint2*>(yp0);
      const uint2 y1v = *reinterpret_cast<const uint2*>(yp1);
      const uint32_t u4 = *reinterpret_cast<const uint32_t*>(up), v4 = *reinterpret_cast<const uint32_t*>(vp);
      *reinterpret_cast<uint2*>(yp0) = make_uint2(blend_luma4(y0v.x, a0.x, Y0), blend_luma4(y0v.y, a0.y, Y0));
      if (has_row1) *reinterpret_cast<uint2*>(yp1) = make_uint2(blend_luma4(y1v.x, a1.x, Y0), blend_luma4(y1v.y, a1.y, Y0));
      // per chroma sample: alpha summed over its 2x2 block (pairs of bytes of both rows)
      uint32_t un = 0, vn = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t w0 = (k < 2) ? a0.x : a0.y, w1 = (k < 2) ? a1.x : a1.y;
        const int sh = 16 * (k & 1);
        const uint32_t asum = ((w0 >> sh) & 0xff) + ((w0 >> (sh + 8)) & 0xff) + ((w1 >> sh) & 0xff) + ((w1 >> (sh + 8)) & 0xff);
        un |= blend10(U0, (u4 >> (8 * k)) & 0xff, asum) << (8 * k);
        vn |= blend10(V0, (v4 >> (8 * k)) & 0xff, asum) << (8 * k);
      }
      *reinterpret_cast<uint32_t*>(up) = un;
      *reinterpret_cast<uint32_t*>(vp) = vn;
      *reinterpret_cast<uint2*>(ap0) = make_uint2(0xffffffffu, 0xffffffffu);
      if (has_row1) *reinterpret_cast<uint2*>(ap1) = make_uint2(0xffffffffu, 0xffffffffu);
    } else {
      for (int k = 0; k < 4; ++k) blend_sample(j, rp, cx0 + k, has_row1, Y0, U0, V0);
    }
  }
}

// ARGB form of WebPBlendAlpha (:253-276) and WebPReplaceTransparentPixels (:95-108):
// unit = (row, 1024-pixel segment), 4 pixels per lane and pass
template <bool BLEND>
__global__ void __launch_bounds__(kThreads, 4)
argb_pixel_kernel(const PicJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix, uint32_t n_jobs,
                  uint32_t uniform, uint32_t total, uint32_t color) {
  const int lane = threadIdx.x & 31;
  const uint32_t red = (color >> 16) & 0xff, green = (color >> 8) & 0xff, blue = color & 0xff;
  auto px = [&](uint32_t p) -> uint32_t {
    const uint32_t alpha = p >> 24;
    if (BLEND) {
      if (alpha == 0xffu) return p;
      if (alpha == 0u) return 0xff000000u | (color & 0xffffffu);
      return 0xff000000u | (blend8(red, (p >> 16) & 0xff, alpha) << 16) | (blend8(green, (p >> 8) & 0xff, alpha) << 8) |
             blend8(blue, p & 0xff, alpha);
    }
    return (alpha == 0u) ? (color & 0xffffffu) : p;
  };
  for (uint32_t unit = blockIdx.x * kWarps + (threadIdx.x >> 5); unit < total; unit += gridDim.x * kWarps) {
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, unit, &local);
    const PicJobDev* __restrict__ j = jobs + img;
    const int width = j->width;
    const int row = (int)(local / j->units_x), seg = (int)(local % j->units_x);
    uint32_t* rowp = j->argb + (ptrdiff_t)row * j->argb_stride;
    const bool al16 = (((uintptr_t)j->argb & 15) == 0) && ((j->argb_stride & 3) == 0);
#pragma unroll
    for (int pass = 0; pass < 8; ++pass) {
      const int x = seg * 1024 + pass * 128 + lane * 4;
      if (x >= width) break;
      if (al16 && x + 4 <= width) {
        uint4 q = *reinterpret_cast<const uint4*>(rowp + x);
        q.x = px(q.x); q.y = px(q.y); q.z = px(q.z); q.w = px(q.w);
        *reinterpret_cast<uint4*>(rowp + x) = q;
      } else {
        for (int k = 0; k < 4 && x + k < width; ++k) rowp[x + k] = px(rowp[x + k]);
      }
    }
  }
}

enum { PIC_CLEAN_YUVA, PIC_CLEAN_ARGB, PIC_BLEND_YUVA, PIC_ARGB_ROWS };

int prepare_pic(b2_ctx* ctx, int stream, const b2_picture_job* jobs, int n_jobs, int kind, const PicJobDev** dj,
                const uint32_t** dp, uint32_t* uniform, uint32_t* total) {
  const size_t jb = (sizeof(PicJobDev) * (size_t)n_jobs + 15) & ~(size_t)15;
  const size_t pb = sizeof(uint32_t) * ((size_t)n_jobs + 1);
  if (jb + pb > B2_DESC_RING_BYTES) return B2_ERR_INVALID_CONFIG;
  uint8_t* host = (uint8_t*)calloc(1, jb + pb);
  if (host == NULL) return B2_ERR_OOM;
  PicJobDev* d = (PicJobDev*)host;
  uint32_t* pre = (uint32_t*)(host + jb);
  uint64_t tot = 0;
  uint32_t uni = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const b2_picture_job& j = jobs[i];
    int st = B2_OK;
    const bool argb_form = (kind == PIC_CLEAN_ARGB || kind == PIC_ARGB_ROWS);
    if (argb_form ? (j.argb == NULL) : (j.y == NULL || j.u == NULL || j.v == NULL || j.a == NULL)) st = B2_ERR_NULL;
    else if (j.width <= 0 || j.height <= 0 || j.width > B2_MAX_DIMENSION || j.height > B2_MAX_DIMENSION) st = B2_ERR_BAD_DIM;
    else if (argb_form ? (j.argb_stride < j.width)
                       : (j.y_stride < j.width || j.a_stride < j.width || j.uv_stride < (j.width + 1) / 2)) st = B2_ERR_BAD_DIM;
    if (st != B2_OK) { free(host); return st; }
    d[i].y = j.y; d[i].u = j.u; d[i].v = j.v; d[i].a = j.a; d[i].argb = j.argb;
    d[i].y_stride = j.y_stride; d[i].uv_stride = j.uv_stride; d[i].a_stride = j.a_stride; d[i].argb_stride = j.argb_stride;
    d[i].width = j.width; d[i].height = j.height;
    uint32_t items = 0;
    switch (kind) {
      case PIC_CLEAN_YUVA: items = (uint32_t)((j.height + 7) / 8); break;
      case PIC_CLEAN_ARGB: items = (uint32_t)(j.height / 8); break;
      case PIC_BLEND_YUVA: d[i].units_x = (uint32_t)((j.width + 255) / 256); items = d[i].units_x * (uint32_t)((j.height + 1) / 2); break;
      default: d[i].units_x = (uint32_t)((j.width + 1023) / 1024); items = d[i].units_x * (uint32_t)j.height; break;
    }
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
  *dj = (const PicJobDev*)dev; *dp = (const uint32_t*)((uint8_t*)dev + jb); *uniform = uni; *total = (uint32_t)tot;
  return B2_OK;
}

// pkg/color/yuv/conversion.go:51-69 on the host, for the background colour of WebPBlendAlpha
int clip_uv(int uv, int rounding) {
  uv = (uv + rounding + (128 << 18)) >> 18;
  return ((uv & ~0xff) == 0) ? uv : (uv < 0) ? 0 : 255;
}

}  // namespace

#define B2_PIC_ENTRY_CHECKS          \
  B2_CHECK_CTX(ctx, stream);          \
  if (n_jobs < 0) return B2_ERR_INVALID_CONFIG; \
  if (n_jobs == 0) return B2_OK;      \
  if (jobs == NULL) return B2_ERR_NULL;

extern "C" int b2_picture_cleanup_transparent(b2_ctx* ctx, int stream, const b2_picture_job* jobs, int n_jobs, int use_argb) {
  B2_PIC_ENTRY_CHECKS
  const PicJobDev* dj; const uint32_t* dp; uint32_t uni, tot;
  const int rc = prepare_pic(ctx, stream, jobs, n_jobs, use_argb ? PIC_CLEAN_ARGB : PIC_CLEAN_YUVA, &dj, &dp, &uni, &tot);
  if (rc != B2_OK) return rc;
  if (tot == 0) return B2_OK;   // images lower than one block row: nothing to clean
  const int grid = b2_grid_for(ctx, 4, (tot + kWarps - 1) / kWarps);
  if (use_argb) cleanup_argb_kernel<<<grid, kThreads, 0, ctx->streams[stream]>>>(dj, dp, (uint32_t)n_jobs, uni, tot);
  else cleanup_yuva_kernel<<<b2_grid_for(ctx, B2_CLEANUP_CTAS, (tot + kWarps - 1) / kWarps), kThreads, 0, ctx->streams[stream]>>>(dj, dp, (uint32_t)n_jobs, uni, tot);
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}

extern "C" int b2_picture_blend_alpha(b2_ctx* ctx, int stream, const b2_picture_job* jobs, int n_jobs, int use_argb,
                                      uint32_t background_rgb) {
  B2_PIC_ENTRY_CHECKS
  const PicJobDev* dj; const uint32_t* dp; uint32_t uni, tot;
  const int rc = prepare_pic(ctx, stream, jobs, n_jobs, use_argb ? PIC_ARGB_ROWS : PIC_BLEND_YUVA, &dj, &dp, &uni, &tot);
  if (rc != B2_OK) return rc;
  const int grid = b2_grid_for(ctx, 8, (tot + kWarps - 1) / kWarps);
  if (use_argb) {
    argb_pixel_kernel<true><<<grid, kThreads, 0, ctx->streams[stream]>>>(dj, dp, (uint32_t)n_jobs, uni, tot, background_rgb & 0xffffffu);
  } else {
    const int r = (background_rgb >> 16) & 0xff, g = (background_rgb >> 8) & 0xff, b = background_rgb & 0xff;
    const int half = 1 << 15;
    const uint32_t Y0 = (uint32_t)((16839 * r + 33059 * g + 6420 * b + half + (16 << 16)) >> 16);
    // VP8RGBToU/V expect the values summed over four pixels (picture_tools_enc.c.go:212-214)
    const uint32_t U0 = (uint32_t)clip_uv(-9719 * 4 * r - 19081 * 4 * g + 28800 * 4 * b, 4 * half);
    const uint32_t V0 = (uint32_t)clip_uv(28800 * 4 * r - 24116 * 4 * g - 4684 * 4 * b, 4 * half);
    blend_yuva_kernel<<<grid, kThreads, 0, ctx->streams[stream]>>>(dj, dp, (uint32_t)n_jobs, uni, tot, Y0, U0, V0);
  }
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}

extern "C" int b2_picture_replace_transparent(b2_ctx* ctx, int stream, const b2_picture_job* jobs, int n_jobs, uint32_t color) {
  B2_PIC_ENTRY_CHECKS
  const PicJobDev* dj; const uint32_t* dp; uint32_t uni, tot;
  const int rc = prepare_pic(ctx, stream, jobs, n_jobs, PIC_ARGB_ROWS, &dj, &dp, &uni, &tot);
  if (rc != B2_OK) return rc;
  const int grid = b2_grid_for(ctx, 8, (tot + kWarps - 1) / kWarps);
  argb_pixel_kernel<false><<<grid, kThreads, 0, ctx->streams[stream]>>>(dj, dp, (uint32_t)n_jobs, uni, tot, color & 0xffffffu);
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}
