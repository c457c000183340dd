// b2_predict.cu -- VP8L predictor transform, both directions:
//   inverse  PredictorInverseTransform_C, PredictorAdd0..13_C, VP8LPredictor0..13_C
//            pkg/libwebp/dsp/lossless.go:31-333
//   forward  PredictorSub0..13_C with GIVEN modes, pkg/libwebp/dsp/lossless_enc.go:339-557
//            (driven as pkg/libwebp/enc/predictor_enc.go:666-712 does once modes are chosen)
//
// Pixels are uint32 0xAARRGGBB, images tightly packed.  data = the mode image:
// ceil(w/2^bits) x ceil(h/2^bits) words, mode = bits 8..11 (lossless.go:313).
//
// The forward direction reads only original pixels and is fully parallel.
//
// The inverse is serial along a row for the modes that use the left neighbour, and
// every row needs the finished row above it up to one pixel to the right (TR).  The
// parallelism that exists is a wavefront with a lag of two pixels per row, so:
//   * a warp owns a band of 32 rows, lane r = row r of the band, and walks the band in
//     SKEWED columns k = c + 2r: at step k every lane has its left / top-left / top /
//     top-right neighbours ready (TR comes from the lane above by one shuffle, the
//     rest are its own history), and all 32 lanes are busy on every step;
//   * ONE persistent launch: warps draw bands from a ticket counter in band-major order
//     (band 0 of every image, then band 1, ...).  Lane 0's row above is the last row of
//     the band above.  The fast bands (aligned images, tiles of 8 pixels or more) hand it
//     over as 8-byte messages {pixel, column + 1}, one relaxed 64-bit store per pixel and
//     relaxed 64-bit loads on the other side: the tag says the pixel next to it is final,
//     so there is no fence at all.  The general bands read it back from the output image
//     behind the band's progress word (st.release.gpu every kPublish columns,
//     ld.acquire.gpu poll).  Liveness does not rest on stream order or on residency: a
//     band only ever waits on the band above it, which holds a LOWER ticket, so it was
//     drawn earlier and is running (or done); by induction the lowest unfinished ticket
//     never waits on an undrawn band.
//   * the fused variants (b2_vp8l_predictor_inverse_chain) apply the cross-colour
//     inverse to the residual before the prediction is added and add-green to the
//     pixel on its way out; the neighbours a lane keeps are the un-add-greened pixels.
#include <stdlib.h>

#include <type_traits>

#include "b2_common.cuh"

namespace {

struct PredJobDev {
  const uint32_t* src;
  uint32_t* dst;
  const uint32_t* modes;
  int32_t width, height, bits;
  uint32_t tiles_per_row;  // mode image row length
  uint32_t bands;          // ceil(height / 32)
  uint32_t flag_off;       // first progress counter of this image (one per band)
  // fused chain (b2_vp8l_predictor_inverse_chain): cross-colour inverse on the way in, add-green and a
  // byte-order change on the way out
  const uint32_t* codes;   // cross-colour tile codes, or NULL
  int32_t cbits;
  uint32_t ctiles_per_row;
  uint32_t chain;          // CH_* bits
  uint32_t pad;
  unsigned long long* msg; // one row of `width` 8-byte messages per band: {column + 1, pixel} of the band's last row (fast bands)
};
enum { CH_CC = 1, CH_GREEN = 2, CH_RGBA = 4, CH_ARGB = 8 };

// ---- pixel arithmetic (lossless_common.go:110-121, lossless.go:31-149) -----------
__device__ __forceinline__ uint32_t add_pixels(uint32_t a, uint32_t b) {
  const uint32_t ag = (a & 0xff00ff00u) + (b & 0xff00ff00u);
  const uint32_t rb = (a & 0x00ff00ffu) + (b & 0x00ff00ffu);
  return (ag & 0xff00ff00u) | (rb & 0x00ff00ffu);
}
__device__ __forceinline__ uint32_t sub_pixels(uint32_t a, uint32_t b) {
  const uint32_t ag = 0x00ff00ffu + (a & 0xff00ff00u) - (b & 0xff00ff00u);
  const uint32_t rb = 0xff00ff00u + (a & 0x00ff00ffu) - (b & 0x00ff00ffu);
  return (ag & 0xff00ff00u) | (rb & 0x00ff00ffu);
}
__device__ __forceinline__ uint32_t avg2(uint32_t a, uint32_t b) {
  return (((a ^ b) & 0xfefefefeu) >> 1) + (a & b);
}
// Select (lossless.go:103-111): sum over channels of |b-c| - |a-c|; <= 0 picks a.
__device__ __forceinline__ uint32_t select_px(uint32_t a, uint32_t b, uint32_t c) {
  const int d = (int)__vsadu4(b, c) - (int)__vsadu4(a, c);  // per-byte sums of absolute differences
  return (d <= 0) ? a : b;
}
// ClampedAddSubtractFull (lossless.go:64-73): per channel clip255(a + b - c), two
// channels per 16-bit-lane word: a + b + (255 - c) needs no borrow, then one packed
// add-min-relu subtracts the 255 again and clamps to [0, 255].
__device__ __forceinline__ uint32_t clamp_add_sub_full(uint32_t a, uint32_t b, uint32_t c) {
  constexpr uint32_t m = 0x00ff00ffu;
  const uint32_t lo = (a & m) + (b & m) + ((c & m) ^ m);
  const uint32_t hi = ((a >> 8) & m) + ((b >> 8) & m) + (((c >> 8) & m) ^ m);
  const uint32_t rl = __viaddmin_s16x2_relu(lo, 0xff01ff01u, m);   // + (-255), min 255, max 0
  const uint32_t rh = __viaddmin_s16x2_relu(hi, 0xff01ff01u, m);
  return rl | (rh << 8);
}
// ClampedAddSubtractHalf (lossless.go:79-89): per channel clip255(ave + (ave - c) / 2)
// with C's truncating division.  With e = ave - c + 256 in [1, 511]:
// trunc((ave - c) / 2) = ((e + 1 - (e >> 8)) >> 1) - 128.
__device__ __forceinline__ uint32_t clamp_add_sub_half(uint32_t a, uint32_t b, uint32_t c) {
  constexpr uint32_t m = 0x00ff00ffu;
  const uint32_t ave = avg2(a, b);
  uint32_t r[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const uint32_t av = (ave >> (8 * k)) & m, cv = (c >> (8 * k)) & m;
    const uint32_t e = av + (cv ^ m) + 0x00010001u;                 // ave - c + 256 per lane
    const uint32_t h = ((e + 0x00010001u - ((e >> 8) & 0x00010001u)) >> 1) & 0x01ff01ffu;
    r[k] = __viaddmin_s16x2_relu(av + h, 0xff80ff80u, m);           // - 128, clamp to [0, 255]
  }
  return r[0] | (r[1] << 8);
}
// VP8LPredictor0..13_C (lossless.go:113-149); modes 14, 15 alias 0 (table of 16).
__device__ __forceinline__ uint32_t predict(uint32_t mode, uint32_t l, uint32_t tl, uint32_t t, uint32_t tr) {
  switch (mode) {
    case 1: return l;
    case 2: return t;
    case 3: return tr;
    case 4: return tl;
    case 5: return avg2(avg2(l, tr), t);
    case 6: return avg2(l, tl);
    case 7: return avg2(l, t);
    case 8: return avg2(tl, t);
    case 9: return avg2(t, tr);
    case 10: return avg2(avg2(l, tl), avg2(t, tr));
    case 11: return select_px(t, l, tl);
    case 12: return clamp_add_sub_full(l, t, tl);
    case 13: return clamp_add_sub_half(l, t, tl);
    default: return 0xff000000u;
  }
}

// predict() with the mode known at compile time, and the 4 pixels of a lane
template <int MODE>
__device__ __forceinline__ uint32_t predict_c(uint32_t l, uint32_t tl, uint32_t t, uint32_t tr) {
  if (MODE == 1) return l;
  if (MODE == 2) return t;
  if (MODE == 3) return tr;
  if (MODE == 4) return tl;
  if (MODE == 5) return avg2(avg2(l, tr), t);
  if (MODE == 6) return avg2(l, tl);
  if (MODE == 7) return avg2(l, t);
  if (MODE == 8) return avg2(tl, t);
  if (MODE == 9) return avg2(t, tr);
  if (MODE == 10) return avg2(avg2(l, tl), avg2(t, tr));
  if (MODE == 11) return select_px(t, l, tl);
  if (MODE == 12) return clamp_add_sub_full(l, t, tl);
  if (MODE == 13) return clamp_add_sub_half(l, t, tl);
  return 0xff000000u;
}
template <int MODE>
__device__ __forceinline__ void forward4(const uint32_t (&cur)[4], const uint32_t (&up)[6], uint32_t l, bool first_col,
                                         uint32_t (&o)[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t pred = predict_c<MODE>(l, up[k], up[k + 1], up[k + 2]);
    if (k == 0 && first_col) pred = up[1];   // first column = T (lossless_enc.go / predictor_enc.go:690)
    o[k] = sub_pixels(cur[k], pred);
    l = cur[k];
  }
}

// Rows y_first, y_first + 4, ... < y_end of one lane's 4-pixel column, all in one mode tile row (mode known at
// compile time); whole groups, 16-byte aligned rows.  NOT unrolled: with the row loop unrolled twice the fourteen
// instantiations no longer fit the instruction cache (0.51 of the peak instead of 0.72, profiles/r02c_predictor_forward.md);
// prefetch.global.L2 of the rows 1..8 passes ahead was measured at 0.57..0.46.
#ifndef B2_FWD_SEGMENTS
#define B2_FWD_SEGMENTS 1
#endif
#ifndef B2_FWD_SEG_UNROLL
#define B2_FWD_SEG_UNROLL 1
#endif
constexpr int kFwdSegUnroll = B2_FWD_SEG_UNROLL;
template <int MODE>
__device__ __forceinline__ void forward_rows(const uint32_t* __restrict__ a, uint32_t* __restrict__ dst, int w, int x,
                                             int y_first, int y_end, size_t step) {
  const uint32_t* ap = a + (size_t)y_first * w + x;
  uint32_t* dp = dst + (size_t)y_first * w + x;
#pragma unroll kFwdSegUnroll
  for (int y = y_first; y < y_end; y += 4, ap += step, dp += step) {
    const uint4 cv = b2_ldg_stream_v4(ap);
    const uint32_t cur[4] = {cv.x, cv.y, cv.z, cv.w};
    uint32_t o[4];
    if (y == 0) {  // first row = L, first pixel = black (lossless_enc.go / predictor_enc.go:683-689)
      uint32_t l = (x == 0) ? 0xff000000u : __ldg(ap - 1);
#pragma unroll
      for (int k = 0; k < 4; ++k) { o[k] = sub_pixels(cur[k], l); l = cur[k]; }
    } else {
      uint32_t up[6];
      const uint4 uv = __ldg(reinterpret_cast<const uint4*>(ap - w));
      up[1] = uv.x; up[2] = uv.y; up[3] = uv.z; up[4] = uv.w;
      // the neighbours a mode does not look at are not loaded
      constexpr bool kNeedL = MODE == 1 || MODE == 5 || MODE == 6 || MODE == 7 || MODE == 10 || MODE == 11 || MODE == 12 || MODE == 13;
      constexpr bool kNeedTL = MODE == 4 || MODE == 6 || MODE == 8 || MODE == 10 || MODE == 11 || MODE == 12 || MODE == 13;
      constexpr bool kNeedTR = MODE == 3 || MODE == 5 || MODE == 9 || MODE == 10;
      up[0] = (kNeedTL && x != 0) ? __ldg(ap - w - 1) : 0u;
      up[5] = kNeedTR ? __ldg(ap - w + 4) : 0u;   // upper[x + 1] of the last column is the first pixel of the current row
      const uint32_t l0 = (kNeedL && x != 0) ? __ldg(ap - 1) : 0u;
      forward4<MODE>(cur, up, l0, x == 0, o);
    }
    b2_stg_stream_v4(dp, make_uint4(o[0], o[1], o[2], o[3]));
  }
}

// ---- forward ----------------------------------------------------------------------
constexpr int kFwdThreads = 256;
constexpr int kFwdBlock = 256;  // item = a 256-pixel wide, kFwdRows tall block of one image: a warp per 32-pixel
#ifndef B2_FWD_ROWS
#define B2_FWD_ROWS 64
#endif
constexpr int kFwdRows = B2_FWD_ROWS;    // column, 4 rows per pass
// A CTA takes a 256-pixel wide block (1 KB of every row it touches), a warp a 32-pixel
// column of it, 4 rows per pass, a lane 4 consecutive pixels of one row: a warp-wide
// access is four whole 128-byte lines.  With the reference's usual tile
// sizes (bits >= 5) the block lies in ONE mode tile, so the predictor switch is uniform
// over the warp; with smaller tiles a lane's 4 aligned pixels still share a tile
// (bits >= 2).  Two 128-bit loads (current and upper row) plus three scalar neighbours,
// one 128-bit store.  Rows that are not 16-byte aligned take the same mapping with scalar
// accesses.
#ifndef B2_FWD_CTAS
#define B2_FWD_CTAS 8
#endif
#ifndef B2_FWD_UNROLL
#define B2_FWD_UNROLL 1
#endif
constexpr int kFwdUnroll = B2_FWD_UNROLL;
__global__ void __launch_bounds__(kFwdThreads, B2_FWD_CTAS)
predictor_forward_kernel(const PredJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix,
                         uint32_t n_jobs, uint32_t uniform, uint32_t total_items) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t item = blockIdx.x; item < total_items; item += gridDim.x) {
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, item, &local);
    const PredJobDev* __restrict__ j = jobs + img;
    const int w = j->width, h = j->height;
    const uint32_t bx = (uint32_t)(w + kFwdBlock - 1) / kFwdBlock;
    const int by0 = (int)(local / bx) * kFwdRows, bx0 = (int)(local % bx) * kFwdBlock;
    const int x = bx0 + warp * 32 + (lane & 7) * 4;
    if (x >= w) continue;
    const uint32_t* __restrict__ a = j->src;
    const uint32_t mode_shift = (uint32_t)j->bits;
    const uint32_t* __restrict__ mode_col = j->modes + (x >> mode_shift);
    const uint32_t tiles_per_row = j->tiles_per_row;
    uint32_t* __restrict__ dst = j->dst;
    const int y_end = min(h, by0 + kFwdRows);
    if (B2_FWD_SEGMENTS && ((w & 3) == 0) && (((((uintptr_t)a) | ((uintptr_t)j->dst)) & 15) == 0)) {
      // Whole 4-pixel groups on 16-byte aligned rows: the rows of the block are walked one mode-tile row at a
      // time, with the mode looked up once per segment and the dispatch OUTSIDE the row loop, so the loop body
      // is straight-line code for one predictor (no mode load, no branch tree and no address rebuild per row)
      // and a row's loads are not held up behind the previous row's dispatch.
      const int ry = lane >> 3;
      const size_t step = 4 * (size_t)w;
      for (int ys = by0; ys < y_end; ys = ((ys >> mode_shift) + 1) << mode_shift) {
        const int ye = min(y_end, ((ys >> mode_shift) + 1) << mode_shift);
        const uint32_t mode = (__ldg(mode_col + (size_t)((uint32_t)ys >> mode_shift) * tiles_per_row) >> 8) & 0xf;
        const int y_first = ys + ry;
        switch (mode) {
#define B2_FWD_CASE(M) case M: forward_rows<M>(a, dst, w, x, y_first, ye, step); break;
          B2_FWD_CASE(1) B2_FWD_CASE(2) B2_FWD_CASE(3) B2_FWD_CASE(4) B2_FWD_CASE(5) B2_FWD_CASE(6) B2_FWD_CASE(7)
          B2_FWD_CASE(8) B2_FWD_CASE(9) B2_FWD_CASE(10) B2_FWD_CASE(11) B2_FWD_CASE(12) B2_FWD_CASE(13)
#undef B2_FWD_CASE
          default: forward_rows<0>(a, dst, w, x, y_first, ye, step); break;
        }
      }
      continue;
    }
#pragma unroll kFwdUnroll
    for (int y = by0 + (lane >> 3); y < y_end; y += 4) {
    const size_t p0 = (size_t)y * w + x;
    const bool vec = ((w & 3) == 0) && (((((uintptr_t)a) | ((uintptr_t)j->dst)) & 15) == 0);
    const int n = min(4, w - x);
    uint32_t cur[4] = {0, 0, 0, 0}, o[4];
    if (vec) {
      const uint4 cv = b2_ldg_stream_v4(a + p0);
      cur[0] = cv.x; cur[1] = cv.y; cur[2] = cv.z; cur[3] = cv.w;
    } else {
      for (int k = 0; k < n; ++k) cur[k] = __ldg(a + p0 + k);
    }
    if (y == 0) {  // first row = L, first pixel = black (lossless_enc.go / predictor_enc.go:683-689)
      uint32_t l = (x == 0) ? 0xff000000u : __ldg(a + p0 - 1);
#pragma unroll
      for (int k = 0; k < 4; ++k) { o[k] = sub_pixels(cur[k], l); l = cur[k]; }
    } else {
      // upper[x + 1] of the last column is the first pixel of the current row (contiguous image)
      uint32_t up[6];
      up[0] = (x == 0) ? 0u : __ldg(a + p0 - w - 1);
      if (vec) {
        const uint4 uv = __ldg(reinterpret_cast<const uint4*>(a + p0 - w));
        up[1] = uv.x; up[2] = uv.y; up[3] = uv.z; up[4] = uv.w;
        up[5] = __ldg(a + p0 - w + 4);
      } else {
#pragma unroll
        for (int k = 0; k < 5; ++k) up[1 + k] = (k <= n) ? __ldg(a + p0 - w + k) : 0u;
      }
      const uint32_t mode = (__ldg(mode_col + (size_t)((uint32_t)y >> mode_shift) * tiles_per_row) >> 8) & 0xf;
      const uint32_t l0 = (x == 0) ? 0u : __ldg(a + p0 - 1);
      // one dispatch for the lane's 4 pixels: the mode is a compile-time constant inside
      switch (mode) {
#define B2_FWD_CASE(M) case M: forward4<M>(cur, up, l0, x == 0, o); break;
        B2_FWD_CASE(1) B2_FWD_CASE(2) B2_FWD_CASE(3) B2_FWD_CASE(4) B2_FWD_CASE(5) B2_FWD_CASE(6) B2_FWD_CASE(7)
        B2_FWD_CASE(8) B2_FWD_CASE(9) B2_FWD_CASE(10) B2_FWD_CASE(11) B2_FWD_CASE(12) B2_FWD_CASE(13)
#undef B2_FWD_CASE
        default: forward4<0>(cur, up, l0, x == 0, o); break;
      }
    }
    if (vec) {
      b2_stg_stream_v4(dst + p0, make_uint4(o[0], o[1], o[2], o[3]));
    } else {
      for (int k = 0; k < n; ++k) dst[p0 + k] = o[k];
    }
    }
  }
}

// ---- inverse ----------------------------------------------------------------------
// Persistent wavefront.  A warp takes a band of 32 rows (lane r = row r of the band) and
// walks it once, left to right, in skewed columns c = k - 2r.  TR comes from the lane above
// by one shuffle; lane 0's row above is the last row of the band above, read back from the
// output image behind a progress counter that the band above publishes (release) and this
// band polls (acquire) every 32 columns.  Bands are handed out through one ticket counter
// in band-major order (band 0 of every image, then band 1, ...): the band a warp waits on
// always holds a lower ticket, so it is already running and the wait cannot deadlock,
// whatever the number of resident warps.  The wavefront needs W + 2H serial steps per image
// (plus the 32-column publishing granularity per band); what fills the GPU is the ~W/100
// bands per image that are in flight at once, times the images of the batch.
//
// Memory traffic is kept to whole sectors although every lane walks its own row: a lane
// fetches 8 residuals at a time with two 16-byte cp.async, 8 steps ahead, into a private
// 16-pixel ring in shared memory indexed by column, and collects 8 outputs in a second
// private ring before writing them as one sector.  A lane's rings are 28 words apart:
// with the 2-column skew between lanes the per-step accesses are at most 2-way bank
// conflicted, the 128-bit flush reads are conflict free.
constexpr int kInvWarps = 4;
constexpr int kInvThreads = kInvWarps * 32;
constexpr int kInSlots = 16, kOutSlots = 8;
constexpr int kLaneWords = 28;   // 16 in + 8 out + 4 pad; 16-byte aligned, 28 l mod 32 has period 8
#ifndef B2_PRED_PUBLISH
#define B2_PRED_PUBLISH 16
#endif
constexpr int kPublish = B2_PRED_PUBLISH;
#ifndef B2_PRED_CHUNK
#define B2_PRED_CHUNK 16
#endif
#ifndef B2_PRED_FAST
#define B2_PRED_FAST 1
#endif
constexpr int kChunk = B2_PRED_CHUNK;   // columns of the row above fetched at a time (power of two <= 32)   // columns between two progress publications

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// ---- fused chain pieces (lossless.go:337-374) ----------------------------------------
__device__ __forceinline__ uint32_t add_green_px(uint32_t argb) {
  const uint32_t green = (argb >> 8) & 0xff;
  uint32_t rb = argb & 0x00ff00ffu;
  rb += (green << 16) | green;
  return (argb & 0xff00ff00u) | (rb & 0x00ff00ffu);
}
__device__ __forceinline__ uint32_t sub_green_px(uint32_t argb) {
  const uint32_t green = (argb >> 8) & 0xff;
  uint32_t rb = (argb & 0x00ff00ffu) | 0x01000100u;
  rb -= (green << 16) | green;
  return (argb & 0xff00ff00u) | (rb & 0x00ff00ffu);
}
__device__ __forceinline__ uint32_t cc_inverse_px(uint32_t code, uint32_t argb) {  // VP8LTransformColorInverse_C
  const int g2r = (int)(int8_t)(code & 0xff), g2b = (int)(int8_t)((code >> 8) & 0xff), r2b = (int)(int8_t)((code >> 16) & 0xff);
  const int green = (int)(int8_t)((argb >> 8) & 0xff);
  int new_red = (int)((argb >> 16) & 0xff);
  int new_blue = (int)(argb & 0xff);
  new_red = (new_red + ((g2r * green) >> 5)) & 0xff;
  new_blue += (g2b * green) >> 5;
  new_blue += (r2b * (int)(int8_t)new_red) >> 5;
  new_blue &= 0xff;
  return (argb & 0xff00ff00u) | ((uint32_t)new_red << 16) | (uint32_t)new_blue;
}
// What leaves the kernel for a reconstructed pixel o (uint32 0xAARRGGBB = bytes B,G,R,A), and its inverse
// (the row above a band is read back from the output image).  chain is warp-uniform.
__device__ __forceinline__ uint32_t chain_out(uint32_t chain, uint32_t o) {
  if (chain & CH_GREEN) o = add_green_px(o);
  if (chain & CH_RGBA) o = __byte_perm(o, 0u, 0x3012);        // B,G,R,A -> R,G,B,A (VP8LConvertBGRAToRGBA_C)
  else if (chain & CH_ARGB) o = __byte_perm(o, 0u, 0x0123);   // -> A,R,G,B (CopyOrSwap, lossless.go:598-614)
  return o;
}
__device__ __forceinline__ uint32_t chain_back(uint32_t chain, uint32_t v) {
  if (chain & CH_RGBA) v = __byte_perm(v, 0u, 0x3012);        // R,G,B,A -> B,G,R,A (the permutation is its own inverse)
  else if (chain & CH_ARGB) v = __byte_perm(v, 0u, 0x0123);
  if (chain & CH_GREEN) v = sub_green_px(v);
  return v;
}

// predict() without divergent branches: the serial chain of the wavefront is one step long,
// and a divergent switch (the lanes of a band sit in 2-3 different mode tiles) costs more
// than computing the candidates.  `need` = warp-wide union of the modes in use (bit m set
// when some lane is in mode m), so whole candidate groups are skipped by UNIFORM branches
// when no lane wants them.  Late candidates go last in the select chain.
__device__ __forceinline__ uint32_t predict_flat(uint32_t mode, uint32_t need, uint32_t l, uint32_t tl, uint32_t t, uint32_t tr) {
  uint32_t p = 0xff000000u;                              // modes 0, 14, 15
  p = (mode == 2) ? t : p;
  p = (mode == 4) ? tl : p;
  p = (mode == 3) ? tr : p;
  p = (mode == 1) ? l : p;
  if (need & 0x07e0u) {                                  // the averaging modes 5..10
    const uint32_t x1 = avg2(l, tr), x2 = avg2(l, tl), x3 = avg2(t, tr);
    const uint32_t a = (mode == 5) ? x1 : (mode == 7) ? l : (mode == 8) ? tl : x2;
    const uint32_t b = (mode == 10) ? x3 : t;
    const uint32_t pav = avg2(a, b);                     // modes 5, 7, 8, 10
    p = (mode == 9) ? x3 : p;
    p = (mode == 6) ? x2 : p;
    p = (mode == 5 || mode == 7 || mode == 8 || mode == 10) ? pav : p;
  }
  if (need & (1u << 11)) p = (mode == 11) ? select_px(t, l, tl) : p;
  if (need & (1u << 12)) p = (mode == 12) ? clamp_add_sub_full(l, t, tl) : p;
  if (need & (1u << 13)) p = (mode == 13) ? clamp_add_sub_half(l, t, tl) : p;
  return p;
}

// One band.  VEC: width % 4 == 0 and 16-byte aligned images, so a lane's 8-pixel groups
// are two aligned 128-bit accesses.  The step is written without divergent branches: what
// happens for only some lanes (fetching the next 8 residuals when c % 8 == 0 -- always an
// even step --, writing 8 finished outputs when c % 8 == 7 -- always an odd step --,
// switching mode tile) is predicated; the branches left are warp-uniform or rare (the
// chunk of the row above every 32 steps, publishing progress, unaligned images).
template <bool VEC>
__device__ __forceinline__ void inverse_band(const PredJobDev* __restrict__ jp, uint32_t i, uint32_t* __restrict__ flags,
                                             uint32_t* ring, int lane) {
  const int w = jp->width, h = jp->height, bits = jp->bits;
  const int row = (int)i * 32 + lane;
  const bool row_ok = row < h;
  const uint32_t* __restrict__ in_row = jp->src + (size_t)row * w;
  uint32_t* __restrict__ out_row = jp->dst + (size_t)row * w;
  const uint32_t* up_row = jp->dst + (size_t)(i * 32 - 1) * w;      // last row of the band above (i > 0)
  const uint32_t* __restrict__ mode_row = jp->modes + (size_t)(row >> bits) * jp->tiles_per_row;
  const uint32_t* flag_up = flags + jp->flag_off + i - 1;
  uint32_t* flag_me = flags + jp->flag_off + i;
  const bool publish = (lane == 31) && (i + 1 < jp->bands);
  const uint32_t chain = jp->chain;                                  // warp-uniform
  const int cbits = jp->cbits;
  const uint32_t* __restrict__ code_row = (chain & CH_CC) ? jp->codes + (size_t)(row >> cbits) * jp->ctiles_per_row : nullptr;
  const int ct_last = (int)jp->ctiles_per_row - 1;
  auto load_code = [&](int ct) -> uint32_t {
    return (row_ok && ct <= ct_last) ? __ldg(code_row + ct) : 0u;
  };
  uint32_t code = 0, code_nxt = (chain & CH_CC) ? load_code(0) : 0u;
  uint32_t avail = 0;            // columns of up_row known to be complete
  uint32_t upc = 0;              // lane q: up_row[k0 + 1 + q] of the current 32-column chunk
  uint32_t left = 0, tl = 0, t = 0, first_px = 0;   // left doubles as the value the lane below takes as TR
  const int tile_mask = (1 << bits) - 1;
  const int mt_last = (int)jp->tiles_per_row - 1;
  auto load_mode = [&](int mt) -> uint32_t {
    return (row_ok && mt <= mt_last) ? (__ldg(mode_row + mt) >> 8) & 0xf : 0u;
  };
  // The mode of a lane's current 8-column group; the next group's is fetched one group
  // ahead, both on the even steps where the lane's c % 8 == 0 (8 aligned columns share a
  // mode tile when bits >= 3; 4-pixel tiles re-read the mode every 4 columns instead).
  uint32_t mode = 0, mode_nxt = load_mode(0);
  uint32_t need = 0;             // warp-wide union of 1 << mode
  uint32_t* const in_l = ring + lane * kLaneWords;
  const uint32_t in_s = (uint32_t)__cvta_generic_to_shared(in_l);
  const uint32_t out_s = in_s + kInSlots * 4;
  const bool row0 = (row == 0);
  const int k_end = w + 62;

  auto step = [&](auto odd_tag, auto steady_tag, int k) {
    constexpr bool ODD = decltype(odd_tag)::value;
    constexpr bool STEADY = decltype(steady_tag)::value;   // every lane at 1 <= c <= w-2 of an existing row > 0
    const int c = k - 2 * lane;
    if (((k + 1) & (kChunk - 1)) == 0 && i > 0) {   // next chunk of the row above: columns k+1 .. k+kChunk (uniform)
      const uint32_t need_cols = (uint32_t)min(w, k + 1 + kChunk);
      while (avail < need_cols) {
        avail = ld_acquire_u32(flag_up);
        if (avail < need_cols) __nanosleep(64);
      }
      const int col = k + 1 + lane;
      upc = (lane < kChunk && col < w) ? __ldcg(up_row + col) : 0u;
      if (chain & ~(uint32_t)CH_CC) upc = chain_back(chain, upc);      // the image holds finished pixels
    }
    // TR = (row-1, c+1): the lane above finished it on the previous step
    uint32_t tr = __shfl_up_sync(0xffffffffu, left, 1);
    const uint32_t u = __shfl_sync(0xffffffffu, upc, (k + 1) & (kChunk - 1));
    tr = (lane == 0) ? u : tr;
    if (!ODD) {
      // lanes with c % 8 == 0 fetch columns c+8 .. c+15 into the other half of their ring.
      // Every lane commits one (possibly empty) group per even step; the group a lane
      // reads from now on was committed 4 even steps ago.
      const int cc = c + 8;
      const bool grp = (c & 7) == 0;
      const bool trig = grp && (STEADY || (row_ok && cc >= 0)) && cc < w;
      const uint32_t dst = in_s + (uint32_t)(cc & 8) * 4u;
      if (VEC) {
        if (trig) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(in_row + cc) : "memory");
        if (trig && cc + 4 < w) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + 16), "l"(in_row + cc + 4) : "memory");
      } else if (trig) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (cc + e < w) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst + 4 * e), "l"(in_row + cc + e) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 4;" ::: "memory");
      if (bits >= 3) {
        if (grp && c >= 0) { mode = mode_nxt; mode_nxt = load_mode((c + 8) >> bits); }
      } else if ((c & 3) == 0 && c >= 0) {
        mode = load_mode(c >> bits);
      }
      need = __reduce_or_sync(0xffffffffu, 1u << mode);
      if (chain & CH_CC) {   // the cross-colour tile of the group, fetched like the mode
        if (cbits >= 3) {
          if (grp && c >= 0) { code = code_nxt; code_nxt = load_code((c + 8) >> cbits); }
        } else if ((c & 3) == 0 && c >= 0) {
          code = load_code(c >> cbits);
        }
      }
    }
    const bool active = STEADY || (row_ok && (unsigned)c < (unsigned)w);
    uint32_t res;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(res) : "r"(in_s + (uint32_t)(c & (kInSlots - 1)) * 4u) : "memory");
    if (chain & CH_CC) res = cc_inverse_px(code, res);   // ApplyInverseTransforms order: cross-colour first (vp8l_dec.c.go:710-729)
    if (!STEADY) tr = (c == w - 1) ? first_px : tr;      // upper[x+1] of the last column wraps to the row start
    uint32_t pred = predict_flat(mode, need, left, tl, t, tr);
    if (!STEADY) {
      pred = (c == 0) ? t : pred;                  // lossless.go:312
      pred = row0 ? ((c == 0) ? 0xff000000u : left) : pred;   // :292-295
    }
    const uint32_t o = add_pixels(res, pred);
    // the neighbours a lane keeps (left, t, tl, first_px) are the predictor's own output; add-green and the
    // byte order only touch what is written out
    const uint32_t ov = (chain & ~(uint32_t)CH_CC) ? chain_out(chain, o) : o;
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(out_s + (uint32_t)(c & 7) * 4u), "r"(ov) : "memory");
    if (!STEADY) first_px = (active && c == 0) ? o : first_px;
    left = active ? o : left;
    tl = t;
    t = tr;
    // a finished group of 8, or the row's tail (an odd step when width % 4 == 0)
    const bool fl = active && ((c & 7) == 7 || (!STEADY && c == w - 1));
    if (VEC) {
      if (ODD) {
        uint4 q0, q1;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w) : "r"(out_s) : "memory");
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q1.x), "=r"(q1.y), "=r"(q1.z), "=r"(q1.w) : "r"(out_s + 16) : "memory");
        uint32_t* g = out_row + (c & ~7);
        if (fl) b2_stg_stream_v4(g, q0);
        if (fl && (c & 4)) b2_stg_stream_v4(g + 4, q1);
      }
    } else if (fl) {
      const int base = c & ~7;
      const uint32_t* out_l = in_l + kInSlots;
      for (int e = 0; e <= c - base; ++e) out_row[base + e] = *reinterpret_cast<const volatile uint32_t*>(out_l + e);
    }
    if ((ODD || !VEC) && publish && fl && (((c + 1) & (kPublish - 1)) == 0 || c == w - 1)) st_release_u32(flag_me, (uint32_t)(c + 1));
  };
  // steady steps: lane 31 at c >= 1 (k >= 63) and lane 0 at c <= w - 2 (k <= w - 2), full band below row 0
  const bool full_band = (i > 0) && ((int)i * 32 + 31 < h);
  const int ks0 = 64, ks1 = full_band ? ((w - 1) & ~1) : 0;   // [ks0, ks1), both even
  int k = -8;
  for (; k < k_end && (k < ks0 || ks1 <= ks0); k += 2) {
    step(std::false_type{}, std::false_type{}, k);
    step(std::true_type{}, std::false_type{}, k + 1);
  }
  for (; k < ks1; k += 2) {
    step(std::false_type{}, std::true_type{}, k);
    step(std::true_type{}, std::true_type{}, k + 1);
  }
  for (; k < k_end; k += 2) {
    step(std::false_type{}, std::false_type{}, k);
    step(std::true_type{}, std::false_type{}, k + 1);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// The same band, restructured so that nothing inside a step is lane-dependent control flow.  In
// inverse_band a lane fetches, flushes and switches tile when ITS column hits a multiple of 8, i.e. at a
// different step than its neighbours (the lanes are 2 columns apart): every step carries the compares,
// branches and convergence barriers of all those events (SASS: 136 of the 521 instructions of a step pair
// are ISETP / BRA / BSSY / BSYNC, ~100 more are moves and ring addressing).  Here time is cut into groups of
// 8 steps and everything periodic happens at group boundaries for ALL lanes at once: a lane's 8 columns of
// a group start at c0 = 8g - 2 lane (8-byte aligned instead of 32-byte aligned), its residuals arrive as
// four 8-byte cp.async two groups ahead, its 8 outputs stay in registers and leave as four 64-bit stores,
// the input ring is indexed by step (immediate offsets in the unrolled body), the tile of the NEXT group's
// last column is fetched one group ahead and switched in by a select when the column crosses a tile
// boundary (tiles are at least 8 wide: at most one boundary per group).  Needs the vector preconditions
// (width % 4 == 0, 16-byte aligned images) and tiles of 8 pixels or more; other jobs take inverse_band.
constexpr int kFastRing = 4 * 8;        // input ring: 4 groups of 8 residuals per lane
#ifndef B2_PRED_FETCH_AT
#define B2_PRED_FETCH_AT 4
#endif
constexpr int kFetchAt = B2_PRED_FETCH_AT;   // step of a group at which the next group's chunk of the row above is fetched
constexpr int kFastLaneWords = 34;      // 8-byte aligned slots; 34 l mod 32 = 2 l: two lanes per bank at most
// The last row of a band goes to the band below as 8-byte messages {pixel, column + 1}: a naturally aligned 64-bit
// access is single-copy atomic, so the tag says that the pixel next to it is the final one -- no fence on the
// writing side (a st.release costs the band a MEMBAR.GPU: every prior store acknowledged by L2), no acquire on
// the reading side, and the row is published pixel by pixel instead of every 16 columns.
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void inverse_band_fast(const PredJobDev* __restrict__ jp, uint32_t i, uint32_t* __restrict__ flags,
                                                  uint32_t* ring, int lane, const bool pairs) {
  const int w = jp->width, h = jp->height, bits = jp->bits;
  const int row = (int)i * 32 + lane;
  const bool row_ok = row < h;
  const uint32_t* __restrict__ in_row = jp->src + (size_t)row * w;
  uint32_t* __restrict__ out_row = jp->dst + (size_t)row * w;
  const uint32_t* __restrict__ mode_row = jp->modes + (size_t)(row >> bits) * jp->tiles_per_row;
  unsigned long long* const msg_me = jp->msg + (size_t)i * w;                        // this band's last row, for the band below
  const unsigned long long* const msg_up = jp->msg + (size_t)(i > 0 ? i - 1 : 0) * w;   // the last row of the band above (i > 0)
  const bool publish = (lane == 31) && (i + 1 < jp->bands);
  const uint32_t chain = jp->chain;                                  // warp-uniform
  const bool cc = (chain & CH_CC) != 0, finish = (chain & ~(uint32_t)CH_CC) != 0;
  const int cbits = jp->cbits;
  const uint32_t* __restrict__ code_row = cc ? jp->codes + (size_t)(row >> cbits) * jp->ctiles_per_row : nullptr;
  const int mt_last = (int)jp->tiles_per_row - 1, ct_last = (int)jp->ctiles_per_row - 1;
  const int tile_mask = (1 << bits) - 1, ctile_mask = (1 << cbits) - 1;
  // raw words: the mode sits in bits 8..11 and is cut out only when the word is needed, a group later, so the
  // load's latency hides behind 8 steps instead of stalling the shift that follows it
  auto load_mode_raw = [&](int col) -> uint32_t {
    const int mt = col >> bits;
    return (row_ok && col >= 0 && mt <= mt_last) ? __ldg(mode_row + mt) : 0u;
  };
  auto load_code = [&](int col) -> uint32_t {
    const int ct = col >> cbits;
    return (cc && row_ok && col >= 0 && ct <= ct_last) ? __ldg(code_row + ct) : 0u;
  };
  uint32_t* const ring_l = ring + lane * kFastLaneWords;
  const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring_l);
  const bool row0 = (row == 0);
  const bool full_band = (i > 0) && ((int)i * 32 + 31 < h);
  const int k_end = w + 62;
  // The row above (lane 0's TR / T / TL source), 8 columns at a time: group g uses columns 8g + 1 .. 8g + 8 (lane q
  // of upc holds column 8g + 1 + q), the next group's chunk is fetched kFetchAt steps into the group -- a few steps
  // cover an L2 round trip -- into upc_nx.  A chunk is 8 messages read by 8 lanes with relaxed 64-bit loads,
  // repeated until every tag says its pixel is the final one.  The messages carry the predictor's own output,
  // before add-green and the byte order, so nothing is undone here.  The band below thus starts
  // 62 (the skew) + 8 + (8 - kFetchAt) columns behind, not 62 + 16 + 8 + the publication period.
  uint32_t upc = 0, upc_nx = 0;
  auto fetch_chunk = [&](int first_col) -> uint32_t {   // columns first_col .. + 7 of the row above (uniform call)
    if (i == 0 || first_col >= w) return 0u;
    const int col = first_col + lane;
    const bool want = lane < 8 && col >= 0 && col < w;
    unsigned long long m = 0ull;
    for (;;) {
      if (want) m = ld_relaxed_u64(msg_up + col);
      const bool ok = !want || (uint32_t)(m >> 32) == (uint32_t)(col + 1);
      if (__all_sync(0xffffffffu, ok)) break;
      __nanosleep(32);
    }
    return want ? (uint32_t)m : 0u;
  };
  uint32_t left = 0, tl = 0, t = 0, first_px = 0;
  // mode / code of the current column, and of the tile that holds the LAST column of the current group
  uint32_t mode = 0, mode_e = (load_mode_raw(7 - 2 * lane) >> 8) & 0xf, code = 0, code_e = load_code(7 - 2 * lane);

  auto issue = [&](int g) {   // residuals of group g: columns 8 g - 2 lane .. + 7, four 8-byte copies
    const int cc0 = 8 * g - 2 * lane;
    const uint32_t dst = ring_s + (uint32_t)((g & 3) * 8) * 4u;
    if (pairs) {   // warp-uniform: even width on an 8-byte aligned image, every pair of columns is 8-byte aligned
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int col = cc0 + 2 * p;
        if (row_ok && (unsigned)col < (unsigned)w)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(dst + 8u * p), "l"(in_row + col) : "memory");
      }
    } else {       // odd widths: the rows alternate between 8- and 4-byte alignment, one pixel per copy
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int col = cc0 + p;
        if (row_ok && (unsigned)col < (unsigned)w)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst + 4u * p), "l"(in_row + col) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  auto group = [&](auto steady_tag, int g) {
    constexpr bool STEADY = decltype(steady_tag)::value;   // every lane at 1 <= c <= w - 2 of an existing row > 0, all 8 steps
    const int k0 = 8 * g, c0 = k0 - 2 * lane;
    // the tiles of the next group's last column, one group ahead
    const uint32_t mode_n_raw = load_mode_raw(c0 + 15), code_n = load_code(c0 + 15);
    const uint32_t need = __reduce_or_sync(0xffffffffu, (1u << mode) | (1u << mode_e));
    uint32_t res[8];
    {
      const uint32_t src = ring_s + (uint32_t)((g & 3) * 8) * 4u;
      asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(res[0]), "=r"(res[1]) : "r"(src) : "memory");
      asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(res[2]), "=r"(res[3]) : "r"(src + 8) : "memory");
      asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(res[4]), "=r"(res[5]) : "r"(src + 16) : "memory");
      asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(res[6]), "=r"(res[7]) : "r"(src + 24) : "memory");
    }
    uint32_t outv[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int k = k0 + s, c = c0 + s;
      if (s == kFetchAt) upc_nx = fetch_chunk(k0 + 9);     // uniform: the next group's columns
      uint32_t tr = __shfl_up_sync(0xffffffffu, left, 1);
      const uint32_t u = __shfl_sync(0xffffffffu, upc, s);
      tr = (lane == 0) ? u : tr;
      mode = ((c & tile_mask) == 0) ? mode_e : mode;     // a new tile starts at this column (at most once per group)
      uint32_t r = res[s];
      if (cc) {
        code = ((c & ctile_mask) == 0) ? code_e : code;
        r = cc_inverse_px(code, r);
      }
      if (!STEADY) tr = (c == w - 1) ? first_px : tr;    // upper[x+1] of the last column wraps to the row start
      uint32_t pred = predict_flat(mode, need, left, tl, t, tr);
      if (!STEADY) {
        pred = (c == 0) ? t : pred;                      // lossless.go:312
        pred = row0 ? ((c == 0) ? 0xff000000u : left) : pred;   // :292-295
      }
      const uint32_t o = add_pixels(r, pred);
      outv[s] = finish ? chain_out(chain, o) : o;
      const bool active = STEADY || (row_ok && (unsigned)c < (unsigned)w);
      if (publish && active) st_relaxed_u64(msg_me + c, ((unsigned long long)(uint32_t)(c + 1) << 32) | o);
      if (!STEADY) first_px = (active && c == 0) ? o : first_px;
      left = active ? o : left;
      tl = t;
      t = tr;
    }
    if (pairs) {
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int col = c0 + 2 * p;
        if (STEADY || (row_ok && (unsigned)col < (unsigned)w)) b2_stg_stream_v2(out_row + col, make_uint2(outv[2 * p], outv[2 * p + 1]));
      }
    } else {
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int col = c0 + p;
        if (STEADY || (row_ok && (unsigned)col < (unsigned)w)) b2_stg_stream_u32(out_row + col, outv[p]);
      }
    }
    mode_e = (mode_n_raw >> 8) & 0xf;
    code_e = code_n;
    upc = upc_nx;
  };

  upc = fetch_chunk(-7);     // group -1: column 0 is read at k = -1
#pragma unroll 1
  for (int g = -3; 8 * g < k_end; ++g) {
    issue(g + 2);
    if (g < -1) continue;
    asm volatile("cp.async.wait_group 2;" ::: "memory");
    const int k0 = 8 * g;
    if (full_band && k0 >= 64 && k0 + 7 <= w - 2) group(std::true_type{}, g);
    else group(std::false_type{}, g);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(kInvThreads, 8)
predictor_inverse_wave_kernel(const PredJobDev* __restrict__ jobs, const uint32_t* __restrict__ order,
                              uint32_t n_bands, uint32_t* __restrict__ ctl) {
  __shared__ __align__(16) uint32_t rings[kInvWarps][(kFastLaneWords > kLaneWords ? kFastLaneWords : kLaneWords) * 32];
  const int lane = threadIdx.x & 31;
  uint32_t* ring = rings[threadIdx.x >> 5];
  uint32_t* flags = ctl + 1;
  for (;;) {
    uint32_t ticket = 0;
    if (lane == 0) ticket = atomicAdd(ctl, 1u);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if (ticket >= n_bands) break;
    const uint32_t e = __ldg(order + ticket);    // image << 10 | band
    const PredJobDev* __restrict__ jp = jobs + (e >> 10);
    const bool vec = ((jp->width & 3) == 0) && (((((uintptr_t)jp->src) | ((uintptr_t)jp->dst)) & 15) == 0);
    // the group-structured band moves pixels in pairs (8-byte copies and stores at even columns) when the width is
    // even and the image 8-byte aligned, one by one otherwise: eight more copy / store instructions per 8-step group
    // of ~960.  (On the generic band a width of 4094 ran 1.9x and a width of 4095 3x slower than 4096.)
    const bool pairs = ((jp->width & 1) == 0) && (((((uintptr_t)jp->src) | ((uintptr_t)jp->dst)) & 7) == 0);
    const bool fast = jp->bits >= 3 && (!(jp->chain & CH_CC) || jp->cbits >= 3) && B2_PRED_FAST;
    if (fast) inverse_band_fast(jp, e & 1023u, flags, ring, lane, pairs);
    else if (vec) inverse_band<true>(jp, e & 1023u, flags, ring, lane);
    else inverse_band<false>(jp, e & 1023u, flags, ring, lane);
    __syncwarp();
  }
}

// Uploads the job table and the flat-item prefix (forward), plus -- with_wave -- the
// band-major ticket order and a zeroed control block {ticket counter, one progress
// counter per band} for the inverse.
// what prepare_pred needs of a job, from either ABI struct
struct PredSrc {
  const uint32_t* src; uint32_t* dst; const uint32_t* modes; const uint32_t* codes;
  int32_t width, height, bits, cbits; uint32_t chain;
};

struct PredPlan {
  const PredJobDev* jobs;
  const uint32_t* prefix;
  const uint32_t* order;
  uint32_t* ctl;
  uint32_t uniform, total_items, total_bands;
};
int prepare_pred(b2_ctx* ctx, int stream, const PredSrc* jobs, int n_jobs, bool with_wave, PredPlan* plan) {
  uint64_t bands_all = 0;
  uint32_t max_bands = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const PredSrc& j = jobs[i];
    if (j.src == NULL || j.dst == NULL || j.modes == NULL) return B2_ERR_NULL;
    if (j.width <= 0 || j.height <= 0 || j.width > B2_MAX_DIMENSION || j.height > B2_MAX_DIMENSION) return B2_ERR_BAD_DIM;
    if (j.bits < 2 || j.bits > 9 || j.src == j.dst) return B2_ERR_INVALID_CONFIG;  // vp8l_dec.c.go:1243
    if ((j.chain & CH_CC) && (j.codes == NULL || j.cbits < 2 || j.cbits > 9)) return B2_ERR_INVALID_CONFIG;
    const uint32_t nb = (uint32_t)((j.height + 31) / 32);
    bands_all += nb;
    if (nb > max_bands) max_bands = nb;
  }
  if (with_wave && (n_jobs > (1 << 22) || max_bands > 1024)) return B2_ERR_INVALID_CONFIG;  // order entry = image << 10 | band
  // message rows of the inverse wavefront (one row of `width` 8-byte messages per band), cleared on every launch:
  // a tag left by an earlier launch must not read as "final"
  unsigned long long* msg_base = NULL;
  if (with_wave) {
    size_t entries = 0;
    for (int i = 0; i < n_jobs; ++i) entries += (size_t)((jobs[i].height + 31) / 32) * (size_t)jobs[i].width;
    void* sp = NULL;
    const int st = b2_stream_scratch(ctx, stream, entries * sizeof(unsigned long long), &sp);
    if (st != B2_OK) return st;
    msg_base = (unsigned long long*)sp;
    cudaError_t e = cudaMemsetAsync(sp, 0, entries * sizeof(unsigned long long), ctx->streams[stream]);
    if (e != cudaSuccess) return b2_fail_cuda(ctx, e, "cudaMemsetAsync(message rows)");
  }
  const size_t jb = (sizeof(PredJobDev) * (size_t)n_jobs + 15) & ~(size_t)15;
  const size_t pb = (sizeof(uint32_t) * ((size_t)n_jobs + 1) + 15) & ~(size_t)15;
  const size_t ob = with_wave ? (sizeof(uint32_t) * (size_t)bands_all + 15) & ~(size_t)15 : 0;
  const size_t cb = with_wave ? sizeof(uint32_t) * ((size_t)bands_all + 1) : 0;
  if (jb + pb + ob + cb > B2_DESC_RING_BYTES) return B2_ERR_INVALID_CONFIG;
  uint8_t* host = (uint8_t*)calloc(1, jb + pb + ob + cb);
  if (host == NULL) return B2_ERR_OOM;
  PredJobDev* dj = (PredJobDev*)host;
  uint32_t* pre = (uint32_t*)(host + jb);
  uint32_t* ord = (uint32_t*)(host + jb + pb);
  uint64_t tot = 0;
  uint32_t uni = 0, flag_off = 0;
  size_t msg_off = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const PredSrc& j = jobs[i];
    PredJobDev& d = dj[i];
    d.src = j.src; d.dst = j.dst; d.modes = j.modes;
    d.width = j.width; d.height = j.height; d.bits = j.bits;
    d.codes = j.codes; d.cbits = j.cbits; d.chain = j.chain; d.pad = 0;
    d.ctiles_per_row = (j.chain & CH_CC) ? (uint32_t)((j.width + (1 << j.cbits) - 1) >> j.cbits) : 0u;
    d.tiles_per_row = (uint32_t)((j.width + (1 << j.bits) - 1) >> j.bits);
    d.bands = (uint32_t)((j.height + 31) / 32);
    d.flag_off = flag_off;
    flag_off += d.bands;
    d.msg = msg_base ? msg_base + msg_off : NULL;
    msg_off += (size_t)d.bands * (size_t)j.width;
    const uint32_t items = (uint32_t)((j.width + kFwdBlock - 1) / kFwdBlock) * (uint32_t)((j.height + kFwdRows - 1) / kFwdRows);
    if (i == 0) uni = items; else if (items != uni) uni = 0;
    pre[i] = (uint32_t)tot;
    tot += items;
    if (tot > 0x7fffffffull) { free(host); return B2_ERR_BAD_DIM; }
  }
  pre[n_jobs] = (uint32_t)tot;
  if (with_wave) {
    size_t q = 0;
    for (uint32_t band = 0; band < max_bands; ++band)
      for (int i = 0; i < n_jobs; ++i)
        if (band < dj[i].bands) ord[q++] = ((uint32_t)i << 10) | band;
  }
  void* dev = NULL;
  const int st = b2_upload_desc(ctx, stream, host, jb + pb + ob + cb, &dev);
  free(host);
  if (st != B2_OK) return st;
  plan->jobs = (const PredJobDev*)dev;
  plan->prefix = (const uint32_t*)((uint8_t*)dev + jb);
  plan->order = (const uint32_t*)((uint8_t*)dev + jb + pb);
  plan->ctl = (uint32_t*)((uint8_t*)dev + jb + pb + ob);
  plan->uniform = uni; plan->total_items = (uint32_t)tot; plan->total_bands = (uint32_t)bands_all;
  return B2_OK;
}

PredSrc* plain_jobs(const b2_argb_job* jobs, int n_jobs) {
  PredSrc* ps = (PredSrc*)malloc(sizeof(PredSrc) * (size_t)n_jobs);
  if (ps == NULL) return NULL;
  for (int i = 0; i < n_jobs; ++i) {
    ps[i].src = jobs[i].src; ps[i].dst = jobs[i].dst; ps[i].modes = jobs[i].data; ps[i].codes = NULL;
    ps[i].width = jobs[i].width; ps[i].height = jobs[i].height; ps[i].bits = jobs[i].bits; ps[i].cbits = 0; ps[i].chain = 0;
  }
  return ps;
}

}  // namespace

extern "C" int b2_vp8l_predictor_forward(b2_ctx* ctx, int stream, const b2_argb_job* jobs, int n_jobs) {
  B2_CHECK_CTX(ctx, stream);
  if (n_jobs < 0) return B2_ERR_INVALID_CONFIG;
  if (n_jobs == 0) return B2_OK;
  if (jobs == NULL) return B2_ERR_NULL;
  PredPlan pl;
  PredSrc* ps = plain_jobs(jobs, n_jobs);
  if (ps == NULL) return B2_ERR_OOM;
  const int st = prepare_pred(ctx, stream, ps, n_jobs, false, &pl);
  free(ps);
  if (st != B2_OK) return st;
  const int grid = b2_grid_for(ctx, B2_FWD_CTAS, pl.total_items);
  predictor_forward_kernel<<<grid, kFwdThreads, 0, ctx->streams[stream]>>>(pl.jobs, pl.prefix, (uint32_t)n_jobs, pl.uniform, pl.total_items);
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}

extern "C" int b2_vp8l_predictor_inverse(b2_ctx* ctx, int stream, const b2_argb_job* jobs, int n_jobs) {
  B2_CHECK_CTX(ctx, stream);
  if (n_jobs < 0) return B2_ERR_INVALID_CONFIG;
  if (n_jobs == 0) return B2_OK;
  if (jobs == NULL) return B2_ERR_NULL;
  PredPlan pl;
  PredSrc* ps = plain_jobs(jobs, n_jobs);
  if (ps == NULL) return B2_ERR_OOM;
  const int st = prepare_pred(ctx, stream, ps, n_jobs, true, &pl);
  free(ps);
  if (st != B2_OK) return st;
  const int grid = b2_grid_for(ctx, 8, (pl.total_bands + kInvWarps - 1) / kInvWarps);
  predictor_inverse_wave_kernel<<<grid, kInvThreads, 0, ctx->streams[stream]>>>(pl.jobs, pl.order, pl.total_bands, pl.ctl);
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}

extern "C" int b2_vp8l_predictor_inverse_chain(b2_ctx* ctx, int stream, const b2_vp8l_chain_job* jobs, int n_jobs, int out_mode) {
  B2_CHECK_CTX(ctx, stream);
  if (n_jobs < 0) return B2_ERR_INVALID_CONFIG;
  if (n_jobs == 0) return B2_OK;
  if (jobs == NULL) return B2_ERR_NULL;
  uint32_t order_bits;
  switch (out_mode) {   // the 4-byte, not premultiplied modes: a byte permutation of the pixel, undone when a band reads the row above back
    case B2_MODE_BGRA: order_bits = 0; break;
    case B2_MODE_RGBA: order_bits = CH_RGBA; break;
    case B2_MODE_ARGB: order_bits = CH_ARGB; break;
    default: return B2_ERR_INVALID_CONFIG;
  }
  PredSrc* ps = (PredSrc*)malloc(sizeof(PredSrc) * (size_t)n_jobs);
  if (ps == NULL) return B2_ERR_OOM;
  for (int i = 0; i < n_jobs; ++i) {
    const b2_vp8l_chain_job& j = jobs[i];
    ps[i].src = j.src; ps[i].dst = j.dst; ps[i].modes = j.predictor_data; ps[i].codes = j.color_data;
    ps[i].width = j.width; ps[i].height = j.height; ps[i].bits = j.predictor_bits; ps[i].cbits = j.color_bits;
    ps[i].chain = (j.color_data ? CH_CC : 0u) | (j.add_green ? CH_GREEN : 0u) | order_bits;
  }
  PredPlan pl;
  const int st = prepare_pred(ctx, stream, ps, n_jobs, true, &pl);
  free(ps);
  if (st != B2_OK) return st;
  const int grid = b2_grid_for(ctx, 8, (pl.total_bands + kInvWarps - 1) / kInvWarps);
  predictor_inverse_wave_kernel<<<grid, kInvThreads, 0, ctx->streams[stream]>>>(pl.jobs, pl.order, pl.total_bands, pl.ctl);
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}
