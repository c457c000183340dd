// b2_filters.cu -- alpha-plane filters:
//   unfilters  HorizontalUnfilter_C / VerticalUnfilter_C / GradientUnfilter_C,
//              pkg/libwebp/dsp/filters.go:130-171, applied row after row against the previous
//              OUTPUT row as ALPHDecode / AlphaApplyFilter do (pkg/vp8/vp8l_dec.c.go:802-814)
//   filters    DoHorizontalFilter_C / DoVerticalFilter_C / DoGradientFilter_C, filters.go:24-112
// Planes are bytes; everything is arithmetic mod 256.
//
// The forward filters read only the unfiltered plane: one thread per 4 pixels.  The
// unfilters are recurrences, restated so that they parallelise:
//   horizontal  out[y][x] = sum_{r<=y} in[r][0] + sum_{1<=k<=x} in[y][k]: a warp per row --
//               the column-0 sum over the rows above, then a warp scan along the row;
//   vertical    out[y][x] = out[0][x] + sum_{1<=r<=y} in[r][x] with out[0] the horizontal
//               prefix of row 0: a thread per 4 columns walks down the rows (whole 128-byte
//               lines per warp), after summing row 0 up to its segment;
//   gradient    clip(left + top - topleft) is not linear: a wavefront.  A band of 32 rows runs inside
//               one warp (lane = row, one 4-pixel word of skew per row), the bands of a plane run as
//               a pipeline over the 16 warps of one CTA; with 4-byte aligned rows everything moves
//               through shared-memory tiles (gradient_band_tiled), unaligned rows keep the
//               byte-skewed loop that reads the row above a band back from the output.
// Planes whose rows are not 16-byte aligned (the reference's packed alpha plane at three widths in
// four) are re-pitched into padded scratch, filtered there by the aligned paths and copied back
// (run_with_repitch at the end of the file); the byte-wise paths remain for planes narrower than 64.
#include <stdlib.h>

#include <type_traits>

#include "b2_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct FiltJobDev {
  const uint8_t* src;
  uint8_t* dst;
  int32_t src_stride, dst_stride, width, height;
  int32_t filter;
  uint32_t units_x;
};

__device__ __forceinline__ int gradient_pred(int a, int b, int c) {   // GradientPredictor_C
  const int g = a + b - c;
  return min(max(g, 0), 255);
}

// ---- forward: unit = (row, 512-pixel segment), 16 pixels per lane ------------------
constexpr int kFwdLane = 16, kFwdUnit = 32 * kFwdLane;

// predictions for the 4 pixels of word c: left = the byte before them, t = the word above,
// up_left = the byte before that one
__device__ __forceinline__ uint32_t filter_pred4(int filter, bool top_row, uint32_t c, uint32_t left, uint32_t t, uint32_t up_left) {
  if (filter == 0) return 0u;
  if (top_row || filter == 1) return (c << 8) | left;      // top row and horizontal: from the left
  if (filter == 2) return t;
  // GradientPredictor_C on four pixels at once, two per register in 16-bit fields (even bytes, odd bytes):
  // L + T - TL + 511 = L + T + (TL ^ 0x1ff) never leaves its field, and min(relu(x - 511), 255) is one
  // VIADDMNMX.S16x2.RELU (the byte-by-byte form cost 15 instructions per pixel and kept the ALU pipe 83 % busy)
  const uint32_t l4 = (c << 8) | left, tl4 = (t << 8) | up_left;
  const uint32_t se = (l4 & 0x00ff00ffu) + (t & 0x00ff00ffu) + ((tl4 & 0x00ff00ffu) ^ 0x01ff01ffu);
  const uint32_t so = __byte_perm(l4, 0u, 0x4341) + __byte_perm(t, 0u, 0x4341) + (__byte_perm(tl4, 0u, 0x4341) ^ 0x01ff01ffu);
  const uint32_t pe = __viaddmin_s16x2_relu(se, 0xfe01fe01u, 0x00ff00ffu);
  const uint32_t po = __viaddmin_s16x2_relu(so, 0xfe01fe01u, 0x00ff00ffu);
  return __byte_perm(pe, po, 0x6240);
}

// unit = (group of kFwdRows rows, 512-pixel segment): the rows of a group are fetched together (one 128-bit load per
// lane and row when the rows are 16-byte aligned), a row's upper neighbour is the previous row's registers, and
// the job table, the division by the segment count and the edge tests are paid once per group, not per row.
#ifndef B2_FWD_ROWS
#define B2_FWD_ROWS 4
#endif
#ifndef B2_FWDF_CTAS
#define B2_FWDF_CTAS 4
#endif
constexpr int kFwdRows = B2_FWD_ROWS;
// A whole 512-pixel unit on 16-byte aligned rows, the filter known at compile time: the loads of the unit's rows (and
// of the row above it for the vertical / gradient filters) first, then straight-line code per row -- the generic body
// below tests the filter, the alignment and the lane's share of the row for every word (10 instructions per pixel,
// issue slots 65 % busy on a kernel that only moves 2 B/px).
#ifndef B2_FWDF_FAST
#define B2_FWDF_FAST 1
#endif
template <int F>
__device__ __forceinline__ void forward_unit_fast(const FiltJobDev& j, int y0, int ny, int x0, int lane) {
  const bool live = x0 < j.width;   // the row may end before this lane's 16 pixels (never inside them: width % 16 == 0)
  const uint8_t* p = j.src + (ptrdiff_t)y0 * j.src_stride + x0;
  uint8_t* d = j.dst + (ptrdiff_t)y0 * j.dst_stride + x0;
  uint4 c[kFwdRows + 1];
  c[0] = make_uint4(0u, 0u, 0u, 0u);
  if (F >= 2 && y0 > 0 && live) c[0] = __ldg(reinterpret_cast<const uint4*>(p - j.src_stride));
#pragma unroll
  for (int r = 0; r < kFwdRows; ++r) {
    c[r + 1] = make_uint4(0u, 0u, 0u, 0u);
    if (r < ny && live) c[r + 1] = __ldg(reinterpret_cast<const uint4*>(p + (ptrdiff_t)r * j.src_stride));
  }
#pragma unroll
  for (int r = 0; r < kFwdRows; ++r) {
    if (r >= ny) break;
    const int y = y0 + r;
    const uint32_t cw[4] = {c[r + 1].x, c[r + 1].y, c[r + 1].z, c[r + 1].w};
    const uint32_t t[4] = {c[r].x, c[r].y, c[r].z, c[r].w};
    uint32_t left = __shfl_up_sync(0xffffffffu, cw[3] >> 24, 1);
    uint32_t up_left = (F == 3) ? __shfl_up_sync(0xffffffffu, t[3] >> 24, 1) : 0u;
    if (lane == 0) {   // the byte before the unit comes from memory (zero at the row start)
      left = x0 > 0 ? p[(ptrdiff_t)r * j.src_stride - 1] : 0u;
      if (F == 3) up_left = (x0 > 0 && y > 0) ? p[(ptrdiff_t)(r - 1) * j.src_stride - 1] : 0u;
    }
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t pred = filter_pred4(F, y == 0, cw[q], left, t[q], up_left);
      if (q == 0 && x0 == 0) pred = (pred & 0xffffff00u) | (y == 0 ? 0u : (uint32_t)p[(ptrdiff_t)(r - 1) * j.src_stride]);   // leftmost pixel: as is / from above
      o[q] = __vsub4(cw[q], pred);
      left = cw[q] >> 24;
      if (F == 3) up_left = t[q] >> 24;
    }
    if (live) *reinterpret_cast<uint4*>(d + (ptrdiff_t)r * j.dst_stride) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void __launch_bounds__(kThreads, B2_FWDF_CTAS)
alpha_filter_kernel(const FiltJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix, uint32_t n_jobs,
                    uint32_t uniform, uint32_t total) {
  const int lane = threadIdx.x & 31;
  for (uint32_t unit = blockIdx.x * kWarps + (threadIdx.x >> 5); unit < total; unit += gridDim.x * kWarps) {
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, unit, &local);
    const FiltJobDev j = jobs[img];
    const uint32_t fwd_units = (uint32_t)((j.width + kFwdUnit - 1) / kFwdUnit);
    const int y0 = (int)(local / fwd_units) * kFwdRows, x0 = (int)(local % fwd_units) * kFwdUnit + lane * kFwdLane;
    const int ny = min(kFwdRows, j.height - y0);
    // whole words when the rows are 4-byte aligned: the left / upper-left neighbours of a
    // lane's first pixel are the last bytes of the previous lane's words (one shuffle each)
    const uintptr_t bits = ((uintptr_t)j.src | (uintptr_t)j.dst) | (uint32_t)j.src_stride | (uint32_t)j.dst_stride;
    const bool aligned = (bits & 3) == 0, al16 = (bits & 15) == 0;
    const bool vec = aligned && x0 + kFwdLane <= j.width;
    if (B2_FWDF_FAST && al16 && (j.width & 15) == 0 && j.filter >= 1 && j.filter <= 3) {   // warp-uniform
      if (j.filter == 1) forward_unit_fast<1>(j, y0, ny, x0, lane);
      else if (j.filter == 2) forward_unit_fast<2>(j, y0, ny, x0, lane);
      else forward_unit_fast<3>(j, y0, ny, x0, lane);
      continue;
    }
    // the group's rows and the row above it (vertical / gradient), all requested before the first is used
    uint32_t c[kFwdRows + 1][4];
    auto row_words = [&](int y, uint32_t (&w)[4]) {
      w[0] = w[1] = w[2] = w[3] = 0u;
      if (!vec || y < 0 || y >= j.height) return;
      const uint8_t* p = j.src + (ptrdiff_t)y * j.src_stride + x0;
      if (al16) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
        w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = __ldg((const uint32_t*)p + q);
      }
    };
    row_words((j.filter >= 2) ? y0 - 1 : -1, c[0]);
#pragma unroll
    for (int r = 0; r < kFwdRows; ++r) row_words(r < ny ? y0 + r : -1, c[r + 1]);
#pragma unroll
    for (int r = 0; r < kFwdRows; ++r) {
      if (r >= ny) break;
      const int y = y0 + r;
      const uint8_t* cur = j.src + (ptrdiff_t)y * j.src_stride;
      const uint8_t* up = cur - j.src_stride;
      uint8_t* dst = j.dst + (ptrdiff_t)y * j.dst_stride;
      const bool use_up = y > 0 && j.filter >= 2;
      const uint32_t (&cw)[4] = c[r + 1];
      const uint32_t (&t)[4] = c[r];              // zeros when the row above is not used
      uint32_t left = __shfl_up_sync(0xffffffffu, cw[3] >> 24, 1), up_left = __shfl_up_sync(0xffffffffu, t[3] >> 24, 1);   // every lane, vector or not
      if (vec) {
        if (lane == 0) {
          left = x0 > 0 ? cur[x0 - 1] : 0u;
          up_left = (x0 > 0 && use_up) ? up[x0 - 1] : 0u;
        }
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t pred = filter_pred4(j.filter, y == 0, cw[q], left, t[q], up_left);
          if (q == 0 && x0 == 0 && j.filter != 0) pred = (pred & 0xffffff00u) | (y == 0 ? 0u : (uint32_t)up[0]);   // leftmost pixel: as is / from above
          o[q] = __vsub4(cw[q], pred);
          left = cw[q] >> 24;
          up_left = t[q] >> 24;
        }
        if (al16) {
          *reinterpret_cast<uint4*>(dst + x0) = make_uint4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) *((uint32_t*)(dst + x0) + q) = o[q];
        }
        continue;
      }
      for (int x = x0; x < min(x0 + kFwdLane, j.width); ++x) {
        int pred;
        if (j.filter == 0) pred = 0;
        else if (x == 0) pred = (y == 0) ? 0 : up[0];                        // leftmost pixel: as is / from above
        else if (y == 0 || j.filter == 1) pred = cur[x - 1];                // top row and horizontal: from the left
        else if (j.filter == 2) pred = up[x];
        else pred = gradient_pred(cur[x - 1], up[x], up[x - 1]);
        dst[x] = (uint8_t)(cur[x] - pred);
      }
    }
  }
}

// bytewise inclusive prefix sum of the 4 bytes of w (mod 256 per byte)
__device__ __forceinline__ uint32_t prefix4(uint32_t w) {
  w = __vadd4(w, w << 8);
  return __vadd4(w, w << 16);
}

// ---- unfilter, NONE / HORIZONTAL: a warp per group of 8 rows -------------------------
// A row is a bytewise running sum started from prev[0] = the sum of column 0 over the rows
// above (each row's first output is the next row's start).  The warp sums column 0 down to
// its first row once and carries it from row to row; along a row every lane takes 16 bytes
// (prefix inside the lane with packed adds, one scan of the lane totals per 512 bytes).
constexpr int kRowsPerWarp = 8;
#ifndef B2_ROWS_AHEAD
#define B2_ROWS_AHEAD 4
#endif
#ifndef B2_ROWS_CTAS
#define B2_ROWS_CTAS 4
#endif
// The fast path of alpha_unfilter_rows_kernel (16-byte aligned rows, width a multiple of 16), one instantiation per
// filter so that the loop holds ONE load site: with the copy and the scan in one loop the compiler merges their loads
// through a spare register and copies it into the ring at the end of the item -- i.e. waits for it there.
template <int FILTER>
__device__ __forceinline__ void unfilter_rows_fast(const FiltJobDev& j, int y0, int y1, uint32_t col0, int lane, uint4* ring) {
      // Whole 16-byte pieces everywhere: the (row, 512-byte chunk) items of the group as one list, fetched
      // kAhead items ahead by 16-byte cp.async into a ring in shared memory (a lane reads back what it copied: no
      // barrier) -- the scan of a chunk waits for nothing but its own carry.  One chunk at a time, the warp had a
      // single 512-byte load in flight (0.42 of the measured peak); a ring of REGISTERS does not survive the
      // compiler, which parks the new load in a spare register and copies it into the ring at the end of the item,
      // i.e. waits for it there (ncu: 44 % of the warp time on that copy).
      constexpr int kAhead = B2_ROWS_AHEAD;
      const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
      const int cpr = (j.width + 511) >> 9, n_items = (y1 - y0) * cpr;
      int pr = y0, pc = 0;                         // the item the next fetch takes
      auto fetch = [&](int slot) {
        const int x = pc * 512 + lane * 16;
        if (pr < y1 && x < j.width)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(ring_s + (uint32_t)slot * 512u), "l"(j.src + (ptrdiff_t)pr * j.src_stride + x) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (++pc == cpr) { pc = 0; ++pr; }
      };
#pragma unroll
      for (int q = 0; q < kAhead; ++q) fetch(q);
      int cr = y0, cc = 0;
      uint32_t carry = 0;
      for (int i0 = 0; i0 < n_items; i0 += kAhead) {
#pragma unroll
        for (int q = 0; q < kAhead; ++q) {
          if (i0 + q < n_items) {                  // warp-uniform
            asm volatile("cp.async.wait_group %0;" :: "n"(kAhead - 1) : "memory");
            const uint4 v = ring[q * 32];
            if (cc == 0) {                         // a new row: the value left of it, then its own first sample joins the column sum
              carry = col0 & 0xffu;
              col0 += __shfl_sync(0xffffffffu, v.x & 0xffu, 0);
            }
            const int x = cc * 512 + lane * 16;
            uint8_t* const dp = j.dst + (ptrdiff_t)cr * j.dst_stride + x;
            if (FILTER != 1) {                     // filter NONE: a copy
              if (x < j.width) b2_stg_stream_v4(dp, v);
              fetch(q);
            } else {
              // The running sums of the lane's 16 bytes in 16-bit fields (no packed byte adds, which cost five
              // instructions each): for a word [b0 b1 b2 b3], e = [b0, b2], o = [b1, b3], x = e + o, the even outputs
              // are e + (x << 16) = [b0, b0+b1+b2] and the odd ones x * 0x10001 = [b0+b1, b0+b1+b2+b3]; what precedes
              // the word (the words before it in the lane, the lanes before it, the chunks before it) is one
              // multiply-add per register, and the low bytes of the fields are the outputs mod 256.
              const uint32_t w[4] = {v.x, v.y, v.z, v.w};
              uint32_t pe[4], po[4], off[4];
              uint32_t run = 0;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t e = w[k] & 0x00ff00ffu, o = __byte_perm(w[k], 0u, 0x4341);
                const uint32_t xs = e + o;
                pe[k] = e + (xs << 16);
                po[k] = xs * 0x00010001u;
                off[k] = run;
                run += po[k] >> 16;
              }
              fetch(q);                            // the slot's words are in registers: refill it
              const uint32_t tot = run;            // lane total, then an exclusive scan over the lanes
              uint32_t scan = tot;
#pragma unroll
              for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, scan, o);
                if (lane >= o) scan += t;
              }
              const uint32_t before = (scan - tot + carry) & 0xffu;
              uint32_t ow[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t sk = (off[k] + before) * 0x00010001u;
                ow[k] = __byte_perm(pe[k] + sk, po[k] + sk, 0x6240);
              }
              carry = (__shfl_sync(0xffffffffu, scan, 31) + carry) & 0xffu;
              if (x < j.width) b2_stg_stream_v4(dp, make_uint4(ow[0], ow[1], ow[2], ow[3]));
            }
            if (++cc == cpr) { cc = 0; ++cr; }
          }
        }
      }
}

__global__ void __launch_bounds__(kThreads, B2_ROWS_CTAS)
alpha_unfilter_rows_kernel(const FiltJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix, uint32_t n_jobs,
                           uint32_t uniform, uint32_t total) {
  __shared__ uint4 ring_sm[kWarps][B2_ROWS_AHEAD][32];   // the fast path's prefetch ring, 512 bytes per warp and stage
  const int lane = threadIdx.x & 31;
  for (uint32_t unit = blockIdx.x * kWarps + (threadIdx.x >> 5); unit < total; unit += gridDim.x * kWarps) {
    uint32_t group;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, unit, &group);
    const FiltJobDev j = jobs[img];
    if (j.filter > 1) continue;                  // vertical / gradient rows are done by their own kernels
    const int y0 = (int)group * kRowsPerWarp, y1 = min(y0 + kRowsPerWarp, j.height);
    uint32_t col0 = 0;                           // sum of column 0 over the rows above y
    if (j.filter == 1 && y0 > 0) {
      uint32_t s = 0;
      int r = lane;
      for (; r + 7 * 32 < y0; r += 8 * 32) {       // eight independent loads at a time: one per iteration makes this walk
        uint32_t t[8];                              // (y0 / 32 round trips to L2) as long as the group's own 8 rows
#pragma unroll
        for (int k = 0; k < 8; ++k) t[k] = __ldg(j.src + (ptrdiff_t)(r + 32 * k) * j.src_stride);
#pragma unroll
        for (int k = 0; k < 8; ++k) s += t[k];
      }
      for (; r < y0; r += 32) s += __ldg(j.src + (ptrdiff_t)r * j.src_stride);
#pragma unroll
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      col0 = s;
    }
    const uintptr_t bits = (uintptr_t)j.src | (uintptr_t)j.dst | (uintptr_t)(uint32_t)j.src_stride | (uintptr_t)(uint32_t)j.dst_stride;
    const bool al4 = (bits & 3) == 0, al16 = (bits & 15) == 0;
    if (al16 && (j.width & 15) == 0) {
      if (j.filter == 1) unfilter_rows_fast<1>(j, y0, y1, col0, lane, &ring_sm[threadIdx.x >> 5][0][lane]);
      else unfilter_rows_fast<0>(j, y0, y1, col0, lane, &ring_sm[threadIdx.x >> 5][0][lane]);
      continue;
    }
    for (int y = y0; y < y1; ++y) {
      const uint8_t* in = j.src + (ptrdiff_t)y * j.src_stride;
      uint8_t* out = j.dst + (ptrdiff_t)y * j.dst_stride;
      uint32_t carry = col0 & 0xffu;             // value left of the current chunk (pred)
      col0 += in[0];
      for (int x0 = 0; x0 < j.width; x0 += 512) {
        const int x = x0 + lane * 16;
        const bool vec16 = al16 && x + 16 <= j.width;
        uint32_t w[4] = {0, 0, 0, 0};
        if (vec16) {
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(in + x));
          w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int xq = x + 4 * q;
            if (al4 && xq + 4 <= j.width) w[q] = __ldg(reinterpret_cast<const uint32_t*>(in + xq));
            else {
#pragma unroll
              for (int k = 0; k < 4; ++k) if (xq + k < j.width) w[q] |= (uint32_t)in[xq + k] << (8 * k);
            }
          }
        }
        if (j.filter == 1) {
          uint32_t p[4];
          p[0] = prefix4(w[0]);                  // within the lane
#pragma unroll
          for (int q = 1; q < 4; ++q) p[q] = __vadd4(prefix4(w[q]), (p[q - 1] >> 24) * 0x01010101u);
          const uint32_t tot = p[3] >> 24;       // lane total, then an exclusive scan over the lanes
          uint32_t scan = tot;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, scan, o);
            if (lane >= o) scan += t;
          }
          const uint32_t before = ((scan - tot + carry) & 0xffu) * 0x01010101u;
#pragma unroll
          for (int q = 0; q < 4; ++q) w[q] = __vadd4(p[q], before);
          carry = (__shfl_sync(0xffffffffu, scan, 31) + carry) & 0xffu;
        }
        if (vec16) {
          *reinterpret_cast<uint4*>(out + x) = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int xq = x + 4 * q;
            if (al4 && xq + 4 <= j.width) *reinterpret_cast<uint32_t*>(out + xq) = w[q];
            else {
#pragma unroll
              for (int k = 0; k < 4; ++k) if (xq + k < j.width) out[xq + k] = (uint8_t)(w[q] >> (8 * k));
            }
          }
        }
      }
    }
  }
}

#ifndef B2_VERT_ROWS
#define B2_VERT_ROWS 32
#endif
#ifndef B2_VERT_DOUBLE
#define B2_VERT_DOUBLE 1
#endif
// ---- unfilter, VERTICAL: unit = 128-column segment of an image -------------------------
#ifndef B2_VERT_CTAS
#define B2_VERT_CTAS 2
#endif
__global__ void __launch_bounds__(kThreads, B2_VERT_CTAS)
alpha_unfilter_vertical_kernel(const FiltJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix, uint32_t n_jobs,
                               uint32_t uniform, uint32_t total) {
  const int lane = threadIdx.x & 31;
  for (uint32_t unit = blockIdx.x * kWarps + (threadIdx.x >> 5); unit < total; unit += gridDim.x * kWarps) {
    uint32_t seg;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, unit, &seg);
    const FiltJobDev j = jobs[img];
    if (j.filter != 2) continue;
    const int x = (int)seg * 128 + lane * 4;
    // row 0: horizontal prefix; the sum of everything left of this segment first
    uint32_t s = 0;
    for (int k = lane; k < (int)seg * 128; k += 32) s += j.src[k];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) if (x + k < j.width) w |= (uint32_t)j.src[x + k] << (8 * k);
    const uint32_t p = prefix4(w);
    const uint32_t tot = p >> 24;
    uint32_t scan = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, scan, o);
      if (lane >= o) scan += t;
    }
    uint32_t acc = __vadd4(p, ((scan - tot + s) & 0xffu) * 0x01010101u);
    const bool al4 = ((((uintptr_t)j.src | (uintptr_t)j.dst | (uintptr_t)(uint32_t)j.src_stride | (uintptr_t)(uint32_t)j.dst_stride) & 3) == 0) &&
                     x + 4 <= j.width;
    if (al4) {
      // B2_VERT_ROWS rows of loads in flight per thread, and the next batch requested before this one is consumed: the
      // walk down a column is a chain of adds, the loads must neither wait for one another nor for the stores.
      // 128 x 1080p planes give an SM only ~13 warps of this kernel, so the bytes in flight have to come from depth:
      // 16 rows, one batch at a time 0.52 of the measured peak; 16 + 16 0.64; 32 + 32 (two CTAs per SM) 0.74
      constexpr int kInFlight = B2_VERT_ROWS;
      *reinterpret_cast<uint32_t*>(j.dst + x) = acc;
      auto fetch = [&](int y0, uint32_t (&v)[kInFlight]) {
#pragma unroll
        for (int q = 0; q < kInFlight; ++q)
          v[q] = (y0 + q < j.height) ? __ldg(reinterpret_cast<const uint32_t*>(j.src + (ptrdiff_t)(y0 + q) * j.src_stride + x)) : 0u;
      };
      uint32_t v[kInFlight], nv[kInFlight];
      fetch(1, v);
      for (int y0 = 1; y0 < j.height; y0 += kInFlight) {
#if B2_VERT_DOUBLE
        fetch(y0 + kInFlight, nv);
#endif
#pragma unroll
        for (int q = 0; q < kInFlight; ++q) {
          if (y0 + q < j.height) {
            acc = __vadd4(acc, v[q]);
            *reinterpret_cast<uint32_t*>(j.dst + (ptrdiff_t)(y0 + q) * j.dst_stride + x) = acc;
          }
        }
#if B2_VERT_DOUBLE
#pragma unroll
        for (int q = 0; q < kInFlight; ++q) v[q] = nv[q];
#else
        fetch(y0 + kInFlight, v);
#endif
      }
    } else {
      for (int y = 0; y < j.height; ++y) {
        const uint8_t* in = j.src + (ptrdiff_t)y * j.src_stride + x;
        uint8_t* out = j.dst + (ptrdiff_t)y * j.dst_stride + x;
        if (y > 0) {
          uint32_t v = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) if (x + k < j.width) v |= (uint32_t)in[k] << (8 * k);
          acc = __vadd4(acc, v);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) if (x + k < j.width) out[k] = (uint8_t)(acc >> (8 * k));
      }
    }
  }
}

// ---- gradient band on aligned rows, through shared-memory tiles -----------------------------
// The wavefront puts lane l on row l of the band, so a lane-indexed global access touches 32
// different rows: one sector per lane, 32 L1 wavefronts per load or store instruction for 128
// useful bytes.  Here global memory is only touched in row-contiguous pieces: time is cut into
// blocks of 8 quad steps; block b's inputs are the parallelogram {row r, words 8b - r .. 8b - r + 7},
// fetched two blocks ahead by eight 4-byte cp.async per lane (lane -> row 4i + lane/8, word lane%8:
// 32 contiguous bytes per row, four rows per instruction) into a [32][9]-word tile; the step reads
// tile[lane][t] (stride 9: conflict free), writes its output word to a second tile, and the block's
// outputs leave the same way the inputs came.
//
// The step itself is latency bound (ncu: a working warp issues on 20 % of its cycles, the rest is
// the dependent chain), so NOTHING but the recurrence sits on the chain:
//  * out = (in + clip(left + top - topleft)) & 0xff (GradientUnfilter_C, filters.go:150-171); what a
//    lane hands to the lane below is not its output word but the four differences d_k = top_k -
//    top_{k-1} that lane needs (differences of its own consecutive outputs), each by its own shuffle
//    issued as soon as the pixel is done, three quarters of a step before it is used.  A pixel is then
//    relu(min(left + d, 255)) -- ONE instruction whose operand is the shuffle's result itself --, an add
//    (the input byte's shift folded in) and a mask: three dependent instructions per pixel;
//  * the shuffle is a rotation: lane 0 reads lane 31, and lane 31 sends the differences of the row
//    above the band instead of its own, so no lane selects anything on the receiving side;
//  * the row above the band never goes through global memory: the band above leaves the words of its
//    last row in a 512-word ring in shared memory (its warp's), the band converts nine of them to
//    differences at the start of a block (lanes 0..8) and lane 31 reads them back a step at a time.
// Progress words count finished blocks: (band << 12) + blocks.  A band needs block b + 4 of the band
// above finished before its block b; a band never runs more than 60 blocks ahead of the band below (the
// ring), and a warp starts a new band only when the band below its previous one is done with the ring.
// 15 bands x 60 blocks cover the longest row (516 blocks), so the three waits cannot form a cycle.
// No special cases in the arithmetic: row 0 has d = 0 (pred = left: the horizontal rule of the first
// row), column 0 starts with left = 0 and d0 = top_0 (pred = clip(top) = top: the reference's prev_line[0]).
#ifndef B2_GRAD_WARPS
#define B2_GRAD_WARPS 16
#endif
constexpr int kGradWarps = B2_GRAD_WARPS;
#ifndef B2_GRAD_CTAS
#define B2_GRAD_CTAS 2
#endif
constexpr int kGradCtasPerSm = B2_GRAD_CTAS;
#ifndef B2_GRAD_SLEEP
#define B2_GRAD_SLEEP 200
#endif
constexpr int kGS = 9;                       // tile row stride, words
constexpr int kGTile = 32 * kGS;             // one tile: 32 rows x 8 words (+1 pad)
constexpr int kGRing = 3;                    // input tiles in flight
constexpr int kGHand = 512;                  // words of the band's last row kept for the band below
constexpr int kGGap = 60;                    // blocks a band may run ahead of the band below
constexpr int kGXSlot = 33 * 2 + 2;          // one exchange slot: 33 entries of two differences (+ pad, keeps 8-byte alignment)
constexpr int kGX = 8 * 2 * kGXSlot;         // 8 steps x {(d0, d1), (d2, d3)}
constexpr int kGWarpWords = (kGRing + 1) * kGTile + kGHand + kGX;
extern __shared__ __align__(16) uint32_t grad_sm[];
#ifdef B2_GRAD_TRACE
__device__ long long g_grad_trace[64 * 128];   // [band][0]: start, [band][1 + b]: end of block b (tools/grad_trace.cu)
#define B2_TRACE(slot) do { if (lane == 0 && bi < 64 && (slot) < 128) g_grad_trace[bi * 128 + (slot)] = clock64(); } while (0)
#else
#define B2_TRACE(slot) do { } while (0)
#endif

__device__ __forceinline__ int ld_acquire_cta_shared(const volatile int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared((const void*)p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta_shared(volatile int* p, int v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared((void*)p)), "r"(v) : "memory");
}

__device__ __forceinline__ void gradient_band_tiled(const FiltJobDev& j, int bi, volatile int* progress, int warp,
                                                    uint32_t* sm_all, int lane) {
  uint32_t* const sm = sm_all + warp * kGWarpWords;
  const int w = j.width, h = j.height;
  const int band = bi * 32, row = band + lane;
  const bool row_ok = row < h;
  const int nq = (w + 3) >> 2, nq_full = w >> 2;
  const int nblk = (nq + 31 + 7) >> 3;       // quad steps 0 .. nq + 30
  const int up_warp = (warp + kGradWarps - 1) % kGradWarps, down_warp = (warp + 1) % kGradWarps;
  const bool has_down = band + 32 < h;
  auto wait_word = [&](int who, int value) {   // lane 0 polls (acquire: a plain shared load on this target), the warp follows
    if (lane == 0) {
      while (ld_acquire_cta_shared(progress + who) < value) {
#if B2_GRAD_SLEEP > 0
        __nanosleep(B2_GRAD_SLEEP);
#endif
      }
    }
    __syncwarp();
  };
  if (bi >= kGradWarps) wait_word(down_warp, ((bi - kGradWarps + 1) << 12) + nblk);   // the ring is free again
  uint32_t* const tout = sm + kGRing * kGTile;
  uint32_t* const hand = tout + kGTile;
  // exchange between neighbouring lanes: slot t, entry e: what lane e - 1 computed in step t (entry 0: the row above)
  uint32_t* const xch = hand + kGHand;
  const uint32_t* const hand_up = sm_all + up_warp * kGWarpWords + (kGRing + 1) * kGTile;
  // this lane's share of a tile transfer: row tr0 + 4 i, word tw of the tile
  const int tr0 = lane >> 3, tw = lane & 7;
  const uint32_t sm_s = (uint32_t)__cvta_generic_to_shared(sm);
  const uint32_t tile_off = (uint32_t)(tr0 * kGS + tw) * 4u;
  const uint8_t* src_l = j.src + (ptrdiff_t)(band + tr0) * j.src_stride + 4 * (tw - tr0);
  uint8_t* dst_l = j.dst + (ptrdiff_t)(band + tr0) * j.dst_stride + 4 * (tw - tr0);
  const ptrdiff_t src_step = 4 * (ptrdiff_t)j.src_stride - 16, dst_step = 4 * (ptrdiff_t)j.dst_stride - 16;
  // rows of this lane's tile items that exist (a short last band): bit i <-> row tr0 + 4 i
  uint32_t rows_ok = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) rows_ok |= (band + tr0 + 4 * i < h) ? (1u << i) : 0u;
  auto issue = [&](int b) {            // the inputs of block b (a whole word is inside the row's stride: strides are multiples of 4)
    const uint32_t d0 = sm_s + (uint32_t)((b % kGRing) * kGTile) * 4u + tile_off;
    const uint8_t* p = src_l + 32 * b;
    if (8 * b >= 31 && 8 * b + 7 < nq) {    // warp-uniform: the whole parallelogram is inside the rows
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t nbytes = ((rows_ok >> i) & 1u) ? 4u : 0u;   // 0: nothing is read (the address is not touched), the word is zero-filled
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(d0 + (uint32_t)(4 * i * kGS) * 4u), "l"(p), "r"(nbytes) : "memory");
        p += src_step;
      }
    } else if (b < nblk) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = 8 * b - (tr0 + 4 * i) + tw;
        const uint32_t nbytes = (((rows_ok >> i) & 1u) && m >= 0 && m < nq) ? 4u : 0u;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(d0 + (uint32_t)(4 * i * kGS) * 4u), "l"(p), "r"(nbytes) : "memory");
        p += src_step;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  uint32_t left = 0;                   // this row's previous output byte
  int d0 = 0, d1 = 0, d2 = 0, d3 = 0;  // the differences of the word this lane takes next
  // words 8 b .. 8 b + 8 of the row above (lanes 0 .. 8) -> their four differences; lane 1 + t puts them where lane 0
  // finds them after step t (entry 0 of slot t); word 0 goes straight into lane 0's registers
  auto above_block = [&](int b) {
    const int m = 8 * b + lane;
    uint32_t aw = 0u;
    if (band > 0 && lane < 9 && m < nq) aw = hand_up[m & (kGHand - 1)];
    const uint32_t prev = __shfl_up_sync(0xffffffffu, aw, 1);
    const int tl = (lane == 0) ? 0 : (int)(prev >> 24);    // lane 0's own numbers are only used for word 0 (b == 0): nothing to its left
    const int a0 = (int)(aw & 0xff), a1 = (int)((aw >> 8) & 0xff), a2 = (int)((aw >> 16) & 0xff), a3 = (int)(aw >> 24);
    if (lane >= 1 && lane < 9) {
      *reinterpret_cast<int2*>(xch + (2 * (lane - 1)) * kGXSlot) = make_int2(a0 - tl, a1 - a0);
      *reinterpret_cast<int2*>(xch + (2 * (lane - 1) + 1) * kGXSlot) = make_int2(a2 - a1, a3 - a2);
    }
    if (b == 0 && lane == 0) { d0 = a0; d1 = a1 - a0; d2 = a2 - a1; d3 = a3 - a2; }
    __syncwarp();
  };
  // One quad step.  The lane's differences leave through shared memory (entry lane + 1 of the step's slot) and the
  // next word's arrive from entry `lane`, in two halves, each behind a warp barrier: the barrier pins the order
  // store -> load in the instruction stream, so the loads are issued half a step and a full step before the chain
  // needs them (shuffles were grouped by ptxas at the end of the step, back on the chain).
  uint32_t* const xw = xch + 2 * (lane + 1);
  const uint32_t* const xr = xch + 2 * lane;
  // No activity test in the step: a lane that is not inside its row yet (word < 0), or whose row does not exist,
  // reads zero-filled inputs and differences that are zero, so it produces zeros, keeps left = 0 and hands zeros
  // down -- exactly the state the first word of its row must start from; past the end of the row the numbers are
  // garbage that only reaches lanes which are past the end too.  Only the stores know about the edges.
  auto step = [&](int t, uint32_t vin) -> uint32_t {
    const uint32_t o0 = (vin + (uint32_t)__viaddmin_s32_relu((int)left, d0, 255)) & 0xffu;
    const uint32_t o1 = ((vin >> 8) + (uint32_t)__viaddmin_s32_relu((int)o0, d1, 255)) & 0xffu;
    *reinterpret_cast<int2*>(xw + (2 * t) * kGXSlot) = make_int2((int)(o0 - left), (int)(o1 - o0));
    __syncwarp();
    const int2 n01 = *reinterpret_cast<const int2*>(xr + (2 * t) * kGXSlot);
    const uint32_t o2 = ((vin >> 16) + (uint32_t)__viaddmin_s32_relu((int)o1, d2, 255)) & 0xffu;
    const uint32_t o3 = ((vin >> 24) + (uint32_t)__viaddmin_s32_relu((int)o2, d3, 255)) & 0xffu;
    *reinterpret_cast<int2*>(xw + (2 * t + 1) * kGXSlot) = make_int2((int)(o2 - o1), (int)(o3 - o2));
    __syncwarp();
    const int2 n23 = *reinterpret_cast<const int2*>(xr + (2 * t + 1) * kGXSlot);
    left = o3;
    d0 = n01.x; d1 = n01.y; d2 = n23.x; d3 = n23.y;
    return __byte_perm(__byte_perm(o0, o1, 0x0040), __byte_perm(o2, o3, 0x0040), 0x5410);
  };
  const int tag_up = (bi - 1) << 12, tag_down = (bi + 1) << 12;
  B2_TRACE(0);
  issue(0);
  issue(1);
  if (band > 0) wait_word(up_warp, tag_up + min(5, nblk));
  above_block(0);
  for (int b = 0; b < nblk; ++b) {
    const int q0 = 8 * b;
    issue(b + 2);
    if (b > 0) {
      if (band > 0) wait_word(up_warp, tag_up + min(b + 5, nblk));
      above_block(b);
    }
    if (has_down && b > kGGap) wait_word(down_warp, tag_down + b - kGGap);
    asm volatile("cp.async.wait_group 2;" ::: "memory");
    __syncwarp();
    const uint32_t* tin = sm + (b % kGRing) * kGTile;
    const bool steady = q0 >= 31 && q0 + 7 < nq_full;   // warp-uniform: every lane inside its row (if the row exists) on whole words: plain stores
    uint32_t vin[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) vin[t] = tin[lane * kGS + t];
    // lane 31's words are the row above the next band: they go to the ring as they are made, and the band below
    // is told before this block's outputs start for global memory
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int m31 = q0 + t - 31;       // lane 31's word
      const uint32_t o = step(t, vin[t]);
      tout[lane * kGS + t] = o;
      if (lane == 31 && has_down && (steady || (m31 >= 0 && m31 < nq))) hand[m31 & (kGHand - 1)] = o;
    }
    __syncwarp();                        // the block's outputs are in the output tile
    if (lane == 31) st_release_cta_shared(progress + warp, (bi << 12) + b + 1);
    uint8_t* p = dst_l + 32 * b;
    uint32_t v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = tout[(tr0 + 4 * i) * kGS + tw];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = q0 - (tr0 + 4 * i) + tw;
      if (((rows_ok >> i) & 1u) && (steady || (m >= 0 && m < nq_full))) *reinterpret_cast<uint32_t*>(p + i * dst_step) = v[i];
    }
    if (!steady && (w & 3)) {            // the row's last, partial word
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = q0 - (tr0 + 4 * i) + tw;
        if (((rows_ok >> i) & 1u) && m == nq_full)
          for (int t2 = 0; 4 * m + t2 < w; ++t2) p[i * dst_step + t2] = (uint8_t)(v[i] >> (8 * t2));
      }
    }
    __syncwarp();                        // the output tile is free again
    B2_TRACE(1 + b);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
}

// ---- unfilter, GRADIENT: one CTA per image, its warps pipelined down the bands ------------
// Warp w takes bands w, w + kGradWarps, ... of its image.  Aligned rows: gradient_band_tiled above.
// Unaligned rows (the byte-skewed loop below): a band is a wavefront inside one warp (lane = row, one
// column of skew per row); band b needs the last row of band b-1 one column ahead of its own lane 0, so
// after every 8 steps a warp publishes how many columns of its band's last row are final (band * 2^14 +
// columns, one word per warp in shared memory, monotonic) and the warp below waits for the columns it is
// about to read.  The last row itself travels through global memory (written by lane 31, CTA-scope
// fence, flag; read back by lane 0 with volatile loads after its own fence).
__global__ void __launch_bounds__(32 * kGradWarps)
alpha_unfilter_gradient_kernel(const FiltJobDev* __restrict__ jobs, uint32_t n_jobs) {
  __shared__ volatile int progress[kGradWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t img = blockIdx.x; img < n_jobs; img += gridDim.x) {
    const FiltJobDev j = jobs[img];
    __syncthreads();                 // the previous image's flags are no longer read
    if (lane == 0) progress[warp] = 0;
    __syncthreads();
    if (j.filter != 3) continue;
    const int w = j.width, h = j.height;
    const int producer = (warp + kGradWarps - 1) % kGradWarps;
    for (int bi = warp; bi * 32 < h; bi += kGradWarps) {
      const int band = bi * 32;
      const int row = band + lane;
      const bool row_ok = row < h;
      const uint8_t* in = j.src + (ptrdiff_t)row * j.src_stride;
      uint8_t* out = j.dst + (ptrdiff_t)row * j.dst_stride;
      const uint8_t* above = j.dst + (ptrdiff_t)(band - 1) * j.dst_stride;   // last row of the band above (band > 0)
      const int need_base = (bi - 1) << 14;
      auto wait_for = [&](int columns) {   // columns of the row above that must be final
        if (band > 0) {
          if (lane == 0) {
            const int need = need_base + min(columns, w);
            while (progress[producer] < need) {}
            __threadfence_block();
          }
          __syncwarp();
        }
      };
      const bool words = ((((uintptr_t)j.src | (uintptr_t)j.dst) | (uint32_t)j.src_stride | (uint32_t)j.dst_stride) & 3) == 0;
      if (words) {                     // 4-byte aligned rows: the band through shared-memory tiles (above)
        gradient_band_tiled(j, bi, progress, warp, grad_sm, lane);
        continue;
      }
      int left = 0, top_left = 0, cur_out = 0;
      // a lane touches a different row than its neighbours, so every access costs the L1 one
      // sector per lane: outputs leave as aligned 32-bit words (a quarter of the lanes per step)
      const bool packed = ((((uintptr_t)j.dst) | (uint32_t)j.dst_stride) & 3) == 0;
      uint32_t acc = 0;
      // inputs (and, for lane 0, the row above the band) are fetched 8 steps ahead: the step
      // is a dependency chain of a shuffle and a handful of ALU operations, nothing else
      uint32_t pre[8], apre[8];
      wait_for(8);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int c = q - lane;
        pre[q] = (row_ok && c >= 0 && c < w) ? in[c] : 0u;
        apre[q] = (lane == 0 && band > 0 && q < w) ? (uint32_t)*(const volatile uint8_t*)(above + q) : 0u;
      }
      for (int k0 = 0; k0 < w + 31; k0 += 8) {
        wait_for(k0 + 16);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int k = k0 + q;
          const int c = k - lane;
          const int v = (int)pre[q];
          const int atop = (int)apre[q];
          const int cn = c + 8;
          pre[q] = (row_ok && cn >= 0 && cn < w) ? in[cn] : 0u;
          apre[q] = (lane == 0 && band > 0 && k + 8 < w) ? (uint32_t)*(const volatile uint8_t*)(above + k + 8) : 0u;
          // top = (row-1, c): the lane above produced it on the previous step
          int top = __shfl_up_sync(0xffffffffu, cur_out, 1);
          if (lane == 0) top = atop;
          if (row_ok && c >= 0 && c < w) {
            int o;
            if (row == 0) o = (v + left) & 0xff;                        // first row: horizontal, pred 0 at c == 0
            else if (c == 0) o = (v + top) & 0xff;                      // left = top = top_left = prev[0]
            else o = (v + gradient_pred(left, top, top_left)) & 0xff;
            acc = (acc >> 8) | ((uint32_t)o << 24);
            if (!packed) {
              out[c] = (uint8_t)o;
            } else if ((c & 3) == 3) {
              *(uint32_t*)(out + c - 3) = acc;
            } else if (c == w - 1) {                                     // ragged end of the row
              const uint32_t tail = acc >> (8 * (3 - (c & 3)));
              for (int t = 0; t <= (c & 3); ++t) out[c - (c & 3) + t] = (uint8_t)(tail >> (8 * t));
            }
            left = o;
            cur_out = o;
          }
          top_left = top;
        }
        if (lane == 31) {            // columns 0 .. k0 + 7 - 31 of the band's last row are written
          int done = min(max(k0 + 8 - 31, 0), w);
          if (packed && done < w) done &= ~3;   // whole words only have left the registers
          __threadfence_block();   // producer and consumer share the CTA; the row is read back with volatile loads
          progress[warp] = (bi << 14) + done;
        }
      }
    }
  }
}

enum { KIND_FORWARD, KIND_ROWS, KIND_VERTICAL };

int prepare_filt(b2_ctx* ctx, int stream, const b2_alpha_filter_job* jobs, int n_jobs, int kind, const FiltJobDev** dj,
                 const uint32_t** dp, uint32_t* uniform, uint32_t* total, int* counts /*[4]*/) {
  const size_t jb = (sizeof(FiltJobDev) * (size_t)n_jobs + 15) & ~(size_t)15;
  const size_t pb = sizeof(uint32_t) * ((size_t)n_jobs + 1);
  if (jb + pb > B2_DESC_RING_BYTES) return B2_ERR_INVALID_CONFIG;
  uint8_t* host = (uint8_t*)calloc(1, jb + pb);
  if (host == NULL) return B2_ERR_OOM;
  FiltJobDev* d = (FiltJobDev*)host;
  uint32_t* pre = (uint32_t*)(host + jb);
  uint64_t tot = 0;
  uint32_t uni = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const b2_alpha_filter_job& j = jobs[i];
    int st = B2_OK;
    if (j.src == NULL || j.dst == NULL) st = B2_ERR_NULL;
    else if (j.width <= 0 || j.height <= 0 || j.width > B2_MAX_DIMENSION || j.height > B2_MAX_DIMENSION) st = B2_ERR_BAD_DIM;
    else if (j.src_stride < j.width || j.dst_stride < j.width) st = B2_ERR_BAD_DIM;
    else if (j.filter < 0 || j.filter > 3 || (const uint8_t*)j.dst == j.src) st = B2_ERR_INVALID_CONFIG;   // WEBP_FILTER_TYPE; out of place
    if (st != B2_OK) { free(host); return st; }
    d[i].src = j.src; d[i].dst = j.dst; d[i].src_stride = j.src_stride; d[i].dst_stride = j.dst_stride;
    d[i].width = j.width; d[i].height = j.height; d[i].filter = j.filter;
    d[i].units_x = (uint32_t)((j.width + 127) / 128);
    if (counts) counts[j.filter]++;
    uint32_t items = 0;
    switch (kind) {
      case KIND_FORWARD: items = (uint32_t)((j.width + kFwdUnit - 1) / kFwdUnit) * (uint32_t)((j.height + kFwdRows - 1) / kFwdRows); break;
      case KIND_ROWS: items = (uint32_t)((j.height + kRowsPerWarp - 1) / kRowsPerWarp); break;
      default: items = d[i].units_x; break;
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
  *dj = (const FiltJobDev*)dev; *dp = (const uint32_t*)((uint8_t*)dev + jb); *uniform = uni; *total = (uint32_t)tot;
  return B2_OK;
}

// ---- rows that miss the fast layout: re-pitch into padded scratch, run the fast kernels there, re-pitch back -------
// The reference's alpha plane is tightly packed (stride = width), so three planes in four have rows of every byte
// alignment, which the kernels above can only walk byte by byte (0.02 - 0.13 of the copy bandwidth against 0.27 - 0.81).
// A job whose pointers or strides are not multiples of 16 is therefore copied into stream scratch with a stride (and
// a working width) rounded up to 16 -- extra columns on the right never influence the columns left of them, in any of
// the filters, either direction --, filtered there by the fast paths, and copied back: two more passes of 2 B/px.
struct PitchJobDev {
  const uint8_t* src;
  uint8_t* dst;
  int32_t src_stride, dst_stride, width, height;
  uint32_t segs, pad;
};
constexpr int kPitchSegWords = 1024;   // destination words of one row per warp item
__global__ void __launch_bounds__(kThreads, 8)
repitch_kernel(const PitchJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix, uint32_t n_jobs, uint32_t uniform,
               uint32_t total) {
  const int lane = threadIdx.x & 31;
  for (uint32_t item = blockIdx.x * kWarps + (threadIdx.x >> 5); item < total; item += gridDim.x * kWarps) {
    uint32_t local;
    const uint32_t img = b2_find_image_warp(prefix, n_jobs, uniform, item, &local, (uint32_t)lane);
    const PitchJobDev* __restrict__ j = jobs + img;
    const int w = j->width;
    const uint32_t y = local / j->segs;
    const int seg = (int)(local - y * j->segs);
    const uint8_t* s0 = j->src + (ptrdiff_t)y * j->src_stride;
    uint8_t* d0 = j->dst + (ptrdiff_t)y * j->dst_stride;
    const int db = (int)((uintptr_t)d0 & 3);                 // the row's first byte inside its first aligned word
    uint32_t* const dw = (uint32_t*)((uintptr_t)d0 & ~(uintptr_t)3);
    const int n_words = (db + w + 3) >> 2;
    const uintptr_t s_first = (uintptr_t)s0 & ~(uintptr_t)3, s_last = (uintptr_t)(s0 + w - 1) & ~(uintptr_t)3;   // words that hold a byte of the row
    const int k_end = min(n_words, (seg + 1) * kPitchSegWords);
    for (int k0 = seg * kPitchSegWords + lane; k0 < k_end + lane; k0 += 4 * 32) {   // warp-uniform trip count
      uint32_t lo[4], hi[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + 32 * u;
        lo[u] = hi[u] = 0u;
        if (k < k_end) {
          const uintptr_t sa = (uintptr_t)(s0 + (4 * k - db));
          const uintptr_t a = sa & ~(uintptr_t)3;
          if (a >= s_first && a <= s_last) lo[u] = __ldg((const uint32_t*)a);
          if ((sa & 3) != 0 && a + 4 >= s_first && a + 4 <= s_last) hi[u] = __ldg((const uint32_t*)(a + 4));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + 32 * u;
        if (k < k_end) {
          const int o = 4 * k - db;                          // offset in the row of the word's first byte
          const uint32_t sm = (uint32_t)((uintptr_t)(s0 + o) & 3);
          const uint32_t v = __funnelshift_r(lo[u], hi[u], 8u * sm);
          if (o >= 0 && o + 4 <= w) {
            dw[k] = v;
          } else {
#pragma unroll
            for (int b = 0; b < 4; ++b)
              if (o + b >= 0 && o + b < w) ((uint8_t*)(dw + k))[b] = (uint8_t)(v >> (8 * b));
          }
        }
      }
    }
  }
}

int launch_repitch(b2_ctx* ctx, int stream, const PitchJobDev* host_jobs, int n) {
  const size_t jb = (sizeof(PitchJobDev) * (size_t)n + 15) & ~(size_t)15;
  const size_t pb = sizeof(uint32_t) * ((size_t)n + 1);
  if (jb + pb > B2_DESC_RING_BYTES) return B2_ERR_INVALID_CONFIG;
  uint8_t* host = (uint8_t*)calloc(1, jb + pb);
  if (host == NULL) return B2_ERR_OOM;
  PitchJobDev* d = (PitchJobDev*)host;
  uint32_t* pre = (uint32_t*)(host + jb);
  uint64_t tot = 0;
  uint32_t uni = 0;
  for (int i = 0; i < n; ++i) {
    d[i] = host_jobs[i];
    d[i].segs = (uint32_t)((((d[i].width + 6) >> 2) + kPitchSegWords - 1) / kPitchSegWords);   // up to (3 + w + 3) / 4 words
    const uint32_t items = d[i].segs * (uint32_t)d[i].height;
    if (i == 0) uni = items; else if (items != uni) uni = 0;
    pre[i] = (uint32_t)tot;
    tot += items;
    if (tot > 0x7fffffffull) { free(host); return B2_ERR_BAD_DIM; }
  }
  pre[n] = (uint32_t)tot;
  void* dev = NULL;
  const int st = b2_upload_desc(ctx, stream, host, jb + pb, &dev);
  free(host);
  if (st != B2_OK) return st;
  repitch_kernel<<<b2_grid_for(ctx, 8, (tot + kWarps - 1) / kWarps), kThreads, 0, ctx->streams[stream]>>>(
      (const PitchJobDev*)dev, (const uint32_t*)((uint8_t*)dev + jb), (uint32_t)n, uni, (uint32_t)tot);
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}

int filter_aligned(b2_ctx* ctx, int stream, const b2_alpha_filter_job* jobs, int n_jobs);
int unfilter_aligned(b2_ctx* ctx, int stream, const b2_alpha_filter_job* jobs, int n_jobs);

// Validates like prepare_filt, then runs `jobs` with the rows of the misaligned ones moved onto padded scratch.
int run_with_repitch(b2_ctx* ctx, int stream, const b2_alpha_filter_job* jobs, int n_jobs, bool unfilter) {
  size_t scratch = 0;
  int n_bad = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const b2_alpha_filter_job& j = jobs[i];
    if (j.src == NULL || j.dst == NULL) return B2_ERR_NULL;
    if (j.width <= 0 || j.height <= 0 || j.width > B2_MAX_DIMENSION || j.height > B2_MAX_DIMENSION) return B2_ERR_BAD_DIM;
    if (j.src_stride < j.width || j.dst_stride < j.width) return B2_ERR_BAD_DIM;
    if (j.filter < 0 || j.filter > 3 || (const uint8_t*)j.dst == j.src) return B2_ERR_INVALID_CONFIG;
    const bool bad = ((((uintptr_t)j.src | (uintptr_t)j.dst) | (uint32_t)j.src_stride | (uint32_t)j.dst_stride) & 15) != 0;
    if (bad && j.width >= 64) {   // narrower planes are not worth three launches
      ++n_bad;
      scratch += 2 * (size_t)((j.width + 15) & ~15) * (size_t)j.height;
    }
  }
  if (n_bad == 0) return unfilter ? unfilter_aligned(ctx, stream, jobs, n_jobs) : filter_aligned(ctx, stream, jobs, n_jobs);
  if (scratch > ((size_t)2 << 30) && n_jobs > 1) {   // keep the (grow-only) stream scratch bounded: halves, one after the other
    const int half = n_jobs / 2;
    const int st0 = run_with_repitch(ctx, stream, jobs, half, unfilter);
    return st0 != B2_OK ? st0 : run_with_repitch(ctx, stream, jobs + half, n_jobs - half, unfilter);
  }
  void* sp = NULL;
  int st = b2_stream_scratch(ctx, stream, scratch + 256, &sp);
  if (st != B2_OK) return st;
  uint8_t* base = (uint8_t*)(((uintptr_t)sp + 255) & ~(uintptr_t)255);
  b2_alpha_filter_job* work = (b2_alpha_filter_job*)malloc(sizeof(b2_alpha_filter_job) * (size_t)n_jobs);
  PitchJobDev* cin = (PitchJobDev*)calloc((size_t)n_bad, sizeof(PitchJobDev));
  PitchJobDev* cout = (PitchJobDev*)calloc((size_t)n_bad, sizeof(PitchJobDev));
  if (work == NULL || cin == NULL || cout == NULL) { free(work); free(cin); free(cout); return B2_ERR_OOM; }
  size_t off = 0;
  int k = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const b2_alpha_filter_job& j = jobs[i];
    work[i] = j;
    const bool bad = ((((uintptr_t)j.src | (uintptr_t)j.dst) | (uint32_t)j.src_stride | (uint32_t)j.dst_stride) & 15) != 0;
    if (!(bad && j.width >= 64)) continue;
    const int pw = (j.width + 15) & ~15;
    uint8_t* s_in = base + off;
    uint8_t* s_out = s_in + (size_t)pw * j.height;
    off += 2 * (size_t)pw * j.height;
    cin[k].src = j.src; cin[k].dst = s_in; cin[k].src_stride = j.src_stride; cin[k].dst_stride = pw; cin[k].width = j.width; cin[k].height = j.height;
    cout[k].src = s_out; cout[k].dst = j.dst; cout[k].src_stride = pw; cout[k].dst_stride = j.dst_stride; cout[k].width = j.width; cout[k].height = j.height;
    work[i].src = s_in; work[i].dst = s_out; work[i].src_stride = pw; work[i].dst_stride = pw; work[i].width = pw;
    ++k;
  }
  st = launch_repitch(ctx, stream, cin, n_bad);
  if (st == B2_OK) st = unfilter ? unfilter_aligned(ctx, stream, work, n_jobs) : filter_aligned(ctx, stream, work, n_jobs);
  if (st == B2_OK) st = launch_repitch(ctx, stream, cout, n_bad);
  free(work); free(cin); free(cout);
  return st;
}

int filter_aligned(b2_ctx* ctx, int stream, const b2_alpha_filter_job* jobs, int n_jobs) {
  const FiltJobDev* dj; const uint32_t* dp; uint32_t uni, tot;
  const int rc = prepare_filt(ctx, stream, jobs, n_jobs, KIND_FORWARD, &dj, &dp, &uni, &tot, NULL);
  if (rc != B2_OK) return rc;
  alpha_filter_kernel<<<b2_grid_for(ctx, 8, (tot + kWarps - 1) / kWarps), kThreads, 0, ctx->streams[stream]>>>(dj, dp, (uint32_t)n_jobs, uni, tot);
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}

int unfilter_aligned(b2_ctx* ctx, int stream, const b2_alpha_filter_job* jobs, int n_jobs) {
  cudaStream_t s = ctx->streams[stream];
  int counts[4] = {0, 0, 0, 0};
  const FiltJobDev* dj; const uint32_t* dp; uint32_t uni, tot;
  int rc = prepare_filt(ctx, stream, jobs, n_jobs, KIND_ROWS, &dj, &dp, &uni, &tot, counts);
  if (rc != B2_OK) return rc;
  if (counts[0] + counts[1] > 0) {
    alpha_unfilter_rows_kernel<<<b2_grid_for(ctx, 8, (tot + kWarps - 1) / kWarps), kThreads, 0, s>>>(dj, dp, (uint32_t)n_jobs, uni, tot);
    ctx->launches++;
  }
  if (counts[3] > 0) {
    const int grad_smem = kGradWarps * kGWarpWords * (int)sizeof(uint32_t);
    if (grad_smem > 0)   // per device: set on every call (a host-side table write)
      B2_CUDA(ctx, cudaFuncSetAttribute(alpha_unfilter_gradient_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, grad_smem));
    alpha_unfilter_gradient_kernel<<<b2_grid_for(ctx, kGradCtasPerSm, (uint64_t)n_jobs), 32 * kGradWarps, grad_smem, s>>>(dj, (uint32_t)n_jobs);
    ctx->launches++;
  }
  if (counts[2] > 0) {
    rc = prepare_filt(ctx, stream, jobs, n_jobs, KIND_VERTICAL, &dj, &dp, &uni, &tot, NULL);
    if (rc != B2_OK) return rc;
    alpha_unfilter_vertical_kernel<<<b2_grid_for(ctx, 8, (tot + kWarps - 1) / kWarps), kThreads, 0, s>>>(dj, dp, (uint32_t)n_jobs, uni, tot);
    ctx->launches++;
  }
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}

}  // namespace

#define B2_FILT_ENTRY_CHECKS          \
  B2_CHECK_CTX(ctx, stream);          \
  if (n_jobs < 0) return B2_ERR_INVALID_CONFIG; \
  if (n_jobs == 0) return B2_OK;      \
  if (jobs == NULL) return B2_ERR_NULL;

extern "C" int b2_alpha_filter(b2_ctx* ctx, int stream, const b2_alpha_filter_job* jobs, int n_jobs) {
  B2_FILT_ENTRY_CHECKS
  return run_with_repitch(ctx, stream, jobs, n_jobs, false);
}

extern "C" int b2_alpha_unfilter(b2_ctx* ctx, int stream, const b2_alpha_filter_job* jobs, int n_jobs) {
  B2_FILT_ENTRY_CHECKS
  return run_with_repitch(ctx, stream, jobs, n_jobs, true);
}
