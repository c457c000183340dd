// b2_rescale.cu -- the fixed-point rescaler:
//   WebPRescalerInit / Import / Export   pkg/libwebp/utils/rescaler_utils.c.go:31-158
//   ImportRowExpand / ImportRowShrink / ExportRowExpand / ExportRowShrink
//                                        pkg/libwebp/dsp/rescaler.c.go:39-197
// driven over a whole plane as RescalePlane does (pkg/libwebp/enc/picture_rescale_enc.c.go:
// 125-136; the decoder's EmitRescaled* loops feed the same machine row batch by row batch).
//
// The reference is a streaming state machine: rows are imported one by one into `frow`,
// accumulated into `irow`, and exported whenever the vertical accumulator runs out.  What
// carries from one output to the next -- horizontally the fractional start of the next
// pixel, vertically the fractional start of the next row -- is a function of the LAST input
// sample and of the accumulator only, never of the running sums.  So every output sample is
// an independent function of a small rectangle of input samples once the accumulator walks
// are known: the host replays the two integer walks (O(width + height), no pixel work) into
// an x table and a y table, and one kernel computes every output sample from them.
#include <stdlib.h>
#include <string.h>

#include <map>
#include <tuple>
#include <vector>

#include "b2_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr uint64_t kOne = 1ull << 32;
constexpr uint64_t kRounder = kOne >> 1;

struct __align__(16) XEnt { int32_t a, b, accum, pad; };   // expand: left, right sample; shrink: [a, b) consumed; accum after
                                                           // (16 bytes: one 128-bit load per entry, consecutive lanes = consecutive entries)
struct YEnt { int32_t prev_last, last, y_accum; uint32_t prev_yscale, yscale; };   // last < 0: row never exported;
                                                                                   // expand: yscale holds B = FRAC(-y_accum, y_sub)

struct RescaleJobDev {
  const uint8_t* src;
  uint8_t* dst;
  int32_t src_stride, dst_stride, dst_width, dst_height, num_channels;
  int32_t x_expand, y_expand, x_add, x_sub, y_sub;
  uint32_t fx_scale, fy_scale, fxy_scale;
  uint32_t xtab, ytab;   // entry offsets into the table pools
  uint32_t units_x;
};

__device__ __forceinline__ uint32_t mult_fix(uint32_t x, uint32_t y) { return (uint32_t)(((uint64_t)x * y + kRounder) >> 32); }
__device__ __forceinline__ uint32_t mult_fix_floor(uint32_t x, uint32_t y) { return (uint32_t)(((uint64_t)x * y) >> 32); }

// frow[xo] of source row `row` (rescaler.c.go:39-105), c = channel
// XE / YE / CH: compile-time x_expand / y_expand / num_channels of a batch whose jobs all agree (-1: read the job)
template <int XE, int CH>
// e = the x-table entry of output sample xo, p = the entry of xo - 1 (shrinking only)
__device__ __forceinline__ uint32_t frow_at(const RescaleJobDev& j, const XEnt& e, const XEnt& p, const uint8_t* __restrict__ row,
                                            int xo, int c) {
  const int ch = CH > 0 ? CH : j.num_channels;
  if (XE < 0 ? j.x_expand : XE) {
    const uint32_t left = row[e.a * ch + c], right = row[e.b * ch + c];
    return right * (uint32_t)j.x_add + (left - right) * (uint32_t)e.accum;
  }
  uint32_t sum = 0;
  if (xo > 0) {   // fractional start left by the previous output sample
    sum = mult_fix((uint32_t)row[(p.b - 1) * ch + c] * (uint32_t)(-p.accum), j.fx_scale);
  }
  uint32_t base = 0;
  for (int xi = e.a; xi < e.b; ++xi) {
    base = row[xi * ch + c];
    sum += base;
  }
  return sum * (uint32_t)j.x_sub - base * (uint32_t)(-e.accum);
}

// one output sample: row dy, sample s = x_out * channels + c
template <int XE, int YE, int CH>
__device__ __forceinline__ uint32_t rescale_sample(const RescaleJobDev& j, const XEnt& ex, const XEnt& px, const YEnt& e, int xo, int c) {
  const uint8_t* last_row = j.src + (ptrdiff_t)e.last * j.src_stride;
  if (YE < 0 ? j.y_expand : YE) {                               // ExportRowExpand, rescaler.c.go:110-136
    const uint32_t f = frow_at<XE, CH>(j, ex, px, last_row, xo, c);
    uint32_t J = f;
    if (e.y_accum != 0) {
      const uint32_t i = (e.last > 0) ? frow_at<XE, CH>(j, ex, px, last_row - j.src_stride, xo, c) : 0u;   // irow = the previous row's frow
      const uint32_t B = e.yscale;   // WEBP_RESCALER_FRAC(-y_accum, y_sub), from the host
      const uint32_t A = (uint32_t)(kOne - B);
      J = (uint32_t)(((uint64_t)A * f + (uint64_t)B * i + kRounder) >> 32);
    }
    return min(mult_fix(J, j.fy_scale), 255u);
  }
  // rows accumulated since the previous export (+ its fractional rest)
  uint32_t irow = 0;
  if (e.prev_last >= 0 && e.prev_yscale != 0u)
    irow = mult_fix_floor(frow_at<XE, CH>(j, ex, px, j.src + (ptrdiff_t)e.prev_last * j.src_stride, xo, c), e.prev_yscale);
  uint32_t f = 0;
  for (int r = e.prev_last + 1; r <= e.last; ++r) {
    f = frow_at<XE, CH>(j, ex, px, j.src + (ptrdiff_t)r * j.src_stride, xo, c);
    irow += f;
  }
  if (j.fxy_scale == 0u) return irow & 0xffu;     // special case, rescaler.c.go:184-191
  if (e.yscale != 0u) return min(mult_fix(irow - mult_fix_floor(f, e.yscale), j.fxy_scale), 255u);   // ExportRowShrink :138-163
  return min(mult_fix(irow, j.fxy_scale), 255u);
}

// unit = (job, group of output rows, 1024 output samples); a thread computes 4 samples of
// every row of the group with the x-table entries it loaded once.  Vertical expansion takes
// 8 rows per unit (neighbouring output rows share their source rows), the others one.
constexpr int kPerThread = 4;
__host__ __device__ constexpr int rows_per_unit(int ye) { return ye == 1 ? 8 : 1; }
template <int XE, int YE, int CH>
__global__ void __launch_bounds__(kThreads, 4)
rescale_kernel(const RescaleJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix, uint32_t n_jobs, uint32_t uniform,
               uint32_t total, const XEnt* __restrict__ xpool, const YEnt* __restrict__ ypool) {
  for (uint32_t unit = blockIdx.x; unit < total; unit += gridDim.x) {
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, unit, &local);
    // every thread reads the job and the row entries itself (uniform addresses: one broadcast
    // transaction per warp, L1 hits after the first): no barrier, warps run ahead independently
    const RescaleJobDev j = jobs[img];
    // consecutive lanes take consecutive samples (table entries and source bytes coalesce);
    // a thread's 4 samples are kThreads apart
    const int sbase = (int)(local % j.units_x) * kThreads * kPerThread + (int)threadIdx.x;
    const int ch = CH > 0 ? CH : j.num_channels, n_s = j.dst_width * ch;
    if (sbase >= n_s) continue;
    const XEnt* __restrict__ xt = xpool + j.xtab;
    constexpr int kRowsPerUnit = rows_per_unit(YE);
    const int dy0 = (int)(local / j.units_x) * kRowsPerUnit, dy1 = min(dy0 + kRowsPerUnit, j.dst_height);
    if (kRowsPerUnit == 1) {       // one row: entries are used once, keep one sample's worth in registers
      const YEnt e = ypool[j.ytab + dy0];
      if (e.last < 0) continue;                     // the state machine never exports this row
      uint8_t* out = j.dst + (ptrdiff_t)dy0 * j.dst_stride;
#pragma unroll
      for (int k = 0; k < kPerThread; ++k) {
        const int s = sbase + k * kThreads;
        if (s >= n_s) break;
        const int xo = (ch == 1) ? s : s / ch, c = s - xo * ch;
        out[s] = (uint8_t)rescale_sample<XE, YE, CH>(j, xt[xo], xt[max(xo - 1, 0)], e, xo, c);
      }
      continue;
    }
    XEnt ex[kPerThread], px[kPerThread];
    int xo[kPerThread], c[kPerThread];
#pragma unroll
    for (int k = 0; k < kPerThread; ++k) {
      const int s = min(sbase + k * kThreads, n_s - 1);
      xo[k] = (ch == 1) ? s : s / ch;
      c[k] = s - xo[k] * ch;
      ex[k] = xt[xo[k]];
      px[k] = xt[max(xo[k] - 1, 0)];
    }
    for (int dy = dy0; dy < dy1; ++dy) {
      const YEnt e = ypool[j.ytab + dy];
      if (e.last < 0) continue;
      uint8_t* out = j.dst + (ptrdiff_t)dy * j.dst_stride;
#pragma unroll
      for (int k = 0; k < kPerThread; ++k) {
        const int s = sbase + k * kThreads;
        if (s >= n_s) break;
        out[s] = (uint8_t)rescale_sample<XE, YE, CH>(j, ex[k], px[k], e, xo[k], c[k]);
      }
    }
  }
}

// ---- one-channel planes: a thread walks DOWN its output columns -----------------------------------
// The per-sample kernel above recomputes the horizontally rescaled value frow[xo] of a source row for
// every output row that touches it (four times per source row when a plane is doubled) and pays a table
// entry, an address computation and byte loads per sample.  For one-channel planes whose jobs share the
// direction, a thread instead owns kCols adjacent output columns and walks a group of kRows output rows
// top to bottom -- the reference's own streaming order (WebPRescalerImport / Export,
// rescaler_utils.c.go:119-158) -- keeping what the state machine keeps: frow of the last imported row and
// either the previous row's frow (vertical expansion: irow of ExportRowExpand) or the running sum
// (shrinking: irow of ExportRowShrink).  Every source row's frow is computed once per column and group,
// the x entries are loaded once per group, four samples leave as one 32-bit store.
constexpr int kCols = 4;
constexpr int kRowsGroup = 32;
template <int XE, int YE>
__global__ void __launch_bounds__(kThreads, 3)
rescale_columns_kernel(const RescaleJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix, uint32_t n_jobs, uint32_t uniform,
                       uint32_t total, const XEnt* __restrict__ xpool, const YEnt* __restrict__ ypool) {
  for (uint32_t unit = blockIdx.x; unit < total; unit += gridDim.x) {
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, unit, &local);
    const RescaleJobDev j = jobs[img];
    const int x0 = (int)(local % j.units_x) * kThreads * kCols + (int)threadIdx.x * kCols;
    if (x0 >= j.dst_width) continue;
    const int dy0 = (int)(local / j.units_x) * kRowsGroup, dy1 = min(dy0 + kRowsGroup, j.dst_height);
    const XEnt* __restrict__ xt = xpool + j.xtab;
    const YEnt* __restrict__ yt = ypool + j.ytab;
    XEnt ex[kCols], px[kCols];
    int xo[kCols];
#pragma unroll
    for (int k = 0; k < kCols; ++k) {
      xo[k] = min(x0 + k, j.dst_width - 1);      // columns past the end recompute the last one; they are not stored
      ex[k] = xt[xo[k]];
      px[k] = xt[max(xo[k] - 1, 0)];
    }
    const int n_valid = min(kCols, j.dst_width - x0);
    const bool word_ok = n_valid == kCols && ((((uintptr_t)j.dst) | (uintptr_t)(uint32_t)j.dst_stride) & 3) == 0;
    auto frow_row = [&](int r, uint32_t (&f)[kCols]) {
      const uint8_t* __restrict__ row = j.src + (ptrdiff_t)r * j.src_stride;
#pragma unroll
      for (int k = 0; k < kCols; ++k) f[k] = frow_at<XE, 1>(j, ex[k], px[k], row, xo[k], 0);
    };
    auto store = [&](int dy, const uint32_t (&v)[kCols]) {
      uint8_t* out = j.dst + (ptrdiff_t)dy * j.dst_stride + x0;
      if (word_ok) {
        *reinterpret_cast<uint32_t*>(out) = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
      } else {
#pragma unroll
        for (int k = 0; k < kCols; ++k) if (k < n_valid) out[k] = (uint8_t)v[k];
      }
    };
    uint32_t f[kCols] = {0, 0, 0, 0}, g[kCols] = {0, 0, 0, 0};   // f: frow of row `have`; g: frow of row have - 1 (YE) / of the previous export's last row (YS)
    int have = -2, have_g = -2;
    for (int dy = dy0; dy < dy1; ++dy) {
      const YEnt e = yt[dy];
      if (e.last < 0) continue;                  // the state machine never exports this row
      uint32_t v[kCols];
      if (YE) {                                  // ExportRowExpand, rescaler.c.go:110-136
        if (e.last != have) {
          if (e.last == have + 1 && have >= 0) {
#pragma unroll
            for (int k = 0; k < kCols; ++k) g[k] = f[k];
          } else if (e.last > 0) {
            frow_row(e.last - 1, g);
          }
          frow_row(e.last, f);
          have = e.last;
        }
        const uint32_t B = e.yscale, A = (uint32_t)(kOne - B);
#pragma unroll
        for (int k = 0; k < kCols; ++k) {
          uint32_t J = f[k];
          if (e.y_accum != 0) {
            const uint32_t i = (e.last > 0) ? g[k] : 0u;
            J = (uint32_t)(((uint64_t)A * f[k] + (uint64_t)B * i + kRounder) >> 32);
          }
          v[k] = min(mult_fix(J, j.fy_scale), 255u);
        }
      } else {                                   // ExportRowShrink :138-163 over the rows imported since the previous export
        uint32_t irow[kCols] = {0, 0, 0, 0};
        if (e.prev_last >= 0 && e.prev_yscale != 0u) {
          if (have_g != e.prev_last) frow_row(e.prev_last, g);
#pragma unroll
          for (int k = 0; k < kCols; ++k) irow[k] = mult_fix_floor(g[k], e.prev_yscale);
        }
        for (int r = e.prev_last + 1; r <= e.last; ++r) {
          frow_row(r, f);
#pragma unroll
          for (int k = 0; k < kCols; ++k) irow[k] += f[k];
        }
#pragma unroll
        for (int k = 0; k < kCols; ++k) {
          if (j.fxy_scale == 0u) v[k] = irow[k] & 0xffu;     // special case, rescaler.c.go:184-191
          else if (e.yscale != 0u) v[k] = min(mult_fix(irow[k] - mult_fix_floor(f[k], e.yscale), j.fxy_scale), 255u);
          else v[k] = min(mult_fix(irow[k], j.fxy_scale), 255u);
          g[k] = f[k];
        }
        have_g = e.last;
      }
      store(dy, v);
    }
  }
}

// ---- one-channel planes, streamed: source rows go through shared memory ----------------------------
// The column walker above still waits for its byte loads: a thread needs row r + 1 only after it has
// finished row r, and 24 warps do not hide a DRAM round trip per source row (measured: 1.3-1.8x the
// per-sample kernel, 0.05-0.12 of the HBM peak).  Here the CTA streams the source rows of its unit
// (1024 output columns x 32 output rows) through a two-stage ring in shared memory, four rows per stage,
// with 16-byte cp.async issued by all threads (whole sectors, every source byte fetched once per unit);
// the threads then run the same state machine on rows that are already on chip.  Needs 16-byte aligned
// source rows and a source segment of at most kSegMax bytes per row and unit (the host checks both;
// otherwise the column walker runs).
constexpr int kRS = 4;
constexpr int kSegMax = 4096 + 64;
constexpr int kStreamSmem = 2 * kRS * kSegMax;
template <int XE, int YE>
__global__ void __launch_bounds__(kThreads, 3)
rescale_stream_kernel(const RescaleJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix, uint32_t n_jobs, uint32_t uniform,
                      uint32_t total, const XEnt* __restrict__ xpool, const YEnt* __restrict__ ypool) {
  extern __shared__ __align__(16) uint8_t ring[];
  __shared__ YEnt ys[kRowsGroup];
  __shared__ int info[4];   // lo_al, seg bytes, r_begin, r_end
  const int tid = (int)threadIdx.x;
  for (uint32_t unit = blockIdx.x; unit < total; unit += gridDim.x) {
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, unit, &local);
    const RescaleJobDev j = jobs[img];
    const int xc0 = (int)(local % j.units_x) * kThreads * kCols, xc1 = min(xc0 + kThreads * kCols, j.dst_width);
    const int dy0 = (int)(local / j.units_x) * kRowsGroup, ng = min(kRowsGroup, j.dst_height - dy0);
    const XEnt* __restrict__ xt = xpool + j.xtab;
    __syncthreads();   // the previous unit is done with ys / info / the ring
    if (tid < ng) ys[tid] = ypool[j.ytab + dy0 + tid];
    __syncthreads();
    if (tid == 0) {
      const XEnt e0 = xt[xc0], e1 = xt[xc1 - 1];
      const int lo = XE ? e0.a : max(e0.a - 1, 0), hi = XE ? e1.b + 1 : e1.b;
      info[0] = lo & ~15;
      info[1] = hi - (lo & ~15);
      int r_begin = 0, r_end = -1;
      for (int k = 0; k < ng; ++k) {
        if (ys[k].last < 0) continue;
        if (r_end < 0) {
          if (YE) r_begin = max(ys[k].last - 1, 0);
          else r_begin = (ys[k].prev_last >= 0 && ys[k].prev_yscale != 0u) ? ys[k].prev_last : ys[k].prev_last + 1;
        }
        r_end = ys[k].last;
      }
      info[2] = r_begin; info[3] = r_end;
    }
    __syncthreads();
    const int lo_al = info[0], seg = info[1], r_begin = info[2], r_end = info[3];
    if (r_end < 0) continue;   // uniform: nothing is exported in this group
    const int x0 = xc0 + tid * kCols;
    const bool live = x0 < j.dst_width;
    XEnt ex[kCols], px[kCols];
    int xo[kCols];
#pragma unroll
    for (int k = 0; k < kCols; ++k) {
      xo[k] = min(x0 + k, j.dst_width - 1);
      ex[k] = xt[xo[k]];
      px[k] = xt[max(xo[k] - 1, 0)];
    }
    const int n_valid = min(kCols, j.dst_width - x0);
    const bool word_ok = n_valid == kCols && ((((uintptr_t)j.dst) | (uintptr_t)(uint32_t)j.dst_stride) & 3) == 0;
    const int n_chunks = (seg + 15) >> 4;
    const int n_rows = r_end - r_begin + 1, n_stage = (n_rows + kRS - 1) / kRS;
    auto issue = [&](int st) {   // all threads: rows r_begin + st * kRS .. of the unit into ring stage st & 1
      const int rows_here = min(kRS, n_rows - st * kRS);
      for (int c = tid; c < rows_here * n_chunks; c += kThreads) {
        const int rr = c / n_chunks, ck = c - rr * n_chunks;
        const int r = r_begin + st * kRS + rr;
        const uint8_t* g = j.src + (ptrdiff_t)r * j.src_stride + lo_al + 16 * ck;
        uint8_t* d = ring + ((st & 1) * kRS + rr) * kSegMax + 16 * ck;
        if (lo_al + 16 * ck + 16 <= j.src_stride) {   // the row's own bytes (and its padding): a whole 16-byte chunk
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(d)), "l"(g) : "memory");
        } else {
          for (int b = 0; lo_al + 16 * ck + b < j.src_stride && b < 16; ++b) d[b] = g[b];
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    uint32_t f[kCols] = {0, 0, 0, 0}, g[kCols] = {0, 0, 0, 0}, acc[kCols] = {0, 0, 0, 0};
    int dyi = 0;
    issue(0);
    for (int st = 0; st < n_stage; ++st) {
      if (st + 1 < n_stage) {
        issue(st + 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();
      const int rows_here = min(kRS, n_rows - st * kRS);
      if (live) {
        for (int rr = 0; rr < rows_here; ++rr) {
          const int r = r_begin + st * kRS + rr;
          const uint8_t* row = ring + ((st & 1) * kRS + rr) * kSegMax - lo_al;   // indexed by source x
          uint32_t fn[kCols];
#pragma unroll
          for (int k = 0; k < kCols; ++k) fn[k] = frow_at<XE, 1>(j, ex[k], px[k], row, xo[k], 0);
          if (YE) {                                  // ExportRowExpand, rescaler.c.go:110-136
#pragma unroll
            for (int k = 0; k < kCols; ++k) { g[k] = f[k]; f[k] = fn[k]; }
            while (dyi < ng && ys[dyi].last <= r) {
              const YEnt e = ys[dyi];
              if (e.last == r) {
                const uint32_t B = e.yscale, A = (uint32_t)(kOne - B);
                uint32_t v[kCols];
#pragma unroll
                for (int k = 0; k < kCols; ++k) {
                  uint32_t J = f[k];
                  if (e.y_accum != 0) {
                    const uint32_t i = (e.last > 0) ? g[k] : 0u;
                    J = (uint32_t)(((uint64_t)A * f[k] + (uint64_t)B * i + kRounder) >> 32);
                  }
                  v[k] = min(mult_fix(J, j.fy_scale), 255u);
                }
                uint8_t* out = j.dst + (ptrdiff_t)(dy0 + dyi) * j.dst_stride + x0;
                if (word_ok) *reinterpret_cast<uint32_t*>(out) = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
                else for (int k = 0; k < n_valid; ++k) out[k] = (uint8_t)v[k];
              }
              ++dyi;
            }
          } else {                                   // ExportRowShrink :138-163
            while (dyi < ng && ys[dyi].last < 0) ++dyi;
            if (dyi >= ng) break;
            const YEnt e = ys[dyi];
            if (r <= e.prev_last) {                  // the row the previous export ended on: only its fractional rest counts
#pragma unroll
              for (int k = 0; k < kCols; ++k) g[k] = fn[k];
              continue;
            }
#pragma unroll
            for (int k = 0; k < kCols; ++k) acc[k] += fn[k];
            if (r == e.last) {
              uint32_t v[kCols];
#pragma unroll
              for (int k = 0; k < kCols; ++k) {
                uint32_t irow = acc[k];
                if (e.prev_last >= 0 && e.prev_yscale != 0u) irow += mult_fix_floor(g[k], e.prev_yscale);
                if (j.fxy_scale == 0u) v[k] = irow & 0xffu;     // special case, rescaler.c.go:184-191
                else if (e.yscale != 0u) v[k] = min(mult_fix(irow - mult_fix_floor(fn[k], e.yscale), j.fxy_scale), 255u);
                else v[k] = min(mult_fix(irow, j.fxy_scale), 255u);
                g[k] = fn[k];
                acc[k] = 0;
              }
              uint8_t* out = j.dst + (ptrdiff_t)(dy0 + dyi) * j.dst_stride + x0;
              if (word_ok) *reinterpret_cast<uint32_t*>(out) = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
              else for (int k = 0; k < n_valid; ++k) out[k] = (uint8_t)v[k];
              ++dyi;
            }
          }
        }
      }
      __syncthreads();   // the stage may be refilled
    }
  }
}

// ---- one-channel planes shrinking both ways by less than 4: no loops over source samples -----------
// ncu on the streamed kernel above, 4K -> 1080p: 45 thread instructions per INPUT pixel, issue slots 57 % busy --
// the per-sample loop over source bytes (a byte load from shared memory, an add, a compare and a branch each),
// a 64-bit product to re-derive the previous sample's rest for every sample, the walk over the y entries for
// every source row and an integer division per staged chunk.  With windows of at most 5 source samples
// (ratio <= 4) a sample's window lies in two aligned words, so here
//   * the window sum is two dot products (dp4a) of those words with 0/1 byte masks computed once per unit,
//     the window's last sample one byte permute; the rest a sample leaves for its right neighbour
//     (MULT_FIX(frac, fx_scale), rescaler.c.go:84-100) is chained through the thread's four columns, so it
//     is computed once per sample, not twice;
//   * what happens to a source row (rest only / accumulate / accumulate and export output row dy with
//     yscale) is a two-word entry of a per-unit schedule in shared memory, built by the first 32 threads;
//   * the rest a row leaves for the next output row (MULT_FIX_FLOOR(frow, yscale), :150-158) is the very
//     number the export subtracts: it stays in the accumulator instead of being recomputed.
// Same staging (16-byte cp.async, four rows per stage, two stages), same results bit for bit.
struct RowAct { uint32_t scale; int32_t act; };   // act >= 0: accumulate, export output row dy0 + act with yscale = scale;
                                                  // -1: accumulate; -2: only its rest counts (scale = the previous export's yscale)
constexpr int kMaxWin = 5;
#ifndef B2_RESCALE_RSF
#define B2_RESCALE_RSF 8
#endif
constexpr int kRSF = B2_RESCALE_RSF;             // source rows per ring stage
#ifndef B2_RESCALE_FT
#define B2_RESCALE_FT 128
#endif
constexpr int kFT = B2_RESCALE_FT;                // threads per CTA: a unit is kFT * kCols output columns
constexpr int kFSegMax = kFT * kCols * 4 + 64;    // source bytes per row and unit (ratio <= 4)
constexpr int kFastSmem = 2 * kRSF * kFSegMax;
constexpr int kFCtas = 768 / kFT;
constexpr int kMaxUnitRows = kRowsGroup * kMaxWin + 8;
__global__ void __launch_bounds__(kFT, kFCtas)
rescale_shrink_fast_kernel(const RescaleJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix, uint32_t n_jobs, uint32_t uniform,
                           uint32_t total, const XEnt* __restrict__ xpool, const YEnt* __restrict__ ypool) {
  extern __shared__ __align__(16) uint8_t ring[];
  __shared__ RowAct acts[kMaxUnitRows];
  __shared__ int info[4];   // lo_al, seg bytes, r_begin, r_end
  const int tid = (int)threadIdx.x;
  for (uint32_t unit = blockIdx.x; unit < total; unit += gridDim.x) {
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, unit, &local);
    const RescaleJobDev j = jobs[img];
    const int xc0 = (int)(local % j.units_x) * kFT * kCols, xc1 = min(xc0 + kFT * kCols, j.dst_width);
    const int dy0 = (int)(local / j.units_x) * kRowsGroup, ng = min(kRowsGroup, j.dst_height - dy0);
    const XEnt* __restrict__ xt = xpool + j.xtab;
    __syncthreads();   // the previous unit is done with acts / info / the ring
    if (tid < 32) {    // the unit's row schedule: lane t owns output row dy0 + t
      YEnt e = YEnt{-1, -1, 0, 0u, 0u};
      if (tid < ng) e = ypool[j.ytab + dy0 + tid];
      const uint32_t valid = __ballot_sync(0xffffffffu, e.last >= 0);
      int r_begin = 0, r_end = -1;
      if (valid) {
        const int first = __ffs(valid) - 1, last = 31 - __clz(valid);
        const int f_prev = __shfl_sync(0xffffffffu, e.prev_last, first);
        const uint32_t f_scale = __shfl_sync(0xffffffffu, e.prev_yscale, first);
        r_end = __shfl_sync(0xffffffffu, e.last, last);
        const bool rest_row = f_prev >= 0 && f_scale != 0u;
        r_begin = rest_row ? f_prev : f_prev + 1;
        if (tid == first && rest_row) acts[0] = RowAct{f_scale, -2};
        if (e.last >= 0) {
          for (int r = e.prev_last + 1; r < e.last; ++r) acts[r - r_begin] = RowAct{0u, -1};
          acts[e.last - r_begin] = RowAct{e.yscale, tid};
        }
      }
      if (tid == 0) {
        const XEnt e0 = xt[xc0], e1 = xt[xc1 - 1];
        const int lo = max(e0.a - 1, 0);
        info[0] = lo & ~15;
        info[1] = e1.b - (lo & ~15);
        info[2] = r_begin; info[3] = r_end;
      }
    }
    __syncthreads();
    const int lo_al = info[0], seg = info[1], r_begin = info[2], r_end = info[3];
    if (r_end < 0) continue;   // uniform: nothing is exported in this group
    const int x0 = xc0 + tid * kCols;
    const bool live = x0 < j.dst_width;
    const int n_valid = min(kCols, j.dst_width - x0);
    const bool word_ok = n_valid == kCols && ((((uintptr_t)j.dst) | (uintptr_t)(uint32_t)j.dst_stride) & 3) == 0;
    // per column: shared-memory address of the window's first word in stage row 0, the 0/1 masks of the window in that
    // word and the next, the byte permute that picks its last sample, -accum
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    uint32_t waddr[kCols], m0[kCols], m1[kCols], sel[kCols], nacc[kCols];
#pragma unroll
    for (int k = 0; k < kCols; ++k) {
      const XEnt e = xt[min(x0 + min(k, max(n_valid - 1, 0)), j.dst_width - 1)];   // columns past the end repeat the last one; they are not stored
      const int rel = e.a - lo_al, s0 = rel & 3, len = e.b - e.a;
      const uint64_t m = (0x0101010101ull & ((1ull << (8 * len)) - 1ull)) << (8 * s0);
      waddr[k] = ring_s + (uint32_t)(rel & ~3);
      m0[k] = (uint32_t)m;
      m1[k] = (uint32_t)(m >> 32);
      sel[k] = 0x4440u | (uint32_t)(s0 + len - 1);   // byte s0 + len - 1 of {next word : word} into byte 0 (the rest is masked off)
      nacc[k] = (uint32_t)(-e.accum);
    }
    // the rest the sample left of the thread's first column hands over: its last source sample times -accum (none: times 0)
    uint32_t prev_addr = ring_s, prev_nacc = 0;
    if (live && x0 > 0) {
      const XEnt p = xt[x0 - 1];
      prev_addr = ring_s + (uint32_t)(p.b - 1 - lo_al);
      prev_nacc = (uint32_t)(-p.accum);
    }
    const int n_chunks = (seg + 15) >> 4;
    const int n_rows = r_end - r_begin + 1, n_stage = (n_rows + kRSF - 1) / kRSF;
    auto issue = [&](int st) {   // all threads: rows r_begin + st * kRSF .. of the unit into ring stage st & 1
      const int rows_here = min(kRSF, n_rows - st * kRSF);
      for (int rr = 0; rr < rows_here; ++rr) {
        const uint8_t* g = j.src + (ptrdiff_t)(r_begin + st * kRSF + rr) * j.src_stride + lo_al;
        const uint32_t d = ring_s + (uint32_t)(((st & 1) * kRSF + rr) * kFSegMax);
        for (int ck = tid; ck < n_chunks; ck += kFT) {
          if (lo_al + 16 * ck + 16 <= j.src_stride) {   // the row's own bytes (and its padding): a whole 16-byte chunk
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d + 16u * ck), "l"(g + 16 * ck) : "memory");
          } else {
            uint8_t* dd = ring + ((st & 1) * kRSF + rr) * kFSegMax;
            for (int b = 0; lo_al + 16 * ck + b < j.src_stride && b < 16; ++b) dd[16 * ck + b] = g[16 * ck + b];
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto lds32 = [](uint32_t a) -> uint32_t { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; };
    auto lds8 = [](uint32_t a) -> uint32_t { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; };
    const uint32_t x_sub = (uint32_t)j.x_sub, fx_scale = j.fx_scale, fxy_scale = j.fxy_scale;
    uint32_t acc[kCols] = {0, 0, 0, 0};
    issue(0);
    for (int st = 0; st < n_stage; ++st) {
      if (st + 1 < n_stage) {
        issue(st + 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();
      const int rows_here = min(kRSF, n_rows - st * kRSF);
      if (live) {
        uint32_t roff = (uint32_t)((st & 1) * kRSF * kFSegMax);
        for (int rr = 0; rr < rows_here; ++rr, roff += kFSegMax) {
          const RowAct ra = acts[st * kRSF + rr];
          uint32_t fn[kCols];
          uint32_t carry = mult_fix(lds8(prev_addr + roff) * prev_nacc, fx_scale);
#pragma unroll
          for (int k = 0; k < kCols; ++k) {
            const uint32_t w0 = lds32(waddr[k] + roff), w1 = lds32(waddr[k] + roff + 4u);
            const uint32_t sum = __dp4a(w1, m1[k], __dp4a(w0, m0[k], carry));
            const uint32_t base = __byte_perm(w0, w1, sel[k]) & 0xffu;
            const uint32_t frac = base * nacc[k];
            fn[k] = sum * x_sub - frac;
            carry = mult_fix(frac, fx_scale);
          }
          // the rest this row would leave for the next output row: used by an exporting row and by the row the
          // previous unit's last export ended on (scale 0 otherwise: no rest)
          const bool only_rest = ra.act == -2, exports = ra.act >= 0;
          uint32_t rest[kCols];
#pragma unroll
          for (int k = 0; k < kCols; ++k) {
            rest[k] = mult_fix_floor(fn[k], ra.scale);
            acc[k] = only_rest ? 0u : acc[k] + fn[k];
          }
          if (exports) {                             // ExportRowShrink, rescaler.c.go:138-163
            uint32_t v[kCols];
#pragma unroll
            for (int k = 0; k < kCols; ++k) v[k] = min(mult_fix(acc[k] - rest[k], fxy_scale), 255u);
            uint8_t* out = j.dst + (ptrdiff_t)(dy0 + ra.act) * j.dst_stride + x0;
            if (word_ok) *reinterpret_cast<uint32_t*>(out) = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
            else for (int k = 0; k < n_valid; ++k) out[k] = (uint8_t)v[k];
          }
          if (only_rest || exports) {
#pragma unroll
            for (int k = 0; k < kCols; ++k) acc[k] = rest[k];
          }
        }
      }
      __syncthreads();   // the stage may be refilled
    }
  }
}

// ---- one-channel planes expanding both ways: the unit's source rows staged at once ------------------
// Expanding, an output row needs the horizontally expanded source row `last` and the one before it
// (ExportRowExpand, rescaler.c.go:110-136), and `last` moves by at most one row per output row: a unit of
// 32 output rows reads at most 34 source rows of at most 512 + a few samples -- all of them go to shared
// memory in one wave of 16-byte cp.async.  A thread's four output columns take their left / right samples
// (ImportRowExpand :39-72) from the two aligned words that hold them (byte permutes computed once per
// unit), the row walk is `advance while have < last`, the vertical blend is two 32 x 32 -> 64 products.
#ifndef B2_RESCALE_ONE_PRODUCT
#define B2_RESCALE_ONE_PRODUCT 1
#endif
struct ExpAct { uint32_t B; int32_t last; };      // last < 0: the row is never exported; B == 0: y_accum == 0 (J = frow)
constexpr int kERows = 40;
constexpr int kESeg = kFT * kCols + 64;
constexpr int kExpSmem = kERows * kESeg;
#ifndef B2_RESCALE_ECTAS
#define B2_RESCALE_ECTAS 8
#endif
constexpr int kECtas = B2_RESCALE_ECTAS;
__global__ void __launch_bounds__(kFT, kECtas)
rescale_expand_fast_kernel(const RescaleJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix, uint32_t n_jobs, uint32_t uniform,
                           uint32_t total, const XEnt* __restrict__ xpool, const YEnt* __restrict__ ypool) {
  extern __shared__ __align__(16) uint8_t ring[];
  __shared__ ExpAct acts[kRowsGroup];
  __shared__ int info[4];   // lo_al, seg bytes, r_begin, r_end
  const int tid = (int)threadIdx.x;
  const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
  for (uint32_t unit = blockIdx.x; unit < total; unit += gridDim.x) {
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, unit, &local);
    const RescaleJobDev j = jobs[img];
    const int xc0 = (int)(local % j.units_x) * kFT * kCols, xc1 = min(xc0 + kFT * kCols, j.dst_width);
    const int dy0 = (int)(local / j.units_x) * kRowsGroup, ng = min(kRowsGroup, j.dst_height - dy0);
    const XEnt* __restrict__ xt = xpool + j.xtab;
    __syncthreads();   // the previous unit is done with acts / info / the ring
    if (tid < 32) {
      YEnt e = YEnt{-1, -1, 0, 0u, 0u};
      if (tid < ng) e = ypool[j.ytab + dy0 + tid];
      acts[tid] = ExpAct{e.y_accum != 0 ? e.yscale : 0u, e.last};
      const uint32_t valid = __ballot_sync(0xffffffffu, e.last >= 0);
      if (tid == 0) { info[2] = 0; info[3] = -1; }
      if (valid) {
        const int first = __ffs(valid) - 1, last = 31 - __clz(valid);
        const int l0 = __shfl_sync(0xffffffffu, e.last, first), l1 = __shfl_sync(0xffffffffu, e.last, last);
        if (tid == 0) { info[2] = max(l0 - 1, 0); info[3] = l1; }
      }
      if (tid == 0) {
        const XEnt e0 = xt[xc0], e1 = xt[xc1 - 1];
        info[0] = e0.a & ~15;
        info[1] = e1.b + 1 - (e0.a & ~15);
      }
    }
    __syncthreads();
    const int lo_al = info[0], seg = info[1], r_begin = info[2], r_end = info[3];
    if (r_end < 0) continue;   // uniform: nothing is exported in this group
    {   // every source row of the unit, 16 bytes per copy
      const int n_chunks = (seg + 15) >> 4, n_rows = r_end - r_begin + 1;
      for (int rr = 0; rr < n_rows; ++rr) {
        const uint8_t* g = j.src + (ptrdiff_t)(r_begin + rr) * j.src_stride + lo_al;
        for (int ck = tid; ck < n_chunks; ck += kFT) {
          if (lo_al + 16 * ck + 16 <= j.src_stride) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(ring_s + (uint32_t)(rr * kESeg + 16 * ck)), "l"(g + 16 * ck) : "memory");
          } else {
            for (int b = 0; lo_al + 16 * ck + b < j.src_stride && b < 16; ++b) ring[rr * kESeg + 16 * ck + b] = g[16 * ck + b];
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const int x0 = xc0 + tid * kCols;
    const bool live = x0 < j.dst_width;
    const int n_valid = min(kCols, j.dst_width - x0);
    const bool word_ok = n_valid == kCols && ((((uintptr_t)j.dst) | (uintptr_t)(uint32_t)j.dst_stride) & 3) == 0;
    // the thread's samples lie in the aligned word that holds its first left sample and the next one
    const XEnt e_first = xt[min(x0, j.dst_width - 1)];
    const uint32_t wbase = ring_s + (uint32_t)((e_first.a - lo_al) & ~3);
    uint32_t sel_l[kCols], sel_r[kCols], c_l[kCols], c_r[kCols];
#pragma unroll
    for (int k = 0; k < kCols; ++k) {
      const XEnt e = xt[min(x0 + min(k, max(n_valid - 1, 0)), j.dst_width - 1)];
      const int base = (e_first.a - lo_al) & ~3;
      sel_l[k] = 0x4440u | (uint32_t)(e.a - lo_al - base);   // bytes 0 .. 7 of {word 1 : word 0}
      sel_r[k] = 0x4440u | (uint32_t)(e.b - lo_al - base);
      // frow = right * x_add + (left - right) * accum = left * accum + right * (x_add - accum)
      c_l[k] = (uint32_t)e.accum;
      c_r[k] = (uint32_t)j.x_add - (uint32_t)e.accum;
    }
    auto lds32 = [](uint32_t a) -> uint32_t { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; };
    const uint32_t fy_scale = j.fy_scale;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (!live) continue;
    uint32_t f[kCols] = {0, 0, 0, 0}, g[kCols] = {0, 0, 0, 0};
    int have = r_begin - 1;
    for (int dy = 0; dy < ng; ++dy) {
      const ExpAct a = acts[dy];
      if (a.last < 0) continue;
      while (have < a.last) {
        ++have;
        const uint32_t ra = wbase + (uint32_t)((have - r_begin) * kESeg);
        const uint32_t w0 = lds32(ra), w1 = lds32(ra + 4u);
#pragma unroll
        for (int k = 0; k < kCols; ++k) {
          g[k] = f[k];
          const uint32_t left = __byte_perm(w0, w1, sel_l[k]) & 0xffu, right = __byte_perm(w0, w1, sel_r[k]) & 0xffu;
          f[k] = left * c_l[k] + right * c_r[k];
        }
      }
      uint32_t v[kCols];
#pragma unroll
      for (int k = 0; k < kCols; ++k) {
        uint32_t J = f[k];
        if (a.B != 0u) {                          // y_accum != 0: blend with the previous row's frow (zero above the plane)
          const uint32_t i = (a.last > 0) ? g[k] : 0u;
#if B2_RESCALE_ONE_PRODUCT
          // A * f + B * i with A = 2^32 - B (rescaler.c.go:121-124) is 2^32 * f + B * (i - f): ONE 32 x 32 -> 64 product
          // instead of two.  i - f wraps when it is negative, which adds B * 2^32 to the product: taken off the high word.
          const uint32_t du = i - f[k];
          const uint32_t hi = (uint32_t)(((uint64_t)a.B * du + kRounder) >> 32);
          J = f[k] + hi - (a.B & (uint32_t)((int32_t)du >> 31));
#else
          J = (uint32_t)(((uint64_t)(uint32_t)(kOne - a.B) * f[k] + (uint64_t)a.B * i + kRounder) >> 32);
#endif
        }
        v[k] = min(mult_fix(J, fy_scale), 255u);
      }
      uint8_t* out = j.dst + (ptrdiff_t)(dy0 + dy) * j.dst_stride + x0;
      if (word_ok) *reinterpret_cast<uint32_t*>(out) = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
      else for (int k = 0; k < n_valid; ++k) out[k] = (uint8_t)v[k];
    }
  }
}

struct Geometry {
  int x_expand, y_expand, x_add, x_sub, y_add, y_sub;
  uint32_t fx_scale, fy_scale, fxy_scale;
};

// WebPRescalerInit, rescaler_utils.c.go:31-86
Geometry init_geometry(int sw, int sh, int dw, int dh) {
  Geometry g;
  g.x_expand = sw < dw;
  g.y_expand = sh < dh;
  g.x_add = g.x_expand ? (dw - 1) : sw;
  g.x_sub = g.x_expand ? (sw - 1) : dw;
  g.fx_scale = g.x_expand ? 0u : (uint32_t)(kOne / (uint64_t)g.x_sub);
  g.y_add = g.y_expand ? (sh - 1) : sh;
  g.y_sub = g.y_expand ? (dh - 1) : dh;
  g.fxy_scale = 0;
  if (!g.y_expand) {
    const uint64_t ratio = ((uint64_t)dh * kOne) / ((uint64_t)g.x_add * (uint64_t)g.y_add);
    g.fxy_scale = (ratio != (uint32_t)ratio) ? 0u : (uint32_t)ratio;
    g.fy_scale = (uint32_t)(kOne / (uint64_t)g.y_sub);
  } else {
    g.fy_scale = (uint32_t)(kOne / (uint64_t)g.x_add);
  }
  return g;
}

// the horizontal accumulator walk of ImportRowExpand / ImportRowShrink
void build_xtab(const Geometry& g, int sw, int dw, std::vector<XEnt>* out) {
  if (g.x_expand) {
    int x_in = 0, accum = g.x_add;
    int left = 0, right = (sw > 1) ? 1 : 0;
    x_in = right;
    for (int xo = 0;; ) {
      out->push_back(XEnt{left, right, accum, 0});
      if (++xo >= dw) break;
      accum -= g.x_sub;
      if (accum < 0) {
        left = right;
        x_in += 1;
        right = x_in;
        accum += g.x_add;
      }
    }
  } else {
    int x_in = 0, accum = 0;
    for (int xo = 0; xo < dw; ++xo) {
      const int begin = x_in;
      accum += g.x_add;
      while (accum > 0) { accum -= g.x_sub; ++x_in; }
      out->push_back(XEnt{begin, x_in, accum, 0});
    }
  }
}

// the vertical walk of WebPRescalerImport / WebPRescalerExport (rescaler_utils.c.go:119-158)
void build_ytab(const Geometry& g, int sh, int dh, std::vector<YEnt>* out) {
  const size_t first = out->size();
  out->resize(first + (size_t)dh, YEnt{-1, -1, 0, 0u, 0u});
  int y_accum = g.y_expand ? g.y_sub : g.y_add;
  int src_y = 0, dst_y = 0, last = -1, prev_last = -1;
  uint32_t prev_yscale = 0;
  while (src_y < sh) {
    while (src_y < sh && !(dst_y < dh && y_accum <= 0)) {
      last = src_y++;
      y_accum -= g.y_sub;
    }
    while (dst_y < dh && y_accum <= 0) {
      const uint32_t yscale = g.y_expand ? ((y_accum != 0) ? (uint32_t)(((uint64_t)(uint32_t)(-y_accum) << 32) / (uint64_t)g.y_sub) : 0u)
                                         : g.fy_scale * (uint32_t)(-y_accum);
      (*out)[first + (size_t)dst_y] = YEnt{prev_last, last, y_accum, prev_yscale, yscale};
      prev_last = last;
      prev_yscale = yscale;
      y_accum += g.y_add;
      ++dst_y;
    }
  }
}

}  // namespace

extern "C" int b2_rescale_planes(b2_ctx* ctx, int stream, const b2_rescale_job* jobs, int n_jobs) {
  B2_CHECK_CTX(ctx, stream);
  if (n_jobs < 0) return B2_ERR_INVALID_CONFIG;
  if (n_jobs == 0) return B2_OK;
  if (jobs == NULL) return B2_ERR_NULL;
  std::vector<RescaleJobDev> dj((size_t)n_jobs);
  std::vector<uint32_t> pre((size_t)n_jobs + 1);
  std::vector<XEnt> xpool;
  std::vector<YEnt> ypool;
  std::map<std::tuple<int, int, int, int>, std::pair<uint32_t, uint32_t>> seen;   // geometry -> table offsets
  uint64_t tot = 0;
  uint32_t uni = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const b2_rescale_job& j = jobs[i];
    if (j.src == NULL || j.dst == NULL) return B2_ERR_NULL;
    if (j.src_width <= 0 || j.src_height <= 0 || j.dst_width <= 0 || j.dst_height <= 0 || j.src_width > B2_MAX_DIMENSION ||
        j.src_height > B2_MAX_DIMENSION || j.dst_width > B2_MAX_DIMENSION || j.dst_height > B2_MAX_DIMENSION) return B2_ERR_BAD_DIM;
    if (j.num_channels < 1 || j.num_channels > 4) return B2_ERR_INVALID_CONFIG;
    if (j.src_stride < j.src_width * j.num_channels || j.dst_stride < j.dst_width * j.num_channels) return B2_ERR_BAD_DIM;
    const Geometry g = init_geometry(j.src_width, j.src_height, j.dst_width, j.dst_height);
    const auto key = std::make_tuple(j.src_width, j.src_height, j.dst_width, j.dst_height);
    auto it = seen.find(key);
    if (it == seen.end()) {
      const uint32_t xo = (uint32_t)xpool.size(), yo = (uint32_t)ypool.size();
      build_xtab(g, j.src_width, j.dst_width, &xpool);
      build_ytab(g, j.src_height, j.dst_height, &ypool);
      it = seen.emplace(key, std::make_pair(xo, yo)).first;
    }
    RescaleJobDev& d = dj[(size_t)i];
    d.src = j.src; d.dst = j.dst; d.src_stride = j.src_stride; d.dst_stride = j.dst_stride;
    d.dst_width = j.dst_width; d.dst_height = j.dst_height; d.num_channels = j.num_channels;
    d.x_expand = g.x_expand; d.y_expand = g.y_expand; d.x_add = g.x_add; d.x_sub = g.x_sub; d.y_sub = g.y_sub;
    d.fx_scale = g.fx_scale; d.fy_scale = g.fy_scale; d.fxy_scale = g.fxy_scale;
    d.xtab = it->second.first; d.ytab = it->second.second;
    d.units_x = (uint32_t)((j.dst_width * j.num_channels + kThreads * kPerThread - 1) / (kThreads * kPerThread));
  }
  // batches whose jobs all share the direction and have one channel take a specialised instantiation
  int cls = dj[0].num_channels == 1 ? (dj[0].x_expand * 2 + dj[0].y_expand) : -1;
  for (int i = 1; i < n_jobs && cls >= 0; ++i)
    if (dj[(size_t)i].num_channels != 1 || dj[(size_t)i].x_expand * 2 + dj[(size_t)i].y_expand != cls) cls = -1;
  const bool columns = cls >= 0 && getenv("B2_RESCALE_SAMPLES") == NULL;   // one-channel planes of one direction: the column walkers
  // the streamed form needs 16-byte aligned source rows and a bounded source segment per 1024-column unit
  bool stream_flag = columns && getenv("B2_RESCALE_COLUMNS") == NULL;
  for (int i = 0; i < n_jobs && stream_flag; ++i) {
    const b2_rescale_job& j = jobs[i];
    if ((((uintptr_t)j.src | (uintptr_t)(uint32_t)j.src_stride) & 15) != 0) { stream_flag = false; break; }
    const XEnt* xt = xpool.data() + dj[(size_t)i].xtab;
    for (int xc0 = 0; xc0 < j.dst_width; xc0 += kThreads * kCols) {
      const int xc1 = (xc0 + kThreads * kCols < j.dst_width) ? xc0 + kThreads * kCols : j.dst_width;
      const int lo = dj[(size_t)i].x_expand ? xt[xc0].a : (xt[xc0].a > 0 ? xt[xc0].a - 1 : 0);
      const int hi = dj[(size_t)i].x_expand ? xt[xc1 - 1].b + 1 : xt[xc1 - 1].b;
      if (hi - (lo & ~15) + 16 > kSegMax) { stream_flag = false; break; }
    }
  }
  // shrinking both ways with windows of at most kMaxWin samples: the loop-free kernel
  bool shrink_fast = stream_flag && cls == 0 && getenv("B2_RESCALE_NOFAST") == NULL;
  for (int i = 0; i < n_jobs && shrink_fast; ++i) {
    const RescaleJobDev& d = dj[(size_t)i];
    if (d.fxy_scale == 0u) { shrink_fast = false; break; }
    const XEnt* xt = xpool.data() + d.xtab;
    for (int x = 0; x < d.dst_width && shrink_fast; ++x) if (xt[x].b - xt[x].a > kMaxWin) shrink_fast = false;
    const YEnt* yt = ypool.data() + d.ytab;
    for (int y = 0; y < d.dst_height && shrink_fast; ++y) if (yt[y].last >= 0 && yt[y].last - yt[y].prev_last > kMaxWin) shrink_fast = false;
  }
  for (int i = 0; i < n_jobs && shrink_fast; ++i) {   // the source segment of a unit (one sample of the previous unit included)
    const RescaleJobDev& d = dj[(size_t)i];
    const XEnt* xt = xpool.data() + d.xtab;
    for (int xc0 = 0; xc0 < d.dst_width && shrink_fast; xc0 += kFT * kCols) {
      const int xc1 = (xc0 + kFT * kCols < d.dst_width) ? xc0 + kFT * kCols : d.dst_width;
      const int lo = xt[xc0].a > 0 ? xt[xc0].a - 1 : 0;
      if (xt[xc1 - 1].b - (lo & ~15) + 16 > kFSegMax) shrink_fast = false;
    }
  }
  // expanding both ways: the kernel that stages a unit's source rows at once
  bool expand_fast = columns && cls == 3 && getenv("B2_RESCALE_NOFAST") == NULL;
  for (int i = 0; i < n_jobs && expand_fast; ++i) {
    const b2_rescale_job& j = jobs[i];
    const RescaleJobDev& d = dj[(size_t)i];
    if ((((uintptr_t)j.src | (uintptr_t)(uint32_t)j.src_stride) & 15) != 0) { expand_fast = false; break; }
    const XEnt* xt = xpool.data() + d.xtab;
    for (int xc0 = 0; xc0 < d.dst_width && expand_fast; xc0 += kFT * kCols) {
      const int xc1 = (xc0 + kFT * kCols < d.dst_width) ? xc0 + kFT * kCols : d.dst_width;
      if (xt[xc1 - 1].b + 1 - (xt[xc0].a & ~15) + 16 > kESeg) expand_fast = false;
      // a thread's four columns must find their samples in two aligned words
      for (int x = xc0; x < xc1 && expand_fast; x += kCols) {
        const int base = xt[x].a & ~3, xe = (x + kCols - 1 < d.dst_width) ? x + kCols - 1 : d.dst_width - 1;
        if (xt[xe].b - base > 7) expand_fast = false;
      }
    }
    const YEnt* yt = ypool.data() + d.ytab;
    for (int y0 = 0; y0 < d.dst_height && expand_fast; y0 += kRowsGroup) {
      int lo = -1, hi = -1;
      for (int y = y0; y < y0 + kRowsGroup && y < d.dst_height; ++y) {
        if (yt[y].last < 0) continue;
        if (lo < 0) lo = yt[y].last;
        if (yt[y].last < hi) expand_fast = false;     // the walk only moves forward
        hi = yt[y].last;
      }
      if (hi >= 0 && hi - (lo > 0 ? lo - 1 : 0) + 1 > kERows) expand_fast = false;
    }
  }
  const int rows = columns ? kRowsGroup : rows_per_unit(cls < 0 ? -1 : (cls & 1));
  const int unit_cols = (shrink_fast || expand_fast) ? kFT * kCols : kThreads * kCols;
  for (int i = 0; i < n_jobs; ++i) {
    if (columns) dj[(size_t)i].units_x = (uint32_t)((dj[(size_t)i].dst_width + unit_cols - 1) / unit_cols);
    const uint32_t items = dj[(size_t)i].units_x * (uint32_t)((dj[(size_t)i].dst_height + rows - 1) / rows);
    if (i == 0) uni = items; else if (items != uni) uni = 0;
    pre[(size_t)i] = (uint32_t)tot;
    tot += items;
    if (tot > 0x7fffffffull) return B2_ERR_BAD_DIM;
  }
  pre[(size_t)n_jobs] = (uint32_t)tot;
  const size_t jb = (sizeof(RescaleJobDev) * (size_t)n_jobs + 15) & ~(size_t)15;
  const size_t pb = (sizeof(uint32_t) * ((size_t)n_jobs + 1) + 15) & ~(size_t)15;
  const size_t xb = (sizeof(XEnt) * xpool.size() + 15) & ~(size_t)15;
  const size_t yb = sizeof(YEnt) * ypool.size();
  if (jb + pb + xb + yb > B2_DESC_RING_BYTES) return B2_ERR_INVALID_CONFIG;   // too many distinct geometries for one call
  std::vector<uint8_t> host(jb + pb + xb + yb);
  memcpy(host.data(), dj.data(), sizeof(RescaleJobDev) * (size_t)n_jobs);
  memcpy(host.data() + jb, pre.data(), sizeof(uint32_t) * ((size_t)n_jobs + 1));
  memcpy(host.data() + jb + pb, xpool.data(), sizeof(XEnt) * xpool.size());
  memcpy(host.data() + jb + pb + xb, ypool.data(), yb);
  void* dev = NULL;
  const int st = b2_upload_desc(ctx, stream, host.data(), host.size(), &dev);
  if (st != B2_OK) return st;
  const uint8_t* base = (const uint8_t*)dev;
#define B2_RESCALE_LAUNCH(XE, YE, CH)                                                                            \
  rescale_kernel<XE, YE, CH><<<b2_grid_for(ctx, 4, tot), kThreads, 0, ctx->streams[stream]>>>( \
      (const RescaleJobDev*)base, (const uint32_t*)(base + jb), (uint32_t)n_jobs, uni, (uint32_t)tot,             \
      (const XEnt*)(base + jb + pb), (const YEnt*)(base + jb + pb + xb))
#define B2_RESCALE_COLUMNS(XE, YE)                                                                              \
  rescale_columns_kernel<XE, YE><<<b2_grid_for(ctx, 3, tot), kThreads, 0, ctx->streams[stream]>>>(                 \
      (const RescaleJobDev*)base, (const uint32_t*)(base + jb), (uint32_t)n_jobs, uni, (uint32_t)tot,             \
      (const XEnt*)(base + jb + pb), (const YEnt*)(base + jb + pb + xb))
#define B2_RESCALE_STREAM(XE, YE)                                                                               \
  do {                                                                                                            \
    B2_CUDA(ctx, cudaFuncSetAttribute(rescale_stream_kernel<XE, YE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmem)); \
    rescale_stream_kernel<XE, YE><<<b2_grid_for(ctx, 3, tot), kThreads, kStreamSmem, ctx->streams[stream_id]>>>(     \
        (const RescaleJobDev*)base, (const uint32_t*)(base + jb), (uint32_t)n_jobs, uni, (uint32_t)tot,           \
        (const XEnt*)(base + jb + pb), (const YEnt*)(base + jb + pb + xb));                                       \
  } while (0)
  const int stream_id = stream;
  if (shrink_fast) {
    B2_CUDA(ctx, cudaFuncSetAttribute(rescale_shrink_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFastSmem));
    rescale_shrink_fast_kernel<<<b2_grid_for(ctx, kFCtas, tot), kFT, kFastSmem, ctx->streams[stream_id]>>>(
        (const RescaleJobDev*)base, (const uint32_t*)(base + jb), (uint32_t)n_jobs, uni, (uint32_t)tot,
        (const XEnt*)(base + jb + pb), (const YEnt*)(base + jb + pb + xb));
    ctx->launches++;
    B2_CUDA(ctx, cudaGetLastError());
    return B2_OK;
  }
  if (expand_fast) {
    B2_CUDA(ctx, cudaFuncSetAttribute(rescale_expand_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kExpSmem));
    rescale_expand_fast_kernel<<<b2_grid_for(ctx, kECtas, tot), kFT, kExpSmem, ctx->streams[stream_id]>>>(
        (const RescaleJobDev*)base, (const uint32_t*)(base + jb), (uint32_t)n_jobs, uni, (uint32_t)tot,
        (const XEnt*)(base + jb + pb), (const YEnt*)(base + jb + pb + xb));
    ctx->launches++;
    B2_CUDA(ctx, cudaGetLastError());
    return B2_OK;
  }
  if (stream_flag) {
    switch (cls) {
      case 0: B2_RESCALE_STREAM(0, 0); break;
      case 1: B2_RESCALE_STREAM(0, 1); break;
      case 2: B2_RESCALE_STREAM(1, 0); break;
      default: B2_RESCALE_STREAM(1, 1); break;
    }
    ctx->launches++;
    B2_CUDA(ctx, cudaGetLastError());
    return B2_OK;
  }
#undef B2_RESCALE_STREAM
  if (columns) {
    switch (cls) {
      case 0: B2_RESCALE_COLUMNS(0, 0); break;
      case 1: B2_RESCALE_COLUMNS(0, 1); break;
      case 2: B2_RESCALE_COLUMNS(1, 0); break;
      default: B2_RESCALE_COLUMNS(1, 1); break;
    }
    ctx->launches++;
    B2_CUDA(ctx, cudaGetLastError());
    return B2_OK;
  }
#undef B2_RESCALE_COLUMNS
  switch (cls) {
    case 0: B2_RESCALE_LAUNCH(0, 0, 1); break;
    case 1: B2_RESCALE_LAUNCH(0, 1, 1); break;
    case 2: B2_RESCALE_LAUNCH(1, 0, 1); break;
    case 3: B2_RESCALE_LAUNCH(1, 1, 1); break;
    default: B2_RESCALE_LAUNCH(-1, -1, -1); break;
  }
#undef B2_RESCALE_LAUNCH
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}
