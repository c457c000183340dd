// b2_dequant.cu -- alpha level smoothing: WebPDequantizeLevels
// (pkg/libwebp/utils/quant_levels_dec_utils.c.go:259-286), which ALPHDecode runs over the
// decoded alpha plane when the ALPH chunk says its levels were quantised and the caller asked
// for alpha dithering (pkg/libwebp/decoder/alpha_dec.go:198-203).
//
// The reference streams rows through a ring of cumulative sums (VFilter :86-111, HFilter
// :115-136, all modulo 2^16).  Restated per pixel: with cs[k] the sum of column k over rows
// y-r .. y+r (rows above the plane replicate row 0; the last r rows of the plane are never
// emitted, so rows below it are never needed), the box value of pixel x is
//     x <= r      : sum cs[0 .. x+r-1] + sum cs[0 .. r-x]          (the reference's left mirror)
//     x >= w - r  : sum cs[2w-1-r-x .. w-1] + sum cs[x-r .. w-1]   (right mirror)
//     otherwise   : sum cs[x-r .. x+r]
// i.e. with cs mirrored about both edges (column -1-k for k < 0, 2w-1-k for k >= w) a plain
// window centred on x, except that columns x <= r take the window centred on max(x-1, 0)
// at most (2r+1)^2 * 255 < 2^16, so the modulo never bites.  average = (box * scale) >> 16,
// then ApplyFilter (:139-157) through the correction table (InitCorrectionLUT :162-184) built
// from the smallest distance between two used levels (CountLevels :186-219).
//
// Two launches: used-level bitmap per plane (256 bits, atomicOr), then the smoother, whose CTAs
// first turn the bitmap into min / max / min distance / count (count_levels): a warp owns 120
// output columns x 32 rows, each lane keeps the vertical sums of its 4 columns in registers
// (sliding down the band), neighbours' sums come by shuffle (r <= 4: one lane each side, the
// two outer lanes of a warp are halo only), the correction table sits in shared memory.
// Out of place: the reference filters in place, but it reads row y+r before it writes row y,
// so reading the untouched source is the same computation.
#include <stdlib.h>

#include "b2_common.cuh"

namespace {

constexpr int kThreads = 256, kWarps = kThreads / 32;
constexpr int kTileCols = 120;   // output columns per warp (lanes 1..30 x 4)
constexpr int kBandRows = 32;
constexpr int kLfix = 2, kFix = 16;
constexpr int kLutSize = (1 << (8 + kLfix)) - 1;

struct DqJobDev {
  const uint8_t* src;
  uint8_t* dst;
  int32_t src_stride, dst_stride, width, height, radius;
  uint32_t tiles_x, units;
};
struct DqStats {          // per plane: 8 words of used-level bits, then CountLevels' results
  uint32_t used[8];
  int32_t min, max, min_dist, num_levels;
};

// 4 bytes of row 'row' at columns x0 .. x0+3, any alignment.  Columns outside the plane are
// mirrored about its edges (column -1-k for k < 0, 2w-1-k for k >= w; zero where that falls
// outside too): the extension on which the reference's edge sums become plain windows.
__device__ __noinline__ uint32_t load4_edge(const uint8_t* __restrict__ p, int width, int x0) {
  uint32_t v = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int c = x0 + k;
    c = c < 0 ? -1 - c : (c >= width ? 2 * width - 1 - c : c);
    if (c >= 0 && c < width) v |= (uint32_t)__ldg(p + c) << (8 * k);
  }
  return v;
}

// Used-level bits of one plane: flags in shared memory, one ballot per warp at the end.
// Alpha planes are mostly flat: a word equal to the previous one, or made of one byte value,
// costs one store or none.
__device__ __forceinline__ void flag_word(uint8_t* flag, uint32_t v, uint32_t& last, bool& have) {
  if (have && v == last) return;
  last = v;
  have = true;
  const uint32_t b0 = v & 0xff;
  flag[b0] = 1;
  if (v != b0 * 0x01010101u) {
    flag[(v >> 8) & 0xff] = 1;
    flag[(v >> 16) & 0xff] = 1;
    flag[v >> 24] = 1;
  }
}

__global__ void __launch_bounds__(kThreads)
dq_levels_kernel(const DqJobDev* __restrict__ jobs, DqStats* __restrict__ stats) {
  __shared__ uint8_t flag[256];
  const DqJobDev j = jobs[blockIdx.y];
  if (j.radius == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  flag[threadIdx.x] = 0;
  __syncthreads();
  for (int row = blockIdx.x * kWarps + warp; row < j.height; row += gridDim.x * kWarps) {
    const uint8_t* p = j.src + (size_t)row * j.src_stride;
    int x = 0;
    if ((((uintptr_t)p) & 15) == 0) {
      const int nvec = j.width >> 4;
      uint32_t last = 0;
      bool have = false;
      for (int i = lane; i < nvec; i += 32) {
        const uint4 v = b2_ldg_stream_v4(p + 16 * i);
        flag_word(flag, v.x, last, have);
        flag_word(flag, v.y, last, have);
        flag_word(flag, v.z, last, have);
        flag_word(flag, v.w, last, have);
      }
      x = nvec << 4;
    }
    for (x += lane; x < j.width; x += 32) flag[__ldg(p + x)] = 1;
  }
  __syncthreads();
  const uint32_t bits = __ballot_sync(0xffffffffu, flag[threadIdx.x] != 0);
  if (lane == 0 && bits) atomicOr(&stats[blockIdx.y].used[warp], bits);
}

// CountLevels (:186-219) from the used-level bitmap, by the whole CTA: thread t owns level t; a used level looks
// for the used level below it (its own word first, then the words below), the smallest gap is reduced over the CTA.
__device__ __forceinline__ void count_levels(const uint32_t* __restrict__ used, int* sh /*[4 + 2 * kWarps]*/) {
  const int t = threadIdx.x, word = t >> 5, bit = t & 31;
  const uint32_t mine = __ldg(used + word);
  int gap = 255;
  const bool is_used = (mine >> bit) & 1u;
  if (is_used) {
    uint32_t below = mine & ((1u << bit) - 1u);
    int wd = word;
    while (below == 0u && wd > 0) below = __ldg(used + --wd);
    if (below != 0u) gap = t - (wd * 32 + 31 - __clz(below));
  }
  const uint32_t ballot = __ballot_sync(0xffffffffu, is_used);
  for (int o = 16; o > 0; o >>= 1) gap = min(gap, __shfl_xor_sync(0xffffffffu, gap, o));
  if (bit == 0) {
    sh[4 + word] = gap;
    sh[4 + kWarps + word] = (int)ballot;
  }
  __syncthreads();
  if (t == 0) {
    int mn = 255, mx = 0, n = 0, dist = 255;
    for (int k = 0; k < kWarps; ++k) {
      const uint32_t bts = (uint32_t)sh[4 + kWarps + k];
      if (bts) {
        if (n == 0) mn = k * 32 + __ffs(bts) - 1;
        mx = k * 32 + 31 - __clz(bts);
        n += __popc(bts);
      }
      dist = min(dist, sh[4 + k]);
    }
    if (n == 0) { mn = 255; mx = 0; }
    if (mx - mn < dist) dist = mx - mn;
    sh[0] = mn; sh[1] = mx; sh[2] = dist; sh[3] = n;
  }
  __syncthreads();
}

template <int DFIX>
__global__ void __launch_bounds__(kThreads)
dq_smooth_kernel(const DqJobDev* __restrict__ jobs, const DqStats* __restrict__ stats) {
  __shared__ int16_t lut[2 * kLutSize + 1];
  const DqJobDev j = jobs[blockIdx.y];
  if (blockIdx.x * kWarps >= j.units) return;
  __shared__ int lv[4 + 2 * kWarps];   // min, max, min distance, number of levels; scratch
  count_levels(stats[blockIdx.y].used, lv);
  const int vmin = lv[0], vmax = lv[1];
  const bool active = j.radius > 0 && lv[3] > 2;
  if (active) {   // InitCorrectionLUT
    const int threshold1 = lv[2] << kLfix;
    const int threshold2 = (3 * threshold1) >> 2;
    const int max_threshold = threshold2 << DFIX;
    const int delta = threshold1 - threshold2;
    for (int i = threadIdx.x; i <= kLutSize; i += kThreads) {
      int c = 0;
      if (i > 0) {
        c = (i <= threshold2) ? (i << DFIX) : (i < threshold1) ? max_threshold * (threshold1 - i) / delta : 0;
        c >>= kLfix;
      }
      lut[kLutSize + i] = (int16_t)c;
      lut[kLutSize - i] = (int16_t)-c;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t unit = blockIdx.x * kWarps + warp;
  if (unit >= j.units) return;
  const int w = j.width, h = j.height, r = j.radius;
  const int x0 = (int)(unit % j.tiles_x) * kTileCols - 4 + 4 * lane;
  const int yb = (int)(unit / j.tiles_x) * kBandRows, ye = min(yb + kBandRows, h);
  const bool writer = lane >= 1 && lane <= 30 && x0 < w;
  const int filt_end = active ? h - r : 0;          // rows [0, filt_end) are smoothed
  const uint32_t scale16 = ((1u << (kFix + kLfix)) / (uint32_t)((2 * r + 1) * (2 * r + 1))) << 16;   // average = umulhi(box, scale16)
  const uint32_t m2 = r >= 2 ? ~0u : 0u, m3 = r >= 3 ? ~0u : 0u, m4 = r >= 4 ? ~0u : 0u;
  const bool inside = x0 >= 0 && x0 + 3 < w;
  const bool fast_src = inside && ((((uintptr_t)(j.src + x0)) | (uint32_t)j.src_stride) & 3) == 0;
  const bool fast_dst = inside && ((((uintptr_t)(j.dst + x0)) | (uint32_t)j.dst_stride) & 3) == 0;
  auto ld = [&](int row) -> uint32_t {
    const uint8_t* p = j.src + (size_t)row * j.src_stride;
    return fast_src ? __ldg((const uint32_t*)(p + x0)) : load4_edge(p, w, x0);
  };
  // keep = vmin < v < vmax as one unsigned compare
  const uint32_t keep_lo = (uint32_t)(vmin + 1), keep_span = (uint32_t)(vmax - vmin - 1);
  const uint32_t lut_mid = (uint32_t)__cvta_generic_to_shared(lut) + 2u * kLutSize;

  // vertical sums of my 4 columns, two 16-bit fields per word (fields stay below 2^16)
  uint32_t cs01 = 0, cs23 = 0;
  if (yb < filt_end) {
    for (int k = -r; k <= r; ++k) {
      const uint32_t v = ld(max(yb + k, 0));
      cs01 += __byte_perm(v, 0, 0x4140);
      cs23 += __byte_perm(v, 0, 0x4342);
    }
  }
  for (int y = yb; y < ye; ++y) {
    const uint32_t pix = ld(y);
    uint32_t out = pix;
    if (y < filt_end) {                              // warp-uniform
      if (y > yb) {
        const uint32_t a = ld(y + r), s = ld(max(y - 1 - r, 0));
        cs01 += __byte_perm(a, 0, 0x4140) - __byte_perm(s, 0, 0x4140);
        cs23 += __byte_perm(a, 0, 0x4342) - __byte_perm(s, 0, 0x4342);
      }
      // column pairs at even offsets from x0 - 4 (P) and at odd offsets (Q)
      const uint32_t P0 = __shfl_up_sync(0xffffffffu, cs01, 1), P1 = __shfl_up_sync(0xffffffffu, cs23, 1);
      const uint32_t P4 = __shfl_down_sync(0xffffffffu, cs01, 1), P5 = __shfl_down_sync(0xffffffffu, cs23, 1);
      // pixels at the lowest or the highest level are never touched (:143): a lane whose four
      // pixels are all there (the flat opaque / transparent areas of a real alpha plane) is done
      bool touched = false;
#pragma unroll
      for (int q = 0; q < 4; ++q) touched |= (((pix >> (8 * q)) & 0xff) - keep_lo) < keep_span;
      if (writer && touched) {
        const uint32_t P2 = cs01, P3 = cs23;
        const uint32_t Q0 = __byte_perm(P0, P1, 0x5432), Q1 = __byte_perm(P1, P2, 0x5432), Q2 = __byte_perm(P2, P3, 0x5432),
                       Q3 = __byte_perm(P3, P4, 0x5432), Q4 = __byte_perm(P4, P5, 0x5432);
        // plain (2r+1)-wide windows on the mirrored extension: outputs 0,1 centred on offsets 4,5; 2,3 on 6,7
        uint32_t box01 = P2 + Q1 + Q2 + ((P1 + P3) & m2) + ((Q0 + Q3) & m3) + ((P0 + P4) & m4);
        uint32_t box23 = P3 + Q2 + Q3 + ((P2 + P4) & m2) + ((Q1 + Q4) & m3) + ((P1 + P5) & m4);
        // ... which is the reference's right mirror exactly; its left mirror (sum cs[0..x+r-1] +
        // sum cs[0..r-x] for x <= r) is the window centred on max(x-1, 0): only the lanes at
        // columns 0 and, for r = 4, 4 of the plane re-arrange their sums.
        if (x0 <= r) {
          if (x0 == 0) {
            const uint32_t n0 = box01 & 0xffff, n1 = box01 >> 16, n2 = box23 & 0xffff, n3 = box23 >> 16;
            box01 = n0 | (n0 << 16);
            box23 = (r >= 2 ? n1 : n2) | ((r >= 3 ? n2 : n3) << 16);
          } else {   // column 4 takes the window centred on 3: columns -1 (= 0) .. 7
            const uint32_t t = P0 + P1 + P2 + P3;
            box01 = (box01 & 0xffff0000u) | ((t & 0xffff) + (t >> 16) + (P0 & 0xffff));
          }
        }
        const uint32_t drow = DFIX ? (uint32_t)(0x5d7f91b36e4ca280ull >> (16 * ((y + r) & 3))) : 0u;   // kOrderedDither row (:39-44), a nibble per column
        uint32_t res[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t box = (q & 1) ? ((q < 2 ? box01 : box23) >> 16) : ((q < 2 ? box01 : box23) & 0xffff);
          const uint32_t pv = (pix >> (8 * q)) & 0xff;
          const uint32_t avg = __umulhi(box, scale16);
          int corr;
          asm volatile("ld.shared.s16 %0, [%1];" : "=r"(corr) : "r"(lut_mid + 2u * avg - 8u * pv));
          int c = (int)(pv << DFIX) + corr;
          if (DFIX) c += (int)((drow >> (4 * q)) & 15);
          c = min(max(c, 0), (256 << DFIX) - 1) >> DFIX;
          res[q] = (pv - keep_lo < keep_span) ? (uint32_t)c : pv;
        }
        out = res[0] | (res[1] << 8) | (res[2] << 16) | (res[3] << 24);
      }
    }
    if (writer) {
      uint8_t* d = j.dst + (size_t)y * j.dst_stride + x0;
      if (fast_dst) {
        *(uint32_t*)d = out;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (x0 + q < w) d[q] = (uint8_t)(out >> (8 * q));
      }
    }
  }
}

}  // namespace

extern "C" int b2_alpha_dequantize_levels(b2_ctx* ctx, int stream, const b2_alpha_dequant_job* jobs, int n_jobs,
                                          int ordered_dither) {
  B2_CHECK_CTX(ctx, stream);
  if (n_jobs < 0 || n_jobs > 65535) return B2_ERR_INVALID_CONFIG;
  if (n_jobs == 0) return B2_OK;
  if (jobs == NULL) return B2_ERR_NULL;
  const size_t jb = sizeof(DqJobDev) * (size_t)n_jobs;
  if (jb > B2_DESC_RING_BYTES) return B2_ERR_INVALID_CONFIG;
  DqJobDev* d = (DqJobDev*)calloc(1, jb);
  if (d == NULL) return B2_ERR_OOM;
  uint32_t max_units = 0, max_rows = 0;
  bool any = false;
  for (int i = 0; i < n_jobs; ++i) {
    const b2_alpha_dequant_job& j = jobs[i];
    int st = B2_OK;
    if (j.src == NULL || j.dst == NULL) st = B2_ERR_NULL;
    else if (j.width <= 0 || j.height <= 0 || j.width > B2_MAX_DIMENSION || j.height > B2_MAX_DIMENSION) st = B2_ERR_BAD_DIM;
    else if (j.src_stride < j.width || j.dst_stride < j.width) st = B2_ERR_BAD_DIM;
    else if (j.strength < 0 || j.strength > 100 || (const uint8_t*)j.dst == j.src) st = B2_ERR_INVALID_CONFIG;   // :263; out of place
    if (st != B2_OK) { free(d); return st; }
    int radius = 4 * j.strength / 100;                                  // :260-273
    if (2 * radius + 1 > j.width) radius = (j.width - 1) >> 1;
    if (2 * radius + 1 > j.height) radius = (j.height - 1) >> 1;
    d[i].src = j.src; d[i].dst = j.dst; d[i].src_stride = j.src_stride; d[i].dst_stride = j.dst_stride;
    d[i].width = j.width; d[i].height = j.height; d[i].radius = radius;
    d[i].tiles_x = (uint32_t)((j.width + kTileCols - 1) / kTileCols);
    d[i].units = d[i].tiles_x * (uint32_t)((j.height + kBandRows - 1) / kBandRows);
    if (d[i].units > max_units) max_units = d[i].units;
    if (radius > 0) { any = true; if ((uint32_t)j.height > max_rows) max_rows = (uint32_t)j.height; }
  }
  void* dev = NULL;
  int st = b2_upload_desc(ctx, stream, d, jb, &dev);
  free(d);
  if (st != B2_OK) return st;
  void* sp = NULL;
  st = b2_stream_scratch(ctx, stream, sizeof(DqStats) * (size_t)n_jobs, &sp);
  if (st != B2_OK) return st;
  cudaStream_t s = ctx->streams[stream];
  const DqJobDev* dj = (const DqJobDev*)dev;
  DqStats* ds = (DqStats*)sp;
  B2_CUDA(ctx, cudaMemsetAsync(ds, 0, sizeof(DqStats) * (size_t)n_jobs, s));
  if (any) {
    uint32_t gx = (max_rows + kWarps - 1) / kWarps;
    const uint32_t cap = (uint32_t)(ctx->sm_count * 8 + n_jobs - 1) / (uint32_t)n_jobs;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    dq_levels_kernel<<<dim3(gx, (unsigned)n_jobs), kThreads, 0, s>>>(dj, ds);
    ctx->launches += 1;
  }
  const dim3 grid((max_units + kWarps - 1) / kWarps, (unsigned)n_jobs);
  if (ordered_dither) dq_smooth_kernel<4><<<grid, kThreads, 0, s>>>(dj, ds);
  else dq_smooth_kernel<0><<<grid, kThreads, 0, s>>>(dj, ds);
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}
