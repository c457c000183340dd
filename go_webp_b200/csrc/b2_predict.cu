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
//   * a band is cut into tiles of 32 skewed columns; tile (i, j) depends on tiles
//     (i, j-1) and (i-1, j .. j+2), so all tiles with 3i + j == d are independent: one
//     launch per diagonal d, no flags, no spinning, every dependency satisfied by
//     stream order.  Tile boundaries read their neighbours back from the output image.
#include <stdlib.h>

#include "b2_common.cuh"

namespace {

struct PredJobDev {
  const uint32_t* src;
  uint32_t* dst;
  const uint32_t* modes;
  int32_t width, height, bits;
  uint32_t tiles_per_row;  // mode image row length
  uint32_t bands;          // ceil(height / 32)
  uint32_t ktiles;         // skewed tiles per band: ceil((width + 62) / kTileK)
};

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

// ---- forward ----------------------------------------------------------------------
constexpr int kFwdThreads = 256;
// item = 1024 consecutive pixels of the flat image, 4 per thread.  With width % 4 == 0
// and aligned images a thread's 4 pixels share a row and a mode tile (bits >= 2): two
// 128-bit loads (current and upper row) plus three scalar neighbours, one 128-bit store.
__global__ void __launch_bounds__(kFwdThreads, 8)
predictor_forward_kernel(const PredJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix,
                         uint32_t n_jobs, uint32_t uniform, uint32_t total_items) {
  for (uint32_t item = blockIdx.x; item < total_items; item += gridDim.x) {
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, item, &local);
    const PredJobDev* __restrict__ j = jobs + img;
    const int w = j->width;
    const size_t n = (size_t)w * j->height;
    const size_t p0 = ((size_t)local * kFwdThreads + threadIdx.x) * 4;
    if (p0 >= n) continue;
    const uint32_t* __restrict__ a = j->src;
    const bool vec = ((w & 3) == 0) && (((((uintptr_t)a) | ((uintptr_t)j->dst)) & 15) == 0);
    if (vec) {
      const int y = (int)(p0 / (uint32_t)w), x = (int)(p0 - (size_t)y * w);
      const uint4 cv = b2_ldg_stream_v4(a + p0);
      const uint32_t cur[4] = {cv.x, cv.y, cv.z, cv.w};
      uint32_t o[4];
      if (y == 0) {  // first row = L, first pixel = black (lossless_enc.go / predictor_enc.go:683-689)
        uint32_t l = (x == 0) ? 0xff000000u : __ldg(a + p0 - 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) { o[k] = sub_pixels(cur[k], l); l = cur[k]; }
      } else {
        const uint4 uv = __ldg(reinterpret_cast<const uint4*>(a + p0 - w));
        // upper[x + 1] of the last column is the first pixel of the current row (contiguous image)
        const uint32_t up[6] = {(x == 0) ? 0u : __ldg(a + p0 - w - 1), uv.x, uv.y, uv.z, uv.w, __ldg(a + p0 - w + 4)};
        const uint32_t mode = (__ldg(j->modes + (size_t)(y >> j->bits) * j->tiles_per_row + (x >> j->bits)) >> 8) & 0xf;
        uint32_t l = (x == 0) ? 0u : __ldg(a + p0 - 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t pred = (x + k == 0) ? up[1] : predict(mode, l, up[k], up[k + 1], up[k + 2]);  // first column = T
          o[k] = sub_pixels(cur[k], pred);
          l = cur[k];
        }
      }
      b2_stg_stream_v4(j->dst + p0, make_uint4(o[0], o[1], o[2], o[3]));
    } else {
      for (size_t p = p0; p < min(p0 + 4, n); ++p) {
        const int y = (int)(p / (uint32_t)w), x = (int)(p - (size_t)y * w);
        uint32_t pred;
        if (y == 0) pred = (x == 0) ? 0xff000000u : __ldg(a + p - 1);
        else if (x == 0) pred = __ldg(a + p - w);
        else {
          const uint32_t mode = (__ldg(j->modes + (size_t)(y >> j->bits) * j->tiles_per_row + (x >> j->bits)) >> 8) & 0xf;
          pred = predict(mode, __ldg(a + p - 1), __ldg(a + p - w - 1), __ldg(a + p - w), __ldg(a + p - w + 1));
        }
        j->dst[p] = sub_pixels(__ldg(a + p), pred);
      }
    }
  }
}

// ---- inverse ----------------------------------------------------------------------
constexpr int kInvWarps = 4;
constexpr int kInvThreads = kInvWarps * 32;
constexpr int kTileK = 32;   // skewed columns per tile: the wavefront needs W + 2H steps in all, and the
                             // diagonal schedule spends (2 H/32 + (W+62)/K) K of them, so small K is optimal

// Lane 31 of a band runs 62 columns behind lane 0, so lane 0 of tile (i, j) reads the
// row above up to skewed tile j + kDep of band i - 1: with d = (kDep + 1) i + j every
// dependency ((i, j-1), (i-1, j .. j+kDep)) lies on an earlier diagonal.
constexpr uint32_t kDep = (kTileK + 62) / kTileK;
constexpr uint32_t kDiag = kDep + 1;
// One launch = one diagonal d of every image; grid.y = image, the warps of grid.x
// enumerate the bands i that have a tile on it.
__global__ void __launch_bounds__(kInvThreads, 8)
predictor_inverse_diag_kernel(const PredJobDev* __restrict__ jobs, uint32_t d) {
  __shared__ uint32_t upper_all[kInvWarps][kTileK + 2];
  const int lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5;
  const PredJobDev* __restrict__ jp = jobs + blockIdx.y;
  const uint32_t bands = jp->bands, ktiles = jp->ktiles;
  // tiles on this diagonal: (kDiag i + j == d), j in [0, ktiles)
  const uint32_t i_hi = min(bands - 1, d / kDiag);
  const uint32_t i_lo = (d + 1 > ktiles) ? (d + 1 - ktiles + kDiag - 1) / kDiag : 0;
  const uint32_t slot = blockIdx.x * kInvWarps + warp;
  if (i_lo > i_hi || slot > i_hi - i_lo) return;
  const uint32_t i = i_lo + slot, j = d - kDiag * i;
  const int w = jp->width, h = jp->height, bits = jp->bits;
  const int row = (int)i * 32 + lane;
  const bool row_ok = row < h;
  const uint32_t* __restrict__ in_row = jp->src + (size_t)row * w;
  uint32_t* __restrict__ out = jp->dst;
  uint32_t* out_row = out + (size_t)row * w;
  const uint32_t* __restrict__ mode_row = jp->modes + (size_t)(row >> bits) * jp->tiles_per_row;
  uint32_t* upper = upper_all[warp];

  const int k0 = (int)j * kTileK;   // first skewed column of the tile
  int c = k0 - 2 * lane;            // this lane's column at the first step
  // lane 0's upper row comes from the band above: out[32 i - 1][k0 - 1 .. k0 + kTileK]
  if (i > 0) {
    const uint32_t* up = out + (size_t)(i * 32 - 1) * w;
    for (int t = lane; t < kTileK + 2; t += 32) {
      const int cc = k0 - 1 + t;
      upper[t] = (cc >= 0 && cc < w) ? up[cc] : 0u;
    }
  }
  __syncwarp();
  // state carried from the tile to the left (or the image edge): left = (row, c-1),
  // tl = (row-1, c-1), t = (row-1, c); and this lane's last output for the lane below
  uint32_t left = 0, tl = 0, t = 0, last_out = 0, first_px = 0;
  if (row_ok && row > 0) {
    const uint32_t* up = out_row - w;
    if (c - 1 >= 0 && c - 1 < w) tl = up[c - 1];
    if (c >= 0 && c < w) t = up[c];
  }
  if (row_ok && c - 1 >= 0 && c - 1 < w) left = out_row[c - 1];
  // what the lane below needs as its first TR: (row, c + 1 - 2) ... = this lane's pixel at c - 1
  last_out = left;
  if (row_ok && c > 0) first_px = out_row[0];  // TR of the last column (lossless.go: upper[x+1] wraps to the row start)
  // Predictor mode of the current 2^bits-wide mode tile and of the next one (fetched a
  // whole tile ahead so its latency never sits on the dependency chain).
  const int tile_mask = (1 << bits) - 1;
  const int mt_last = (int)jp->tiles_per_row - 1;
  auto load_mode = [&](int mt) -> uint32_t {
    return (row_ok && mt >= 0 && mt <= mt_last) ? (__ldg(mode_row + mt) >> 8) & 0xf : 0u;
  };
  // (initialised for column c - 1, so that a tile starting exactly on a mode-tile
  // boundary takes the regular switch on its first step)
  uint32_t mode = load_mode(max(c - 1, 0) >> bits);
  uint32_t mode_nxt = load_mode((max(c - 1, 0) >> bits) + 1);
  // Residuals: a 16-deep ring of scalar prefetches per lane (each lane walks its own
  // row, so a warp-wide load touches 32 lines; 16 steps of lead hide the DRAM latency).
  constexpr int kRing = 16;
  uint32_t ring[kRing];
#pragma unroll
  for (int q = 0; q < kRing; ++q) {
    const int cq = c + q;
    ring[q] = (row_ok && cq >= 0 && cq < w) ? __ldg(in_row + cq) : 0u;
  }

  for (int s0 = 0; s0 < kTileK; s0 += kRing) {
#pragma unroll
    for (int q = 0; q < kRing; ++q, ++c) {
      const int s = s0 + q;
      // TR = (row-1, c+1): the lane above finished it on the previous step
      uint32_t tr = __shfl_up_sync(0xffffffffu, last_out, 1);
      if (lane == 0) tr = upper[s + 2];
      const uint32_t res = ring[q];
      const int ca = c + kRing;
      ring[q] = (row_ok && ca >= 0 && ca < w) ? __ldg(in_row + ca) : 0u;
      if (c > 0 && (c & tile_mask) == 0) {  // entering the next mode tile
        mode = mode_nxt;
        mode_nxt = load_mode((c >> bits) + 1);
      }
      if (row_ok && c >= 0 && c < w) {
        if (c == w - 1) tr = first_px;
        uint32_t pred;
        if (row == 0) pred = (c == 0) ? 0xff000000u : left;   // lossless.go:292-295
        else if (c == 0) pred = t;                               // :312
        else pred = predict(mode, left, tl, t, tr);
        const uint32_t o = add_pixels(res, pred);
        out_row[c] = o;
        if (c == 0) first_px = o;
        left = o;
        last_out = o;
      }
      tl = t;
      t = tr;
    }
  }
}

int prepare_pred(b2_ctx* ctx, int stream, const b2_argb_job* jobs, int n_jobs, const PredJobDev** djobs,
                 const uint32_t** dprefix, uint32_t* uniform, uint32_t* total, uint32_t* max_bands, uint32_t* n_diag) {
  const size_t jb = (sizeof(PredJobDev) * (size_t)n_jobs + 15) & ~(size_t)15;
  const size_t pb = sizeof(uint32_t) * ((size_t)n_jobs + 1);
  if (jb + pb > B2_DESC_RING_BYTES) return B2_ERR_INVALID_CONFIG;
  uint8_t* host = (uint8_t*)malloc(jb + pb);
  if (host == NULL) return B2_ERR_OOM;
  PredJobDev* dj = (PredJobDev*)host;
  uint32_t* pre = (uint32_t*)(host + jb);
  uint64_t tot = 0;
  uint32_t uni = 0, mb = 0, nd = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const b2_argb_job& j = jobs[i];
    int st = B2_OK;
    if (j.src == NULL || j.dst == NULL || j.data == NULL) st = B2_ERR_NULL;
    else if (j.width <= 0 || j.height <= 0 || j.width > B2_MAX_DIMENSION || j.height > B2_MAX_DIMENSION) st = B2_ERR_BAD_DIM;
    else if (j.bits < 2 || j.bits > 9 || j.src == j.dst) st = B2_ERR_INVALID_CONFIG;  // vp8l_dec.c.go:1243
    if (st != B2_OK) { free(host); return st; }
    PredJobDev& d = dj[i];
    d.src = j.src; d.dst = j.dst; d.modes = j.data;
    d.width = j.width; d.height = j.height; d.bits = j.bits;
    d.tiles_per_row = (uint32_t)((j.width + (1 << j.bits) - 1) >> j.bits);
    d.bands = (uint32_t)((j.height + 31) / 32);
    d.ktiles = (uint32_t)((j.width + 62 + kTileK - 1) / kTileK);
    const uint64_t n = (uint64_t)j.width * (uint64_t)j.height;
    const uint32_t items = (uint32_t)((n + 4 * kFwdThreads - 1) / (4 * kFwdThreads));
    if (i == 0) uni = items; else if (items != uni) uni = 0;
    pre[i] = (uint32_t)tot;
    tot += items;
    if (tot > 0x7fffffffull) { free(host); return B2_ERR_BAD_DIM; }
    if (d.bands > mb) mb = d.bands;
    const uint32_t diags = kDiag * (d.bands - 1) + d.ktiles;
    if (diags > nd) nd = diags;
  }
  pre[n_jobs] = (uint32_t)tot;
  void* dev = NULL;
  const int st = b2_upload_desc(ctx, stream, host, jb + pb, &dev);
  free(host);
  if (st != B2_OK) return st;
  *djobs = (const PredJobDev*)dev;
  *dprefix = (const uint32_t*)((uint8_t*)dev + jb);
  *uniform = uni; *total = (uint32_t)tot; *max_bands = mb; *n_diag = nd;
  return B2_OK;
}

}  // namespace

extern "C" int b2_vp8l_predictor_forward(b2_ctx* ctx, int stream, const b2_argb_job* jobs, int n_jobs) {
  B2_CHECK_CTX(ctx, stream);
  if (n_jobs < 0) return B2_ERR_INVALID_CONFIG;
  if (n_jobs == 0) return B2_OK;
  if (jobs == NULL) return B2_ERR_NULL;
  const PredJobDev* dj; const uint32_t* dp; uint32_t uni, tot, mb, nd;
  const int st = prepare_pred(ctx, stream, jobs, n_jobs, &dj, &dp, &uni, &tot, &mb, &nd);
  if (st != B2_OK) return st;
  const int grid = b2_grid_for(ctx, 8, tot);
  predictor_forward_kernel<<<grid, kFwdThreads, 0, ctx->streams[stream]>>>(dj, dp, (uint32_t)n_jobs, uni, tot);
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}

extern "C" int b2_vp8l_predictor_inverse(b2_ctx* ctx, int stream, const b2_argb_job* jobs, int n_jobs) {
  B2_CHECK_CTX(ctx, stream);
  if (n_jobs < 0) return B2_ERR_INVALID_CONFIG;
  if (n_jobs == 0) return B2_OK;
  if (jobs == NULL) return B2_ERR_NULL;
  if (n_jobs > 65535) return B2_ERR_INVALID_CONFIG;  // grid.y = image
  const PredJobDev* dj; const uint32_t* dp; uint32_t uni, tot, mb, nd;
  const int st = prepare_pred(ctx, stream, jobs, n_jobs, &dj, &dp, &uni, &tot, &mb, &nd);
  if (st != B2_OK) return st;
  // at most min(bands, ceil(ktiles / 2) + 1) tiles of one image share a diagonal
  const dim3 grid((mb + kInvWarps - 1) / kInvWarps, (unsigned)n_jobs);
  for (uint32_t d = 0; d < nd; ++d) {
    predictor_inverse_diag_kernel<<<grid, kInvThreads, 0, ctx->streams[stream]>>>(dj, d);
    ctx->launches++;
  }
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}
