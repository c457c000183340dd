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
//   * a band is cut into tiles of 128 skewed columns; tile (i, j) depends on tiles
//     (i, j-1), (i-1, j) and (i-1, j+1), so all tiles with 2i + j == d are independent:
//     one launch per diagonal d, no flags, no spinning, every dependency satisfied by
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
  uint32_t ktiles;         // skewed tiles per band: ceil((width + 62) / 128)
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
// ClampedAddSubtractFull (lossless.go:64-73): per channel clip255(a + b - c)
__device__ __forceinline__ uint32_t clamp_add_sub_full(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t out = 0;
#pragma unroll
  for (int sh = 0; sh < 32; sh += 8) {
    const int v = (int)((a >> sh) & 0xff) + (int)((b >> sh) & 0xff) - (int)((c >> sh) & 0xff);
    out |= (uint32_t)min(max(v, 0), 255) << sh;
  }
  return out;
}
// ClampedAddSubtractHalf (lossless.go:79-89): per channel clip255(ave + (ave - c) / 2), C division
__device__ __forceinline__ uint32_t clamp_add_sub_half(uint32_t a, uint32_t b, uint32_t c) {
  const uint32_t ave = avg2(a, b);
  uint32_t out = 0;
#pragma unroll
  for (int sh = 0; sh < 32; sh += 8) {
    const int av = (int)((ave >> sh) & 0xff), cv = (int)((c >> sh) & 0xff);
    const int v = av + (av - cv) / 2;
    out |= (uint32_t)min(max(v, 0), 255) << sh;
  }
  return out;
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
__global__ void __launch_bounds__(kFwdThreads, 8)
predictor_forward_kernel(const PredJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix,
                         uint32_t n_jobs, uint32_t uniform, uint32_t total_items) {
  // item = 256 consecutive pixels of the flat image, one per thread
  for (uint32_t item = blockIdx.x; item < total_items; item += gridDim.x) {
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, item, &local);
    const PredJobDev* __restrict__ j = jobs + img;
    const int w = j->width;
    const size_t n = (size_t)w * j->height;
    const size_t p = (size_t)local * kFwdThreads + threadIdx.x;
    if (p >= n) continue;
    const uint32_t* __restrict__ a = j->src;
    const int y = (int)(p / (uint32_t)w), x = (int)(p - (size_t)y * w);
    const uint32_t cur = __ldg(a + p);
    uint32_t pred;
    if (y == 0) {
      pred = (x == 0) ? 0xff000000u : __ldg(a + p - 1);          // lossless_enc.go: first row = L
    } else if (x == 0) {
      pred = __ldg(a + p - w);                                     // first column = T
    } else {
      const uint32_t mode = (__ldg(j->modes + (size_t)(y >> j->bits) * j->tiles_per_row + (x >> j->bits)) >> 8) & 0xf;
      // upper[x + 1] of the last column is the first pixel of the current row (contiguous image)
      pred = predict(mode, __ldg(a + p - 1), __ldg(a + p - w - 1), __ldg(a + p - w), __ldg(a + p - w + 1));
    }
    j->dst[p] = sub_pixels(cur, pred);
  }
}

// ---- inverse ----------------------------------------------------------------------
constexpr int kInvWarps = 4;
constexpr int kInvThreads = kInvWarps * 32;
constexpr int kTileK = 128;  // skewed columns per tile

// One launch = one diagonal d: tile (band i, skewed tile j) with 2 i + j == d of every
// image; grid.y = image, the warps of grid.x enumerate i.
__global__ void __launch_bounds__(kInvThreads, 8)
predictor_inverse_diag_kernel(const PredJobDev* __restrict__ jobs, uint32_t d) {
  __shared__ uint32_t upper_all[kInvWarps][kTileK + 2];
  const int lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5;
  const PredJobDev* __restrict__ jp = jobs + blockIdx.y;
  const uint32_t bands = jp->bands, ktiles = jp->ktiles;
  // bands on this diagonal: i in [i_lo, i_hi], j = d - 2 i in [0, ktiles)
  const uint32_t i_hi = min(bands - 1, d >> 1);
  const uint32_t i_lo = (d + 1 > ktiles) ? (d + 1 - ktiles + 1) >> 1 : 0;
  const uint32_t slot = blockIdx.x * kInvWarps + warp;
  if (i_lo > i_hi || slot > i_hi - i_lo) return;
  const uint32_t i = i_lo + slot, j = d - 2 * i;
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
  // lane 0's upper row comes from the band above: out[32 i - 1][k0 - 1 .. k0 + 128]
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
  uint32_t mode = 0;
  if (row_ok && c > 0 && c < w) mode = (__ldg(mode_row + (c >> bits)) >> 8) & 0xf;
  const int tile_mask = (1 << bits) - 1;
  uint32_t res = (row_ok && c >= 0 && c < w) ? __ldg(in_row + c) : 0u;

  for (int s = 0; s < kTileK; ++s, ++c) {
    // TR = (row-1, c+1): the lane above finished it on the previous step
    uint32_t tr = __shfl_up_sync(0xffffffffu, last_out, 1);
    if (lane == 0) tr = upper[s + 2];
    const bool live = row_ok && c >= 0 && c < w;
    // prefetch the next residual and, at a mode-tile boundary, the next mode
    uint32_t res_next = 0, mode_next = mode;
    const int cn = c + 1;
    if (row_ok && cn >= 0 && cn < w) {
      res_next = __ldg(in_row + cn);
      if ((cn & tile_mask) == 0 || cn == 1) mode_next = (__ldg(mode_row + (cn >> bits)) >> 8) & 0xf;
    }
    if (live) {
      if (c == w - 1) tr = (c == 0) ? 0u : first_px;
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
    res = res_next;
    mode = mode_next;
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
    const uint32_t items = (uint32_t)((n + kFwdThreads - 1) / kFwdThreads);
    if (i == 0) uni = items; else if (items != uni) uni = 0;
    pre[i] = (uint32_t)tot;
    tot += items;
    if (tot > 0x7fffffffull) { free(host); return B2_ERR_BAD_DIM; }
    if (d.bands > mb) mb = d.bands;
    const uint32_t diags = 2 * (d.bands - 1) + d.ktiles;
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
