// b2_extra.cu -- the remaining rows of the hot-path table:
//   a14  non-gamma ARGB -> Y / U,V (+A) of the VP8L -> YUVA output path
//        WebPConvertARGBToY / WebPConvertARGBToUV, pkg/libwebp/dsp/yuv.go:74-120,
//        driven per row by ConvertToYUVA, pkg/vp8/vp8l_dec.c.go:543-562
//   a28  colour-index inverse on an 8-bit (alpha) plane
//        MapAlpha_C / VP8LColorIndexInverseTransformAlpha, pkg/libwebp/dsp/lossless.go:461-505
//   a29  ExpandColorMap, pkg/vp8/vp8l_dec.c.go:1196-1219 (byte-wise prefix sum of a
//        delta-coded palette, zero padded to 1 << (8 >> bits) entries)
#include <stdlib.h>

#include "b2_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// ---- a14 -------------------------------------------------------------------------
struct PlainJobDev {
  const uint32_t* argb;
  uint8_t* y;
  uint8_t* u;
  uint8_t* v;
  uint8_t* a;
  int32_t argb_stride;  // pixels
  int32_t y_stride, uv_stride, a_stride;
  int32_t width, height;
  int32_t y_odd;  // 1: the first row sits at an odd y position and averages into the stored U/V
  uint32_t tiles_x;
};

// RGBToY(r, g, b, YUV_HALF), pkg/color/yuv/conversion.go:56-59
__device__ __forceinline__ uint32_t luma(uint32_t p) {
  const uint32_t r = (p >> 16) & 0xff, g = (p >> 8) & 0xff, b = p & 0xff;
  return (16839u * r + 33059u * g + 6420u * b + (1u << 15) + (16u << 16)) >> 16;
}
// RGBToU / RGBToV on "sum of four" channel values with rounding YUV_HALF << 2 and
// ClipUV (conversion.go:51-54,61-69)
__device__ __forceinline__ int clip_uv(int uv) {
  uv = (uv + (1 << 17) + (128 << 18)) >> 18;
  return min(max(uv, 0), 255);
}
__device__ __forceinline__ void chroma_of(uint32_t v0, uint32_t v1, bool pair, int* u, int* v) {
  int r, g, b;
  if (pair) {  // yuv.go:88-93: two pixels, each channel scaled by 2
    r = (int)(((v0 >> 15) & 0x1fe) + ((v1 >> 15) & 0x1fe));
    g = (int)(((v0 >> 7) & 0x1fe) + ((v1 >> 7) & 0x1fe));
    b = (int)(((v0 << 1) & 0x1fe) + ((v1 << 1) & 0x1fe));
  } else {     // yuv.go:106-109: last pixel of an odd row, scaled by 4
    r = (int)((v0 >> 14) & 0x3fc);
    g = (int)((v0 >> 6) & 0x3fc);
    b = (int)((v0 << 2) & 0x3fc);
  }
  *u = clip_uv(-9719 * r - 19081 * g + 28800 * b);
  *v = clip_uv(28800 * r - 24116 * g - 4684 * b);
}

// Work item = 64 chroma samples (128 pixels) x one row pair; a thread owns 2 chroma
// samples = 4 pixels x 2 rows: 128-bit loads when the rows allow it.
__global__ void __launch_bounds__(kThreads, 8)
argb_to_yuva_plain_kernel(const PlainJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix,
                          uint32_t n_jobs, uint32_t uniform, uint32_t total_items) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5;
  for (uint32_t item = blockIdx.x * kWarps + warp; item < total_items; item += gridDim.x * kWarps) {
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, item, &local);
    const PlainJobDev j = jobs[img];
    const uint32_t rp = local / j.tiles_x;
    const int x0 = (int)(local - rp * j.tiles_x) * 128 + lane * 4;
    if (x0 >= j.width) continue;
    // rows r0 (even position) and r0 + 1 (odd position) of this chunk; r0 == -1 when the
    // chunk starts on an odd row, whose even partner was converted by an earlier call
    const int r0 = 2 * (int)rp - j.y_odd;
    const bool row0 = (r0 >= 0), row1 = (r0 + 1 < j.height);
    const uint32_t* s0 = j.argb + (ptrdiff_t)(row0 ? r0 : r0 + 1) * j.argb_stride + x0;
    const uint32_t* s1 = j.argb + (ptrdiff_t)(row1 ? r0 + 1 : r0) * j.argb_stride + x0;
    const int nv = min(4, j.width - x0);
    uint32_t p0[4], p1[4];
    if (nv == 4 && ((((uintptr_t)s0) | ((uintptr_t)s1)) & 15) == 0) {
      const uint4 a = b2_ldg_stream_v4(s0), b = b2_ldg_stream_v4(s1);
      p0[0] = a.x; p0[1] = a.y; p0[2] = a.z; p0[3] = a.w;
      p1[0] = b.x; p1[1] = b.y; p1[2] = b.z; p1[3] = b.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) { p0[k] = (k < nv) ? __ldg(s0 + k) : 0u; p1[k] = (k < nv) ? __ldg(s1 + k) : 0u; }
    }
    uint8_t* yr0 = j.y + (ptrdiff_t)r0 * j.y_stride + x0;
    uint8_t* yr1 = yr0 + j.y_stride;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < nv) {
        if (row0) yr0[k] = (uint8_t)luma(p0[k]);
        if (row1) yr1[k] = (uint8_t)luma(p1[k]);
      }
    }
    if (j.a != nullptr) {  // WebPExtractAlpha on byte 3 of every pixel (vp8l_dec.c.go:557-560)
      uint8_t* ar0 = j.a + (ptrdiff_t)r0 * j.a_stride + x0;
      uint8_t* ar1 = ar0 + j.a_stride;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < nv) {
          if (row0) ar0[k] = (uint8_t)(p0[k] >> 24);
          if (row1) ar1[k] = (uint8_t)(p1[k] >> 24);
        }
      }
    }
    uint8_t* ur = j.u + (ptrdiff_t)rp * j.uv_stride + (x0 >> 1);
    uint8_t* vr = j.v + (ptrdiff_t)rp * j.uv_stride + (x0 >> 1);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int xs = 2 * c;
      if (xs >= nv) break;
      const bool pair = (xs + 1 < nv);
      int u0, v0, u1, v1;
      if (row0) chroma_of(p0[xs], p0[xs + 1 < 4 ? xs + 1 : xs], pair, &u0, &v0);  // even row: store (yuv.go:96-98)
      else { u0 = ur[c]; v0 = vr[c]; }                                              // stored by the earlier call
      if (row1) {  // odd row: average with the stored value (yuv.go:100-102)
        chroma_of(p1[xs], p1[xs + 1 < 4 ? xs + 1 : xs], pair, &u1, &v1);
        u0 = (u0 + u1 + 1) >> 1;
        v0 = (v0 + v1 + 1) >> 1;
      }
      ur[c] = (uint8_t)u0;
      vr[c] = (uint8_t)v0;
    }
  }
}

// ---- a28 -------------------------------------------------------------------------
struct AlphaIndexJobDev {
  const uint8_t* src;
  uint8_t* dst;
  const uint32_t* palette;
  int32_t width, height, bits;
  uint32_t src_row;   // bytes per source row = ceil(width / 2^bits)
  uint32_t segs;      // 4096-pixel segments per row
};
constexpr int kAlphaSeg = 4096;

// A warp owns a 4096-pixel segment of one row; a lane produces 4 consecutive output
// bytes per step (one 32-bit store) from the source byte(s) that hold their indices.
// The palette's green bytes (VP8GetAlphaValue, lossless_common.go:31) sit in shared
// memory, one copy per warp.
template <int BITS>
__global__ void __launch_bounds__(kThreads, 8)
alpha_index_kernel(const AlphaIndexJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix,
                   uint32_t n_jobs, uint32_t uniform, uint32_t total_items) {
  constexpr int kPal = 1 << (8 >> BITS);
  constexpr int kBpp = 8 >> BITS;
  constexpr uint32_t kMask = (1u << kBpp) - 1;
  __shared__ uint8_t green_all[kWarps * kPal];
  const int lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5;
  uint8_t* green = green_all + warp * kPal;
  uint32_t cur_img = 0xffffffffu;
  for (uint32_t item = blockIdx.x * kWarps + warp; item < total_items; item += gridDim.x * kWarps) {
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, item, &local);
    const AlphaIndexJobDev j = jobs[img];
    if (img != cur_img) {
      __syncwarp();
      for (int i = lane; i < kPal; i += 32) green[i] = (uint8_t)(__ldg(j.palette + i) >> 8);
      __syncwarp();
      cur_img = img;
    }
    const uint32_t y = local / j.segs;
    const int xseg = (int)(local - y * j.segs) * kAlphaSeg;
    const int xend = min(j.width, xseg + kAlphaSeg);
    const uint8_t* srow = j.src + (size_t)y * j.src_row;
    uint8_t* drow = j.dst + (size_t)y * (uint32_t)j.width;
    const bool vec = ((((uintptr_t)drow) & 3) == 0);
    for (int x4 = xseg + lane * 4; x4 < xend; x4 += 128) {
      const int nv = min(4, xend - x4);
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int x = x4 + k;
        uint32_t idx = 0;
        if (k < nv) {
          const uint32_t packed = __ldg(srow + (x >> BITS));  // lossless.go:476-481
          idx = (packed >> (kBpp * (x & ((1 << BITS) - 1)))) & kMask;
        }
        o[k] = green[idx];
      }
      if (nv == 4 && vec) {
        *reinterpret_cast<uint32_t*>(drow + x4) = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) if (k < nv) drow[x4 + k] = (uint8_t)o[k];
      }
    }
  }
}

// ---- a29 -------------------------------------------------------------------------
struct ExpandJobDev {
  const uint32_t* src;
  uint32_t* dst;
  int32_t num_colors, bits;
};
// byte-wise a + b (mod 256 per byte): VP8LAddPixels, lossless_common.go:110-114
__device__ __forceinline__ uint32_t add_pixels(uint32_t a, uint32_t b) {
  const uint32_t ag = (a & 0xff00ff00u) + (b & 0xff00ff00u);
  const uint32_t rb = (a & 0x00ff00ffu) + (b & 0x00ff00ffu);
  return (ag & 0xff00ff00u) | (rb & 0x00ff00ffu);
}
// One warp per palette: lane l scans entries 8l..8l+7, a warp scan joins the lanes.
__global__ void __launch_bounds__(kThreads, 8)
expand_color_map_kernel(const ExpandJobDev* __restrict__ jobs, uint32_t n_jobs) {
  const int lane = threadIdx.x & 31;
  const uint32_t job = blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (job >= n_jobs) return;
  const ExpandJobDev j = jobs[job];
  const int final_colors = 1 << (8 >> j.bits);
  uint32_t e[8];
  uint32_t run = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = lane * 8 + k;
    const uint32_t d = (i < j.num_colors) ? __ldg(j.src + i) : 0u;
    run = add_pixels(run, d);
    e[k] = run;
  }
  uint32_t incl = run;  // inclusive scan of the lane totals
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl = add_pixels(incl, t);
  }
  uint32_t excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = lane * 8 + k;
    if (i < final_colors) j.dst[i] = (i < j.num_colors) ? add_pixels(e[k], excl) : 0u;  // zero padding, :1213-1215
  }
}

}  // namespace

extern "C" int b2_argb_to_yuva_plain(b2_ctx* ctx, int stream, const b2_plain_job* jobs, int n_jobs) {
  B2_CHECK_CTX(ctx, stream);
  if (n_jobs < 0) return B2_ERR_INVALID_CONFIG;
  if (n_jobs == 0) return B2_OK;
  if (jobs == NULL) return B2_ERR_NULL;
  const size_t jb = (sizeof(PlainJobDev) * (size_t)n_jobs + 15) & ~(size_t)15;
  const size_t pb = sizeof(uint32_t) * ((size_t)n_jobs + 1);
  if (jb + pb > B2_DESC_RING_BYTES) return B2_ERR_INVALID_CONFIG;
  uint8_t* host = (uint8_t*)malloc(jb + pb);
  if (host == NULL) return B2_ERR_OOM;
  PlainJobDev* dj = (PlainJobDev*)host;
  uint32_t* pre = (uint32_t*)(host + jb);
  uint64_t total = 0;
  uint32_t uniform = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const b2_plain_job& j = jobs[i];
    if (j.argb == NULL || j.y == NULL || j.u == NULL || j.v == NULL) { free(host); return B2_ERR_NULL; }
    if (j.width <= 0 || j.height <= 0 || j.width > B2_MAX_DIMENSION || j.height > B2_MAX_DIMENSION || j.y_pos < 0 ||
        (j.argb_stride < 0 ? -(int64_t)j.argb_stride : j.argb_stride) < j.width) { free(host); return B2_ERR_BAD_DIM; }
    PlainJobDev& d = dj[i];
    d.argb = j.argb; d.y = j.y; d.u = j.u; d.v = j.v; d.a = j.a;
    d.argb_stride = j.argb_stride; d.y_stride = j.y_stride; d.uv_stride = j.uv_stride; d.a_stride = j.a_stride;
    d.width = j.width; d.height = j.height;
    d.y_odd = j.y_pos & 1;
    d.tiles_x = (uint32_t)((j.width + 127) / 128);
    const uint32_t items = d.tiles_x * (uint32_t)((j.height + d.y_odd + 1) / 2);
    if (i == 0) uniform = items; else if (items != uniform) uniform = 0;
    pre[i] = (uint32_t)total;
    total += items;
    if (total > 0x7fffffffull) { free(host); return B2_ERR_BAD_DIM; }
  }
  pre[n_jobs] = (uint32_t)total;
  void* dev = NULL;
  const int st = b2_upload_desc(ctx, stream, host, jb + pb, &dev);
  free(host);
  if (st != B2_OK) return st;
  const int grid = b2_grid_for(ctx, 8, (total + kWarps - 1) / kWarps);
  argb_to_yuva_plain_kernel<<<grid, kThreads, 0, ctx->streams[stream]>>>(
      (const PlainJobDev*)dev, (const uint32_t*)((uint8_t*)dev + jb), (uint32_t)n_jobs, uniform, (uint32_t)total);
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}

extern "C" int b2_vp8l_color_index_inverse_alpha(b2_ctx* ctx, int stream, const b2_alpha_index_job* jobs, int n_jobs) {
  B2_CHECK_CTX(ctx, stream);
  if (n_jobs < 0) return B2_ERR_INVALID_CONFIG;
  if (n_jobs == 0) return B2_OK;
  if (jobs == NULL) return B2_ERR_NULL;
  for (int i = 0; i < n_jobs; ++i) {
    const b2_alpha_index_job& j = jobs[i];
    if (j.src == NULL || j.dst == NULL || j.palette == NULL) return B2_ERR_NULL;
    if (j.width <= 0 || j.height <= 0 || j.width > B2_MAX_DIMENSION || j.height > B2_MAX_DIMENSION) return B2_ERR_BAD_DIM;
    if (j.bits < 0 || j.bits > 3 || j.src == j.dst) return B2_ERR_INVALID_CONFIG;
  }
  // one launch per 'bits' value present (the kernel is specialised on it)
  for (int bits = 0; bits <= 3; ++bits) {
    int cnt = 0;
    for (int i = 0; i < n_jobs; ++i) cnt += (jobs[i].bits == bits);
    if (cnt == 0) continue;
    const size_t jb = (sizeof(AlphaIndexJobDev) * (size_t)cnt + 15) & ~(size_t)15;
    const size_t pb = sizeof(uint32_t) * ((size_t)cnt + 1);
    if (jb + pb > B2_DESC_RING_BYTES) return B2_ERR_INVALID_CONFIG;
    uint8_t* host = (uint8_t*)malloc(jb + pb);
    if (host == NULL) return B2_ERR_OOM;
    AlphaIndexJobDev* dj = (AlphaIndexJobDev*)host;
    uint32_t* pre = (uint32_t*)(host + jb);
    uint64_t total = 0;
    uint32_t uniform = 0;
    int k = 0;
    for (int i = 0; i < n_jobs; ++i) {
      const b2_alpha_index_job& j = jobs[i];
      if (j.bits != bits) continue;
      AlphaIndexJobDev& d = dj[k];
      d.src = j.src; d.dst = j.dst; d.palette = j.palette;
      d.width = j.width; d.height = j.height; d.bits = j.bits;
      d.src_row = (uint32_t)((j.width + (1 << bits) - 1) >> bits);  // VP8LSubSampleSize, lossless_common.go:34-36
      d.segs = (uint32_t)((j.width + kAlphaSeg - 1) / kAlphaSeg);
      const uint32_t items = d.segs * (uint32_t)j.height;
      if (k == 0) uniform = items; else if (items != uniform) uniform = 0;
      pre[k++] = (uint32_t)total;
      total += items;
    }
    pre[cnt] = (uint32_t)total;
    void* dev = NULL;
    const int st = b2_upload_desc(ctx, stream, host, jb + pb, &dev);
    free(host);
    if (st != B2_OK) return st;
    const AlphaIndexJobDev* djobs = (const AlphaIndexJobDev*)dev;
    const uint32_t* dpre = (const uint32_t*)((uint8_t*)dev + jb);
    const int grid = b2_grid_for(ctx, 8, (total + kWarps - 1) / kWarps);
    cudaStream_t s = ctx->streams[stream];
    switch (bits) {
      case 0: alpha_index_kernel<0><<<grid, kThreads, 0, s>>>(djobs, dpre, (uint32_t)cnt, uniform, (uint32_t)total); break;
      case 1: alpha_index_kernel<1><<<grid, kThreads, 0, s>>>(djobs, dpre, (uint32_t)cnt, uniform, (uint32_t)total); break;
      case 2: alpha_index_kernel<2><<<grid, kThreads, 0, s>>>(djobs, dpre, (uint32_t)cnt, uniform, (uint32_t)total); break;
      default: alpha_index_kernel<3><<<grid, kThreads, 0, s>>>(djobs, dpre, (uint32_t)cnt, uniform, (uint32_t)total); break;
    }
    ctx->launches++;
    B2_CUDA(ctx, cudaGetLastError());
  }
  return B2_OK;
}

extern "C" int b2_vp8l_expand_color_map(b2_ctx* ctx, int stream, const b2_expand_job* jobs, int n_jobs) {
  B2_CHECK_CTX(ctx, stream);
  if (n_jobs < 0) return B2_ERR_INVALID_CONFIG;
  if (n_jobs == 0) return B2_OK;
  if (jobs == NULL) return B2_ERR_NULL;
  const size_t jb = sizeof(ExpandJobDev) * (size_t)n_jobs;
  if (jb > B2_DESC_RING_BYTES) return B2_ERR_INVALID_CONFIG;
  ExpandJobDev* host = (ExpandJobDev*)malloc(jb);
  if (host == NULL) return B2_ERR_OOM;
  for (int i = 0; i < n_jobs; ++i) {
    const b2_expand_job& j = jobs[i];
    if (j.src == NULL || j.dst == NULL) { free(host); return B2_ERR_NULL; }
    // bits follows the colour count (vp8l_dec.c.go:1248-1254); the map holds 1 << (8 >> bits) entries
    if (j.bits < 0 || j.bits > 3 || j.num_colors < 1 || j.num_colors > (1 << (8 >> j.bits))) {
      free(host); return B2_ERR_INVALID_CONFIG;
    }
    host[i].src = j.src; host[i].dst = j.dst; host[i].num_colors = j.num_colors; host[i].bits = j.bits;
  }
  void* dev = NULL;
  const int st = b2_upload_desc(ctx, stream, host, jb, &dev);
  free(host);
  if (st != B2_OK) return st;
  expand_color_map_kernel<<<(n_jobs + kWarps - 1) / kWarps, kThreads, 0, ctx->streams[stream]>>>(
      (const ExpandJobDev*)dev, (uint32_t)n_jobs);
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}
