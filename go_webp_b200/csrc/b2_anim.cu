// b2_anim.cu -- animation pixel work:
//   canvas blending   BlendPixelRowNonPremult / BlendPixelRowPremult,
//                     pkg/libwebp/demux/anim_decode.go:195-270
//   frame comparison  ComparePixelsLossless / ComparePixelsLossy (PixelsAreSimilar),
//                     pkg/libwebp/mux/anim_encode.go:318-367
// Pixels are uint32 with alpha in bits 24-31 (CHANNEL_SHIFT(3) on little endian; the three
// colour channels are treated alike, so RGBA and BGRA canvases use the same code).
#include <stdlib.h>

#include "b2_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct BlendJobDev {
  uint32_t* src;
  const uint32_t* dst;
  int32_t src_stride, dst_stride, width, height;
  uint32_t units_x, pad;
};

// unit = (row, 1024-pixel segment), 4 pixels per lane and pass
template <bool PREMULT>
__global__ void __launch_bounds__(kThreads, 8)
anim_blend_kernel(const BlendJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix, uint32_t n_jobs,
                  uint32_t uniform, uint32_t total) {
  __shared__ uint32_t scale_tab[256];   // (1 << 24) / blend_a, anim_decode.go:216
  if (!PREMULT) {
    scale_tab[threadIdx.x] = threadIdx.x ? (1u << 24) / threadIdx.x : 0u;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  auto px = [&](uint32_t s, uint32_t d) -> uint32_t {
    const uint32_t src_a = s >> 24;
    if (src_a == 0xffu) return s;                       // the row functions skip opaque pixels (:238, :264)
    if (PREMULT) {                                      // BlendPixelPremult :253-257, ChannelwiseMultiply :245-251
      const uint32_t scale = 256u - src_a, mask = 0x00ff00ffu;
      const uint32_t rb = ((d & mask) * scale) >> 8;
      const uint32_t ag = ((d >> 8) & mask) * scale;
      return s + ((rb & mask) | (ag & ~mask));
    }
    if (src_a == 0u) return d;                          // BlendPixelNonPremult :205-232
    const uint32_t dst_a = d >> 24;
    const uint32_t dst_factor_a = (dst_a * (256u - src_a)) >> 8;
    const uint32_t blend_a = src_a + dst_factor_a;
    const uint32_t scale = scale_tab[blend_a & 0xffu];
    uint32_t out = blend_a << 24;
#pragma unroll
    for (int sh = 0; sh < 24; sh += 8) {
      const uint32_t unscaled = ((s >> sh) & 0xff) * src_a + ((d >> sh) & 0xff) * dst_factor_a;
      out |= (((unscaled * scale) >> 24) & 0xffu) << sh;
    }
    return out;
  };
  for (uint32_t unit = blockIdx.x * kWarps + (threadIdx.x >> 5); unit < total; unit += gridDim.x * kWarps) {
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, unit, &local);
    const BlendJobDev* __restrict__ j = jobs + img;
    const int width = j->width;
    const int row = (int)(local / j->units_x), seg = (int)(local % j->units_x);
    uint32_t* srow = j->src + (ptrdiff_t)row * j->src_stride;
    const uint32_t* drow = j->dst + (ptrdiff_t)row * j->dst_stride;
    const bool al16 = ((((uintptr_t)j->src | (uintptr_t)j->dst) & 15) == 0) && (((j->src_stride | j->dst_stride) & 3) == 0);
#pragma unroll
    for (int pass = 0; pass < 8; ++pass) {
      const int x = seg * 1024 + pass * 128 + lane * 4;
      if (x >= width) break;
      if (al16 && x + 4 <= width) {
        uint4 s = *reinterpret_cast<const uint4*>(srow + x);
        const uint4 d = b2_ldg_stream_v4(drow + x);
        s.x = px(s.x, d.x); s.y = px(s.y, d.y); s.z = px(s.z, d.z); s.w = px(s.w, d.w);
        *reinterpret_cast<uint4*>(srow + x) = s;
      } else {
        for (int k = 0; k < 4 && x + k < width; ++k) srow[x + k] = px(srow[x + k], drow[x + k]);
      }
    }
  }
}

struct CompareJobDev {
  const uint32_t* src;
  const uint32_t* dst;
  int32_t src_step, dst_step, length, pad;
};

// PixelsAreSimilar, anim_encode.go:336-351
__device__ __forceinline__ bool similar(uint32_t s, uint32_t d, int max_diff) {
  const int da = (int)(d >> 24);
  if ((s >> 24) != (d >> 24)) return false;
  const int lim = max_diff * 255;
#pragma unroll
  for (int sh = 0; sh < 24; sh += 8) {
    const int a = (int)((s >> sh) & 0xff), b = (int)((d >> sh) & 0xff);
    if (abs(a - b) * da > lim) return false;
  }
  return true;
}

// one warp per comparison: result = 1 iff every pixel pair is equal / similar
__global__ void __launch_bounds__(kThreads, 8)
anim_compare_kernel(const CompareJobDev* __restrict__ jobs, uint32_t n_jobs, int lossless, int max_diff,
                    int32_t* __restrict__ result) {
  const int lane = threadIdx.x & 31;
  for (uint32_t job = blockIdx.x * kWarps + (threadIdx.x >> 5); job < n_jobs; job += gridDim.x * kWarps) {
    const CompareJobDev j = jobs[job];
    bool ok = true;
    for (int i = lane; i < j.length; i += 32) {
      const uint32_t s = __ldg(j.src + (ptrdiff_t)i * j.src_step), d = __ldg(j.dst + (ptrdiff_t)i * j.dst_step);
      ok = ok && (lossless ? (s == d) : similar(s, d, max_diff));
    }
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) result[job] = ok ? 1 : 0;
  }
}

}  // namespace

extern "C" int b2_anim_blend_rows(b2_ctx* ctx, int stream, const b2_anim_blend_job* jobs, int n_jobs, int premultiplied) {
  B2_CHECK_CTX(ctx, stream);
  if (n_jobs < 0) return B2_ERR_INVALID_CONFIG;
  if (n_jobs == 0) return B2_OK;
  if (jobs == NULL) return B2_ERR_NULL;
  const size_t jb = (sizeof(BlendJobDev) * (size_t)n_jobs + 15) & ~(size_t)15;
  const size_t pb = sizeof(uint32_t) * ((size_t)n_jobs + 1);
  if (jb + pb > B2_DESC_RING_BYTES) return B2_ERR_INVALID_CONFIG;
  uint8_t* host = (uint8_t*)calloc(1, jb + pb);
  if (host == NULL) return B2_ERR_OOM;
  BlendJobDev* d = (BlendJobDev*)host;
  uint32_t* pre = (uint32_t*)(host + jb);
  uint64_t tot = 0;
  uint32_t uni = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const b2_anim_blend_job& j = jobs[i];
    int st = B2_OK;
    if (j.src == NULL || j.dst == NULL) st = B2_ERR_NULL;
    else if (j.width <= 0 || j.height <= 0 || j.width > B2_MAX_DIMENSION || j.height > B2_MAX_DIMENSION) st = B2_ERR_BAD_DIM;
    else if (j.src_stride < j.width || j.dst_stride < j.width) st = B2_ERR_BAD_DIM;
    if (st != B2_OK) { free(host); return st; }
    d[i].src = j.src; d[i].dst = j.dst; d[i].src_stride = j.src_stride; d[i].dst_stride = j.dst_stride;
    d[i].width = j.width; d[i].height = j.height;
    d[i].units_x = (uint32_t)((j.width + 1023) / 1024);
    const uint32_t items = d[i].units_x * (uint32_t)j.height;
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
  const BlendJobDev* dj = (const BlendJobDev*)dev;
  const uint32_t* dp = (const uint32_t*)((uint8_t*)dev + jb);
  const int grid = b2_grid_for(ctx, 8, (tot + kWarps - 1) / kWarps);
  if (premultiplied) anim_blend_kernel<true><<<grid, kThreads, 0, ctx->streams[stream]>>>(dj, dp, (uint32_t)n_jobs, uni, (uint32_t)tot);
  else anim_blend_kernel<false><<<grid, kThreads, 0, ctx->streams[stream]>>>(dj, dp, (uint32_t)n_jobs, uni, (uint32_t)tot);
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}

extern "C" int b2_anim_compare_pixels(b2_ctx* ctx, int stream, const b2_anim_compare_job* jobs, int n_jobs, int lossless,
                                      int max_allowed_diff, int32_t* result_dev) {
  B2_CHECK_CTX(ctx, stream);
  if (n_jobs < 0) return B2_ERR_INVALID_CONFIG;
  if (n_jobs == 0) return B2_OK;
  if (jobs == NULL || result_dev == NULL) return B2_ERR_NULL;
  const size_t jb = sizeof(CompareJobDev) * (size_t)n_jobs;
  if (jb > B2_DESC_RING_BYTES) return B2_ERR_INVALID_CONFIG;
  CompareJobDev* host = (CompareJobDev*)calloc(1, jb);
  if (host == NULL) return B2_ERR_OOM;
  for (int i = 0; i < n_jobs; ++i) {
    const b2_anim_compare_job& j = jobs[i];
    if (j.src == NULL || j.dst == NULL) { free(host); return B2_ERR_NULL; }
    if (j.length <= 0) { free(host); return B2_ERR_BAD_DIM; }   // assert(length > 0), anim_encode.go:321
    host[i].src = j.src; host[i].dst = j.dst; host[i].src_step = j.src_step; host[i].dst_step = j.dst_step; host[i].length = j.length;
  }
  void* dev = NULL;
  const int st = b2_upload_desc(ctx, stream, host, jb, &dev);
  free(host);
  if (st != B2_OK) return st;
  const int grid = b2_grid_for(ctx, 8, ((uint64_t)n_jobs + kWarps - 1) / kWarps);
  anim_compare_kernel<<<grid, kThreads, 0, ctx->streams[stream]>>>((const CompareJobDev*)dev, (uint32_t)n_jobs, lossless ? 1 : 0,
                                                                   max_allowed_diff, result_dev);
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}
