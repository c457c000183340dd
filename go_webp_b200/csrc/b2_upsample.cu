// b2_upsample.cu -- part 2: batched YUV420(+A) -> RGBA "fancy" upsampling, sm_100a.
//
// What it computes (bit-exact with the reference):
//   chroma: the 9-3-3-1 bilinear interpolation of UPSAMPLE_FUNC
//           (pkg/libwebp/dsp/upsampling.c.go:43-103) with its exact rounding
//           sequence, including the mirrored first/last rows and columns of
//           WebPPictureYUVAToARGB (pkg/libwebp/enc/picture_csp_enc.c.go:62-118)
//           == EmitFancyRGB (pkg/libwebp/decoder/io_dec.c.go:65-115);
//   pixel : VP8YuvToR/G/B (pkg/color/yuv/conversion.go:28-48) and the byte
//           orders of the VP8YuvTo* writers (upsampling.c.go:107-114);
//   alpha : EmitAlphaRGB / WebPDispatchAlpha (io_dec.c.go:175-195,
//           alpha_processing.go:165-180) fused into the store.
//
// Mapping: work item = 256 pixels x one output row pair (rows 2q-1, 2q, which
// share chroma rows q-1 and q; q = 0 and the row past an even height are the
// reference's single mirrored rows and fall out of clamping the chroma row).
// A thread owns 8 pixels x 2 rows: two 64-bit Y loads, four 32-bit chroma loads
// plus eight halo bytes (L1 hits), four 128-bit RGBA stores.  The interpolation
// is the reference's own 2x16-bit SWAR, here packing two columns (or U and V)
// per 32-bit register.
#include "b2_common.cuh"

namespace {

struct UpsampleJobDev {
  const uint8_t* y;
  const uint8_t* u;
  const uint8_t* v;
  const uint8_t* a;
  uint8_t* dst;
  int32_t y_stride, uv_stride, a_stride, dst_stride;
  int32_t width, height;
  uint32_t tiles_x;
  uint32_t flags;
};
enum { JOB_FAST = 1 };

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kItemW = 256;

// Canonical store modes after aliasing the premultiplied ones.
enum { ST_RGBA = 0, ST_BGRA = 1, ST_ARGB = 2, ST_RGB = 3, ST_BGR = 4, ST_4444 = 5, ST_565 = 6 };

__device__ __forceinline__ uint32_t pack4_sat(int x0, int x1, int x2, int x3) {
  uint32_t out;
  asm("{ .reg .u32 t;\n"
      "cvt.pack.sat.u8.s32.b32 t, %4, %3, 0;\n"
      "cvt.pack.sat.u8.s32.b32 %0, %2, %1, t;\n"
      "}" : "=r"(out) : "r"(x0), "r"(x1), "r"(x2), "r"(x3));
  return out;
}

// VP8YuvToRgb core on pre-shifted operands (value << 24): __umulhi(x<<24, c) ==
// (x*c) >> 8 == MultHi (conversion.go:41-43).  Returns the three 14-bit
// pre-clip values >> 6; Clip8 (conversion.go:46-48) == saturation to [0,255].
__device__ __forceinline__ void yuv_to_rgb6(uint32_t y24, uint32_t u24, uint32_t v24, int* r, int* g,
                                            int* b) {
  const int yy = (int)__umulhi(y24, 19077u);
  *r = (yy + (int)__umulhi(v24, 26149u) - 14234) >> 6;
  *g = (yy - (int)__umulhi(u24, 6419u) - (int)__umulhi(v24, 13320u) + 8708) >> 6;
  *b = (yy + (int)__umulhi(u24, 33050u) - 17685) >> 6;
}

template <int ST>
__device__ __forceinline__ uint32_t make_px4(uint32_t y24, uint32_t u24, uint32_t v24, int alpha) {
  int r, g, b;
  yuv_to_rgb6(y24, u24, v24, &r, &g, &b);
  if (ST == ST_RGBA) return pack4_sat(r, g, b, alpha);
  if (ST == ST_BGRA) return pack4_sat(b, g, r, alpha);
  return pack4_sat(alpha, r, g, b);  // ST_ARGB
}

// One interpolation step of UPSAMPLE_FUNC (upsampling.c.go:69-86) on SWAR words
// holding two independent 8-bit samples at bits 0 and 16.
struct Quad { uint32_t tf, ts, bf, bs; };  // top first/second, bottom first/second
__device__ __forceinline__ Quad interp(uint32_t tl, uint32_t t, uint32_t l, uint32_t c) {
  const uint32_t avg = tl + t + l + c + 0x00080008u;
  const uint32_t d12 = (avg + 2 * (t + l)) >> 3;
  const uint32_t d03 = (avg + 2 * (tl + c)) >> 3;
  Quad q;
  q.tf = (d12 + tl) >> 1;  // top row, pixel 2x-1
  q.ts = (d03 + t) >> 1;   // top row, pixel 2x
  q.bf = (d03 + l) >> 1;   // bottom row, pixel 2x-1
  q.bs = (d12 + c) >> 1;   // bottom row, pixel 2x
  return q;
}
__device__ __forceinline__ uint32_t lo24(uint32_t w) { return __byte_perm(w, 0u, 0x0444); }
__device__ __forceinline__ uint32_t hi24(uint32_t w) { return __byte_perm(w, 0u, 0x2444); }
template <int K> __device__ __forceinline__ uint32_t byte24(uint32_t w) {
  return __byte_perm(w, 0u, 0x0444 | (K << 12));
}

// Scalar reference pixel (any alignment, any edge): SURVEY Appendix A.2 uniform
// form of the same arithmetic, c = (((N + F + 3(A+B) + 8) >> 3) + N) >> 1.
template <int ST>
__device__ __forceinline__ void pixel_generic(const UpsampleJobDev& j, int r, int x) {
  const int uw = (j.width + 1) >> 1, uh = (j.height + 1) >> 1;
  const int rn = r >> 1, cn = x >> 1;
  const int rf = (r & 1) ? min(rn + 1, uh - 1) : max(rn - 1, 0);
  const int cf = (x & 1) ? min(cn + 1, uw - 1) : max(cn - 1, 0);
  const uint8_t* un = j.u + (ptrdiff_t)rn * j.uv_stride;
  const uint8_t* uf = j.u + (ptrdiff_t)rf * j.uv_stride;
  const uint8_t* vn = j.v + (ptrdiff_t)rn * j.uv_stride;
  const uint8_t* vf = j.v + (ptrdiff_t)rf * j.uv_stride;
  const uint32_t u = ((((uint32_t)un[cn] + uf[cf] + 3u * ((uint32_t)un[cf] + uf[cn]) + 8u) >> 3) + un[cn]) >> 1;
  const uint32_t v = ((((uint32_t)vn[cn] + vf[cf] + 3u * ((uint32_t)vn[cf] + vf[cn]) + 8u) >> 3) + vn[cn]) >> 1;
  const uint32_t yv = j.y[(ptrdiff_t)r * j.y_stride + x];
  const int alpha = j.a ? j.a[(ptrdiff_t)r * j.a_stride + x] : 0xff;
  int R, G, B;
  yuv_to_rgb6(yv << 24, u << 24, v << 24, &R, &G, &B);
  R = min(max(R, 0), 255); G = min(max(G, 0), 255); B = min(max(B, 0), 255);
  uint8_t* d = j.dst + (ptrdiff_t)r * j.dst_stride;
  if (ST == ST_RGBA) { d += 4 * x; d[0] = R; d[1] = G; d[2] = B; d[3] = alpha; }
  else if (ST == ST_BGRA) { d += 4 * x; d[0] = B; d[1] = G; d[2] = R; d[3] = alpha; }
  else if (ST == ST_ARGB) { d += 4 * x; d[0] = alpha; d[1] = R; d[2] = G; d[3] = B; }
  else if (ST == ST_RGB) { d += 3 * x; d[0] = R; d[1] = G; d[2] = B; }
  else if (ST == ST_BGR) { d += 3 * x; d[0] = B; d[1] = G; d[2] = R; }
  else if (ST == ST_4444) { d += 2 * x; d[0] = (R & 0xf0) | (G >> 4); d[1] = (B & 0xf0) | 0x0f; }
  else { d += 2 * x; d[0] = (R & 0xf8) | (G >> 5); d[1] = ((G << 3) & 0xe0) | (B >> 3); }
}

template <int ST>
__global__ void __launch_bounds__(kThreads, 4)
upsample_kernel(const UpsampleJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix,
                uint32_t n_jobs, uint32_t uniform, uint32_t total_items) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5;
  for (uint32_t item = blockIdx.x * kWarps + warp; item < total_items; item += gridDim.x * kWarps) {
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, item, &local);
    const UpsampleJobDev j = jobs[img];
    const uint32_t q = local / j.tiles_x;
    const uint32_t tx = local - q * j.tiles_x;
    const int x0 = (int)tx * kItemW + lane * 8;
    if (x0 >= j.width) continue;
    const int rt = 2 * (int)q - 1, rb = 2 * (int)q;
    const bool top_ok = (rt >= 0), bot_ok = (rb < j.height);
    constexpr bool kVec = (ST == ST_RGBA || ST == ST_BGRA || ST == ST_ARGB);
    if (!kVec || !(j.flags & JOB_FAST) || x0 + 8 > j.width) {
      const int xe = min(x0 + 8, j.width);
      for (int x = x0; x < xe; ++x) {
        if (top_ok) pixel_generic<ST>(j, rt, x);
        if (bot_ok) pixel_generic<ST>(j, rb, x);
      }
      continue;
    }
    if (kVec) {
      const int uw = (j.width + 1) >> 1, uh = (j.height + 1) >> 1;
      const int ct = max((int)q - 1, 0), cc = min((int)q, uh - 1);
      const int c0 = x0 >> 1;
      const int cl = max(c0 - 1, 0), cr = min(c0 + 4, uw - 1);
      const uint8_t* ut = j.u + (ptrdiff_t)ct * j.uv_stride;
      const uint8_t* uc = j.u + (ptrdiff_t)cc * j.uv_stride;
      const uint8_t* vt = j.v + (ptrdiff_t)ct * j.uv_stride;
      const uint8_t* vc = j.v + (ptrdiff_t)cc * j.uv_stride;
      // x0 + 8 <= width  =>  c0 + 3 <= uw - 1: the 32-bit chroma loads stay in the row
      const uint32_t wtu = __ldg((const uint32_t*)(ut + c0)), wcu = __ldg((const uint32_t*)(uc + c0));
      const uint32_t wtv = __ldg((const uint32_t*)(vt + c0)), wcv = __ldg((const uint32_t*)(vc + c0));
      const uint32_t ltu = __ldg(ut + cl), rtu = __ldg(ut + cr), lcu = __ldg(uc + cl), rcu = __ldg(uc + cr);
      const uint32_t ltv = __ldg(vt + cl), rtv = __ldg(vt + cr), lcv = __ldg(vc + cl), rcv = __ldg(vc + cr);
      uint2 yt = make_uint2(0, 0), yb = make_uint2(0, 0);
      if (top_ok) yt = b2_ldg_stream_v2(j.y + (ptrdiff_t)rt * j.y_stride + x0);
      if (bot_ok) yb = b2_ldg_stream_v2(j.y + (ptrdiff_t)rb * j.y_stride + x0);
      uint2 at = make_uint2(0xffffffffu, 0xffffffffu), ab = at;
      if (j.a != nullptr) {
        if (top_ok) at = b2_ldg_stream_v2(j.a + (ptrdiff_t)rt * j.a_stride + x0);
        if (bot_ok) ab = b2_ldg_stream_v2(j.a + (ptrdiff_t)rb * j.a_stride + x0);
      }
      // even / odd columns of each chroma word, two samples per register
      const uint32_t etu = wtu & 0x00ff00ffu, otu = (wtu >> 8) & 0x00ff00ffu;
      const uint32_t ecu = wcu & 0x00ff00ffu, ocu = (wcu >> 8) & 0x00ff00ffu;
      const uint32_t etv = wtv & 0x00ff00ffu, otv = (wtv >> 8) & 0x00ff00ffu;
      const uint32_t ecv = wcv & 0x00ff00ffu, ocv = (wcv >> 8) & 0x00ff00ffu;
      // column pairs (1,3): (c0,c0+1),(c0+2,c0+3)   pixels x0+1,x0+2 | x0+5,x0+6
      const Quad au = interp(etu, otu, ecu, ocu);
      const Quad av = interp(etv, otv, ecv, ocv);
      // column pairs (2,4): (c0+1,c0+2),(c0+3,right) pixels x0+3,x0+4 | x0+7,(x0+8)
      const Quad bu = interp(otu, (etu >> 16) | (rtu << 16), ocu, (ecu >> 16) | (rcu << 16));
      const Quad bv = interp(otv, (etv >> 16) | (rtv << 16), ocv, (ecv >> 16) | (rcv << 16));
      // column pair 0: (left,c0) for pixel x0, U in the low lane and V in the high lane
      const Quad cq = interp(ltu | (ltv << 16), (etu & 0xffu) | (etv << 16), lcu | (lcv << 16),
                             (ecu & 0xffu) | (ecv << 16));
      if (top_ok) {
        uint4 o0, o1;
        o0.x = make_px4<ST>(byte24<0>(yt.x), lo24(cq.ts), hi24(cq.ts), at.x & 0xff);
        o0.y = make_px4<ST>(byte24<1>(yt.x), lo24(au.tf), lo24(av.tf), (at.x >> 8) & 0xff);
        o0.z = make_px4<ST>(byte24<2>(yt.x), lo24(au.ts), lo24(av.ts), (at.x >> 16) & 0xff);
        o0.w = make_px4<ST>(byte24<3>(yt.x), lo24(bu.tf), lo24(bv.tf), at.x >> 24);
        o1.x = make_px4<ST>(byte24<0>(yt.y), lo24(bu.ts), lo24(bv.ts), at.y & 0xff);
        o1.y = make_px4<ST>(byte24<1>(yt.y), hi24(au.tf), hi24(av.tf), (at.y >> 8) & 0xff);
        o1.z = make_px4<ST>(byte24<2>(yt.y), hi24(au.ts), hi24(av.ts), (at.y >> 16) & 0xff);
        o1.w = make_px4<ST>(byte24<3>(yt.y), hi24(bu.tf), hi24(bv.tf), at.y >> 24);
        uint8_t* d = j.dst + (ptrdiff_t)rt * j.dst_stride + 4 * x0;
        b2_stg_stream_v4(d, o0);
        b2_stg_stream_v4(d + 16, o1);
      }
      if (bot_ok) {
        uint4 o0, o1;
        o0.x = make_px4<ST>(byte24<0>(yb.x), lo24(cq.bs), hi24(cq.bs), ab.x & 0xff);
        o0.y = make_px4<ST>(byte24<1>(yb.x), lo24(au.bf), lo24(av.bf), (ab.x >> 8) & 0xff);
        o0.z = make_px4<ST>(byte24<2>(yb.x), lo24(au.bs), lo24(av.bs), (ab.x >> 16) & 0xff);
        o0.w = make_px4<ST>(byte24<3>(yb.x), lo24(bu.bf), lo24(bv.bf), ab.x >> 24);
        o1.x = make_px4<ST>(byte24<0>(yb.y), lo24(bu.bs), lo24(bv.bs), ab.y & 0xff);
        o1.y = make_px4<ST>(byte24<1>(yb.y), hi24(au.bf), hi24(av.bf), (ab.y >> 8) & 0xff);
        o1.z = make_px4<ST>(byte24<2>(yb.y), hi24(au.bs), hi24(av.bs), (ab.y >> 16) & 0xff);
        o1.w = make_px4<ST>(byte24<3>(yb.y), hi24(bu.bf), hi24(bv.bf), ab.y >> 24);
        uint8_t* d = j.dst + (ptrdiff_t)rb * j.dst_stride + 4 * x0;
        b2_stg_stream_v4(d, o0);
        b2_stg_stream_v4(d + 16, o1);
      }
    }
  }
}

int canonical_store_mode(int mode) {
  switch (mode) {
    case B2_MODE_RGBA: case B2_MODE_rgbA: return ST_RGBA;   // upsampling.c.go:278,280
    case B2_MODE_BGRA: case B2_MODE_bgrA: return ST_BGRA;   // :279,281
    case B2_MODE_ARGB: case B2_MODE_Argb: return ST_ARGB;   // :284,287
    case B2_MODE_RGB: return ST_RGB;
    case B2_MODE_BGR: return ST_BGR;
    case B2_MODE_RGBA_4444: case B2_MODE_rgbA_4444: return ST_4444;  // :285,288
    case B2_MODE_RGB_565: return ST_565;
    default: return -1;
  }
}

}  // namespace

extern "C" int b2_yuv420a_to_rgba_fancy(b2_ctx* ctx, int stream, const b2_upsample_job* jobs,
                                        int n_jobs, int mode) {
  B2_CHECK_CTX(ctx, stream);
  if (n_jobs < 0) return B2_ERR_INVALID_CONFIG;
  if (n_jobs == 0) return B2_OK;
  if (jobs == NULL) return B2_ERR_NULL;
  const int st_mode = canonical_store_mode(mode);
  if (st_mode < 0) return B2_ERR_INVALID_CONFIG;
  static const int kBpp[7] = {4, 4, 4, 3, 3, 2, 2};  // kModeBpp, buffer_dec.c.go:28-32
  const int bpp = kBpp[st_mode];

  const size_t jb = sizeof(UpsampleJobDev) * (size_t)n_jobs;
  const size_t pb = sizeof(uint32_t) * ((size_t)n_jobs + 1);
  const size_t pre_off = (jb + 15) & ~(size_t)15;
  if (pre_off + pb > B2_DESC_RING_BYTES) return B2_ERR_INVALID_CONFIG;
  uint8_t* host = (uint8_t*)malloc(pre_off + pb);
  if (host == NULL) return B2_ERR_OOM;
  UpsampleJobDev* dj = (UpsampleJobDev*)host;
  uint32_t* pre = (uint32_t*)(host + pre_off);
  uint64_t total = 0;
  uint32_t uniform = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const b2_upsample_job& j = jobs[i];
    if (j.y == NULL || j.u == NULL || j.v == NULL || j.dst == NULL) { free(host); return B2_ERR_NULL; }
    if (j.width <= 0 || j.height <= 0 || j.width > B2_MAX_DIMENSION || j.height > B2_MAX_DIMENSION) {
      free(host); return B2_ERR_BAD_DIM;
    }
    const int64_t ds = j.dst_stride < 0 ? -(int64_t)j.dst_stride : j.dst_stride;
    if (ds < (int64_t)bpp * j.width) { free(host); return B2_ERR_BAD_DIM; }  // buffer_dec.c.go:58-61
    UpsampleJobDev& d = dj[i];
    d.y = j.y; d.u = j.u; d.v = j.v; d.a = j.a; d.dst = j.dst;
    d.y_stride = j.y_stride; d.uv_stride = j.uv_stride; d.a_stride = j.a_stride;
    d.dst_stride = j.dst_stride; d.width = j.width; d.height = j.height;
    d.tiles_x = (uint32_t)((j.width + kItemW - 1) / kItemW);
    bool fast = (((uintptr_t)j.y | (uintptr_t)(uint32_t)j.y_stride) & 7) == 0 &&
                (((uintptr_t)j.u | (uintptr_t)j.v | (uintptr_t)(uint32_t)j.uv_stride) & 3) == 0 &&
                (((uintptr_t)j.dst | (uintptr_t)(uint32_t)j.dst_stride) & 15) == 0;
    if (j.a != NULL && (((uintptr_t)j.a | (uintptr_t)(uint32_t)j.a_stride) & 7) != 0) fast = false;
    d.flags = fast ? JOB_FAST : 0;
    const uint32_t items = d.tiles_x * (uint32_t)((j.height >> 1) + 1);
    if (i == 0) uniform = items; else if (items != uniform) uniform = 0;
    pre[i] = (uint32_t)total;
    total += items;
    if (total > 0x7fffffffull) { free(host); return B2_ERR_BAD_DIM; }
  }
  pre[n_jobs] = (uint32_t)total;
  void* dev = NULL;
  int st = b2_upload_desc(ctx, stream, host, pre_off + pb, &dev);
  free(host);
  if (st != B2_OK) return st;
  const UpsampleJobDev* djobs = (const UpsampleJobDev*)dev;
  const uint32_t* dpre = (const uint32_t*)((uint8_t*)dev + pre_off);
  cudaStream_t s = ctx->streams[stream];
  const int grid = b2_grid_for(ctx, 8, (total + kWarps - 1) / kWarps);
  const uint32_t tot = (uint32_t)total;
  switch (st_mode) {
    case ST_RGBA: upsample_kernel<ST_RGBA><<<grid, kThreads, 0, s>>>(djobs, dpre, n_jobs, uniform, tot); break;
    case ST_BGRA: upsample_kernel<ST_BGRA><<<grid, kThreads, 0, s>>>(djobs, dpre, n_jobs, uniform, tot); break;
    case ST_ARGB: upsample_kernel<ST_ARGB><<<grid, kThreads, 0, s>>>(djobs, dpre, n_jobs, uniform, tot); break;
    case ST_RGB: upsample_kernel<ST_RGB><<<grid, kThreads, 0, s>>>(djobs, dpre, n_jobs, uniform, tot); break;
    case ST_BGR: upsample_kernel<ST_BGR><<<grid, kThreads, 0, s>>>(djobs, dpre, n_jobs, uniform, tot); break;
    case ST_4444: upsample_kernel<ST_4444><<<grid, kThreads, 0, s>>>(djobs, dpre, n_jobs, uniform, tot); break;
    default: upsample_kernel<ST_565><<<grid, kThreads, 0, s>>>(djobs, dpre, n_jobs, uniform, tot); break;
  }
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}
