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
#include <stdlib.h>

#include "b2_common.cuh"
#include "b2_upsample_core.cuh"

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
  int32_t row_begin, row_end;  // output rows to produce (whole image: 0, height)
  uint32_t q_lo;               // first row-pair unit: (row_begin + 1) >> 1
  uint32_t pad;
};
enum { JOB_FAST = 1, JOB_SWAP16 = 2 };   // JOB_SWAP16: ba / gb byte first (b2_ctx_set_swap_16bit_csp)

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
  else if (ST == ST_4444) {  // VP8YuvToRgba4444 + the alpha nibble of EmitAlphaRGBA4444 (io_dec.c.go:197-226: byte 2i in the Go text's order)
    const int sw = (j.flags & JOB_SWAP16) ? 1 : 0;
    d += 2 * x; d[sw] = (R & 0xf0) | (G >> 4); d[sw ^ 1] = (B & 0xf0) | (alpha >> 4);
  } else {
    const int sw = (j.flags & JOB_SWAP16) ? 1 : 0;
    d += 2 * x; d[sw] = (R & 0xf8) | (G >> 5); d[sw ^ 1] = ((G << 3) & 0xe0) | (B >> 3);
  }
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
    const uint32_t ql = local / j.tiles_x;
    const uint32_t tx = local - ql * j.tiles_x;
    const uint32_t q = ql + j.q_lo;
    const int x0 = (int)tx * kItemW + lane * 8;
    if (x0 >= j.width) continue;
    const int rt = 2 * (int)q - 1, rb = 2 * (int)q;
    const bool top_ok = (rt >= j.row_begin) && (rt < j.row_end), bot_ok = (rb >= j.row_begin) && (rb < j.row_end);
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

// Packed arithmetic (b2_upsample_core.cuh): converts and stores 8 pixels of one output
// row from four pixel PAIRS, two pixels per 32-bit register all the way to the
// store.  Pair p holds pixels (1,5), (2,6), (3,7), (0,4) -- the lane layout the
// interpolation produces.  u2/v2 = interpolated chroma, clean 8-bit lanes.
template <int ST, bool HAS_ALPHA>
__device__ __forceinline__ void emit_row8_packed(uint2 yw, uint2 aw, const uint32_t (&u2)[4],
                                                 const uint32_t (&v2)[4], uint8_t* dp, bool store_ok,
                                                 const B2YuvConsts& kc, uint8_t* xs, uint32_t xw0, uint32_t xw1,
                                                 uint32_t sel16 = 0x7351u) {
  const uint32_t ye_x = b2_even(yw.x), yo_x = b2_odd(yw.x), ye_y = b2_even(yw.y), yo_y = b2_odd(yw.y);
  const uint32_t y2[4] = {b2_prmt(yo_x, yo_y, 0x5410), b2_prmt(ye_x, ye_y, 0x7632),
                          b2_prmt(yo_x, yo_y, 0x7632), b2_prmt(ye_x, ye_y, 0x5410)};
  uint32_t ap01 = 0xffffffffu, ap23 = 0xffffffffu;
  if (HAS_ALPHA) {
    ap01 = b2_prmt(aw.x, aw.y, 0x6251);  // a1 a5 a2 a6
    ap23 = b2_prmt(aw.x, aw.y, 0x4073);  // a3 a7 a0 a4
  }
  uint32_t o[8];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const B2Rgb14 c = b2_yuv_pair(b2_luma14(y2[p], kc), u2[p], v2[p], kc);
    const uint32_t r4 = b2_to_byte1(c.r), g4 = b2_to_byte1(c.g), b4 = b2_to_byte1(c.b);
    const uint32_t ap = (p < 2) ? ap01 : ap23;
    const bool k0 = (p & 1) == 0;  // alpha bytes (0,1) or (2,3) of ap
    uint32_t lo, hi;
    if (ST == ST_RGBA || ST == ST_BGRA) {
      const uint32_t c0 = (ST == ST_RGBA) ? r4 : b4, c2 = (ST == ST_RGBA) ? b4 : r4;
      const uint32_t cg = b2_prmt(c0, g4, 0x7351);                  // c0 g c0' g'
      const uint32_t ca = b2_prmt(c2, ap, k0 ? 0x5341 : 0x7361);    // c2 a c2' a'
      lo = b2_prmt(cg, ca, 0x5410);
      hi = b2_prmt(cg, ca, 0x7632);
    } else if (ST == ST_ARGB) {
      const uint32_t ar = b2_prmt(r4, ap, k0 ? 0x3514 : 0x3716);    // a r a' r'
      const uint32_t gb = b2_prmt(g4, b4, 0x7351);                  // g b g' b'
      lo = b2_prmt(ar, gb, 0x5410);
      hi = b2_prmt(ar, gb, 0x7632);
    } else if (ST == ST_565 || ST == ST_4444) {
      // VP8YuvToRgb565 / VP8YuvToRgba4444 (conversion.go) on the pair: the 8-bit values sit in bytes 1 and 3.
      // lo = the pair's two 16-bit pixels; sel16 puts the rg byte first (stock libwebp) or second (go-webp text)
      uint32_t rg, gb;
      if (ST == ST_565) {
        rg = (r4 & 0xf800f800u) | ((g4 >> 5) & 0x07000700u);
        gb = ((g4 << 3) & 0xe000e000u) | ((b4 >> 3) & 0x1f001f00u);
      } else {
        const uint32_t a2 = HAS_ALPHA ? b2_prmt(ap, 0u, k0 ? 0x1404 : 0x3424) : 0xff00ff00u;   // EmitAlphaRGBA4444's nibble
        rg = (r4 & 0xf000f000u) | ((g4 >> 4) & 0x0f000f00u);
        gb = (b4 & 0xf000f000u) | ((a2 >> 4) & 0x0f000f00u);
      }
      lo = b2_prmt(rg, gb, sel16);
      hi = 0u;
    } else {  // ST_RGB / ST_BGR: three bytes per pixel, byte 3 is dropped below
      const uint32_t c0 = (ST == ST_RGB) ? r4 : b4, c2 = (ST == ST_RGB) ? b4 : r4;
      const uint32_t cg = b2_prmt(c0, g4, 0x7351);
      lo = b2_prmt(cg, c2, 0x5510);
      hi = b2_prmt(cg, c2, 0x7732);
    }
    const int first = (p == 3) ? 0 : p + 1;
    o[first] = lo;
    o[first + 4] = hi;
  }
  if (ST == ST_565 || ST == ST_4444) {
    // o[k] (k = 0..3) = pixels k and k + 4 as 16-bit halves; a lane's 8 pixels are 16 contiguous bytes and the
    // lanes of a pass follow each other: one 128-bit store per lane covers 512 contiguous bytes per instruction
    if (store_ok) b2_stg_stream_v4(dp, make_uint4(b2_prmt(o[0], o[1], 0x5410), b2_prmt(o[2], o[3], 0x5410),
                                                   b2_prmt(o[0], o[1], 0x7632), b2_prmt(o[2], o[3], 0x7632)));
  } else if (ST == ST_RGB || ST == ST_BGR) {
    uint32_t w[6];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t* q = o + 4 * h;
      w[3 * h + 0] = b2_prmt(q[0], q[1], 0x4210);
      w[3 * h + 1] = b2_prmt(q[1], q[2], 0x5421);
      w[3 * h + 2] = b2_prmt(q[2], q[3], 0x6542);
    }
    if (xs != nullptr) {  // warp-uniform: a full pass goes through the transposition buffer, 24 bytes per lane
      *reinterpret_cast<uint2*>(xs + xw0) = make_uint2(w[0], w[1]);
      *reinterpret_cast<uint2*>(xs + xw0 + 8) = make_uint2(w[2], w[3]);
      *reinterpret_cast<uint2*>(xs + xw0 + 16) = make_uint2(w[4], w[5]);
    } else if (store_ok) {
      b2_stg_stream_v2(dp, make_uint2(w[0], w[1]));
      b2_stg_stream_v2(dp + 8, make_uint2(w[2], w[3]));
      b2_stg_stream_v2(dp + 16, make_uint2(w[4], w[5]));
    }
  } else {  // 4-byte modes: through the warp's transposition buffer (see XposeLane)
    *reinterpret_cast<uint4*>(xs + xw0) = make_uint4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<uint4*>(xs + xw1) = make_uint4(o[4], o[5], o[6], o[7]);
  }
}

template <int ST> __host__ __device__ constexpr int bulk_bpp() {
  return (ST == ST_RGB || ST == ST_BGR) ? 3 : (ST == ST_565 || ST == ST_4444) ? 2 : 4;
}

// ---- bulk-copy (TMA) staged kernel -------------------------------------------------
// Same packed arithmetic, cheaper plumbing.  ncu on the cp.async kernel: 535 warp
// instructions per 256-pixel step, ~85 of them loader bookkeeping executed by every
// lane, issue slots 73% busy -- the kernel is bound by instruction issue, not by HBM.
// Here a warp owns a 2-stage ring whose stages hold a 1024- (512-) pixel x 2-row tile; each
// input row of a tile is ONE cp.async.bulk (6 per tile, 8 with an alpha plane), all
// issued by lane 0 and completing on the stage's mbarrier: the copy engine takes its
// operands from uniform registers, so a single issuing lane is the cheap form.
// Preconditions (checked on the host): y/u/v(/a) pointers and strides are multiples
// of 16 bytes, strides positive, so every copy is 16-byte aligned and sized.
constexpr int kBulkThreads = 128;
constexpr int kBulkWarps = kBulkThreads / 32;
#ifndef B2_BULK_CTAS
#define B2_BULK_CTAS 6
#endif
#ifndef B2_BULK_TW
#define B2_BULK_TW 1024
#endif
#ifndef B2_PASS_UNROLL
#define B2_PASS_UNROLL 2
#endif
constexpr int kBulkCtasPerSm = B2_BULK_CTAS;
constexpr int kPassUnroll = B2_PASS_UNROLL;
// Stage = TW pixels x 2 rows: 1024 pixels without an alpha plane, 512 with one (shared
// memory: 4 warps x 2 stages x 4.1 KB resp. 3.1 KB per CTA, six CTAs per SM).
template <bool HAS_ALPHA> struct BulkLayout {
  static constexpr int kTW = HAS_ALPHA ? 512 : B2_BULK_TW;
  static constexpr int kY = 0;               // yT[TW] yB[TW]
  static constexpr int kC = 2 * kTW;         // uT uC vT vC, each [16 halo][TW/2][16 halo]
  static constexpr int kCRow = kTW / 2 + 32;
  static constexpr int kA = kC + 4 * kCRow;  // aT[TW] aB[TW]
  static constexpr int kBytes = HAS_ALPHA ? kA + 2 * kTW : kA;
};
// Warp-uniform values: a shuffle from lane 0 tells ptxas the value is the same in every
// lane, so the producer's arithmetic runs on the uniform datapath and the bulk copies
// take their operands from uniform registers without a per-copy R2UR/ELECT loop.
__device__ __forceinline__ uint32_t uni(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ int uni(int v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ const uint8_t* uni(const uint8_t* p) {
  const uint64_t v = (uint64_t)p;
  const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, 0), hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), 0);
  return (const uint8_t*)(((uint64_t)hi << 32) | lo);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\t"
               "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@!p bra W;\n\t}" :: "r"(bar), "r"(parity) : "memory");
}

// Converts 8 pixels x 2 rows from one stage.  sy/sa -> this lane's 8 luma / alpha
// bytes of the top row (bottom row 512 bytes further); sc -> this lane's 4 chroma
// bytes in row uT (uC, vT, vC follow at kCRow steps); loff/roff select the left /
// right neighbour sample (mirrored at the image edges: upsampling.c.go:57,61,93,98).
// Store coalescing.  A lane produces 32 contiguous bytes per row; storing them as two
// 128-bit stores makes every store instruction write HALF of each 32-byte sector, and
// ncu shows the price: twice the ideal number of L2 write sectors, the SM->L2 write
// path 64-69% busy and the kernel stuck at ~4.1 TB/s whatever else changes.  So a full
// 256-pixel pass goes through a per-warp 2 KB buffer instead: each lane writes its two
// 16-byte chunks, reads back chunk `lane` and chunk `32 + lane`, and the two store
// instructions each cover 512 contiguous bytes.  Chunk c lives at c ^ ((c >> 3) & 1),
// which keeps both the writes (chunks 2*lane, 2*lane+1) and the reads bank-conflict free.
struct XposeLane { uint32_t w0, w1, r; };
__device__ __forceinline__ XposeLane xpose_lane(int lane) {
  const uint32_t b = (uint32_t)(lane >> 2) & 1u;
  XposeLane x;
  x.w0 = ((2u * lane) ^ b) * 16u;
  x.w1 = ((2u * lane + 1u) ^ b) * 16u;
  x.r = ((uint32_t)lane ^ (((uint32_t)lane >> 3) & 1u)) * 16u;
  return x;
}

// One 256-pixel pass of a warp: every lane converts 8 pixels x 2 rows (lanes past the
// end of the row compute on stale stage bytes and store nothing).  xpass = first pixel
// of the pass, wv = number of pixels of the row the vector path covers (the width for
// the 4-byte modes, whose stores are masked per pixel; the width rounded down to 8 for
// the 3-byte modes); drt/drb = start of the two output rows; xs = this warp's
// transposition buffer (top row at 0, bottom row at 1024).
template <int ST, bool HAS_ALPHA, bool EDGE>
__device__ __forceinline__ void convert_pass(const uint8_t* sy, const uint8_t* sa, const uint8_t* sc, int lane,
                                             int xpass, int wv, int uw, bool top_ok, bool bot_ok, uint8_t* drt,
                                             uint8_t* drb, const B2YuvConsts& kc, uint8_t* xs, const XposeLane& xl,
                                             uint32_t sel16) {
  constexpr int R = BulkLayout<HAS_ALPHA>::kCRow;
  constexpr int kTileW = BulkLayout<HAS_ALPHA>::kTW;
  constexpr int kBpp = bulk_bpp<ST>();
  const int xh = xpass + lane * 8;
  constexpr int loff = -1, roff = 4;  // left / right neighbour samples; the image edges are fixed up below
  const uint2 yt = *reinterpret_cast<const uint2*>(sy);
  const uint2 yb = *reinterpret_cast<const uint2*>(sy + kTileW);
  uint2 at = make_uint2(0, 0), ab = make_uint2(0, 0);
  if (HAS_ALPHA) {
    at = *reinterpret_cast<const uint2*>(sa);
    ab = *reinterpret_cast<const uint2*>(sa + kTileW);
  }
  uint32_t wtu = *reinterpret_cast<const uint32_t*>(sc);
  uint32_t wcu = *reinterpret_cast<const uint32_t*>(sc + R);
  uint32_t wtv = *reinterpret_cast<const uint32_t*>(sc + 2 * R);
  uint32_t wcv = *reinterpret_cast<const uint32_t*>(sc + 3 * R);
  const uint8_t* lp = sc + loff;
  const uint8_t* rp = sc + roff;
  uint32_t ltu = lp[0], lcu = lp[R], ltv = lp[2 * R], lcv = lp[3 * R];
  uint32_t rtu = rp[0], rcu = rp[R], rtv = rp[2 * R], rcv = rp[3 * R];
  // Image edges mirror the chroma samples (upsampling.c.go:57,61,93,98).  Only the EDGE
  // instantiation (first and last pass of a row) carries this code and the masked
  // stores below; interior passes are straight-line.
  if (EDGE && xpass == 0 && lane == 0) { ltu = wtu & 0xff; lcu = wcu & 0xff; ltv = wtv & 0xff; lcv = wcv & 0xff; }
  if (EDGE && xpass + 256 >= 2 * uw) {
    // kk = position of the last chroma column in this lane's 4-sample word: 3 -> the right
    // neighbour is its copy; 0..2 (a row that ends inside this lane's 8 pixels) -> the
    // sample after it inside the word is its copy
    const int kk = uw - 1 - (xh >> 1);
    if (kk == 3) { rtu = wtu >> 24; rcu = wcu >> 24; rtv = wtv >> 24; rcv = wcv >> 24; }
    else if (kk >= 0 && kk < 3) {
      const uint32_t sel = (kk == 0) ? 0x3200u : (kk == 1) ? 0x3110u : 0x2210u;
      wtu = b2_prmt(wtu, 0u, sel); wcu = b2_prmt(wcu, 0u, sel); wtv = b2_prmt(wtv, 0u, sel); wcv = b2_prmt(wcv, 0u, sel);
    }
  }
  const uint32_t etu = b2_even(wtu), otu = b2_odd(wtu), mtu = b2_prmt(etu, rtu, 0x5432);
  const uint32_t ecu = b2_even(wcu), ocu = b2_odd(wcu), mcu = b2_prmt(ecu, rcu, 0x5432);
  const uint32_t etv = b2_even(wtv), otv = b2_odd(wtv), mtv = b2_prmt(etv, rtv, 0x5432);
  const uint32_t ecv = b2_even(wcv), ocv = b2_odd(wcv), mcv = b2_prmt(ecv, rcv, 0x5432);
  // column pairs (c0,c0+1),(c0+2,c0+3) -> pixels 1,2 | 5,6 ; (c0+1,c0+2),(c0+3,c0+4) -> 3,4 | 7,(8)
  const B2Quad16 au = b2_interp16(etu, otu, ecu, ocu), bu = b2_interp16(otu, mtu, ocu, mcu);
  const B2Quad16 av = b2_interp16(etv, otv, ecv, ocv), bv = b2_interp16(otv, mtv, ocv, mcv);
  // column pair (left,c0) -> pixel 0, U in the low lane and V in the high lane
  const B2Quad16 cq = b2_interp16(b2_prmt(ltu, ltv, 0x5410), b2_prmt(etu, etv, 0x5410),
                                  b2_prmt(lcu, lcv, 0x5410), b2_prmt(ecu, ecv, 0x5410));
  const bool lane_ok = !EDGE || xh + 8 <= wv;
  // 3-byte modes: only full passes are transposed (768 contiguous bytes per row); the
  // last, partial pass of a row stores straight from the lanes
  const bool full3 = (kBpp == 3) && (!EDGE || xpass + 256 <= wv);
  uint8_t* const xst = (kBpp == 4 || full3) ? xs : nullptr;
  const uint32_t xw0 = (kBpp == 4) ? xl.w0 : (uint32_t)lane * 24u;
  if (top_ok) {
    const uint32_t xc = b2_hi8(cq.ts);
    const uint32_t u2[4] = {b2_hi8(au.tf), b2_hi8(au.ts), b2_hi8(bu.tf), b2_prmt(xc, b2_hi8(bu.ts), 0x5410)};
    const uint32_t v2[4] = {b2_hi8(av.tf), b2_hi8(av.ts), b2_hi8(bv.tf), b2_prmt(xc, b2_hi8(bv.ts), 0x5432)};
    emit_row8_packed<ST, HAS_ALPHA>(yt, at, u2, v2, drt + xh * kBpp, lane_ok, kc, xst, xw0, xl.w1, sel16);
  }
  if (bot_ok) {
    const uint32_t xc = b2_hi8(cq.bs);
    const uint32_t u2[4] = {b2_hi8(au.bf), b2_hi8(au.bs), b2_hi8(bu.bf), b2_prmt(xc, b2_hi8(bu.bs), 0x5410)};
    const uint32_t v2[4] = {b2_hi8(av.bf), b2_hi8(av.bs), b2_hi8(bv.bf), b2_prmt(xc, b2_hi8(bv.bs), 0x5432)};
    emit_row8_packed<ST, HAS_ALPHA>(yb, ab, u2, v2, drb + xh * kBpp, lane_ok, kc, xst ? xst + 1024 : nullptr, xw0, xl.w1, sel16);
  }
  if (kBpp == 4) {
    __syncwarp();
    const int xa = xpass + lane * 4;  // this lane stores pixels xa..xa+3 and xa+128..xa+131
    // wv is the image width here: the last group of a row may hold 1..3 pixels
    auto put4 = [&](uint8_t* d, const uint8_t* src, int n) {
      const uint4 v = *reinterpret_cast<const uint4*>(src);
      if (n >= 4) { b2_stg_stream_v4(d, v); return; }
      if (n > 0) b2_stg_stream_u32(d, v.x);
      if (n > 1) b2_stg_stream_u32(d + 4, v.y);
      if (n > 2) b2_stg_stream_u32(d + 8, v.z);
    };
    const bool full = !EDGE || xpass + 256 <= wv;  // warp-uniform: every pass but the last one of a row
    const int n0 = full ? 4 : wv - xa, n1 = full ? 4 : wv - xa - 128;
    if (top_ok) {
      uint8_t* d = drt + xa * 4;
      if (full) {
        b2_stg_stream_v4(d, *reinterpret_cast<const uint4*>(xs + xl.r));
        b2_stg_stream_v4(d + 512, *reinterpret_cast<const uint4*>(xs + 512 + xl.r));
      } else {
        put4(d, xs + xl.r, n0);
        put4(d + 512, xs + 512 + xl.r, n1);
      }
    }
    if (bot_ok) {
      uint8_t* d = drb + xa * 4;
      if (full) {
        b2_stg_stream_v4(d, *reinterpret_cast<const uint4*>(xs + 1024 + xl.r));
        b2_stg_stream_v4(d + 512, *reinterpret_cast<const uint4*>(xs + 1536 + xl.r));
      } else {
        put4(d, xs + 1024 + xl.r, n0);
        put4(d + 512, xs + 1536 + xl.r, n1);
      }
    }
    __syncwarp();
  } else if (full3) {
    __syncwarp();
    const uint32_t ro = (uint32_t)lane * 16u;  // 768 bytes per row: chunk `lane`, and chunk 32 + lane for lanes < 16
    if (top_ok) {
      uint8_t* d = drt + xpass * 3 + ro;
      b2_stg_stream_v4(d, *reinterpret_cast<const uint4*>(xs + ro));
      if (lane < 16) b2_stg_stream_v4(d + 512, *reinterpret_cast<const uint4*>(xs + 512 + ro));
    }
    if (bot_ok) {
      uint8_t* d = drb + xpass * 3 + ro;
      b2_stg_stream_v4(d, *reinterpret_cast<const uint4*>(xs + 1024 + ro));
      if (lane < 16) b2_stg_stream_v4(d + 512, *reinterpret_cast<const uint4*>(xs + 1536 + ro));
    }
    __syncwarp();
  }
}

// The last pass of a row when at most 128 pixels are left (1920 = 7.5 passes of 256): instead of
// half the lanes idling through a whole pass, lanes 0-15 convert the top row and lanes 16-31 the
// bottom row, 8 pixels each.  The bottom row's interpolation is the top row's with the two chroma
// rows swapped (tf/ts of b2_interp16 with (tl,t) <-> (l,c) are its bf/bs), so every lane runs the
// same single-row code on its own operands: about half the work of a full pass.
template <int ST, bool HAS_ALPHA>
__device__ __forceinline__ void convert_pass_split(const uint8_t* sp_y, const uint8_t* sp_a, const uint8_t* sp_c, int lane,
                                                   int xpass, int wv, int uw, uint8_t* drt, uint8_t* drb,
                                                   const B2YuvConsts& kc, uint8_t* xs, const XposeLane& xl, uint32_t sel16) {
  constexpr int R = BulkLayout<HAS_ALPHA>::kCRow;
  constexpr int kTileW = BulkLayout<HAS_ALPHA>::kTW;
  constexpr int kBpp = bulk_bpp<ST>();
  const int l = lane & 15;
  const bool bot = lane >= 16;
  const int xh = xpass + l * 8;
  const int row_off = bot ? kTileW : 0;
  const uint2 yw = *reinterpret_cast<const uint2*>(sp_y + l * 8 + row_off);
  uint2 aw = make_uint2(0, 0);
  if (HAS_ALPHA) aw = *reinterpret_cast<const uint2*>(sp_a + l * 8 + row_off);
  const uint8_t* sc = sp_c + l * 4;
  const int on = bot ? R : 0, of = bot ? 0 : R;   // chroma row next to this lane's output row / the other one
  uint32_t wtu = *reinterpret_cast<const uint32_t*>(sc + on);
  uint32_t wcu = *reinterpret_cast<const uint32_t*>(sc + of);
  uint32_t wtv = *reinterpret_cast<const uint32_t*>(sc + 2 * R + on);
  uint32_t wcv = *reinterpret_cast<const uint32_t*>(sc + 2 * R + of);
  uint32_t ltu = sc[on - 1], lcu = sc[of - 1], ltv = sc[2 * R + on - 1], lcv = sc[2 * R + of - 1];
  uint32_t rtu = sc[on + 4], rcu = sc[of + 4], rtv = sc[2 * R + on + 4], rcv = sc[2 * R + of + 4];
  if (xpass == 0 && l == 0) { ltu = wtu & 0xff; lcu = wcu & 0xff; ltv = wtv & 0xff; lcv = wcv & 0xff; }
  {  // this is the last pass of the row: mirror the last chroma column (upsampling.c.go:93,98)
    const int kk = uw - 1 - (xh >> 1);
    if (kk == 3) { rtu = wtu >> 24; rcu = wcu >> 24; rtv = wtv >> 24; rcv = wcv >> 24; }
    else if (kk >= 0 && kk < 3) {
      const uint32_t sel = (kk == 0) ? 0x3200u : (kk == 1) ? 0x3110u : 0x2210u;
      wtu = b2_prmt(wtu, 0u, sel); wcu = b2_prmt(wcu, 0u, sel); wtv = b2_prmt(wtv, 0u, sel); wcv = b2_prmt(wcv, 0u, sel);
    }
  }
  const uint32_t etu = b2_even(wtu), otu = b2_odd(wtu), mtu = b2_prmt(etu, rtu, 0x5432);
  const uint32_t ecu = b2_even(wcu), ocu = b2_odd(wcu), mcu = b2_prmt(ecu, rcu, 0x5432);
  const uint32_t etv = b2_even(wtv), otv = b2_odd(wtv), mtv = b2_prmt(etv, rtv, 0x5432);
  const uint32_t ecv = b2_even(wcv), ocv = b2_odd(wcv), mcv = b2_prmt(ecv, rcv, 0x5432);
  const B2Quad16 au = b2_interp16(etu, otu, ecu, ocu), bu = b2_interp16(otu, mtu, ocu, mcu);
  const B2Quad16 av = b2_interp16(etv, otv, ecv, ocv), bv = b2_interp16(otv, mtv, ocv, mcv);
  const B2Quad16 cq = b2_interp16(b2_prmt(ltu, ltv, 0x5410), b2_prmt(etu, etv, 0x5410),
                                  b2_prmt(lcu, lcv, 0x5410), b2_prmt(ecu, ecv, 0x5410));
  const bool lane_ok = xh + 8 <= wv;
  const uint32_t xc = b2_hi8(cq.ts);
  const uint32_t u2[4] = {b2_hi8(au.tf), b2_hi8(au.ts), b2_hi8(bu.tf), b2_prmt(xc, b2_hi8(bu.ts), 0x5410)};
  const uint32_t v2[4] = {b2_hi8(av.tf), b2_hi8(av.ts), b2_hi8(bv.tf), b2_prmt(xc, b2_hi8(bv.ts), 0x5432)};
  // 16 lanes x 32 bytes = the first 512 bytes of the row's half of the transposition buffer
  const uint32_t sw = (uint32_t)(l >> 2) & 1u;
  const uint32_t xw0 = ((2u * l) ^ sw) * 16u, xw1 = ((2u * l + 1u) ^ sw) * 16u;
  uint8_t* const drow = bot ? drb : drt;
  emit_row8_packed<ST, HAS_ALPHA>(yw, aw, u2, v2, drow + xh * kBpp, lane_ok, kc,
                                  (kBpp == 4) ? xs + (bot ? 1024 : 0) : nullptr, xw0, xw1, sel16);
  if (kBpp == 4) {
    __syncwarp();
    const int xa = xpass + lane * 4;   // every lane stores pixels xa .. xa+3 of BOTH rows: 512 contiguous bytes per store
    const int n = wv - xa;
    const uint4 vt = *reinterpret_cast<const uint4*>(xs + xl.r);
    const uint4 vb = *reinterpret_cast<const uint4*>(xs + 1024 + xl.r);
    if (n >= 4) {
      b2_stg_stream_v4(drt + xa * 4, vt);
      b2_stg_stream_v4(drb + xa * 4, vb);
    } else {
      if (n > 0) { b2_stg_stream_u32(drt + xa * 4, vt.x); b2_stg_stream_u32(drb + xa * 4, vb.x); }
      if (n > 1) { b2_stg_stream_u32(drt + xa * 4 + 4, vt.y); b2_stg_stream_u32(drb + xa * 4 + 4, vb.y); }
      if (n > 2) { b2_stg_stream_u32(drt + xa * 4 + 8, vt.z); b2_stg_stream_u32(drb + xa * 4 + 8, vb.z); }
    }
    __syncwarp();
  }
}

template <int ST, bool HAS_ALPHA>
__global__ void __launch_bounds__(kBulkThreads, kBulkCtasPerSm)
upsample_bulk_kernel(const UpsampleJobDev* __restrict__ jobs, const uint32_t* __restrict__ unit_prefix,
                     uint32_t n_jobs, uint32_t uniform_units, uint32_t total_units, const B2YuvConsts kc,
                     uint32_t* __restrict__ tickets) {
  using BL = BulkLayout<HAS_ALPHA>;
  constexpr int kTileW = BL::kTW;
  __shared__ __align__(128) uint8_t ring[kBulkWarps * 2 * BL::kBytes];
  __shared__ __align__(16) uint8_t xpose[kBulkWarps * 2048];
  __shared__ __align__(8) uint64_t bars[kBulkWarps * 2];
  const int lane = threadIdx.x & 31;
  const uint32_t warp = uni(threadIdx.x >> 5);
  uint8_t* wring = ring + warp * (2 * BL::kBytes);
  const uint32_t wring_s = (uint32_t)__cvta_generic_to_shared(wring);
  const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(bars + 2 * warp);
  if (lane == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_s + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const uint32_t first_unit = blockIdx.x * kBulkWarps + warp;
  const uint32_t unit_stride = gridDim.x * kBulkWarps;
  total_units = uni(total_units);

  // Row pairs are handed out by a ticket counter (tickets[0]; tickets[1] counts the warps that are done,
  // the last one clears both for the next launch on the stream): a warp starts on row pair first_unit and
  // draws the following ones, so in a batch of mixed sizes no warp is left with more wide rows than the
  // others.  The ticket after the producer's row pair is always drawn one row pair early (ahead, lane 0).
  // The consumer walks the same sequence at most two row pairs behind the producer (c_next1, c_next2).
  uint32_t ahead = 0;
  if (lane == 0) ahead = unit_stride + atomicAdd(tickets, 1u);
  uint32_t c_next1 = 0, c_next2 = 0;
  int n_pending = 0;

  // ---- producer: lane 0 issues every copy of a tile (one mbarrier arrival) ----
  uint32_t p_unit = first_unit;
  int p_t = 0, p_tiles = 0, p_ylen = 0, p_clen = 0;
  const uint8_t* p_y = nullptr; const uint8_t* p_u = nullptr; const uint8_t* p_v = nullptr; const uint8_t* p_a = nullptr;
  int p_ydelta = 0, p_cdelta = 0, p_adelta = 0;  // byte distance from the first to the second row of a pair
  auto producer_setup = [&]() {
    if (p_unit >= total_units) return;
    uint32_t q;
    const uint32_t img = b2_find_image_warp(unit_prefix, n_jobs, uniform_units, p_unit, &q, (uint32_t)lane);
    const UpsampleJobDev* __restrict__ jp = jobs + img;
    q += jp->q_lo;
    const int width = jp->width, height = jp->height;
    const int uh = (height + 1) >> 1;
    const int rt = 2 * (int)q - 1, rb = 2 * (int)q;
    const int ry_t = (rt >= 0) ? rt : rb, ry_b = (rb < height) ? rb : rt;
    const int ct = max((int)q - 1, 0), cc = min((int)q, uh - 1);
    p_tiles = uni((width + kTileW - 1) / kTileW);
    p_t = 0;
    p_y = uni(jp->y + (ptrdiff_t)ry_t * jp->y_stride);
    p_ydelta = uni((ry_b - ry_t) * jp->y_stride);
    p_u = uni(jp->u + (ptrdiff_t)ct * jp->uv_stride);
    p_v = uni(jp->v + (ptrdiff_t)ct * jp->uv_stride);
    p_cdelta = uni((cc - ct) * jp->uv_stride);
    if (HAS_ALPHA) {
      p_a = uni(jp->a + (ptrdiff_t)ry_t * jp->a_stride);
      p_adelta = uni((ry_b - ry_t) * jp->a_stride);
    }
    p_ylen = uni((width + 15) & ~15);
    p_clen = uni((((width + 1) >> 1) + 15) & ~15);
  };
  auto producer_issue = [&](uint32_t stage) {
    if (p_unit >= total_units) return;
    if (lane == 0) {
      const uint32_t bar = bar_s + 8 * stage;
      const uint32_t sb = wring_s + stage * BL::kBytes;
      const int ys0 = p_t * kTileW;
      const uint32_t yb = (uint32_t)min(kTileW, p_ylen - ys0);
      const int g0 = p_t * (kTileW / 2) - 16;     // first chroma byte of the stage row (halo included)
      const int cs = max(g0, 0);
      const uint32_t cb = (uint32_t)(min(g0 + BL::kCRow, p_clen) - cs);
      const uint32_t cd = sb + BL::kC + (uint32_t)(cs - g0);
      mbar_expect_tx(bar, (HAS_ALPHA ? 4u : 2u) * yb + 4u * cb);
      bulk_g2s(sb + BL::kY, p_y + ys0, yb, bar);
      bulk_g2s(sb + BL::kY + kTileW, p_y + p_ydelta + ys0, yb, bar);
      bulk_g2s(cd, p_u + cs, cb, bar);
      bulk_g2s(cd + BL::kCRow, p_u + p_cdelta + cs, cb, bar);
      bulk_g2s(cd + 2 * BL::kCRow, p_v + cs, cb, bar);
      bulk_g2s(cd + 3 * BL::kCRow, p_v + p_cdelta + cs, cb, bar);
      if (HAS_ALPHA) {
        bulk_g2s(sb + BL::kA, p_a + ys0, yb, bar);
        bulk_g2s(sb + BL::kA + kTileW, p_a + p_adelta + ys0, yb, bar);
      }
    }
    if (++p_t == p_tiles) {
      p_unit = __shfl_sync(0xffffffffu, ahead, 0);
      if (lane == 0 && p_unit < total_units) ahead = unit_stride + atomicAdd(tickets, 1u);
      if (n_pending == 0) c_next1 = p_unit; else c_next2 = p_unit;
      ++n_pending;
      producer_setup();
    }
  };

  // ---- consumer ----
  // kc: the three non-immediate constants of the conversion arrive as kernel
  // parameters, so they are constant-bank operands instead of registers that ptxas
  // would rematerialise with a move before every use.
  const XposeLane xl = xpose_lane(lane);
  uint8_t* const xbuf = xpose + warp * 2048;
  uint32_t c_unit = first_unit;
  producer_setup();
  producer_issue(0);
  uint32_t stage = 0, phase = 0;  // phase bit s = parity to wait for on stage s
  while (c_unit < total_units) {
    uint32_t q;
    const uint32_t img = b2_find_image_warp(unit_prefix, n_jobs, uniform_units, c_unit, &q, (uint32_t)lane);
    const UpsampleJobDev* __restrict__ jp = jobs + img;
    q += jp->q_lo;
    const int width = jp->width;
    const int uw = (width + 1) >> 1;
    const int rt = 2 * (int)q - 1, rb = 2 * (int)q;
    const int row_begin = jp->row_begin, row_end = jp->row_end;
    const bool top_ok = (rt >= row_begin) && (rt < row_end), bot_ok = (rb >= row_begin) && (rb < row_end);
    const int dstride = jp->dst_stride;
    constexpr int kBppK = bulk_bpp<ST>();
    const int wv = (kBppK == 4) ? width : (width & ~7);  // the vector path covers pixels [0, wv)
    const uint32_t sel16 = (jp->flags & JOB_SWAP16) ? 0x3715u : 0x7351u;   // 16-bit modes: which byte of a pixel comes first
    uint8_t* const drt = jp->dst + (ptrdiff_t)(top_ok ? rt : rb) * dstride;
    uint8_t* const drb = jp->dst + (ptrdiff_t)(bot_ok ? rb : rt) * dstride;
    const int tiles = (width + kTileW - 1) / kTileW;
    for (int t = 0; t < tiles; ++t) {
      producer_issue(stage ^ 1);  // the other stage was released by the __syncwarp below
      mbar_wait(bar_s + 8 * stage, (phase >> stage) & 1);
      phase ^= 1u << stage;
      const uint8_t* sp = wring + stage * BL::kBytes;
#pragma unroll kPassUnroll
      for (int p = 0; p < kTileW / 256; ++p) {
        const int xpass = t * kTileW + p * 256;
        if (xpass >= wv) break;
        const uint8_t* sy = sp + BL::kY + p * 256 + lane * 8;
        const uint8_t* sa = sp + BL::kA + p * 256 + lane * 8;
        const uint8_t* sc = sp + BL::kC + 16 + p * 128 + lane * 4;
        if (top_ok && bot_ok && wv - xpass <= 128)   // warp-uniform: a last pass of at most 128 pixels, both rows wanted
          convert_pass_split<ST, HAS_ALPHA>(sp + BL::kY + p * 256, sp + BL::kA + p * 256, sp + BL::kC + 16 + p * 128, lane,
                                            xpass, wv, uw, drt, drb, kc, xbuf, xl, sel16);
        else if (xpass == 0 || xpass + 256 >= 2 * uw || xpass + 256 > wv)  // warp-uniform: first / last pass of the row
          convert_pass<ST, HAS_ALPHA, true>(sy, sa, sc, lane, xpass, wv, uw, top_ok, bot_ok, drt, drb, kc, xbuf, xl, sel16);
        else
          convert_pass<ST, HAS_ALPHA, false>(sy, sa, sc, lane, xpass, wv, uw, top_ok, bot_ok, drt, drb, kc, xbuf, xl, sel16);
      }
      __syncwarp();  // every lane is done reading this stage before it is refilled
      stage ^= 1;
    }
    if (kBppK != 4 && wv + lane < width) {  // 3- and 2-byte modes: ragged right edge (width % 8 pixels), per pixel from global
      const UpsampleJobDev j = *jp;
      if (top_ok) pixel_generic<ST>(j, rt, wv + lane);
      if (bot_ok) pixel_generic<ST>(j, rb, wv + lane);
    }
    c_unit = c_next1;   // the producer moved on when it issued this row pair's last tile
    c_next1 = c_next2;
    --n_pending;
  }
  if (lane == 0) {
    if (atomicAdd(tickets + 1, 1u) == unit_stride - 1) { tickets[0] = 0; tickets[1] = 0; }
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
  const size_t pb = (sizeof(uint32_t) * ((size_t)n_jobs + 1) + 15) & ~(size_t)15;
  const size_t pre_off = (jb + 15) & ~(size_t)15;
  if (pre_off + 2 * pb > B2_DESC_RING_BYTES) return B2_ERR_INVALID_CONFIG;
  uint8_t* host = (uint8_t*)malloc(pre_off + 2 * pb);
  if (host == NULL) return B2_ERR_OOM;
  UpsampleJobDev* dj = (UpsampleJobDev*)host;
  uint32_t* pre = (uint32_t*)(host + pre_off);
  uint32_t* upre = (uint32_t*)(host + pre_off + pb);  // prefix in row pairs (tuned kernel)
  uint64_t total = 0, utotal = 0;
  uint32_t uniform = 0, uuniform = 0;
  bool tuned = (st_mode <= ST_565);  // every mode has a staged vector kernel (4-, 3- and 2-byte pixels)
  bool staged = true;  // cp.async kernel: 16-byte aligned planes with positive strides
  int n_with_a = 0;
  int max_width = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const b2_upsample_job& j = jobs[i];
    if (j.width > max_width) max_width = j.width;
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
    d.flags = (fast ? JOB_FAST : 0) | (ctx->swap16 ? JOB_SWAP16 : 0);
    if (!fast && bpp == 4) tuned = false;
    if (bpp != 4 && ((((uintptr_t)j.dst | (uintptr_t)(uint32_t)j.dst_stride) & 15) != 0)) tuned = false;
    if ((((uintptr_t)j.y | (uintptr_t)j.u | (uintptr_t)j.v | (uintptr_t)j.a | (uintptr_t)(uint32_t)j.y_stride |
          (uintptr_t)(uint32_t)j.uv_stride | (uintptr_t)(uint32_t)(j.a ? j.a_stride : 0)) & 15) != 0 ||
        j.y_stride < 0 || j.uv_stride < 0 || (j.a && j.a_stride < 0)) staged = false;
    n_with_a += (j.a != NULL);
    // EmitFancyRGB-style partial output (io_dec.c.go:65-115): rows [row_begin, row_end)
    int rb0 = 0, re0 = j.height;
    if (j.row_end != 0 || j.row_begin != 0) {
      if (j.row_begin < 0 || j.row_end <= j.row_begin || j.row_end > j.height) { free(host); return B2_ERR_INVALID_CONFIG; }
      rb0 = j.row_begin; re0 = j.row_end;
    }
    d.row_begin = rb0; d.row_end = re0;
    d.q_lo = (uint32_t)((rb0 + 1) >> 1);
    d.pad = 0;
    const uint32_t units = (uint32_t)(re0 >> 1) - d.q_lo + 1;
    const uint32_t items = d.tiles_x * units;
    if (i == 0) uniform = items; else if (items != uniform) uniform = 0;
    if (i == 0) uuniform = units; else if (units != uuniform) uuniform = 0;
    pre[i] = (uint32_t)total;
    upre[i] = (uint32_t)utotal;
    total += items;
    utotal += units;
    if (total > 0x7fffffffull) { free(host); return B2_ERR_BAD_DIM; }
  }
  pre[n_jobs] = (uint32_t)total;
  upre[n_jobs] = (uint32_t)utotal;
  if (n_with_a != 0 && n_with_a != n_jobs && (bpp == 4 || st_mode == ST_4444)) tuned = false;  // HAS_ALPHA is a kernel-wide switch
  void* dev = NULL;
  int st = b2_upload_desc(ctx, stream, host, pre_off + 2 * pb, &dev);
  free(host);
  if (st != B2_OK) return st;
  const UpsampleJobDev* djobs = (const UpsampleJobDev*)dev;
  const uint32_t* dpre = (const uint32_t*)((uint8_t*)dev + pre_off);
  const uint32_t* dupre = (const uint32_t*)((uint8_t*)dev + pre_off + pb);
  cudaStream_t s = ctx->streams[stream];
  const int grid = b2_grid_for(ctx, 8, (total + kWarps - 1) / kWarps);
  const uint32_t tot = (uint32_t)total;
  if (tuned && staged) {
    // Persistent CTAs per SM (of the 6 that fit).  Measured (profiles/r01d_grid_sweeps.md):
    // 4 per SM is best for 4K rows (0.85 of the measured HBM peak vs 0.82 at 6), 5 per SM for
    // 1080p rows (0.785 vs 0.755); the address window the warps sweep together matters
    // more than the extra latency hiding.
    static const int grid_override = getenv("B2_UP_GRID") ? atoi(getenv("B2_UP_GRID")) : 0;
    const int bulk_grid_ctas = grid_override ? grid_override : (max_width >= 2560 ? 4 : 5);
    const int bgrid = b2_grid_for(ctx, bulk_grid_ctas, (utotal + kBulkWarps - 1) / kBulkWarps);
    const uint32_t ut = (uint32_t)utotal;
    const bool ha = n_with_a != 0;
#define B2_UPK(ST_, HA_) upsample_bulk_kernel<ST_, HA_><<<bgrid, kBulkThreads, 0, s>>>(djobs, dupre, n_jobs, uuniform, ut, b2_yuv_consts(), b2_tickets(ctx, stream, 1))
    switch (st_mode) {
      case ST_RGBA: if (ha) B2_UPK(ST_RGBA, true); else B2_UPK(ST_RGBA, false); break;
      case ST_BGRA: if (ha) B2_UPK(ST_BGRA, true); else B2_UPK(ST_BGRA, false); break;
      case ST_ARGB: if (ha) B2_UPK(ST_ARGB, true); else B2_UPK(ST_ARGB, false); break;
      case ST_RGB: B2_UPK(ST_RGB, false); break;
      case ST_BGR: B2_UPK(ST_BGR, false); break;
      case ST_4444: if (ha) B2_UPK(ST_4444, true); else B2_UPK(ST_4444, false); break;   // the alpha nibble merged (EmitAlphaRGBA4444)
      default: B2_UPK(ST_565, false); break;
    }
#undef B2_UPK
    ctx->launches++;
    B2_CUDA(ctx, cudaGetLastError());
    return B2_OK;
  }
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
