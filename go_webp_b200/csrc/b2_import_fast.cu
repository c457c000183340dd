// b2_import_fast.cu -- tuned RGBA -> YUV420(+A) import for the common case: 4-byte
// pixels, 4-byte aligned source rows (128-bit loads when they are 16-byte aligned),
// any width, 4-byte aligned Y/A rows.
//
// Same arithmetic as b2_import.cu (pkg/libwebp/dsp/yuv.go:124-382,
// pkg/color/yuv/conversion.go:51-69); what changes is the instruction budget.  The
// first version spent ~410 warp instructions per 4x2-pixel strip and was issue
// bound at 48% of HBM peak (profiles/r01_*).  Here:
//   * a warp walks one whole row pair (the reference's own unit of work,
//     yuv.go:316), so image lookup, descriptor loads and 64-bit address math are
//     paid once per row pair, and the next strip is prefetched while the current
//     one is computed;
//   * luma is two dp2a per pixel (16-bit coefficients x 8-bit channels);
//   * the gamma tables sit in shared memory with a 256-byte row per entry and one
//     word per lane, so a lookup address is ONE byte-permute, (value << 8) | lane*4,
//     and every lookup is bank-conflict free.  Linear values are stored << 7 so the
//     interpolation index of LinearToGamma is a whole byte of the sum;
//   * U and V are scaled by 64 so the result is the top byte of the accumulator;
//   * the translucent-block division 4*sum/a (yuv.go:231) is an exact
//     multiply-high by a per-divisor magic number (verified for every divisor in
//     tests/test_oracle_properties.py) instead of a hardware integer division.
#include "b2_import.cuh"

namespace {

#ifndef B2_IMPORT_L2_PREFETCH
#define B2_IMPORT_L2_PREFETCH 0
#endif
constexpr int kFastThreads = 512;
constexpr int kFastWarps = kFastThreads / 32;
constexpr int kFastCtasPerSm = 2;

// shared-memory map (byte offsets)
//   [0, 64K)    g2l rows: row v = 256 bytes, lower half = g2l[v] << 7, one word per lane
//   [64K, 72K)  LinearToGamma: entry pos = 256 bytes, lane l at +8 l: {base' << 7, slope d}
//   [72K, 80K)  divide table, entry a at +8 a: exact mode {magic << 3, shift}, binary-
//               compatible mode {kInvAlpha[a], 0}
constexpr uint32_t kOffL2G = 65536;
constexpr uint32_t kOffDiv = kOffL2G + 32 * 256;
constexpr int kSmemBytes = kOffDiv + 1024 * 8;  // 80 KB

template <int ORDER> struct Ch;
template <> struct Ch<0> {  // R,G,B,A bytes
  static constexpr uint32_t kLoPair = 16839u | (33059u << 16);  // bytes 0,1
  static constexpr uint32_t kHiPair = 6420u;                    // bytes 2,3
  static constexpr int kR = 0, kB = 2;
};
template <> struct Ch<1> {  // B,G,R,A bytes
  static constexpr uint32_t kLoPair = 6420u | (33059u << 16);
  static constexpr uint32_t kHiPair = 16839u;
  static constexpr int kR = 2, kB = 0;
};

__device__ __forceinline__ uint32_t lds32(const uint8_t* smem, uint32_t off) {
  return *reinterpret_cast<const uint32_t*>(smem + off);
}
// RGBToY (conversion.go:56-59) as two 16x8-bit dot products; Y = byte 2 of the sum
template <int ORDER>
__device__ __forceinline__ uint32_t luma_sum(uint32_t p) {
  return __dp2a_hi(Ch<ORDER>::kHiPair, p, __dp2a_lo(Ch<ORDER>::kLoPair, p, (1u << 15) + (16u << 16)));
}
__device__ __forceinline__ uint32_t pack_b2(uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3) {
  return __byte_perm(__byte_perm(s0, s1, 0x0062), __byte_perm(s2, s3, 0x0062), 0x5410);
}
__device__ __forceinline__ uint32_t pack_b3(uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3) {
  return __byte_perm(__byte_perm(s0, s1, 0x0073), __byte_perm(s2, s3, 0x0073), 0x5410);
}
// g2l[channel K of p] << 7 for this lane: address = (byte << 8) | lane*4
template <int K>
__device__ __forceinline__ uint32_t g2l(const uint8_t* smem, uint32_t p, uint32_t lane4) {
  return lds32(smem, __byte_perm(p, lane4, 0x5504 | (K << 4)));
}
// LinearToGamma (yuv.go:187-202) on a sum held << 7: index = byte 2 of the sum, base
// and slope come in one 64-bit load.  Returns the 10-bit gamma-domain value.
__device__ __forceinline__ int l2g(const uint8_t* smem, uint32_t s7, uint32_t lane8) {
  const uint32_t off = __byte_perm(s7, lane8, 0x5524);
  const int2 bd = *reinterpret_cast<const int2*>(smem + kOffL2G + off);
  return (bd.x + bd.y * (int)s7) >> 14;
}
// RGBToU / RGBToV on gamma-domain sums (conversion.go:51-69, rounding YUV_HALF<<2),
// scaled by 64: the result is byte 3.  Inputs are <= 1020 so ClipUV never clamps.
__device__ __forceinline__ uint32_t u_sum(int r, int g, int b) {
  constexpr uint32_t kRnd = ((1u << 17) + (128u << 18)) << 6;
  return (uint32_t)(-9719 * 64) * (uint32_t)r + (uint32_t)(-19081 * 64) * (uint32_t)g +
         (uint32_t)(28800 * 64) * (uint32_t)b + kRnd;
}
__device__ __forceinline__ uint32_t v_sum(int r, int g, int b) {
  constexpr uint32_t kRnd = ((1u << 17) + (128u << 18)) << 6;
  return (uint32_t)(28800 * 64) * (uint32_t)r + (uint32_t)(-24116 * 64) * (uint32_t)g +
         (uint32_t)(-4684 * 64) * (uint32_t)b + kRnd;
}

// DIVIDE_BY_ALPHA (yuv.go:231) on w7 = (sum of alpha_i * linear_i) << 7; returns the
// quotient << 7 ready for l2g().  dv = this divisor's table entry.
template <bool TABLE>
__device__ __forceinline__ uint32_t div_alpha7(uint32_t w7, uint2 dv) {
  if (TABLE) return (((w7 >> 7) * dv.x) >> 17) << 7;   // libwebp-binary kInvAlpha form
  return (__umulhi(w7, dv.x) >> dv.y) << 7;            // exact 4*sum/a: magic multiply-high
}

// Linear sums (<< 7) of one 2x2 block for the three channels.
// Exact-divide mode is branch-free across the warp: when any lane holds a translucent
// block every lane takes the weighted form, which is exact for the other blocks too
// (four alphas of 255: 4*255*S/1020 == S; an all-transparent block is given equal
// weights, any equal weights reproduce SUM4, yuv.go:249-252).  The binary-compatible
// table divide is not exact for opaque blocks, so that mode branches per lane.
template <int ORDER, bool ALPHA, bool TABLE>
__device__ __forceinline__ void block_sums(const uint8_t* smem, uint32_t lane4, uint32_t pa, uint32_t pb,
                                           uint32_t pc, uint32_t pd, uint32_t al4,
                                           uint32_t* sr, uint32_t* sg, uint32_t* sb) {
  const uint32_t ra = g2l<Ch<ORDER>::kR>(smem, pa, lane4), rb = g2l<Ch<ORDER>::kR>(smem, pb, lane4);
  const uint32_t rc = g2l<Ch<ORDER>::kR>(smem, pc, lane4), rd = g2l<Ch<ORDER>::kR>(smem, pd, lane4);
  const uint32_t ga = g2l<1>(smem, pa, lane4), gb = g2l<1>(smem, pb, lane4);
  const uint32_t gc = g2l<1>(smem, pc, lane4), gd = g2l<1>(smem, pd, lane4);
  const uint32_t ba = g2l<Ch<ORDER>::kB>(smem, pa, lane4), bb = g2l<Ch<ORDER>::kB>(smem, pb, lane4);
  const uint32_t bc = g2l<Ch<ORDER>::kB>(smem, pc, lane4), bd = g2l<Ch<ORDER>::kB>(smem, pd, lane4);
  const bool plain = !ALPHA || al4 == 0xffffffffu || al4 == 0u;  // yuv.go:249 / :270
  if (!ALPHA || (TABLE ? plain : __all_sync(0xffffffffu, plain))) {
    *sr = ra + rb + rc + rd;
    *sg = ga + gb + gc + gd;
    *sb = ba + bb + bc + bd;
  } else {                                           // yuv.go:233-241,254-256
    const uint32_t al = (!TABLE && al4 == 0u) ? 0xffffffffu : al4;
    const uint32_t a0 = al & 0xff, a1 = (al >> 8) & 0xff, a2 = (al >> 16) & 0xff, a3 = al >> 24;
    const uint32_t asum = __dp4a(al, 0x01010101u, 0u);
    const uint2 dv = *reinterpret_cast<const uint2*>(smem + kOffDiv + asum * 8u);
    *sr = div_alpha7<TABLE>(a0 * ra + a1 * rb + a2 * rc + a3 * rd, dv);
    *sg = div_alpha7<TABLE>(a0 * ga + a1 * gb + a2 * gc + a3 * gd, dv);
    *sb = div_alpha7<TABLE>(a0 * ba + a1 * bb + a2 * bc + a3 * bd, dv);
  }
}

template <int ORDER, bool ALPHA, bool WRITE_A, bool TABLE>
__global__ void __launch_bounds__(kFastThreads, kFastCtasPerSm)
import_fast_kernel(const ImportJobDev* __restrict__ jobs, const uint32_t* __restrict__ unit_prefix,
                   uint32_t n_jobs, uint32_t uniform_units, uint32_t total_units,
                   const uint32_t* __restrict__ tab, int32_t* __restrict__ has_alpha) {
  extern __shared__ __align__(256) uint8_t smem[];
  for (int i = threadIdx.x; i < 256 * 32; i += kFastThreads) {
    *reinterpret_cast<uint32_t*>(smem + (i >> 5) * 256 + (i & 31) * 4) = __ldg(tab + B2_TAB_G2L + (i >> 5)) << 7;
  }
  for (int i = threadIdx.x; i < 32 * 32; i += kFastThreads) {
    const int pos = i >> 5;
    const int base = (int)__ldg(tab + B2_TAB_L2G + 2 * pos) * 128;  // base' << 7
    const int d = (int)__ldg(tab + B2_TAB_L2G + 2 * pos + 1);
    *reinterpret_cast<int2*>(smem + kOffL2G + pos * 256 + (i & 31) * 8) = make_int2(base, d);
  }
  for (int i = threadIdx.x; i < 1024; i += kFastThreads) {
    uint2 e;
    if (TABLE) {
      e = make_uint2(__ldg(tab + B2_TAB_INVA + i), 0u);
    } else {
      const uint32_t m = __ldg(tab + B2_TAB_MAGIC + i);  // M | L << 27, M < 2^26: umulhi(w << 3, M) == umulhi(w, M << 3)
      e = make_uint2((m & 0x07ffffffu) << 3, m >> 27);
    }
    *reinterpret_cast<uint2*>(smem + kOffDiv + i * 8) = e;
  }
  __syncthreads();

  const uint32_t lane = threadIdx.x & 31;
  const uint32_t lane4 = lane * 4;
  const uint32_t lane8 = lane * 8;
  const uint32_t warp = threadIdx.x >> 5;

  for (uint32_t unit = blockIdx.x * kFastWarps + warp; unit < total_units; unit += gridDim.x * kFastWarps) {
    uint32_t rp;
    const uint32_t img = b2_find_image(unit_prefix, n_jobs, uniform_units, unit, &rp);
    const ImportJobDev* __restrict__ jp = jobs + img;
    const int width = jp->width;
    const int y0 = 2 * (int)rp;
    const bool has_row1 = (y0 + 1 < jp->height);
    const int sstride = jp->src_stride;
    const int ystride = jp->y_stride;
    const uint8_t* s0 = jp->src + (ptrdiff_t)y0 * sstride + lane * 16;
    const uint8_t* s1 = has_row1 ? s0 + sstride : s0;  // odd last row duplicated (yuv.go:377)
    uint8_t* yp = jp->y + (ptrdiff_t)y0 * ystride + lane4;
    uint8_t* ap = nullptr;
    int astride = 0;
    if (WRITE_A) { astride = jp->a_stride; ap = jp->a + (ptrdiff_t)y0 * astride + lane4; }
    const ptrdiff_t uvo = (ptrdiff_t)rp * jp->uv_stride + lane * 2;
    uint8_t* up = jp->u + uvo;
    uint8_t* vp = jp->v + uvo;
    uint32_t and_acc = 0xffffffffu;

    // One 128-pixel strip per iteration, the next strip's loads in flight meanwhile.
    // Lanes past the end of the row run on zeros (every lane takes part in the vote of
    // block_sums) and store nothing.
    int x = (int)lane4;
    // rows that are only 4-byte aligned (a tight 4*width stride with width % 4 != 0) are
    // read with 32-bit loads of the valid pixels; 16-byte aligned rows with one 128-bit load
    const bool al16 = ((((uintptr_t)jp->src) | (uintptr_t)(uint32_t)sstride) & 15) == 0;  // warp-uniform
    auto load4 = [&](const uint8_t* p, int xx) -> uint4 {
      if (al16) return b2_ldg_stream_v4(p);
      const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
      uint4 r;
      r.x = __ldg(q);
      r.y = (xx + 1 < width) ? __ldg(q + 1) : 0u;
      r.z = (xx + 2 < width) ? __ldg(q + 2) : 0u;
      r.w = (xx + 3 < width) ? __ldg(q + 3) : 0u;
      return r;
    };
    uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0;
    bool cur_ok = x < width;
    if (cur_ok) { c0 = load4(s0, x); c1 = load4(s1, x); }
    const int tiles_x = (width + 127) >> 7;
    for (int t = 0; t < tiles_x; ++t) {
      const int xn = x + 128;
      const bool nxt_ok = xn < width;
      uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0;
      if (nxt_ok) { n0 = load4(s0 + 512, xn); n1 = load4(s1 + 512, xn); }
#if B2_IMPORT_L2_PREFETCH
      // Optional: pull the strips after that into L2.  Measured (profiles/r01d_*): with
      // opaque input the kernel waits on load latency and distance 5 gains 3%, but with
      // translucent input it is issue bound and the two extra instructions cost 6%, so
      // the default is off.
      if (xn + 128 * (B2_IMPORT_L2_PREFETCH - 1) < width) {
        asm volatile("prefetch.global.L2 [%0];" :: "l"(s0 + 512 * B2_IMPORT_L2_PREFETCH));
        asm volatile("prefetch.global.L2 [%0];" :: "l"(s1 + 512 * B2_IMPORT_L2_PREFETCH));
      }
#endif
      // Last group of a row whose width is not a multiple of 4: a 128-bit load stays inside
      // the row's 16-byte-aligned stride span, the 32-bit loads fetch valid pixels only;
      // the missing pixels become copies of valid ones.  Duplicating the odd last column is exactly the reference's SUM2 / SUM2ALPHA
      // forms (yuv.go:211-213,270-277: twice a 2-pixel sum == the 4-pixel sum of the
      // duplicated column), and copies cannot disturb the has-alpha reduction.
      const int nvalid = width - x;
      if (nvalid < 4 && nvalid > 0) {
        if (nvalid == 1) { c0.y = c0.x; c1.y = c1.x; }
        if (nvalid <= 2) { c0.z = c0.x; c1.z = c1.x; c0.w = c0.y; c1.w = c1.y; }
        else { c0.w = c0.z; c1.w = c1.z; }
      }
      const uint32_t yw0 = pack_b2(luma_sum<ORDER>(c0.x), luma_sum<ORDER>(c0.y), luma_sum<ORDER>(c0.z), luma_sum<ORDER>(c0.w));
      const uint32_t yw1 = pack_b2(luma_sum<ORDER>(c1.x), luma_sum<ORDER>(c1.y), luma_sum<ORDER>(c1.z), luma_sum<ORDER>(c1.w));
      uint32_t ap0 = 0xffffffffu, ap1 = 0xffffffffu;
      if (ALPHA) {
        ap0 = pack_b3(c0.x, c0.y, c0.z, c0.w);
        ap1 = pack_b3(c1.x, c1.y, c1.z, c1.w);
        if (cur_ok) and_acc &= ap0 & ap1;
      }
      uint32_t sr0, sg0, sb0, sr1, sg1, sb1;
      block_sums<ORDER, ALPHA, TABLE>(smem, lane4, c0.x, c0.y, c1.x, c1.y, __byte_perm(ap0, ap1, 0x5410), &sr0, &sg0, &sb0);
      block_sums<ORDER, ALPHA, TABLE>(smem, lane4, c0.z, c0.w, c1.z, c1.w, __byte_perm(ap0, ap1, 0x7632), &sr1, &sg1, &sb1);
      const int r0 = l2g(smem, sr0, lane8), g0 = l2g(smem, sg0, lane8), b0 = l2g(smem, sb0, lane8);
      const int r1 = l2g(smem, sr1, lane8), g1 = l2g(smem, sg1, lane8), b1 = l2g(smem, sb1, lane8);
      const uint32_t uu = __byte_perm(u_sum(r0, g0, b0), u_sum(r1, g1, b1), 0x0073);
      const uint32_t vv = __byte_perm(v_sum(r0, g0, b0), v_sum(r1, g1, b1), 0x0073);
      if (nvalid >= 4) {
        b2_stg_stream_u32(yp, yw0);
        if (has_row1) b2_stg_stream_u32(yp + ystride, yw1);
        if (WRITE_A) {
          b2_stg_stream_u32(ap, ap0);
          if (has_row1) b2_stg_stream_u32(ap + astride, ap1);
        }
        b2_stg_stream_u16(up, uu);
        b2_stg_stream_u16(vp, vv);
      } else if (nvalid > 0) {  // partial group: byte stores of the valid samples only
        for (int k = 0; k < nvalid; ++k) {
          yp[k] = (uint8_t)(yw0 >> (8 * k));
          if (has_row1) yp[ystride + k] = (uint8_t)(yw1 >> (8 * k));
          if (WRITE_A) {
            ap[k] = (uint8_t)(ap0 >> (8 * k));
            if (has_row1) ap[astride + k] = (uint8_t)(ap1 >> (8 * k));
          }
        }
        up[0] = (uint8_t)uu; vp[0] = (uint8_t)vv;
        if (nvalid > 2) { up[1] = (uint8_t)(uu >> 8); vp[1] = (uint8_t)(vv >> 8); }
      }
      s0 += 512; s1 += 512; yp += 128; up += 64; vp += 64;
      if (WRITE_A) ap += 128;
      c0 = n0; c1 = n1; cur_ok = nxt_ok; x = xn;
    }
    if (ALPHA && has_alpha != nullptr) {
      if (__any_sync(0xffffffffu, and_acc != 0xffffffffu) && lane == 0) atomicOr(has_alpha + img, 1);
    }
  }
}

template <int ORDER, bool ALPHA, bool WRITE_A, bool TABLE>
cudaError_t launch1(int grid, cudaStream_t s, const ImportJobDev* jobs, const uint32_t* unit_prefix,
                    uint32_t n_jobs, uint32_t uniform_units, uint32_t total_units, const uint32_t* tab,
                    int32_t* has_alpha) {
  auto k = import_fast_kernel<ORDER, ALPHA, WRITE_A, TABLE>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e != cudaSuccess) return e;
  k<<<grid, kFastThreads, kSmemBytes, s>>>(jobs, unit_prefix, n_jobs, uniform_units, total_units, tab, has_alpha);
  return cudaGetLastError();
}
template <int ORDER, bool ALPHA, bool WRITE_A>
cudaError_t launch(int grid, cudaStream_t s, const ImportJobDev* jobs, const uint32_t* unit_prefix,
                   uint32_t n_jobs, uint32_t uniform_units, uint32_t total_units, const uint32_t* tab,
                   int div_table, int32_t* has_alpha) {
  // the divide mode only matters when alpha is imported
  if (ALPHA && div_table) return launch1<ORDER, ALPHA, WRITE_A, true>(grid, s, jobs, unit_prefix, n_jobs, uniform_units, total_units, tab, has_alpha);
  return launch1<ORDER, ALPHA, WRITE_A, false>(grid, s, jobs, unit_prefix, n_jobs, uniform_units, total_units, tab, has_alpha);
}

}  // namespace

cudaError_t b2_import_fast_launch(const b2_ctx* ctx, cudaStream_t s, const ImportJobDev* jobs,
                                  const uint32_t* unit_prefix, uint32_t n_jobs,
                                  uint32_t uniform_units, uint32_t total_units, int order,
                                  bool alpha, bool write_a, int div_table, int32_t* has_alpha) {
  const int grid = b2_grid_for(ctx, kFastCtasPerSm, (total_units + kFastWarps - 1) / kFastWarps);
#define B2_GO(O, A, W) return launch<O, A, W>(grid, s, jobs, unit_prefix, n_jobs, uniform_units, total_units, ctx->tab_dev, div_table, has_alpha)
  if (order == B2_ORDER_RGBA) {
    if (alpha) { if (write_a) B2_GO(0, true, true); else B2_GO(0, true, false); }
    else       { if (write_a) B2_GO(0, false, true); else B2_GO(0, false, false); }
  } else {
    if (alpha) { if (write_a) B2_GO(1, true, true); else B2_GO(1, true, false); }
    else       { if (write_a) B2_GO(1, false, true); else B2_GO(1, false, false); }
  }
#undef B2_GO
}
