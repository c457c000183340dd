// b2_import_fast.cu -- tuned RGBA -> YUV420(+A) import: 4-byte pixels on 4-byte aligned
// source rows (UNAL instantiations when a row is not 16-byte aligned) and 3-byte pixels on
// rows of any alignment (STEP3), any width, 4-byte aligned Y/A rows.  Row pairs are handed
// out by a ticket counter (b2_ctx::tickets_dev).
//
// Same arithmetic as b2_import.cu (pkg/libwebp/dsp/yuv.go:124-382,
// pkg/color/yuv/conversion.go:51-69); what changes is the instruction budget.  The
// first version spent ~410 warp instructions per 4x2-pixel strip and was issue
// bound at 48% of HBM peak (profiles/r01a_ncu_full_first_kernels.csv).  Here:
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

#include <type_traits>

namespace {

// One CTA of 1024 threads per SM: the tables are held once, every warp owns a ring of
// kRingStages source strips in shared memory that per-thread cp.async keeps kRingStages-1
// strips ahead of the computation.  (Measured alternatives, profiles/r01d_* and r01e_*:
// one strip ahead in registers with 2 x 512 threads; L2 prefetch instructions.)
#ifndef B2_IMPORT_RING
#define B2_IMPORT_RING 2
#endif
#ifndef B2_IMPORT_THREADS
#define B2_IMPORT_THREADS 896
#endif
#ifndef B2_IMPORT_UNROLL
#define B2_IMPORT_UNROLL 2
#endif
constexpr int kFastThreads = B2_IMPORT_THREADS;
constexpr int kUnroll = B2_IMPORT_UNROLL;
// kSplit warps share one row pair, taking every kSplit-th strip: fewer concurrent DRAM
// streams for the same number of warps.
#ifndef B2_IMPORT_CHUNKED
#define B2_IMPORT_CHUNKED 0
#endif
#ifndef B2_IMPORT_SPLIT
#define B2_IMPORT_SPLIT 1
#endif
constexpr int kSplit = B2_IMPORT_SPLIT;  // power of two
constexpr int kFastWarps = kFastThreads / 32;
constexpr int kFastCtasPerSm = 1;
constexpr int kRingStages = B2_IMPORT_RING;
constexpr int kDist = kRingStages - 1;              // strips in flight ahead of the computation
// A stage holds one strip of both rows: 32 chunks of 16 bytes (one per lane) + a 33rd chunk, which
// rows that are not 16-byte aligned need (UNAL kernels: their 128 pixels start up to 12 bytes into
// chunk 0 and end as far into chunk 32).  Row 0 at +0, row 1 at +kRow1.
constexpr uint32_t kRow1 = 33 * 16;
constexpr uint32_t kStageBytes = 2 * kRow1;
constexpr uint32_t kRingBytes = kRingStages * kStageBytes;   // per warp
static_assert(kRingStages == 2, "the ring toggles between two stages (next = base + base + stage - current)");

// Shared-memory map, in shared-window addresses (what LDS takes).  T = the first 64 KB
// boundary at or after the start of dynamic shared memory; with the tables on 64 KB
// boundaries the lookup byte lands in address byte 1 with nothing to add, so a lookup
// address -- window base included -- is ONE byte permute:
//   [T, T+64K)          row v = 256 bytes, one word per lane: lower half g2l[v] << 7;
//                       upper half of rows 0..31 LinearToGamma segment v, (base' << 7) + slope
//   [T+64K, T+72K)      divide table, entry a at +8 a: exact mode {magic << 3, shift}, binary-
//                       compatible mode {kInvAlpha[a], 0}
//   rings               as many as fit below T, the rest from T+72K
constexpr uint32_t kOffL2G = 128;
constexpr uint32_t kOffDiv = 65536;
constexpr uint32_t kOffRingHi = kOffDiv + 1024 * 8;
// worst case over the start address: slack below T (up to 64 KB) + tables + the rings that
// do not fit in the slack
constexpr uint32_t kRingWorst = kFastWarps * kRingBytes + 2 * kRingBytes;
constexpr int kSmemBytes = (int)(kOffRingHi + (kRingWorst > 65536u ? kRingWorst : 65536u));
static_assert(kSmemBytes <= 232448, "dynamic shared memory limit");

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

// table reads by shared-window address; the tables are read-only once filled, so these
// may be scheduled freely
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
// RGBToY (conversion.go:56-59) as two 16x8-bit dot products; Y = byte 2 of the sum
// rnd = YUV_HALF, or the VP8RandomBits(rg, YUV_FIX) draw of the dithered import
template <int ORDER>
__device__ __forceinline__ uint32_t luma_sum(uint32_t p, uint32_t rnd = 1u << 15) {
  return __dp2a_hi(Ch<ORDER>::kHiPair, p, __dp2a_lo(Ch<ORDER>::kLoPair, p, rnd + (16u << 16)));
}
__device__ __forceinline__ uint32_t pack_b2(uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3) {
  return __byte_perm(__byte_perm(s0, s1, 0x0062), __byte_perm(s2, s3, 0x0062), 0x5410);
}
__device__ __forceinline__ uint32_t pack_b3(uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3) {
  return __byte_perm(__byte_perm(s0, s1, 0x0073), __byte_perm(s2, s3, 0x0073), 0x5410);
}
// g2l[channel K of p] << 7 for this lane.  gbase = T | lane*4; address = gbase | (byte << 8)
template <int K>
__device__ __forceinline__ uint32_t g2l(uint32_t p, uint32_t gbase) {
  return lds32(__byte_perm(p, gbase, 0x7604 | (K << 4)));
}
// LinearToGamma (yuv.go:187-202) on a sum held << 7: index = byte 2 of the sum.  The
// segment's base (<< 7) and slope (< 128) share one word; the low 7 bits of base << 7
// and of slope * s7 are zero, so the slope rides through the sum without a carry and is
// shifted out.  lbase = (T + kOffL2G) | lane*4.  Returns the 10-bit gamma-domain value.
__device__ __forceinline__ int l2g(uint32_t s7, uint32_t lbase) {
  const uint32_t w = lds32(__byte_perm(s7, lbase, 0x7624));
  return ((int)w + (int)(w & 127u) * (int)s7) >> 14;
}
// RGBToU / RGBToV on gamma-domain sums (conversion.go:51-69, rounding YUV_HALF<<2),
// scaled by 64: the result is byte 3.  Inputs are <= 1020 so ClipUV never clamps.
// rnd = YUV_HALF << 2, or the VP8RandomBits(rg, YUV_FIX + 2) draw (< 2^18: still no clamp, no overflow)
__device__ __forceinline__ uint32_t u_sum(int r, int g, int b, uint32_t rnd = 1u << 17) {
  const uint32_t k = (rnd + (128u << 18)) << 6;
  return (uint32_t)(-9719 * 64) * (uint32_t)r + (uint32_t)(-19081 * 64) * (uint32_t)g +
         (uint32_t)(28800 * 64) * (uint32_t)b + k;
}
__device__ __forceinline__ uint32_t v_sum(int r, int g, int b, uint32_t rnd = 1u << 17) {
  const uint32_t k = (rnd + (128u << 18)) << 6;
  return (uint32_t)(28800 * 64) * (uint32_t)r + (uint32_t)(-24116 * 64) * (uint32_t)g +
         (uint32_t)(-4684 * 64) * (uint32_t)b + k;
}
// VP8RandomBits2 (pkg/vp8/random.go:68-95) on a raw 31-bit draw
template <int NUM_BITS>
__device__ __forceinline__ uint32_t dither_value(uint32_t raw, int amp) {
  int diff;                                          // sign-extend, 0-center: bits 30 .. 31-NUM_BITS of raw
  asm("bfe.s32 %0, %1, %2, %3;" : "=r"(diff) : "r"(raw), "n"(31 - NUM_BITS), "n"(NUM_BITS));
  // restrict range and shift back to 0.5-center: floor(x / 256) + c == floor((x + 256 c) / 256)
  return (uint32_t)((diff * amp + (1 << (NUM_BITS - 1 + 8))) >> 8);
}
constexpr uint32_t kDitherWarpBytes = 3 * 1024;     // three 256-entry rings of raw draws per warp
static_assert(kSmemBytes + kFastWarps * (int)kDitherWarpBytes <= 232448 && kSmemBytes % 16 == 0 && kSplit == 1,
              "dithered import: shared memory / one warp per row pair");

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
__device__ __forceinline__ void block_sums(uint32_t gbase, uint32_t dbase, uint32_t pa, uint32_t pb,
                                           uint32_t pc, uint32_t pd, uint32_t al4,
                                           uint32_t* sr, uint32_t* sg, uint32_t* sb) {
  const uint32_t ra = g2l<Ch<ORDER>::kR>(pa, gbase), rb = g2l<Ch<ORDER>::kR>(pb, gbase);
  const uint32_t rc = g2l<Ch<ORDER>::kR>(pc, gbase), rd = g2l<Ch<ORDER>::kR>(pd, gbase);
  const uint32_t ga = g2l<1>(pa, gbase), gb = g2l<1>(pb, gbase);
  const uint32_t gc = g2l<1>(pc, gbase), gd = g2l<1>(pd, gbase);
  const uint32_t ba = g2l<Ch<ORDER>::kB>(pa, gbase), bb = g2l<Ch<ORDER>::kB>(pb, gbase);
  const uint32_t bc = g2l<Ch<ORDER>::kB>(pc, gbase), bd = g2l<Ch<ORDER>::kB>(pd, gbase);
  const bool plain = !ALPHA || al4 == 0xffffffffu || al4 == 0u;  // yuv.go:249 / :270
  if (!ALPHA || (TABLE ? plain : __all_sync(0xffffffffu, plain))) {
    *sr = ra + rb + rc + rd;
    *sg = ga + gb + gc + gd;
    *sb = ba + bb + bc + bd;
  } else {                                           // yuv.go:233-241,254-256
    const uint32_t al = (!TABLE && al4 == 0u) ? 0xffffffffu : al4;
    const uint32_t a0 = al & 0xff, a1 = (al >> 8) & 0xff, a2 = (al >> 16) & 0xff, a3 = al >> 24;
    const uint32_t asum = __dp4a(al, 0x01010101u, 0u);
    const uint2 dv = lds64(dbase + asum * 8u);
    *sr = div_alpha7<TABLE>(a0 * ra + a1 * rb + a2 * rc + a3 * rd, dv);
    *sg = div_alpha7<TABLE>(a0 * ga + a1 * gb + a2 * gc + a3 * gd, dv);
    *sb = div_alpha7<TABLE>(a0 * ba + a1 * bb + a2 * bc + a3 * bd, dv);
  }
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t nbytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_async16_full(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, uint32_t nbytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(dst), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ uint32_t lds32_ring(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint4 lds128_ring(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// UNAL: source rows that are only 4-byte aligned (a tight 4 * width stride with width % 4 != 0).  A lane
// still fetches whole aligned 16-byte chunks -- chunk 32 t + lane of the row's enclosing aligned span, lane 0
// chunk 32 t + 32 as well -- and takes its four pixels out of its own chunk and the next one (two 128-bit
// shared loads and a register shift by the row's misalignment, warp-uniform) instead of fetching 4 bytes at
// a time: 0.55 -> 0.8 of the measured peak on such images (profiles/r02c_mixed_sizes.md).
// STEP3: 3-byte pixels (WebPPictureImportRGB / BGR, picture_csp_enc.c.go step = 3), rows of any byte alignment.  The
// same enclosing-chunk fetch -- a strip of 128 pixels is 384 bytes = 24 chunks (+ 1 when the row does not start on a
// 16-byte boundary), lanes 0 .. 24 fetch one each -- and a lane takes its 12 bytes from the stage with four aligned
// 32-bit shared loads (lane stride 3 words: conflict-free), three funnel shifts by the row's byte misalignment and
// three byte permutes that spread them into four 4-byte pixels; from there on the code is the 4-byte kernel's (the
// fourth byte of a pixel is never looked at when no alpha is imported).  Was import_generic_kernel: a byte load per
// channel, 0.32 of the measured peak.
template <int ORDER, bool ALPHA, bool WRITE_A, bool TABLE, bool DITHER, bool UNAL, bool STEP3 = false>
__global__ void __launch_bounds__(kFastThreads, kFastCtasPerSm)
import_fast_kernel(const ImportJobDev* __restrict__ jobs, const uint32_t* __restrict__ unit_prefix,
                   uint32_t n_jobs, uint32_t uniform_units, uint32_t total_units,
                   const uint32_t* __restrict__ tab, int32_t* __restrict__ has_alpha, const ImportDither dither,
                   uint32_t* __restrict__ tickets, const uint32_t ticket_batch) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t tbase = (sbase + 0xffffu) & ~0xffffu;       // T
  uint8_t* const tabs = smem + (tbase - sbase);
  for (int i = threadIdx.x; i < 256 * 32; i += kFastThreads) {
    *reinterpret_cast<uint32_t*>(tabs + (i >> 5) * 256 + (i & 31) * 4) = __ldg(tab + B2_TAB_G2L + (i >> 5)) << 7;
  }
  for (int i = threadIdx.x; i < 32 * 32; i += kFastThreads) {
    const int pos = i >> 5;
    const int base = (int)__ldg(tab + B2_TAB_L2G + 2 * pos) * 128;  // base' << 7
    const int d = (int)__ldg(tab + B2_TAB_L2G + 2 * pos + 1);
    *reinterpret_cast<int*>(tabs + kOffL2G + pos * 256 + (i & 31) * 4) = base + d;
  }
  for (int i = threadIdx.x; i < 1024; i += kFastThreads) {
    uint2 e;
    if (TABLE) {
      e = make_uint2(__ldg(tab + B2_TAB_INVA + i), 0u);
    } else {
      const uint32_t m = __ldg(tab + B2_TAB_MAGIC + i);  // M | L << 27, M < 2^26: umulhi(w << 3, M) == umulhi(w, M << 3)
      e = make_uint2((m & 0x07ffffffu) << 3, m >> 27);
    }
    *reinterpret_cast<uint2*>(tabs + kOffDiv + i * 8) = e;
  }
  __syncthreads();

  const uint32_t lane = threadIdx.x & 31;
  const uint32_t lane4 = lane * 4;
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t gbase = tbase | lane4;
  const uint32_t lbase = (tbase + kOffL2G) | lane4;
  const uint32_t dbase = tbase + kOffDiv;
  // this warp's ring: below the tables while there is room, else above them
  const uint32_t ring_lo = (sbase + 15u) & ~15u;
  const uint32_t n_lo = (tbase - ring_lo) / kRingBytes;
  const uint32_t slot0 = (warp < n_lo ? ring_lo + warp * kRingBytes
                                      : tbase + kOffRingHi + (warp - n_lo) * kRingBytes) + lane * 16;
  const uint32_t slot_sum = 2 * slot0 + kStageBytes;   // the other stage of a slot address a: slot_sum - a

  // Row pairs are handed out by a ticket counter (tickets[0]; tickets[1] counts the warps that
  // are done, the last one clears both for the next launch on this stream): in a batch of mixed
  // sizes a row pair costs anything from 1 to 64 strips, and with a fixed stride the warp with the
  // most wide rows decides when the kernel ends (0.50 -> 0.6x of the measured peak on configs[4]).
  // The next ticket is drawn before the current row pair is worked on, so its latency is hidden.
  const uint32_t n_warps_total = gridDim.x * kFastWarps;
  // A ticket stands for ticket_batch consecutive row pairs (1 unless the rows are narrow: a batch of 128-pixel-wide
  // images has one strip per row pair, and half a million tickets drawn on one address took as long as the conversion).
  const uint32_t n_units = total_units * kSplit;
  uint32_t chunk = blockIdx.x * kFastWarps + warp;
  while ((uint64_t)chunk * ticket_batch < n_units) {
    uint32_t next_chunk = 0;
    if (lane == 0) next_chunk = n_warps_total + atomicAdd(tickets, 1u);
    const uint32_t unit_end = (uint32_t)min((uint64_t)(chunk + 1) * ticket_batch, (uint64_t)n_units);
    for (uint32_t unit = chunk * ticket_batch; unit < unit_end; ++unit) {
    uint32_t rp;
    const uint32_t img = b2_find_image_warp(unit_prefix, n_jobs, uniform_units, unit / kSplit, &rp, lane);
    const int phase = (int)(unit % kSplit);
    const ImportJobDev* __restrict__ jp = jobs + img;
    const int width = jp->width;
    const int y0 = 2 * (int)rp;
    const bool has_row1 = (y0 + 1 < jp->height);
    const int sstride = jp->src_stride;
    // odd last row: the source row is duplicated (yuv.go:377), so row 1 of every output is
    // the same as row 0 and is written over it
    const int ystride = has_row1 ? jp->y_stride : 0;
    const uint8_t* const s0 = jp->src + (ptrdiff_t)y0 * sstride + (STEP3 ? 0u : lane * 16 + phase * 512);
    const uint8_t* const s1 = has_row1 ? s0 + sstride : s0;
    uint8_t* yp = jp->y + (ptrdiff_t)y0 * jp->y_stride + lane4 + phase * 128;
    uint8_t* ap = nullptr;
    int astride = 0;
    if (WRITE_A) { astride = has_row1 ? jp->a_stride : 0; ap = jp->a + (ptrdiff_t)y0 * jp->a_stride + lane4 + phase * 128; }
    const ptrdiff_t uvo = (ptrdiff_t)rp * jp->uv_stride + lane * 2 + phase * 64;
    uint8_t* up = jp->u + uvo;
    uint8_t* vp = jp->v + uvo;
    uint32_t and_acc = 0xffffffffu;
    // Dithered import: three rings of raw VP8Random draws (first luma row, second luma row,
    // chroma row), each 128 draws per strip.  Their start windows: the row pair's own state,
    // then width draws later, then width draws later again (a last odd row has no second
    // luma row: its chroma draws start after the first).
    uint32_t* const rng = reinterpret_cast<uint32_t*>(smem + kSmemBytes + warp * kDitherWarpBytes);
    uint32_t rpos = 0;
    if (DITHER) {
      uint32_t* ext = rng + 512;          // scratch inside the third ring until it is filled
      uint32_t* cw = rng + 512 + 128;
      const uint32_t* st = dither.rng_states + (size_t)(unit / kSplit) * B2_RNG_N;
      const uint32_t* pw = dither.rng_poly_w + (size_t)img * B2_RNG_N;
      __syncwarp();
      for (int i = (int)lane; i < B2_RNG_N; i += 32) { ext[i] = __ldg(st + i); cw[i] = __ldg(pw + i); }
      __syncwarp();
      for (int i = (int)lane; i < B2_RNG_N; i += 32) rng[(256 - B2_RNG_N) + i] = ext[i];
      b2_rng_jump_warp(ext, cw, (int)lane);
      for (int i = (int)lane; i < B2_RNG_N; i += 32) rng[256 + (256 - B2_RNG_N) + i] = ext[i];
      if (has_row1) b2_rng_jump_warp(ext, cw, (int)lane);
      const uint32_t e0 = ext[lane], e1 = ext[min((int)lane + 32, B2_RNG_N - 1)];
      __syncwarp();
      rng[512 + (256 - B2_RNG_N) + lane] = e0;
      if (lane + 32 < B2_RNG_N) rng[512 + (256 - B2_RNG_N) + lane + 32] = e1;
      __syncwarp();
    }

    // Strips of 128 pixels.  Rows that are 16-byte aligned are fetched with one 16-byte
    // cp.async per lane and row; whole strips of such rows take the check-free code.  The
    // last partial strip fetches the valid pixels only and takes the general code, where
    // lanes past the end of the row run on zeros (every lane takes part in the vote of
    // block_sums) and store nothing.  Rows that are only 4-byte aligned (a tight 4 * width
    // stride with width % 4 != 0): the UNAL kernels fetch the aligned chunks that enclose them;
    // the others (dithered import only) fetch 4 bytes at a time and take the general code.
    const bool al16 = ((((uintptr_t)jp->src) | (uintptr_t)(uint32_t)sstride) & 15) == 0;  // warp-uniform
    const int tiles_x = (width + 127) >> 7;
    const int full = (UNAL || STEP3 || al16) ? (width >> 7) : 0;
    const uint8_t* g0 = s0;
    const uint8_t* g1 = s1;
    // UNAL: m = pixels between the enclosing aligned address and the row start (0..3), g = this lane's
    // chunk of that aligned span, rem = bytes from the start of the chunk to the end of the row
    uint32_t m0 = 0, m1 = 0;
    int rem0 = 0, rem1 = 0;
    if (STEP3) {   // m = BYTES between the enclosing aligned address and the row start (0..15)
      m0 = (uint32_t)(uintptr_t)s0 & 15u;
      m1 = (uint32_t)(uintptr_t)s1 & 15u;
      g0 = s0 - m0 + 16 * lane;
      g1 = s1 - m1 + 16 * lane;
      rem0 = 3 * width + (int)m0 - 16 * (int)lane;
      rem1 = 3 * width + (int)m1 - 16 * (int)lane;
    } else if (UNAL) {
      m0 = ((uint32_t)(uintptr_t)s0 >> 2) & 3u;
      m1 = ((uint32_t)(uintptr_t)s1 >> 2) & 3u;
      g0 = s0 - 4 * m0;
      g1 = s1 - 4 * m1;
      rem0 = 4 * width + 4 * (int)m0 - 16 * (int)lane;
      rem1 = 4 * width + 4 * (int)m1 - 16 * (int)lane;
    }
    uint32_t wr = slot0, rd = slot0;
    auto issue = [&](int t) {  // one commit group per call, possibly empty
      if (STEP3) {
        if (t < tiles_x) {   // chunk 24 t + lane of both rows, lanes 0 .. 24; only bytes inside the row are read
          if (lane < 25) {
            if (rem0 > 0) cp_async16(wr, g0, (uint32_t)min(rem0, 16));
            if (rem1 > 0) cp_async16(wr + kRow1, g1, (uint32_t)min(rem1, 16));
          }
          rem0 -= 384; rem1 -= 384;
        }
      } else if (UNAL) {
        if (t < tiles_x) {
          // chunk 32 t + lane of both rows, lane 0 chunk 32 t + 32 too; only bytes inside the row are read
          // (the rest of a chunk is zero-filled).  A chunk past the row is not fetched at all: its lanes
          // compute on whatever an earlier strip left in the slot and store nothing (nvalid <= 0).
          if (rem0 > 0) cp_async16(wr, g0, (uint32_t)min(rem0, 16));
          if (rem1 > 0) cp_async16(wr + kRow1, g1, (uint32_t)min(rem1, 16));
          if (lane == 0) {
            if (rem0 > 512) cp_async16(wr + 512, g0 + 512, (uint32_t)min(rem0 - 512, 16));
            if (rem1 > 512) cp_async16(wr + kRow1 + 512, g1 + 512, (uint32_t)min(rem1 - 512, 16));
          }
          rem0 -= 512; rem1 -= 512;
        }
      } else if (t < full) {
        cp_async16_full(wr, g0);
        cp_async16_full(wr + kRow1, g1);
      } else if (t < tiles_x) {
        const int nv = min(max(width - ((int)lane4 + t * 128), 0), 4);
        const uint8_t* q0 = nv ? g0 : s0 - phase * 512;   // lanes past the row: zero fill from a valid address
        const uint8_t* q1 = nv ? g1 : s1 - phase * 512;
        if (al16) {
          cp_async16(wr, q0, 4u * nv);
          cp_async16(wr + kRow1, q1, 4u * nv);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t nb = (k < nv) ? 4u : 0u;
            cp_async4(wr + 4 * k, q0 + (nb ? 4 * k : 0), nb);
            cp_async4(wr + kRow1 + 4 * k, q1 + (nb ? 4 * k : 0), nb);
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      wr = slot_sum - wr;
      g0 += (STEP3 ? 384 : 512) * kSplit; g1 += (STEP3 ? 384 : 512) * kSplit;
    };
    // STEP3: the lane's 12 bytes start m + 12 lane bytes into the stage row
    auto unpack3 = [&](uint32_t row_addr, uint32_t m) -> uint4 {
      const uint32_t off = m + 12u * lane;
      const uint32_t a = row_addr + (off & ~3u), sh = 8u * (off & 3u);
      const uint32_t w0 = lds32_ring(a), w1 = lds32_ring(a + 4u), w2 = lds32_ring(a + 8u), w3 = lds32_ring(a + 12u);
      const uint32_t b0 = __funnelshift_r(w0, w1, sh), b1 = __funnelshift_r(w1, w2, sh), b2 = __funnelshift_r(w2, w3, sh);
      // pixels = bytes 0-2, 3-5, 6-8, 9-11; the top byte of each word is whatever comes next (never used)
      return make_uint4(b0, __byte_perm(b0, b1, 0x6543), __byte_perm(b1, b2, 0x5432), __byte_perm(b2, 0u, 0x0321));
    };
    // UNAL: the lane's four pixels start m words into its chunk and end in the next one
    auto shifted = [&](uint4& c, uint32_t next_addr, uint32_t m) {
      if (m != 0) {            // warp-uniform
        const uint4 n = lds128_ring(next_addr);
        if (m == 1) c = make_uint4(c.y, c.z, c.w, n.x);
        else if (m == 2) c = make_uint4(c.z, c.w, n.x, n.y);
        else c = make_uint4(c.w, n.x, n.y, n.z);
      }
    };
    auto strip = [&](auto full_tag, int t) {
      constexpr bool FULL = decltype(full_tag)::value;
      issue(t + kDist * kSplit);
      asm volatile("cp.async.wait_group %0;" :: "n"(kDist) : "memory");
      uint4 c0, c1;
      if (STEP3) {
        c0 = unpack3(rd - 16u * lane, m0);
        c1 = unpack3(rd - 16u * lane + kRow1, m1);
      } else {
        c0 = lds128_ring(rd);
        c1 = lds128_ring(rd + kRow1);
      }
      if (UNAL) {
        shifted(c0, rd + 16, m0);
        shifted(c1, rd + kRow1 + 16, m1);
      }
      rd = slot_sum - rd;
      uint4 dq0 = make_uint4(0, 0, 0, 0), dq1 = dq0, dq2 = dq0;
      if (DITHER) {
        // the next 128 draws of the three streams, 24 per step: s_n = s_(n-55) - s_(n-24) mod 2^31
#pragma unroll
        for (int b = 0; b < 128; b += 24) {
          const uint32_t i = rpos + b + lane;
          if (lane < 24 && b + (int)lane < 128) {
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              uint32_t* r = rng + 256 * q;
              r[i & 255u] = (r[(i - 55u) & 255u] - r[(i - 24u) & 255u]) & 0x7fffffffu;
            }
          }
          __syncwarp();
        }
        const uint32_t o = (rpos + lane4) & 255u;
        dq0 = *reinterpret_cast<const uint4*>(rng + o);
        dq1 = has_row1 ? *reinterpret_cast<const uint4*>(rng + 256 + o) : dq0;
        dq2 = *reinterpret_cast<const uint4*>(rng + 512 + o);
        rpos = (rpos + 128u) & 255u;
      }
      int nvalid = 4;
      if (!FULL) {
        // Last group of a row whose width is not a multiple of 4: only valid pixels were
        // fetched; the missing pixels become copies of valid ones.  Duplicating the odd
        // last column is exactly the reference's SUM2 / SUM2ALPHA forms (yuv.go:211-213,
        // 270-277: twice a 2-pixel sum == the 4-pixel sum of the duplicated column), and
        // copies cannot disturb the has-alpha reduction.
        nvalid = width - ((int)lane4 + t * 128);
        if (nvalid < 4 && nvalid > 0) {
          if (nvalid == 1) { c0.y = c0.x; c1.y = c1.x; }
          if (nvalid <= 2) { c0.z = c0.x; c1.z = c1.x; c0.w = c0.y; c1.w = c1.y; }
          else { c0.w = c0.z; c1.w = c1.z; }
        }
      }
      uint32_t yw0, yw1;
      if (DITHER) {
        const int amp = dither.amp;
        yw0 = pack_b2(luma_sum<ORDER>(c0.x, dither_value<16>(dq0.x, amp)), luma_sum<ORDER>(c0.y, dither_value<16>(dq0.y, amp)),
                      luma_sum<ORDER>(c0.z, dither_value<16>(dq0.z, amp)), luma_sum<ORDER>(c0.w, dither_value<16>(dq0.w, amp)));
        yw1 = pack_b2(luma_sum<ORDER>(c1.x, dither_value<16>(dq1.x, amp)), luma_sum<ORDER>(c1.y, dither_value<16>(dq1.y, amp)),
                      luma_sum<ORDER>(c1.z, dither_value<16>(dq1.z, amp)), luma_sum<ORDER>(c1.w, dither_value<16>(dq1.w, amp)));
      } else {
        yw0 = pack_b2(luma_sum<ORDER>(c0.x), luma_sum<ORDER>(c0.y), luma_sum<ORDER>(c0.z), luma_sum<ORDER>(c0.w));
        yw1 = pack_b2(luma_sum<ORDER>(c1.x), luma_sum<ORDER>(c1.y), luma_sum<ORDER>(c1.z), luma_sum<ORDER>(c1.w));
      }
      uint32_t ap0 = 0xffffffffu, ap1 = 0xffffffffu;
      if (ALPHA) {
        ap0 = pack_b3(c0.x, c0.y, c0.z, c0.w);
        ap1 = pack_b3(c1.x, c1.y, c1.z, c1.w);
        if (FULL || nvalid > 0) and_acc &= ap0 & ap1;
      }
      uint32_t sr0, sg0, sb0, sr1, sg1, sb1;
      block_sums<ORDER, ALPHA, TABLE>(gbase, dbase, c0.x, c0.y, c1.x, c1.y, __byte_perm(ap0, ap1, 0x5410), &sr0, &sg0, &sb0);
      block_sums<ORDER, ALPHA, TABLE>(gbase, dbase, c0.z, c0.w, c1.z, c1.w, __byte_perm(ap0, ap1, 0x7632), &sr1, &sg1, &sb1);
      const int r0 = l2g(sr0, lbase), gg0 = l2g(sg0, lbase), b0 = l2g(sb0, lbase);
      const int r1 = l2g(sr1, lbase), gg1 = l2g(sg1, lbase), b1 = l2g(sb1, lbase);
      uint32_t uu, vv;
      if (DITHER) {   // draws alternate u, v per chroma sample (picture_csp_enc.c.go:44-46)
        const int amp = dither.amp;
        uu = __byte_perm(u_sum(r0, gg0, b0, dither_value<18>(dq2.x, amp)), u_sum(r1, gg1, b1, dither_value<18>(dq2.z, amp)), 0x0073);
        vv = __byte_perm(v_sum(r0, gg0, b0, dither_value<18>(dq2.y, amp)), v_sum(r1, gg1, b1, dither_value<18>(dq2.w, amp)), 0x0073);
      } else {
        uu = __byte_perm(u_sum(r0, gg0, b0), u_sum(r1, gg1, b1), 0x0073);
        vv = __byte_perm(v_sum(r0, gg0, b0), v_sum(r1, gg1, b1), 0x0073);
      }
      if (FULL || nvalid >= 4) {
        b2_stg_stream_u32(yp, yw0);
        b2_stg_stream_u32(yp + ystride, yw1);
        if (WRITE_A) {
          b2_stg_stream_u32(ap, ap0);
          b2_stg_stream_u32(ap + astride, ap1);
        }
        b2_stg_stream_u16(up, uu);
        b2_stg_stream_u16(vp, vv);
      } else if (nvalid > 0) {  // partial group: byte stores of the valid samples only
        for (int k = 0; k < nvalid; ++k) {
          yp[k] = (uint8_t)(yw0 >> (8 * k));
          yp[ystride + k] = (uint8_t)(yw1 >> (8 * k));
          if (WRITE_A) {
            ap[k] = (uint8_t)(ap0 >> (8 * k));
            ap[astride + k] = (uint8_t)(ap1 >> (8 * k));
          }
        }
        up[0] = (uint8_t)uu; vp[0] = (uint8_t)vv;
        if (nvalid > 2) { up[1] = (uint8_t)(uu >> 8); vp[1] = (uint8_t)(vv >> 8); }
      }
      yp += 128 * kSplit; up += 64 * kSplit; vp += 64 * kSplit;
      if (WRITE_A) ap += 128 * kSplit;
    };
#pragma unroll
    for (int k = 0; k < kDist; ++k) issue(phase + k * kSplit);
    int t = phase;
#pragma unroll kUnroll
    for (; t < full; t += kSplit) strip(std::true_type{}, t);
    for (; t < tiles_x; t += kSplit) strip(std::false_type{}, t);
    if (ALPHA && has_alpha != nullptr) {
      if (__any_sync(0xffffffffu, and_acc != 0xffffffffu) && lane == 0) atomicOr(has_alpha + img, 1);
    }
    }
    chunk = __shfl_sync(0xffffffffu, next_chunk, 0);
  }
  if (lane == 0) {
    if (atomicAdd(tickets + 1, 1u) == n_warps_total - 1) { tickets[0] = 0; tickets[1] = 0; }
  }
}

struct FastArgs {
  int grid;
  cudaStream_t s;
  const ImportJobDev* jobs;
  const uint32_t* unit_prefix;
  uint32_t n_jobs, uniform_units, total_units;
  const uint32_t* tab;
  int32_t* has_alpha;
  ImportDither dither;
  uint32_t* tickets;
  bool unaligned;
  uint32_t ticket_batch;
};
template <int ORDER, bool ALPHA, bool WRITE_A, bool TABLE, bool DITHER, bool UNAL, bool STEP3 = false>
cudaError_t launch3(const FastArgs& a) {
  auto k = import_fast_kernel<ORDER, ALPHA, WRITE_A, TABLE, DITHER, UNAL, STEP3>;
  const int smem = kSmemBytes + (DITHER ? kFastWarps * (int)kDitherWarpBytes : 0);
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  k<<<a.grid, kFastThreads, smem, a.s>>>(a.jobs, a.unit_prefix, a.n_jobs, a.uniform_units, a.total_units, a.tab,
                                         a.has_alpha, a.dither, a.tickets, a.ticket_batch);
  return cudaGetLastError();
}
template <int ORDER, bool ALPHA, bool WRITE_A, bool TABLE>
cudaError_t launch1(const FastArgs& a) {
  // the dithered import keeps the 4-bytes-at-a-time fetch for rows that are not 16-byte aligned
  if (a.dither.rng_states != nullptr) return launch3<ORDER, ALPHA, WRITE_A, TABLE, true, false>(a);
  if (a.unaligned) return launch3<ORDER, ALPHA, WRITE_A, TABLE, false, true>(a);
  return launch3<ORDER, ALPHA, WRITE_A, TABLE, false, false>(a);
}
template <int ORDER, bool ALPHA, bool WRITE_A>
cudaError_t launch(const FastArgs& a, int div_table) {
  // the divide mode only matters when alpha is imported
  if (ALPHA && div_table) return launch1<ORDER, ALPHA, WRITE_A, true>(a);
  return launch1<ORDER, ALPHA, WRITE_A, false>(a);
}

}  // namespace

cudaError_t b2_import_fast_launch(const b2_ctx* ctx, cudaStream_t s, const ImportJobDev* jobs,
                                  const uint32_t* unit_prefix, uint32_t n_jobs,
                                  uint32_t uniform_units, uint32_t total_units, int order,
                                  bool alpha, bool write_a, int div_table, int32_t* has_alpha,
                                  ImportDither dither, uint32_t* tickets, bool unaligned, uint32_t ticket_batch) {
  FastArgs a;
  a.ticket_batch = ticket_batch < 1 ? 1u : ticket_batch;
  a.grid = b2_grid_for(ctx, kFastCtasPerSm, ((uint64_t)total_units * kSplit + kFastWarps - 1) / kFastWarps);
  a.s = s; a.jobs = jobs; a.unit_prefix = unit_prefix; a.n_jobs = n_jobs; a.uniform_units = uniform_units;
  a.total_units = total_units; a.tab = ctx->tab_dev; a.has_alpha = has_alpha; a.dither = dither;
  a.tickets = tickets; a.unaligned = unaligned;
  if (order == B2_ORDER_RGB || order == B2_ORDER_BGR) {   // 3-byte pixels: no alpha to import, no dithering here
    if (alpha || dither.rng_states != nullptr) return cudaErrorInvalidValue;
    if (order == B2_ORDER_RGB) return write_a ? launch3<0, false, true, false, false, false, true>(a) : launch3<0, false, false, false, false, false, true>(a);
    return write_a ? launch3<1, false, true, false, false, false, true>(a) : launch3<1, false, false, false, false, false, true>(a);
  }
#define B2_GO(O, A, W) return launch<O, A, W>(a, div_table)
  if (order == B2_ORDER_RGBA) {
    if (alpha) { if (write_a) B2_GO(0, true, true); else B2_GO(0, true, false); }
    else       { if (write_a) B2_GO(0, false, true); else B2_GO(0, false, false); }
  } else {
    if (alpha) { if (write_a) B2_GO(1, true, true); else B2_GO(1, true, false); }
    else       { if (write_a) B2_GO(1, false, true); else B2_GO(1, false, false); }
  }
#undef B2_GO
}
