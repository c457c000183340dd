// b2_palette.cu -- ApplyPalette for sm_100a: the colour -> index half of row a33, the
// forward counterpart of colour indexing (pkg/vp8/vp8l_encoder.go:1093-1212, caller
// MapImageFromPalette :1220-1244), fused with VP8LBundleColorMap
// (pkg/libwebp/dsp/lossless_enc.go:293-312) so the u8 index row the reference keeps in
// tmp_row never exists: 4 B/px in, 4 / 2^xbits B/px out.
//
// The reference picks, per image, the first strategy that works: greedy compares below 4
// colours, else the first of three hashes that maps the palette into a 2048-entry table
// without collision, else a binary search in the sorted palette
// (pkg/libwebp/utils/palette.go:74-119).  Here a CTA does the same choice when it meets a
// new image and keeps ONE byte table in shared memory for all of them: a thread per colour
// writes its index into its slot, reads it back, and a collision shows as somebody else's
// index.  Every hashed strategy is the same three instructions, slot = ((pix & mask) * mult)
// >> shift, with (mask, mult, shift) = (0xff00, 1, 8), (0xffffff, 4222244071, 21),
// (0xffffff, 2^31 - 1, 21) for ApplyPaletteHash0/1/2.  Strategies 0..3 answer exactly what the
// reference answers for EVERY pixel value (an empty slot reads 0xff = the reference's 0xffff
// cut to tmp_row's uint8).  The last resort replaces the binary search (8 dependent halvings
// per pixel) by a multiplicative hash of the whole pixel into 16384 byte slots, re-drawn until
// it is collision-free on the palette (1 draw in 7 at 256 colours, a few microseconds per
// image): the same index for every colour of the palette, and for a colour that is not in it
// the reference's search does not terminate at all.  Palettes with repeated colours
// (GetColorPalette never makes one) fall back to a linear scan.
//
// Mapping: a warp owns a segment of up to 8192 pixels of one row and streams it through its
// own ring of a few 512-pixel stages (5 with 8-bit packing, 2 with the sub-byte ones), each filled by ONE cp.async.bulk (TMA) that lane 0 issues
// on the stage's mbarrier -- 8 KB in flight per warp whatever the lanes are doing; a lane
// converts 4 pixels per 128-pixel chunk (one 128-bit shared load -> 4 / 2 / 1 / half a word).
// Rows that miss the 16-byte alignment take the same conversion on direct loads.
#include "b2_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kSegPx = 8192;
constexpr int kBlkPx = 512;
// Ring depth and CTAs per SM by packing (measured, profiles/r02_palette_variants.txt): with 8-bit
// packing the kernel writes as much as it reads and wants deep rings (5 stages, 2 CTAs: 0.93-0.96 of the
// measured peak vs 0.89 with 2 stages, 3 CTAs); the sub-byte packings are almost pure reads and want
// more warps with shallow rings (2 colours: 0.71 vs 0.58).
#ifndef B2_PAL_STAGES0
#define B2_PAL_STAGES0 5
#endif
#ifndef B2_PAL_CTAS0
#define B2_PAL_CTAS0 2
#endif
#ifndef B2_PAL_STAGES1
#define B2_PAL_STAGES1 2
#endif
#ifndef B2_PAL_CTAS1
#define B2_PAL_CTAS1 3
#endif
constexpr int kMaxStages = 8;
__host__ __device__ constexpr int stages_of(int xbits) { return xbits == 0 ? B2_PAL_STAGES0 : B2_PAL_STAGES1; }
__host__ __device__ constexpr int ctas_of(int xbits) { return xbits == 0 ? B2_PAL_CTAS0 : B2_PAL_CTAS1; }
constexpr int kStageBytes = kBlkPx * 4;
constexpr int kTabBytes = 16384;
__host__ __device__ constexpr int smem_of(int xbits) { return kTabBytes + kWarps * stages_of(xbits) * kStageBytes; }
constexpr int kMaxDraws = 256;
enum { S_GREEDY = 0, S_HASHED = 1, S_SCAN = 2 };

struct PalJobDev {
  const uint32_t* src;
  uint32_t* dst;
  const uint32_t* palette;
  int32_t src_stride, dst_stride, width, height, palette_size;
  uint32_t segs;   // segments per row
  uint32_t flags;  // bit 0: rows take the bulk-copy path
};

struct PalShared {
  uint32_t pal[256];
  int kind, collide, size, strategy;   // strategy: the reference's numbering (0 greedy, 1..3 hashes, 4 exact search), for the record
  uint32_t mask, mult, shift;
  uint64_t bars[kWarps * kMaxStages];
};
struct Hash { uint32_t mask, mult, shift, g0, g1, g2; };   // g0..g2: the first three colours, for the greedy compares

__device__ __forceinline__ Hash hash_of_attempt(int attempt) {  // vp8l_encoder.go:1111-1127, then the wide draws
  Hash h;
  if (attempt == 0) { h.mask = 0xff00u; h.mult = 1u; h.shift = 8u; }
  else if (attempt == 1) { h.mask = 0x00ffffffu; h.mult = 4222244071u; h.shift = 21u; }
  else if (attempt == 2) { h.mask = 0x00ffffffu; h.mult = 0x7fffffffu; h.shift = 21u; }
  else { h.mask = 0xffffffffu; h.mult = 0x9E3779B1u + 0x7F4A7C15u * 2u * (uint32_t)(attempt - 3); h.shift = 18u; }
  h.g0 = h.g1 = h.g2 = 0u;
  return h;
}
__device__ __forceinline__ uint32_t slot_of(const Hash& h, uint32_t pix) { return ((pix & h.mask) * h.mult) >> h.shift; }

// CTA-wide: choose the strategy for this image's palette and build its table.  my_slot = the slot this
// thread filled for the previous image (-1: none); the table is all 0xff apart from those.
__device__ void build_tables(PalShared& sh, uint8_t* __restrict__ tab, const PalJobDev* __restrict__ jp, int& my_slot) {
  const int tid = threadIdx.x;
  const int P = jp->palette_size;
  __syncthreads();   // everyone is done with the previous image's table
  if (my_slot >= 0) { tab[my_slot] = 0xff; my_slot = -1; }
  const uint32_t c = (tid < P) ? __ldg(jp->palette + tid) : 0u;   // kThreads == 256 == MAX_PALETTE_SIZE
  sh.pal[tid] = c;
  if (tid == 0) { sh.kind = (P < 4) ? S_GREEDY : S_SCAN; sh.strategy = (P < 4) ? 0 : 4; sh.size = P; sh.collide = 0; }
  __syncthreads();
  if (P < 4) return;
  for (int attempt = 0; attempt < 3 + kMaxDraws; ++attempt) {
    Hash h = hash_of_attempt(attempt);
    const int slot = (int)slot_of(h, c);
    if (tid < P) tab[slot] = (uint8_t)tid;
    __syncthreads();
    if (tid < P && tab[slot] != (uint8_t)tid) sh.collide = 1;
    __syncthreads();
    const int collide = sh.collide;
    __syncthreads();
    if (!collide) {
      if (tid < P) my_slot = slot;
      if (tid == 0) { sh.kind = S_HASHED; sh.strategy = attempt < 3 ? attempt + 1 : 4; sh.mask = h.mask; sh.mult = h.mult; sh.shift = h.shift; }
      __syncthreads();
      return;
    }
    if (tid < P) tab[slot] = 0xff;
    if (tid == 0) sh.collide = 0;
    __syncthreads();
  }
}

template <int KIND>
__device__ __forceinline__ uint32_t lookup(const PalShared& sh, const uint8_t* __restrict__ tab, const Hash& h, uint32_t pix) {
  if (KIND == S_GREEDY) {  // SearchColorGreedy, vp8l_encoder.go:1095-1109
    return pix == h.g0 ? 0u : pix == h.g1 ? 1u : pix == h.g2 ? 2u : 3u;
  }
  if (KIND == S_HASHED) return tab[slot_of(h, pix)];
  for (int j = sh.size - 1; j >= 0; --j) if (sh.pal[j] == pix) return (uint32_t)j;   // repeated colours: the later entry
  return 0xffu;
}

// Converts and stores one 512-pixel block held in pv.  FULL: all 512 pixels exist and the rows are
// aligned (no predicates).
template <int XBITS, int KIND, bool FULL>
__device__ __forceinline__ void emit_block(const PalShared& sh, const uint8_t* __restrict__ tab, const Hash& h, const uint4 (&pv)[4],
                                           uint32_t* __restrict__ drow, int xb, int xend, int width, int lane) {
  constexpr int kDepth = 8 >> XBITS;  // bits per index in the bundle
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int x4 = xb + u * 128 + lane * 4;
    const bool live = FULL || x4 < xend;   // lanes past the end still take part in the shuffle below
    uint32_t i0 = 0, i1 = 0, i2 = 0, i3 = 0;
    if (live) {
      i0 = lookup<KIND>(sh, tab, h, pv[u].x);
      if (FULL || x4 + 1 < width) i1 = lookup<KIND>(sh, tab, h, pv[u].y);
      if (FULL || x4 + 2 < width) i2 = lookup<KIND>(sh, tab, h, pv[u].z);
      if (FULL || x4 + 3 < width) i3 = lookup<KIND>(sh, tab, h, pv[u].w);
    }
    if (XBITS == 0) {       // 0xff000000 | idx << 8, lossless_enc.go:308-310
      if (live) {
        const uint4 o = make_uint4(0xff000000u | (i0 << 8), 0xff000000u | (i1 << 8), 0xff000000u | (i2 << 8), 0xff000000u | (i3 << 8));
        if (FULL) {
          b2_stg_stream_v4(drow + x4, o);
        } else {
          drow[x4] = o.x;
          if (x4 + 1 < width) drow[x4 + 1] = o.y;
          if (x4 + 2 < width) drow[x4 + 2] = o.z;
          if (x4 + 3 < width) drow[x4 + 3] = o.w;
        }
      }
    } else if (XBITS == 1) {  // two 4-bit indices per word
      if (live) {
        const uint32_t w0 = 0xff000000u | (i0 << 8) | (i1 << (8 + kDepth));
        const uint32_t w1 = 0xff000000u | (i2 << 8) | (i3 << (8 + kDepth));
        if (FULL) {
          b2_stg_stream_v2(drow + (x4 >> 1), make_uint2(w0, w1));
        } else {
          drow[x4 >> 1] = w0;
          if (x4 + 2 < width) drow[(x4 >> 1) + 1] = w1;
        }
      }
    } else if (XBITS == 2) {  // four 2-bit indices per word
      if (live) drow[x4 >> 2] = 0xff000000u | (i0 << 8) | (i1 << (8 + kDepth)) | (i2 << (8 + 2 * kDepth)) | (i3 << (8 + 3 * kDepth));
    } else {                  // eight 1-bit indices per word: two lanes share one
      const int sh0 = 8 + (x4 & 4);
      uint32_t part = (i0 << sh0) | (i1 << (sh0 + 1)) | (i2 << (sh0 + 2)) | (i3 << (sh0 + 3));
      part |= __shfl_xor_sync(0xffffffffu, part, 1);
      if (live && (lane & 1) == 0) drow[x4 >> 3] = 0xff000000u | part;
    }
  }
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

// One segment through the warp's ring.  phase: bit s = parity to wait for on stage s (kept across
// segments).  Every stage of the segment is requested as soon as its slot is free.
template <int XBITS, int KIND>
__device__ __forceinline__ void convert_segment_bulk(const PalShared& sh, const uint8_t* __restrict__ tab, const Hash& h,
                                                     const uint32_t* __restrict__ srow, uint32_t* __restrict__ drow, int xseg,
                                                     int xend, int width, int lane, uint8_t* ring, uint32_t ring_s,
                                                     uint32_t bar_s, uint32_t& phase) {
  constexpr int kStages = stages_of(XBITS);
  const int n_blk = (xend - xseg + kBlkPx - 1) / kBlkPx;
  auto issue = [&](int b) {   // lane 0: request block b of the segment into stage b % kStages
    const int xb = xseg + b * kBlkPx;
    const uint32_t bytes = (uint32_t)(min(kBlkPx, xend - xb) * 4);   // width % 4 == 0: a multiple of 16
    const uint32_t st = (uint32_t)(b % kStages);
    mbar_expect_tx(bar_s + 8 * st, bytes);
    bulk_g2s(ring_s + st * kStageBytes, srow + xb, bytes, bar_s + 8 * st);
  };
  if (lane == 0) for (int b = 0; b < min(n_blk, kStages); ++b) issue(b);
#pragma unroll 1
  for (int b = 0; b < n_blk; ++b) {
    const uint32_t st = (uint32_t)(b % kStages);
    mbar_wait(bar_s + 8 * st, (phase >> st) & 1u);
    phase ^= 1u << st;
    const int xb = xseg + b * kBlkPx;
    const uint8_t* sp = ring + st * kStageBytes + lane * 16;
    uint4 pv[4];
    const bool full = xb + kBlkPx <= xend;   // warp-uniform
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      pv[u] = make_uint4(0, 0, 0, 0);
      if (full || xb + u * 128 + lane * 4 < xend) pv[u] = *reinterpret_cast<const uint4*>(sp + u * 512);
    }
    __syncwarp();   // every lane has taken its pixels: the stage can be refilled
    if (lane == 0 && b + kStages < n_blk) issue(b + kStages);
    if (full) emit_block<XBITS, KIND, true>(sh, tab, h, pv, drow, xb, xend, width, lane);
    else emit_block<XBITS, KIND, false>(sh, tab, h, pv, drow, xb, xend, width, lane);
  }
}

// Rows that miss the alignment of the bulk copies: direct loads, one block at a time.
template <int XBITS, int KIND>
__device__ __forceinline__ void convert_segment_direct(const PalShared& sh, const uint8_t* __restrict__ tab, const Hash& h,
                                                       const uint32_t* __restrict__ srow, uint32_t* __restrict__ drow, int xseg,
                                                       int xend, int width, int lane) {
#pragma unroll 1
  for (int xb = xseg; xb < xend; xb += kBlkPx) {
    uint4 pv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int x4 = xb + u * 128 + lane * 4;
      pv[u] = make_uint4(0, 0, 0, 0);
      if (x4 < xend) {
        pv[u].x = __ldg(srow + x4);
        if (x4 + 1 < width) pv[u].y = __ldg(srow + x4 + 1);
        if (x4 + 2 < width) pv[u].z = __ldg(srow + x4 + 2);
        if (x4 + 3 < width) pv[u].w = __ldg(srow + x4 + 3);
      }
    }
    emit_block<XBITS, KIND, false>(sh, tab, h, pv, drow, xb, xend, width, lane);
  }
}

template <int XBITS, int KIND>
__device__ __forceinline__ void convert_unit(const PalShared& sh, const uint8_t* __restrict__ tab, const PalJobDev* __restrict__ jp,
                                             uint32_t local, uint32_t warp, int lane, uint8_t* ring, uint32_t ring_s,
                                             uint32_t bar_s, uint32_t& phase) {
  const uint32_t unit = local * kWarps + warp;
  const uint32_t segs = jp->segs;
  if (unit >= segs * (uint32_t)jp->height) return;
  const uint32_t y = unit / segs;
  const int xseg = (int)(unit - y * segs) * kSegPx;
  const int width = jp->width;
  const uint32_t* __restrict__ srow = jp->src + (size_t)y * (uint32_t)jp->src_stride;
  uint32_t* __restrict__ drow = jp->dst + (size_t)y * (uint32_t)jp->dst_stride;
  const int xend = min(width, xseg + kSegPx);
  Hash h;
  h.mask = sh.mask; h.mult = sh.mult; h.shift = sh.shift;
  h.g0 = sh.pal[0]; h.g1 = sh.pal[1]; h.g2 = sh.pal[2];
  if (jp->flags & 1u) convert_segment_bulk<XBITS, KIND>(sh, tab, h, srow, drow, xseg, xend, width, lane, ring, ring_s, bar_s, phase);
  else convert_segment_direct<XBITS, KIND>(sh, tab, h, srow, drow, xseg, xend, width, lane);
}

template <int XBITS>
__global__ void __launch_bounds__(kThreads, ctas_of(XBITS))
apply_palette_kernel(const PalJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix, uint32_t n_jobs,
                     uint32_t uniform, uint32_t total_items) {
  extern __shared__ __align__(128) uint8_t smem[];   // byte table (palette index, 0xff = empty), then the warps' rings
  __shared__ PalShared sh;
  uint8_t* const tab = smem;
  const int lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5;
  constexpr int kStages = stages_of(XBITS);
  uint8_t* const ring = smem + kTabBytes + warp * (kStages * kStageBytes);
  const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
  const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(sh.bars + warp * kMaxStages);
  if (lane == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(bar_s + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < kTabBytes / 16; i += kThreads) reinterpret_cast<uint4*>(tab)[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
  __syncwarp();
  uint32_t phase = 0;
  int my_slot = -1;
  uint32_t cur_img = 0xffffffffu;
#pragma unroll 1
  for (uint32_t item = blockIdx.x; item < total_items; item += gridDim.x) {   // CTA items: one image per item
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, item, &local);
    const PalJobDev* __restrict__ jp = jobs + img;
    if (img != cur_img) {  // CTA-uniform
      build_tables(sh, tab, jp, my_slot);
      cur_img = img;
    }
    const int kind = sh.kind;  // CTA-uniform
    if (kind == S_HASHED) convert_unit<XBITS, S_HASHED>(sh, tab, jp, local, warp, lane, ring, ring_s, bar_s, phase);
    else if (kind == S_GREEDY) convert_unit<XBITS, S_GREEDY>(sh, tab, jp, local, warp, lane, ring, ring_s, bar_s, phase);
    else convert_unit<XBITS, S_SCAN>(sh, tab, jp, local, warp, lane, ring, ring_s, bar_s, phase);
  }
}

}  // namespace

extern "C" int b2_vp8l_apply_palette(b2_ctx* ctx, int stream, const b2_palette_job* jobs, int n_jobs) {
  B2_CHECK_CTX(ctx, stream);
  if (n_jobs < 0) return B2_ERR_INVALID_CONFIG;
  if (n_jobs == 0) return B2_OK;
  if (jobs == NULL) return B2_ERR_NULL;
  for (int i = 0; i < n_jobs; ++i) {
    const b2_palette_job& j = jobs[i];
    if (j.src == NULL || j.dst == NULL || j.palette == NULL) return B2_ERR_NULL;
    if (j.width <= 0 || j.height <= 0 || j.width > B2_MAX_DIMENSION || j.height > B2_MAX_DIMENSION) return B2_ERR_BAD_DIM;
    if (j.xbits < 0 || j.xbits > 3 || j.palette_size < 1 || j.palette_size > 256) return B2_ERR_INVALID_CONFIG;
    if (j.src_stride < j.width || j.dst_stride < ((j.width + (1 << j.xbits) - 1) >> j.xbits)) return B2_ERR_INVALID_CONFIG;
    // "src[] can be the same as dst[]" (vp8l_encoder.go:1151): safe in parallel only when every pixel
    // lands on its own word
    if ((const uint32_t*)j.dst == j.src && !(j.xbits == 0 && j.src_stride == j.dst_stride)) return B2_ERR_INVALID_CONFIG;
  }
  for (int xbits = 0; xbits <= 3; ++xbits) {   // the kernel is specialised on xbits: one launch per value present
    int cnt = 0;
    for (int i = 0; i < n_jobs; ++i) cnt += (jobs[i].xbits == xbits);
    if (cnt == 0) continue;
    const size_t jb = sizeof(PalJobDev) * (size_t)cnt;
    const size_t pre_off = (jb + 15) & ~(size_t)15;
    const size_t pb = sizeof(uint32_t) * ((size_t)cnt + 1);
    if (pre_off + pb > B2_DESC_RING_BYTES) return B2_ERR_INVALID_CONFIG;
    uint8_t* host = (uint8_t*)malloc(pre_off + pb);
    if (host == NULL) return B2_ERR_OOM;
    PalJobDev* dj = (PalJobDev*)host;
    uint32_t* pre = (uint32_t*)(host + pre_off);
    uint64_t total = 0;
    uint32_t uniform = 0;
    int k = 0;
    for (int i = 0; i < n_jobs; ++i) {
      const b2_palette_job& j = jobs[i];
      if (j.xbits != xbits) continue;
      PalJobDev& d = dj[k];
      d.src = j.src; d.dst = j.dst; d.palette = j.palette;
      d.src_stride = j.src_stride; d.dst_stride = j.dst_stride; d.width = j.width; d.height = j.height;
      d.palette_size = j.palette_size;
      // 8-bit indices on tightly packed images are a per-pixel map: seen as ONE row of width * height pixels every
      // row segment is 16-byte aligned whatever the width (4095-wide images took the direct-load path: 0.68 of the
      // peak against 0.91)
      const uint64_t n_px = (uint64_t)j.width * (uint64_t)j.height;
      if (xbits == 0 && (j.width & 3) != 0 && j.src_stride == j.width && j.dst_stride == j.width && (n_px & 3) == 0 && j.height > 1) {
        d.width = (int32_t)n_px; d.height = 1;
        d.src_stride = d.dst_stride = (int32_t)n_px;
      }
      d.segs = (uint32_t)((d.width + kSegPx - 1) / kSegPx);
      // vector path: 16-byte source rows and destination rows aligned for the widest store of this depth
      const int dst_al = xbits == 0 ? 15 : (xbits == 1 ? 7 : 3);
      const bool vec = (d.width & 3) == 0 && ((((uintptr_t)j.src) | ((uintptr_t)(uint32_t)d.src_stride * 4u)) & 15) == 0 &&
                       ((((uintptr_t)j.dst) | ((uintptr_t)(uint32_t)d.dst_stride * 4u)) & (uintptr_t)dst_al) == 0;
      d.flags = vec ? 1u : 0u;
      const uint32_t items = (uint32_t)(((uint64_t)d.segs * (uint32_t)d.height + kWarps - 1) / kWarps);
      if (k == 0) uniform = items; else if (items != uniform) uniform = 0;
      pre[k] = (uint32_t)total;
      total += items;
      ++k;
    }
    pre[cnt] = (uint32_t)total;
    void* dev = NULL;
    const int st = b2_upload_desc(ctx, stream, host, pre_off + pb, &dev);
    free(host);
    if (st != B2_OK) return st;
    const PalJobDev* djobs = (const PalJobDev*)dev;
    const uint32_t* dpre = (const uint32_t*)((uint8_t*)dev + pre_off);
    const int grid = b2_grid_for(ctx, ctas_of(xbits), total);
    cudaStream_t s = ctx->streams[stream];
    static bool smem_set = false;   // 80 KB of dynamic shared memory needs the opt-in, once per process
    if (!smem_set) {
      B2_CUDA(ctx, cudaFuncSetAttribute(apply_palette_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_of(0)));
      B2_CUDA(ctx, cudaFuncSetAttribute(apply_palette_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_of(1)));
      B2_CUDA(ctx, cudaFuncSetAttribute(apply_palette_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_of(2)));
      B2_CUDA(ctx, cudaFuncSetAttribute(apply_palette_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_of(3)));
      smem_set = true;
    }
    switch (xbits) {
      case 0: apply_palette_kernel<0><<<grid, kThreads, smem_of(0), s>>>(djobs, dpre, (uint32_t)cnt, uniform, (uint32_t)total); break;
      case 1: apply_palette_kernel<1><<<grid, kThreads, smem_of(1), s>>>(djobs, dpre, (uint32_t)cnt, uniform, (uint32_t)total); break;
      case 2: apply_palette_kernel<2><<<grid, kThreads, smem_of(2), s>>>(djobs, dpre, (uint32_t)cnt, uniform, (uint32_t)total); break;
      default: apply_palette_kernel<3><<<grid, kThreads, smem_of(3), s>>>(djobs, dpre, (uint32_t)cnt, uniform, (uint32_t)total); break;
    }
    ctx->launches++;
    B2_CUDA(ctx, cudaGetLastError());
  }
  return B2_OK;
}
