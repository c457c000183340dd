// b2_vp8l.cu -- part 3: batched VP8L per-pixel transforms for sm_100a.
//
//   add-green / sub-green       pkg/libwebp/dsp/lossless.go:337-347,
//                               pkg/libwebp/dsp/lossless_enc.go:207-216
//   cross-colour inverse        lossless.go:349-413 (tile walk + per-pixel)
//   cross-colour forward        lossless_enc.go:218-239 over given tile codes
//                               (pkg/libwebp/enc/predictor_enc.go:879-889)
//   fused chains                the ApplyInverseTransforms order of
//                               pkg/vp8/vp8l_dec.c.go:710-729 (cross-colour^-1
//                               then add-green) and the encoder order of
//                               pkg/vp8/vp8l_encoder.go:926-961, one pass each
//   colour-index inverse        lossless.go:415-459 (8/4/2/1-bit packed indices)
//   bundle colour map           lossless_enc.go:293-312
//
// Pixels are uint32 0xAARRGGBB in tightly packed images.  The same-size
// transforms treat an image as a flat array: a thread owns 4 consecutive pixels
// (one 128-bit load, one 128-bit store); with tile bits >= 2 and width % 4 == 0
// those 4 pixels always share one tile code.  Colour-index works per row because
// the sub-byte packing restarts at every row; every thread produces 4 output
// pixels (one 128-bit store) from an 8-bit index source held in the green byte.
#include "b2_common.cuh"

namespace {

struct ArgbJobDev {
  const uint32_t* src;
  uint32_t* dst;
  const uint32_t* data;
  int32_t width, height, bits;
  uint32_t tiles_per_row;  // cross-colour: ceil(width / 2^bits); colour index: src row length
  uint32_t chunks_x;       // colour-index / bundle: chunks per row
  uint32_t flags;
};
enum { JOB_VEC = 1 };

enum Op { OP_ADD_GREEN, OP_SUB_GREEN, OP_CC_INV, OP_CC_FWD, OP_CHAIN_INV, OP_CHAIN_FWD };

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
#ifndef B2_FLAT_ITEM
#define B2_FLAT_ITEM 2048
#endif
#ifndef B2_FLAT_GROUPS
#define B2_FLAT_GROUPS 4
#endif
constexpr int kItemPx = B2_FLAT_ITEM;         // flat kernels: pixels per warp item

__device__ __forceinline__ uint32_t add_green(uint32_t argb) {
  const uint32_t green = (argb >> 8) & 0xff;
  uint32_t rb = argb & 0x00ff00ffu;
  rb += (green << 16) | green;
  rb &= 0x00ff00ffu;
  return (argb & 0xff00ff00u) | rb;
}
__device__ __forceinline__ uint32_t sub_green(uint32_t argb) {
  const uint32_t green = (argb >> 8) & 0xff;
  // (r - g) & 0xff and (b - g) & 0xff in one subtraction with borrow guards
  uint32_t rb = (argb & 0x00ff00ffu) | 0x01000100u;
  rb -= (green << 16) | green;
  rb &= 0x00ff00ffu;
  return (argb & 0xff00ff00u) | rb;
}
__device__ __forceinline__ int s8(uint32_t v) { return (int)(int8_t)(v & 0xff); }
// ColorTransformDelta with both operands promoted to int
__device__ __forceinline__ int ctd(int pred, int color) { return (pred * color) >> 5; }

struct Mult { int g2r, g2b, r2b; };
__device__ __forceinline__ Mult code_to_mult(uint32_t code) {  // lossless.go:353-357
  Mult m; m.g2r = s8(code); m.g2b = s8(code >> 8); m.r2b = s8(code >> 16); return m;
}
__device__ __forceinline__ uint32_t cc_inverse(const Mult& m, uint32_t argb) {  // lossless.go:359-374
  const int green = s8(argb >> 8);
  int new_red = (argb >> 16) & 0xff;
  int new_blue = argb & 0xff;
  new_red = (new_red + ctd(m.g2r, green)) & 0xff;
  new_blue += ctd(m.g2b, green);
  new_blue += ctd(m.r2b, s8((uint32_t)new_red));
  new_blue &= 0xff;
  return (argb & 0xff00ff00u) | ((uint32_t)new_red << 16) | (uint32_t)new_blue;
}
__device__ __forceinline__ uint32_t cc_forward(const Mult& m, uint32_t argb) {  // lossless_enc.go:224-239
  const int green = s8(argb >> 8);
  const int red = s8(argb >> 16);
  int new_red = red & 0xff;
  int new_blue = argb & 0xff;
  new_red = (new_red - ctd(m.g2r, green)) & 0xff;
  new_blue -= ctd(m.g2b, green);
  new_blue -= ctd(m.r2b, red);
  new_blue &= 0xff;
  return (argb & 0xff00ff00u) | ((uint32_t)new_red << 16) | (uint32_t)new_blue;
}

template <int OP>
__device__ __forceinline__ uint32_t apply_px(const Mult& m, uint32_t p) {
  if (OP == OP_ADD_GREEN) return add_green(p);
  if (OP == OP_SUB_GREEN) return sub_green(p);
  if (OP == OP_CC_INV) return cc_inverse(m, p);
  if (OP == OP_CC_FWD) return cc_forward(m, p);
  if (OP == OP_CHAIN_INV) return add_green(cc_inverse(m, p));
  return cc_forward(m, sub_green(p));
}
template <int OP> __host__ __device__ constexpr bool needs_codes() { return OP != OP_ADD_GREEN && OP != OP_SUB_GREEN; }

template <int OP>
__device__ __forceinline__ Mult fetch_mult(const ArgbJobDev& j, uint32_t i) {
  Mult m; m.g2r = m.g2b = m.r2b = 0;
  if (needs_codes<OP>()) {
    const uint32_t y = i / (uint32_t)j.width;
    const uint32_t x = i - y * (uint32_t)j.width;
    m = code_to_mult(__ldg(j.data + (y >> j.bits) * j.tiles_per_row + (x >> j.bits)));
  }
  return m;
}
template <int OP>
__device__ __forceinline__ uint4 apply4(uint4 codes, uint4 p, bool one_tile) {
  uint4 o;
  if (one_tile) {  // warp-uniform (width % 4 == 0): one set of multipliers for the group
    const Mult m = code_to_mult(codes.x);
    o.x = apply_px<OP>(m, p.x); o.y = apply_px<OP>(m, p.y); o.z = apply_px<OP>(m, p.z); o.w = apply_px<OP>(m, p.w);
  } else {
    o.x = apply_px<OP>(code_to_mult(codes.x), p.x); o.y = apply_px<OP>(code_to_mult(codes.y), p.y);
    o.z = apply_px<OP>(code_to_mult(codes.z), p.z); o.w = apply_px<OP>(code_to_mult(codes.w), p.w);
  }
  return o;
}

#ifndef B2_FLAT_CTAS
#define B2_FLAT_CTAS 4
#endif
#ifndef B2_FLAT_GRID_AL
#define B2_FLAT_GRID_AL 3
#endif
#ifndef B2_FLAT_GRID_UN
#define B2_FLAT_GRID_UN 4
#endif
#ifndef B2_FLAT_GROUPS_PLAIN
#define B2_FLAT_GROUPS_PLAIN 2
#endif
// Tile codes of the 4 consecutive flat pixels that start at (y, x).  Nearly always they lie
// in one tile of one row (always when width % 4 == 0): one load.  A group that crosses a
// tile boundary or the end of its row looks every pixel up.
template <int OP>
__device__ __forceinline__ uint4 codes_at(const uint32_t* __restrict__ data, uint32_t tpr, uint32_t bits,
                                          uint32_t w, uint32_t y, uint32_t x, bool* same) {
  uint4 c = make_uint4(0, 0, 0, 0);
  *same = true;
  if (!needs_codes<OP>()) return c;
  const uint32_t* __restrict__ row = data + (y >> bits) * tpr;
  const uint32_t tx = x >> bits;
  c.x = __ldg(row + tx);
  c.y = c.z = c.w = c.x;
  *same = (x + 3 < w) && (((x + 3) >> bits) == tx);
  if (!*same) {
#pragma unroll
    for (uint32_t k = 1; k < 4; ++k) {
      uint32_t xx = x + k, yy = y;
      while (xx >= w) { xx -= w; ++yy; }
      (k == 1 ? c.y : k == 2 ? c.z : c.w) = __ldg(data + (yy >> bits) * tpr + (xx >> bits));
    }
  }
  return c;
}

// A warp item = kItemPx consecutive flat pixels of one image, walked in steps of kSubPx (four
// 128-bit groups per lane, all loads issued before the first store).  The image lookup, the
// descriptor and the one division (flat index -> row, column) are paid once per item; from
// there a lane follows its (row, column) by additions.
template <int OP>
__global__ void __launch_bounds__(kThreads, needs_codes<OP>() ? B2_FLAT_CTAS : 4)
flat_kernel(const ArgbJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix,
            uint32_t n_jobs, uint32_t uniform, uint32_t total_items) {
  // groups per lane and step: the transforms with tile codes want four loads in flight, the
  // plain copies (add / sub green) run best with two (measured: 0.98 against 0.93 of the peak)
  constexpr int kSubGroups = needs_codes<OP>() ? B2_FLAT_GROUPS : B2_FLAT_GROUPS_PLAIN;
  constexpr int kSubPx = 128 * kSubGroups;
  static_assert(kItemPx % kSubPx == 0, "flat item");
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5;
  for (uint32_t item = blockIdx.x * kWarps + warp; item < total_items; item += gridDim.x * kWarps) {
    uint32_t local;
    const uint32_t img = b2_find_image_warp(prefix, n_jobs, uniform, item, &local, lane);
    const ArgbJobDev* __restrict__ jp = jobs + img;
    const uint32_t w = (uint32_t)jp->width;
    const uint32_t n = w * (uint32_t)jp->height;
    const uint32_t* src = jp->src;   // not restrict: dst may alias src
    uint32_t* dst = jp->dst;
    const uint32_t* __restrict__ data = jp->data;
    const uint32_t bits = (uint32_t)jp->bits, tpr = jp->tiles_per_row;
    const bool vec = (jp->flags & JOB_VEC) != 0;
    const uint32_t base = local * kItemPx;
    const uint32_t end = min(n, base + (uint32_t)kItemPx);
    uint32_t i = base + lane * 4;
    uint32_t y = 0, x = 0;
    if (needs_codes<OP>()) { y = i / w; x = i - y * w; }
    for (uint32_t s0 = base; s0 < end; s0 += kSubPx, i += kSubPx) {
      if (vec && s0 + kSubPx <= n) {
        uint4 p[kSubGroups], c[kSubGroups];
        bool same = true;
#pragma unroll
        for (int g = 0; g < kSubGroups; ++g) p[g] = b2_ld_stream_v4(src + i + 128 * g);  // not .nc: dst may alias src
#pragma unroll
        for (int g = 0; g < kSubGroups; ++g) {
          bool sg;
          c[g] = codes_at<OP>(data, tpr, bits, w, y, x, &sg);
          same = same && sg;
          if (needs_codes<OP>()) { x += 128; while (x >= w) { x -= w; ++y; } }
        }
        const bool one_tile = __all_sync(0xffffffffu, same);
#pragma unroll
        for (int g = 0; g < kSubGroups; ++g) b2_stg_stream_v4(dst + i + 128 * g, apply4<OP>(c[g], p[g], one_tile));
      } else {
        for (int g = 0; g < kSubGroups; ++g) {
          const uint32_t gi = i + 128 * g;
          if (vec && gi + 4 <= n) {
            bool sg;
            const uint4 cg = codes_at<OP>(data, tpr, bits, w, y, x, &sg);
            b2_stg_stream_v4(dst + gi, apply4<OP>(cg, b2_ld_stream_v4(src + gi), false));
          } else {
            for (uint32_t k = 0; k < 4; ++k) {
              if (gi + k < n) {
                const Mult m = fetch_mult<OP>(*jp, gi + k);
                dst[gi + k] = apply_px<OP>(m, src[gi + k]);
              }
            }
          }
          if (needs_codes<OP>()) { x += 128; while (x >= w) { x -= w; ++y; } }
        }
      }
    }
  }
}

// ---- colour-index inverse ------------------------------------------------------
// Item = 128 output pixels of one row (4 per thread).  CTA items never straddle
// images so the palette can sit in shared memory.
template <int BITS>
__global__ void __launch_bounds__(kThreads, 8)
color_index_kernel(const ArgbJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix,
                   uint32_t n_jobs, uint32_t uniform, uint32_t total_cta_items) {
  constexpr int kPal = 1 << (8 >> BITS);
  constexpr int kBpp = 8 >> BITS;
  constexpr uint32_t kBitMask = (1u << kBpp) - 1;
  __shared__ uint32_t pal[kPal];
  const int lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5;
  uint32_t cur_img = 0xffffffffu;
  for (uint32_t citem = blockIdx.x; citem < total_cta_items; citem += gridDim.x) {
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, citem, &local);
    const ArgbJobDev j = jobs[img];
    if (img != cur_img) {  // block-uniform
      __syncthreads();
      for (int i = threadIdx.x; i < kPal; i += kThreads) pal[i] = __ldg(j.data + i);
      __syncthreads();
      cur_img = img;
    }
    const uint32_t chunk = local * kWarps + warp;
    const uint32_t y = chunk / j.chunks_x;
    if (y >= (uint32_t)j.height) continue;
    const uint32_t cx = chunk - y * j.chunks_x;
    const int x4 = (int)cx * 128 + lane * 4;
    if (x4 >= j.width) continue;
    const uint32_t* srow = j.src + (size_t)y * j.tiles_per_row;
    uint32_t* drow = j.dst + (size_t)y * (uint32_t)j.width;
    const int nv = min(4, j.width - x4);
    uint32_t o[4];
    if (BITS == 0) {
      uint32_t s[4];
      if (nv == 4 && (((uintptr_t)(srow + x4)) & 15) == 0) {
        const uint4 v = b2_ldg_stream_v4(srow + x4);
        s[0] = v.x; s[1] = v.y; s[2] = v.z; s[3] = v.w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) s[k] = (k < nv) ? __ldg(srow + x4 + k) : 0u;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = pal[(s[k] >> 8) & 0xff];  // MapARGB_C, lossless.go:415-426
    } else {
      // x4 % 4 == 0, so the 4 pixels use at most two source pixels (BITS == 1)
      // or one (BITS >= 2); lossless.go:437-455
      const uint32_t s0 = (__ldg(srow + (x4 >> BITS)) >> 8) & 0xff;
      uint32_t s1 = s0;
      if (BITS == 1 && nv > 2) s1 = (__ldg(srow + ((x4 + 2) >> BITS)) >> 8) & 0xff;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t packed = (BITS == 1 && k >= 2) ? s1 : s0;
        const int sub = (x4 + k) & ((1 << BITS) - 1);
        o[k] = pal[(packed >> (kBpp * sub)) & kBitMask];
      }
    }
    if (nv == 4 && (((uintptr_t)(drow + x4)) & 15) == 0) {
      b2_stg_stream_v4(drow + x4, make_uint4(o[0], o[1], o[2], o[3]));
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) if (k < nv) drow[x4 + k] = o[k];
    }
  }
}

// Tuned colour-index inverse for the common layout (width % 4 == 0, 16-byte aligned
// dst, 8-byte aligned src).  The kernel above pays an image lookup, a descriptor load
// and two divisions for every 4 output pixels of a thread, and has one dependent
// load -> lookup -> store chain in flight per thread (0.60-0.72 of HBM peak for the
// sub-byte depths).  Here a warp owns a 2048-pixel segment of one row: descriptor work
// is paid once per segment, four 128-pixel chunks are in flight per iteration (all
// loads issued before the first lookup), and every warp keeps its own palette copy in
// shared memory.  Loads and stores stay output-centric: lane l of chunk c produces
// pixels 128 c + 4 l .. +3 (one 128-bit store) from the source word(s) that hold them.
constexpr int kSegPx = 2048;
template <int BITS>
__global__ void __launch_bounds__(kThreads, 8)
color_index_rows_kernel(const ArgbJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix,
                        uint32_t n_jobs, uint32_t uniform, uint32_t total_items) {
  constexpr int kPal = 1 << (8 >> BITS);
  constexpr int kBpp = 8 >> BITS;
  constexpr uint32_t kMask4 = ((1u << kBpp) - 1) << 2;  // index mask, pre-scaled to a byte offset
  __shared__ uint32_t pal_all[kWarps * kPal];
  const int lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5;
  uint32_t* pal = pal_all + warp * kPal;
  const uint8_t* palb = reinterpret_cast<const uint8_t*>(pal);
  uint32_t cur_img = 0xffffffffu;
  for (uint32_t item = blockIdx.x * kWarps + warp; item < total_items; item += gridDim.x * kWarps) {
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, item, &local);
    const ArgbJobDev* __restrict__ jp = jobs + img;
    const int width = jp->width;
    if (img != cur_img) {  // warp-uniform
      __syncwarp();
      const uint32_t* __restrict__ src_pal = jp->data;
      for (int i = lane; i < kPal; i += 32) pal[i] = __ldg(src_pal + i);
      __syncwarp();
      cur_img = img;
    }
    const uint32_t segs = jp->chunks_x;  // segments per row
    const uint32_t y = local / segs;
    const int xseg = (int)(local - y * segs) * kSegPx;
    const uint32_t* __restrict__ srow = jp->src + (size_t)y * jp->tiles_per_row;
    uint32_t* __restrict__ drow = jp->dst + (size_t)y * (uint32_t)width;
    const int xend = min(width, xseg + kSegPx);
    // rows of any width and alignment: vector accesses where this ROW allows them (warp-uniform), word by word
    // otherwise; the last group of a row may hold fewer than 4 pixels
    const bool dal = (((uintptr_t)drow) & 15) == 0;
    const bool sal = (((uintptr_t)srow) & (BITS == 0 ? 15 : 7)) == 0;
    const int src_words = (int)jp->tiles_per_row;
    for (int xc = xseg + lane * 4; xc < xend + lane * 4; xc += 4 * 128) {  // warp-uniform trip count
      uint32_t s0[4], s1[4];
      uint4 sv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int x4 = xc + u * 128;
        s0[u] = s1[u] = 0u;
        sv[u] = make_uint4(0u, 0u, 0u, 0u);
        if (x4 < xend) {
          if (BITS == 0) {
            if (sal && x4 + 4 <= xend) sv[u] = b2_ldg_stream_v4(srow + x4);
            else {
              sv[u].x = __ldg(srow + x4);
              if (x4 + 1 < xend) sv[u].y = __ldg(srow + x4 + 1);
              if (x4 + 2 < xend) sv[u].z = __ldg(srow + x4 + 2);
              if (x4 + 3 < xend) sv[u].w = __ldg(srow + x4 + 3);
            }
          } else if (BITS == 1) {
            const int sw = x4 >> 1;
            if (sal && sw + 2 <= src_words) { const uint2 t = b2_ldg_stream_v2(srow + sw); s0[u] = t.x; s1[u] = t.y; }
            else { s0[u] = __ldg(srow + sw); if (sw + 1 < src_words) s1[u] = __ldg(srow + sw + 1); }
          } else s0[u] = __ldg(srow + (x4 >> BITS));   // shared by 1 (BITS == 2) or 2 lanes
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int x4 = xc + u * 128;
        if (x4 < xend) {
          uint4 o;
          if (BITS == 0) {  // MapARGB_C, lossless.go:415-426
            o.x = *reinterpret_cast<const uint32_t*>(palb + ((sv[u].x >> 6) & 0x3fc));
            o.y = *reinterpret_cast<const uint32_t*>(palb + ((sv[u].y >> 6) & 0x3fc));
            o.z = *reinterpret_cast<const uint32_t*>(palb + ((sv[u].z >> 6) & 0x3fc));
            o.w = *reinterpret_cast<const uint32_t*>(palb + ((sv[u].w >> 6) & 0x3fc));
          } else if (BITS == 1) {  // lossless.go:437-455, two 4-bit indices per source pixel
            o.x = *reinterpret_cast<const uint32_t*>(palb + ((s0[u] >> 6) & kMask4));
            o.y = *reinterpret_cast<const uint32_t*>(palb + ((s0[u] >> 10) & kMask4));
            o.z = *reinterpret_cast<const uint32_t*>(palb + ((s1[u] >> 6) & kMask4));
            o.w = *reinterpret_cast<const uint32_t*>(palb + ((s1[u] >> 10) & kMask4));
          } else {  // four (BITS == 2) or eight (BITS == 3) indices per source pixel
            const uint32_t g = (BITS == 3) ? (s0[u] >> (x4 & 4)) : s0[u];  // x4 & 7 is 0 or 4
            o.x = *reinterpret_cast<const uint32_t*>(palb + ((g >> 6) & kMask4));
            o.y = *reinterpret_cast<const uint32_t*>(palb + ((g >> (6 + kBpp)) & kMask4));
            o.z = *reinterpret_cast<const uint32_t*>(palb + ((g >> (6 + 2 * kBpp)) & kMask4));
            o.w = *reinterpret_cast<const uint32_t*>(palb + ((g >> (6 + 3 * kBpp)) & kMask4));
          }
          if (dal && x4 + 4 <= xend) {
            b2_stg_stream_v4(drow + x4, o);
          } else {
            drow[x4] = o.x;
            if (x4 + 1 < xend) drow[x4 + 1] = o.y;
            if (x4 + 2 < xend) drow[x4 + 2] = o.z;
            if (x4 + 3 < xend) drow[x4 + 3] = o.w;
          }
        }
      }
    }
  }
}

// ---- bundle colour map (forward) -------------------------------------------------
template <int BITS>
__global__ void __launch_bounds__(kThreads, 8)
bundle_kernel(const ArgbJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix,
              uint32_t n_jobs, uint32_t uniform, uint32_t total_items) {
  constexpr int kPer = 1 << BITS;
  constexpr int kDepth = 8 >> BITS;  // bit_depth = 1 << (3 - xbits)
  const int lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5;
  for (uint32_t item = blockIdx.x * kWarps + warp; item < total_items; item += gridDim.x * kWarps) {
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, item, &local);
    const ArgbJobDev j = jobs[img];
    const uint32_t y = local / j.chunks_x;
    const uint32_t cx = local - y * j.chunks_x;
    const uint8_t* irow = (const uint8_t*)j.src + (size_t)y * (uint32_t)j.width;
    uint32_t* drow = j.dst + (size_t)y * j.tiles_per_row;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t d = cx * 256 + k * 32 + lane;
      if (d < j.tiles_per_row) {
        uint32_t code = 0xff000000u;
#pragma unroll
        for (int s = 0; s < kPer; ++s) {
          const uint32_t x = d * kPer + s;
          if (x < (uint32_t)j.width) code |= (uint32_t)irow[x] << (8 + kDepth * s);
        }
        drow[d] = code;
      }
    }
  }
}

struct Prepared {
  const ArgbJobDev* jobs;
  const uint32_t* prefix;
  uint32_t uniform;
  uint32_t total;
};

enum Kind { KIND_FLAT, KIND_FLAT_CODES, KIND_INDEX, KIND_INDEX_ROWS, KIND_BUNDLE };

int prepare(b2_ctx* ctx, int stream, const b2_argb_job* jobs, int n_jobs, Kind kind, Prepared* out) {
  const size_t jb = sizeof(ArgbJobDev) * (size_t)n_jobs;
  const size_t pb = sizeof(uint32_t) * ((size_t)n_jobs + 1);
  const size_t pre_off = (jb + 15) & ~(size_t)15;
  if (pre_off + pb > B2_DESC_RING_BYTES) return B2_ERR_INVALID_CONFIG;
  uint8_t* host = (uint8_t*)malloc(pre_off + pb);
  if (host == NULL) return B2_ERR_OOM;
  ArgbJobDev* dj = (ArgbJobDev*)host;
  uint32_t* pre = (uint32_t*)(host + pre_off);
  uint64_t total = 0;
  uint32_t uniform = 0;
  int st = B2_OK;
  for (int i = 0; i < n_jobs && st == B2_OK; ++i) {
    const b2_argb_job& j = jobs[i];
    if (j.src == NULL || j.dst == NULL) { st = B2_ERR_NULL; break; }
    if (j.width <= 0 || j.height <= 0 || j.width > B2_MAX_DIMENSION || j.height > B2_MAX_DIMENSION) {
      st = B2_ERR_BAD_DIM; break;
    }
    ArgbJobDev& d = dj[i];
    d.src = j.src; d.dst = j.dst; d.data = j.data;
    d.width = j.width; d.height = j.height; d.bits = j.bits;
    d.tiles_per_row = 0; d.chunks_x = 0; d.flags = 0;
    uint32_t items = 0;
    const uint64_t n = (uint64_t)j.width * (uint64_t)j.height;
    switch (kind) {
      case KIND_FLAT_CODES:
        if (j.data == NULL) { st = B2_ERR_NULL; break; }
        // MIN_TRANSFORM_BITS + 3 read bits: vp8l_dec.c.go:1243 -> 2..9
        if (j.bits < 2 || j.bits > 9) { st = B2_ERR_INVALID_CONFIG; break; }
        d.tiles_per_row = (uint32_t)((j.width + (1 << j.bits) - 1) >> j.bits);
        // fallthrough
      case KIND_FLAT:
        if ((((uintptr_t)j.src | (uintptr_t)j.dst) & 15) == 0) d.flags |= JOB_VEC;
        items = (uint32_t)((n + kItemPx - 1) / kItemPx);
        break;
      case KIND_INDEX:
        if (j.data == NULL) { st = B2_ERR_NULL; break; }
        if (j.bits < 0 || j.bits > 3) { st = B2_ERR_INVALID_CONFIG; break; }
        if (j.src == j.dst) { st = B2_ERR_INVALID_CONFIG; break; }
        d.tiles_per_row = (uint32_t)((j.width + (1 << j.bits) - 1) >> j.bits);
        d.chunks_x = (uint32_t)((j.width + 127) / 128);
        if (j.bits == 0) {
          // 8-bit indices are a per-pixel map of a tightly packed image onto a tightly packed image: seen as ONE row
          // of width * height pixels, every 4-pixel group is a 16-byte aligned access whatever the width
          // (4095- and 4094-wide images ran at 0.69 / 0.75 of the peak against 0.90, word by word)
          d.width = (int32_t)n; d.height = 1;
          d.tiles_per_row = (uint32_t)n;
          d.chunks_x = (uint32_t)((n + 127) / 128);
        }
        items = (uint32_t)(((uint64_t)d.chunks_x * (uint32_t)d.height + kWarps - 1) / kWarps);  // CTA items
        break;
      case KIND_INDEX_ROWS:  // preconditions were checked by the caller (index_rows_ok)
        d.tiles_per_row = (uint32_t)((j.width + (1 << j.bits) - 1) >> j.bits);
        d.chunks_x = (uint32_t)((j.width + kSegPx - 1) / kSegPx);
        items = d.chunks_x * (uint32_t)j.height;
        break;
      case KIND_BUNDLE:
        if (j.bits < 0 || j.bits > 3) { st = B2_ERR_INVALID_CONFIG; break; }
        d.tiles_per_row = (uint32_t)((j.width + (1 << j.bits) - 1) >> j.bits);
        d.chunks_x = (d.tiles_per_row + 255) / 256;
        items = d.chunks_x * (uint32_t)j.height;
        break;
    }
    if (st != B2_OK) break;
    if (i == 0) uniform = items; else if (items != uniform) uniform = 0;
    pre[i] = (uint32_t)total;
    total += items;
    if (total > 0x7fffffffull) st = B2_ERR_BAD_DIM;
  }
  if (st != B2_OK) { free(host); return st; }
  pre[n_jobs] = (uint32_t)total;
  void* dev = NULL;
  st = b2_upload_desc(ctx, stream, host, pre_off + pb, &dev);
  free(host);
  if (st != B2_OK) return st;
  out->jobs = (const ArgbJobDev*)dev;
  out->prefix = (const uint32_t*)((uint8_t*)dev + pre_off);
  out->uniform = uniform;
  out->total = (uint32_t)total;
  return B2_OK;
}

// The tuned colour-index kernel takes any width and alignment (it picks vector or word accesses per row).
bool index_rows_ok(const b2_argb_job& j) {
  return j.src != NULL && j.dst != NULL && j.data != NULL && j.src != j.dst && j.width > 0 && j.height > 0 &&
         j.width <= B2_MAX_DIMENSION && j.height <= B2_MAX_DIMENSION;
}

template <int OP>
int run_flat(b2_ctx* ctx, int stream, const b2_argb_job* jobs, int n_jobs) {
  B2_CHECK_CTX(ctx, stream);
  if (n_jobs < 0) return B2_ERR_INVALID_CONFIG;
  if (n_jobs == 0) return B2_OK;
  if (jobs == NULL) return B2_ERR_NULL;
  Prepared p;
  const int st = prepare(ctx, stream, jobs, n_jobs, needs_codes<OP>() ? KIND_FLAT_CODES : KIND_FLAT, &p);
  if (st != B2_OK) return st;
  // Persistent CTAs per SM.  Measured on 4096^2 batches (profiles/r01d_grid_sweeps.md):
  // the pure copies (add/sub-green) peak at 4 per SM (0.95 of the measured HBM peak vs
  // 0.88 at 3, 5 or 6), the cross-colour kernels need the extra warps of 6 per SM.
  // Resident CTAs per SM (the grid is persistent; the register budget allows B2_FLAT_CTAS).  Measured on
  // 64 x 4096^2 and on the mixed-size batch of configs[4] (profiles/r02c_mixed_sizes.md): images whose
  // rows are whole 4-pixel groups stream best with 3 CTAs per SM (0.99 of the measured peak against 0.90
  // with 4: fewer concurrent DRAM streams), the others carry more instructions per pixel and want 4.
  bool whole_groups = true;
  for (int i = 0; i < n_jobs; ++i) whole_groups = whole_groups && (jobs[i].width & 3) == 0;
  const int grid_ctas = !needs_codes<OP>() ? 4 : (whole_groups ? B2_FLAT_GRID_AL : B2_FLAT_GRID_UN);
  const int grid = b2_grid_for(ctx, grid_ctas, (p.total + kWarps - 1) / kWarps);
  flat_kernel<OP><<<grid, kThreads, 0, ctx->streams[stream]>>>(p.jobs, p.prefix, n_jobs, p.uniform, p.total);
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}

}  // namespace

extern "C" int b2_vp8l_add_green(b2_ctx* c, int s, const b2_argb_job* j, int n) { return run_flat<OP_ADD_GREEN>(c, s, j, n); }
extern "C" int b2_vp8l_sub_green(b2_ctx* c, int s, const b2_argb_job* j, int n) { return run_flat<OP_SUB_GREEN>(c, s, j, n); }
extern "C" int b2_vp8l_color_inverse(b2_ctx* c, int s, const b2_argb_job* j, int n) { return run_flat<OP_CC_INV>(c, s, j, n); }
extern "C" int b2_vp8l_color_forward(b2_ctx* c, int s, const b2_argb_job* j, int n) { return run_flat<OP_CC_FWD>(c, s, j, n); }
extern "C" int b2_vp8l_inverse_chain(b2_ctx* c, int s, const b2_argb_job* j, int n) { return run_flat<OP_CHAIN_INV>(c, s, j, n); }
extern "C" int b2_vp8l_forward_chain(b2_ctx* c, int s, const b2_argb_job* j, int n) { return run_flat<OP_CHAIN_FWD>(c, s, j, n); }

extern "C" int b2_vp8l_color_index_inverse(b2_ctx* ctx, int stream, const b2_argb_job* jobs, int n_jobs) {
  B2_CHECK_CTX(ctx, stream);
  if (n_jobs < 0) return B2_ERR_INVALID_CONFIG;
  if (n_jobs == 0) return B2_OK;
  if (jobs == NULL) return B2_ERR_NULL;
  // one launch per distinct 'bits' value present in the batch (the kernel is
  // specialised on it); jobs keep their relative order inside each launch.
  for (int bits = 0; bits <= 3; ++bits) {
    int cnt = 0;
    for (int i = 0; i < n_jobs; ++i) {
      if (jobs[i].bits < 0 || jobs[i].bits > 3) return B2_ERR_INVALID_CONFIG;
      cnt += (jobs[i].bits == bits);
    }
    if (cnt == 0) continue;
    const b2_argb_job* sel = jobs;
    b2_argb_job* tmp = NULL;
    if (cnt != n_jobs) {
      tmp = (b2_argb_job*)malloc(sizeof(b2_argb_job) * (size_t)cnt);
      if (tmp == NULL) return B2_ERR_OOM;
      int k = 0;
      for (int i = 0; i < n_jobs; ++i) if (jobs[i].bits == bits) tmp[k++] = jobs[i];
      sel = tmp;
    }
    // 8-bit indices (bits == 0) stay on the per-CTA-palette kernel: with 256 entries the
    // lookups are bank-conflict bound and it measures faster (0.88 vs 0.78 of HBM peak)
    bool rows = bits != 0;
    for (int i = 0; i < cnt; ++i) rows = rows && index_rows_ok(sel[i]);
    Prepared p;
    const int st = prepare(ctx, stream, sel, cnt, rows ? KIND_INDEX_ROWS : KIND_INDEX, &p);
    free(tmp);
    if (st != B2_OK) return st;
    cudaStream_t s = ctx->streams[stream];
    if (rows) {
      // 3 persistent CTAs per SM measured best for the sub-byte depths (0.85-0.92 of the
      // measured HBM peak vs 0.80-0.84 at 8; profiles/r01d_grid_sweeps.md)
      const int rgrid = b2_grid_for(ctx, 3, (p.total + kWarps - 1) / kWarps);
      switch (bits) {
        case 0: color_index_rows_kernel<0><<<rgrid, kThreads, 0, s>>>(p.jobs, p.prefix, cnt, p.uniform, p.total); break;
        case 1: color_index_rows_kernel<1><<<rgrid, kThreads, 0, s>>>(p.jobs, p.prefix, cnt, p.uniform, p.total); break;
        case 2: color_index_rows_kernel<2><<<rgrid, kThreads, 0, s>>>(p.jobs, p.prefix, cnt, p.uniform, p.total); break;
        default: color_index_rows_kernel<3><<<rgrid, kThreads, 0, s>>>(p.jobs, p.prefix, cnt, p.uniform, p.total); break;
      }
      ctx->launches++;
      B2_CUDA(ctx, cudaGetLastError());
      continue;
    }
    const int grid = b2_grid_for(ctx, 8, p.total);
    switch (bits) {
      case 0: color_index_kernel<0><<<grid, kThreads, 0, s>>>(p.jobs, p.prefix, cnt, p.uniform, p.total); break;
      case 1: color_index_kernel<1><<<grid, kThreads, 0, s>>>(p.jobs, p.prefix, cnt, p.uniform, p.total); break;
      case 2: color_index_kernel<2><<<grid, kThreads, 0, s>>>(p.jobs, p.prefix, cnt, p.uniform, p.total); break;
      default: color_index_kernel<3><<<grid, kThreads, 0, s>>>(p.jobs, p.prefix, cnt, p.uniform, p.total); break;
    }
    ctx->launches++;
    B2_CUDA(ctx, cudaGetLastError());
  }
  return B2_OK;
}

extern "C" int b2_vp8l_bundle_color_map(b2_ctx* ctx, int stream, const b2_argb_job* jobs, int n_jobs) {
  B2_CHECK_CTX(ctx, stream);
  if (n_jobs < 0) return B2_ERR_INVALID_CONFIG;
  if (n_jobs == 0) return B2_OK;
  if (jobs == NULL) return B2_ERR_NULL;
  for (int bits = 0; bits <= 3; ++bits) {
    int cnt = 0;
    for (int i = 0; i < n_jobs; ++i) {
      if (jobs[i].bits < 0 || jobs[i].bits > 3) return B2_ERR_INVALID_CONFIG;
      cnt += (jobs[i].bits == bits);
    }
    if (cnt == 0) continue;
    const b2_argb_job* sel = jobs;
    b2_argb_job* tmp = NULL;
    if (cnt != n_jobs) {
      tmp = (b2_argb_job*)malloc(sizeof(b2_argb_job) * (size_t)cnt);
      if (tmp == NULL) return B2_ERR_OOM;
      int k = 0;
      for (int i = 0; i < n_jobs; ++i) if (jobs[i].bits == bits) tmp[k++] = jobs[i];
      sel = tmp;
    }
    Prepared p;
    const int st = prepare(ctx, stream, sel, cnt, KIND_BUNDLE, &p);
    free(tmp);
    if (st != B2_OK) return st;
    const int grid = b2_grid_for(ctx, 8, (p.total + kWarps - 1) / kWarps);
    cudaStream_t s = ctx->streams[stream];
    switch (bits) {
      case 0: bundle_kernel<0><<<grid, kThreads, 0, s>>>(p.jobs, p.prefix, cnt, p.uniform, p.total); break;
      case 1: bundle_kernel<1><<<grid, kThreads, 0, s>>>(p.jobs, p.prefix, cnt, p.uniform, p.total); break;
      case 2: bundle_kernel<2><<<grid, kThreads, 0, s>>>(p.jobs, p.prefix, cnt, p.uniform, p.total); break;
      default: bundle_kernel<3><<<grid, kThreads, 0, s>>>(p.jobs, p.prefix, cnt, p.uniform, p.total); break;
    }
    ctx->launches++;
    B2_CUDA(ctx, cudaGetLastError());
  }
  return B2_OK;
}
