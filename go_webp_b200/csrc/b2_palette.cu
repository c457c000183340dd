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
// new image -- the three hashes are tried with atomicCAS insertion, a collision shows as a
// failed CAS -- and keeps the table in shared memory.  Strategies 0..3 answer exactly what
// the reference answers for EVERY pixel value; the last resort is an open-addressing table
// on the full 32-bit colour (one multiply, one 64-bit shared load, rarely a second probe)
// instead of 8 dependent halvings: the same index for every colour of the palette, and for
// a colour that is not in it the reference's search does not terminate at all.
//
// Mapping: a warp owns a 2048-pixel segment of one row, four 128-pixel chunks in flight,
// a lane converts 4 pixels per chunk (one 128-bit load -> 4 / 2 / 1 / half a word).
#include "b2_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kSegPx = 2048;
constexpr int kLut = 2048;  // PALETTE_INV_SIZE, vp8l_encoder.go:1116-1117
constexpr uint32_t kEmpty16 = 0xffffu;
constexpr unsigned long long kEmpty64 = ~0ull;

struct PalJobDev {
  const uint32_t* src;
  uint32_t* dst;
  const uint32_t* palette;
  int32_t src_stride, dst_stride, width, height, palette_size;
  uint32_t segs;   // segments per row
  uint32_t flags;  // bit 0: rows take the vector path
};

struct PalShared {
  unsigned long long open[kLut];  // strategy 4: colour << 32 | index, open addressing
  uint32_t lut[kLut];             // strategies 1..3: index or 0xffff
  uint32_t pal[256];
  int strategy, collide;
};

__device__ __forceinline__ uint32_t pal_hash(int which, uint32_t c) {  // vp8l_encoder.go:1111-1127
  if (which == 0) return (c >> 8) & 0xffu;
  if (which == 1) return ((c & 0x00ffffffu) * 4222244071u) >> 21;
  return ((c & 0x00ffffffu) * 0x7fffffffu) >> 21;
}
__device__ __forceinline__ uint32_t open_hash(uint32_t c) { return (c * 0x9E3779B1u) >> 21; }

// CTA-wide: choose the strategy for this image's palette and build its table.
__device__ void build_tables(PalShared& sh, const PalJobDev* __restrict__ jp) {
  const int tid = threadIdx.x;
  const int P = jp->palette_size;
  __syncthreads();   // everyone is done with the previous image's tables
  sh.pal[tid] = (tid < P) ? __ldg(jp->palette + tid) : 0u;   // kThreads == 256 == MAX_PALETTE_SIZE
  if (tid == 0) sh.strategy = (P < 4) ? 0 : 4;
  __syncthreads();
  if (P < 4) return;
  for (int which = 0; which < 3; ++which) {
    for (int i = tid; i < kLut; i += kThreads) sh.lut[i] = kEmpty16;
    if (tid == 0) sh.collide = 0;
    __syncthreads();
    if (tid < P && atomicCAS(&sh.lut[pal_hash(which, sh.pal[tid])], kEmpty16, (uint32_t)tid) != kEmpty16) sh.collide = 1;
    __syncthreads();
    const int collide = sh.collide;
    __syncthreads();
    if (!collide) {
      if (tid == 0) sh.strategy = which + 1;
      __syncthreads();
      return;
    }
  }
  for (int i = tid; i < kLut; i += kThreads) sh.open[i] = kEmpty64;
  __syncthreads();
  if (tid < P) {
    const uint32_t c = sh.pal[tid];
    const unsigned long long e = ((unsigned long long)c << 32) | (uint32_t)tid;
    uint32_t h = open_hash(c);
    for (;;) {
      const unsigned long long old = atomicCAS(&sh.open[h], kEmpty64, e);
      if (old == kEmpty64) break;
      if ((uint32_t)(old >> 32) == c) { atomicMax(&sh.open[h], e); break; }  // duplicate colour: the later entry, as idx_map ends up
      h = (h + 1) & (kLut - 1);
    }
  }
  __syncthreads();
}

template <int STRATEGY>
__device__ __forceinline__ uint32_t lookup(const PalShared& sh, uint32_t pix) {
  if (STRATEGY == 0) {  // SearchColorGreedy, vp8l_encoder.go:1095-1109
    return pix == sh.pal[0] ? 0u : pix == sh.pal[1] ? 1u : pix == sh.pal[2] ? 2u : 3u;
  }
  if (STRATEGY <= 3) return sh.lut[pal_hash(STRATEGY - 1, pix)] & 0xffu;  // tmp_row is uint8
  uint32_t h = open_hash(pix);
  for (int probe = 0; probe < kLut; ++probe) {
    const unsigned long long e = sh.open[h];
    if ((uint32_t)(e >> 32) == pix && e != kEmpty64) return (uint32_t)e & 0xffu;
    if (e == kEmpty64) break;
    h = (h + 1) & (kLut - 1);
  }
  return 0xffu;  // not a palette colour: undefined in the reference (its search never ends)
}

// One row segment [xseg, xend) of a row; XBITS and the strategy are compile-time.
template <int XBITS, int STRATEGY, bool VEC>
__device__ __forceinline__ void convert_segment(const PalShared& sh, const uint32_t* __restrict__ srow,
                                                uint32_t* __restrict__ drow, int xseg, int xend, int width, int lane) {
  constexpr int kDepth = 8 >> XBITS;  // bits per index in the bundle
  for (int xc = xseg + lane * 4; xc < xend + lane * 4; xc += 4 * 128) {  // warp-uniform trip count
    uint4 pv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int x4 = xc + u * 128;
      pv[u] = make_uint4(0, 0, 0, 0);
      if (x4 < xend) {
        if (VEC && x4 + 3 < width) {
          pv[u] = b2_ldg_stream_v4(srow + x4);
        } else {
          pv[u].x = __ldg(srow + x4);
          if (x4 + 1 < width) pv[u].y = __ldg(srow + x4 + 1);
          if (x4 + 2 < width) pv[u].z = __ldg(srow + x4 + 2);
          if (x4 + 3 < width) pv[u].w = __ldg(srow + x4 + 3);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int x4 = xc + u * 128;
      const bool live = x4 < xend;   // lanes past the end still take part in the shuffle below
      uint32_t i0 = 0, i1 = 0, i2 = 0, i3 = 0;
      if (live) {
        i0 = lookup<STRATEGY>(sh, pv[u].x);
        if (x4 + 1 < width) i1 = lookup<STRATEGY>(sh, pv[u].y);
        if (x4 + 2 < width) i2 = lookup<STRATEGY>(sh, pv[u].z);
        if (x4 + 3 < width) i3 = lookup<STRATEGY>(sh, pv[u].w);
      }
      if (XBITS == 0) {       // 0xff000000 | idx << 8, lossless_enc.go:308-310
        if (live) {
          const uint4 o = make_uint4(0xff000000u | (i0 << 8), 0xff000000u | (i1 << 8), 0xff000000u | (i2 << 8), 0xff000000u | (i3 << 8));
          if (VEC && x4 + 3 < width) {
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
          if (VEC && x4 + 3 < width) {
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
}

template <int XBITS, int STRATEGY>
__device__ __forceinline__ void convert_units(const PalShared& sh, const PalJobDev* __restrict__ jp, uint32_t local,
                                              uint32_t warp, int lane) {
  const uint32_t unit = local * kWarps + warp;
  const uint32_t segs = jp->segs;
  if (unit >= segs * (uint32_t)jp->height) return;
  const uint32_t y = unit / segs;
  const int xseg = (int)(unit - y * segs) * kSegPx;
  const int width = jp->width;
  const uint32_t* __restrict__ srow = jp->src + (size_t)y * (uint32_t)jp->src_stride;
  uint32_t* __restrict__ drow = jp->dst + (size_t)y * (uint32_t)jp->dst_stride;
  const int xend = min(width, xseg + kSegPx);
  if (jp->flags & 1u) convert_segment<XBITS, STRATEGY, true>(sh, srow, drow, xseg, xend, width, lane);
  else convert_segment<XBITS, STRATEGY, false>(sh, srow, drow, xseg, xend, width, lane);
}

template <int XBITS>
__global__ void __launch_bounds__(kThreads, 3)
apply_palette_kernel(const PalJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix, uint32_t n_jobs,
                     uint32_t uniform, uint32_t total_items) {
  __shared__ PalShared sh;
  const int lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5;
  uint32_t cur_img = 0xffffffffu;
  for (uint32_t item = blockIdx.x; item < total_items; item += gridDim.x) {   // CTA items: one image per item
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, item, &local);
    const PalJobDev* __restrict__ jp = jobs + img;
    if (img != cur_img) {  // CTA-uniform
      build_tables(sh, jp);
      cur_img = img;
    }
    switch (sh.strategy) {  // CTA-uniform
      case 0: convert_units<XBITS, 0>(sh, jp, local, warp, lane); break;
      case 1: convert_units<XBITS, 1>(sh, jp, local, warp, lane); break;
      case 2: convert_units<XBITS, 2>(sh, jp, local, warp, lane); break;
      case 3: convert_units<XBITS, 3>(sh, jp, local, warp, lane); break;
      default: convert_units<XBITS, 4>(sh, jp, local, warp, lane); break;
    }
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
      d.segs = (uint32_t)((j.width + kSegPx - 1) / kSegPx);
      // vector path: 16-byte source rows and destination rows aligned for the widest store of this depth
      const int dst_al = xbits == 0 ? 15 : (xbits == 1 ? 7 : 3);
      const bool vec = ((((uintptr_t)j.src) | ((uintptr_t)j.src_stride * 4u)) & 15) == 0 &&
                       ((((uintptr_t)j.dst) | ((uintptr_t)j.dst_stride * 4u)) & (uintptr_t)dst_al) == 0;
      d.flags = vec ? 1u : 0u;
      const uint32_t items = (uint32_t)(((uint64_t)d.segs * (uint32_t)j.height + kWarps - 1) / kWarps);
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
    const int grid = b2_grid_for(ctx, 3, total);
    cudaStream_t s = ctx->streams[stream];
    switch (xbits) {
      case 0: apply_palette_kernel<0><<<grid, kThreads, 0, s>>>(djobs, dpre, (uint32_t)cnt, uniform, (uint32_t)total); break;
      case 1: apply_palette_kernel<1><<<grid, kThreads, 0, s>>>(djobs, dpre, (uint32_t)cnt, uniform, (uint32_t)total); break;
      case 2: apply_palette_kernel<2><<<grid, kThreads, 0, s>>>(djobs, dpre, (uint32_t)cnt, uniform, (uint32_t)total); break;
      default: apply_palette_kernel<3><<<grid, kThreads, 0, s>>>(djobs, dpre, (uint32_t)cnt, uniform, (uint32_t)total); break;
    }
    ctx->launches++;
    B2_CUDA(ctx, cudaGetLastError());
  }
  return B2_OK;
}
