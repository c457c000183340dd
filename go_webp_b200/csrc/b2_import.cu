// b2_import.cu -- part 1: batched RGBA -> YUV420(+A) import for sm_100a.
//
// What it computes (all integer fixed point, bit-exact with the reference):
//   Y  : pkg/color/yuv/conversion.go:56-59 via pkg/libwebp/dsp/yuv.go:124-140
//   U,V: gamma-corrected (and alpha-weighted) 2x2 mean, yuv.go:155-302, then
//        conversion.go:51-69 with rounding YUV_HALF<<2 (yuv.go:142-153)
//   A  : byte 3 of every pixel (alpha_processing.go:193-207) + the image-level
//        has-alpha reduction of pkg/picture/conversions.go:56 in the same pass
// driven over whole images the way yuv.go:304-382 walks row pairs (the odd last
// row / column are the duplicated row / column the reference feeds SUM2 and
// rgb_stride=0 with, which equals clamping the coordinates).
//
// Mapping: one warp owns a 128-pixel x 2-row strip (a "work item"); a thread
// owns 4 pixels x 2 rows = two 2x2 chroma blocks: two 128-bit loads in, two
// 32-bit Y stores, two 32-bit A stores, one 16-bit U and one 16-bit V store out,
// all fully coalesced across the warp.  Persistent CTAs stride over the flat
// item list of the whole batch, so mixed image sizes share one launch.
// The two gamma tables live in shared memory replicated per lane (bank ==
// lane), which makes every data-dependent lookup conflict-free.
#include "b2_common.cuh"
#include "b2_import.cuh"

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kItemW = 128;  // pixels per warp item (x)

struct Tables {
  uint32_t g2l[256 * 32];  // [value][lane]  kGammaToLinearTab
  int2 l2g[32 * 32];       // [pos][lane]    {base', d}: LinearToGamma(v) = (base' + d*v) >> 7
};

__device__ __forceinline__ void load_tables(Tables& t, const uint32_t* __restrict__ tab) {
  for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) t.g2l[i] = __ldg(tab + B2_TAB_G2L + (i >> 5));
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
    const int pos = i >> 5;
    t.l2g[i] = make_int2((int)__ldg(tab + B2_TAB_L2G + 2 * pos), (int)__ldg(tab + B2_TAB_L2G + 2 * pos + 1));
  }
  __syncthreads();
}

// yuv.go:187-202 folded: v = pos*512 + frac; y = v0*512 + (v1-v0)*frac;
// (y + 64) >> 7 == (base' + d*v) >> 7 with base' = v0*512 + 64 - d*pos*512.
__device__ __forceinline__ int linear_to_gamma(const Tables& t, uint32_t v, int lane) {
  const int2 e = t.l2g[((v >> 9) << 5) + lane];
  return (e.x + e.y * (int)v) >> 7;
}

template <int ORDER> struct Px;
template <> struct Px<0> {  // bytes R,G,B,A
  static constexpr uint32_t kLo = 199u | (35u << 8) | (20u << 16);
  static constexpr uint32_t kHi = 65u | (129u << 8) | (25u << 16);
  __device__ static __forceinline__ uint32_t r(uint32_t p) { return p & 0xff; }
  __device__ static __forceinline__ uint32_t b(uint32_t p) { return (p >> 16) & 0xff; }
};
template <> struct Px<1> {  // bytes B,G,R,A
  static constexpr uint32_t kLo = 20u | (35u << 8) | (199u << 16);
  static constexpr uint32_t kHi = 25u | (129u << 8) | (65u << 16);
  __device__ static __forceinline__ uint32_t r(uint32_t p) { return (p >> 16) & 0xff; }
  __device__ static __forceinline__ uint32_t b(uint32_t p) { return p & 0xff; }
};
__device__ __forceinline__ uint32_t px_g(uint32_t p) { return (p >> 8) & 0xff; }

// RGBToY with rounding YUV_HALF (conversion.go:56-59): coefficients split into
// high and low bytes so two dp4a do the three multiplies on the packed pixel.
// Returns the 24-bit sum; Y is its byte 2.
template <int ORDER>
__device__ __forceinline__ uint32_t luma_sum(uint32_t p) {
  const uint32_t lo = __dp4a(p, Px<ORDER>::kLo, (uint32_t)((1 << 15) + (16 << 16)));
  const uint32_t hi = __dp4a(p, Px<ORDER>::kHi, 0u);
  return lo + (hi << 8);
}
__device__ __forceinline__ uint32_t pack_byte2(uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3) {
  const uint32_t lo = __byte_perm(s0, s1, 0x0062);  // {s0.b2, s1.b2, -, -}
  const uint32_t hi = __byte_perm(s2, s3, 0x0062);
  return __byte_perm(lo, hi, 0x5410);
}
__device__ __forceinline__ uint32_t pack_byte3(uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3) {
  const uint32_t lo = __byte_perm(s0, s1, 0x0073);  // {s0.b3, s1.b3, -, -}
  const uint32_t hi = __byte_perm(s2, s3, 0x0073);
  return __byte_perm(lo, hi, 0x5410);
}

// DIVIDE_BY_ALPHA (yuv.go:231): exact 4*sum/a, or the libwebp-binary table form.
__device__ __forceinline__ uint32_t divide_by_alpha(uint32_t sum, uint32_t a, bool table,
                                                    const uint32_t* __restrict__ tab) {
  if (table) return (sum * __ldg(tab + B2_TAB_INVA + a)) >> (19 - 2);
  return (4u * sum) / a;
}

// One 2x2 block (pa pb / pc pd) -> U,V.  al4 = the four alpha bytes packed.
// yuv.go:243-302 (WebPAccumulateRGBA / RGB) + yuv.go:142-153.
template <int ORDER, bool ALPHA>
__device__ __forceinline__ void block_uv(const Tables& t, int lane, uint32_t pa, uint32_t pb,
                                         uint32_t pc, uint32_t pd, uint32_t al4, bool div_table,
                                         const uint32_t* __restrict__ tab, uint32_t* u, uint32_t* v) {
  const uint32_t* g = t.g2l + lane;
  const uint32_t ra = g[Px<ORDER>::r(pa) << 5], rb = g[Px<ORDER>::r(pb) << 5];
  const uint32_t rc = g[Px<ORDER>::r(pc) << 5], rd = g[Px<ORDER>::r(pd) << 5];
  const uint32_t ga = g[px_g(pa) << 5], gb = g[px_g(pb) << 5];
  const uint32_t gc = g[px_g(pc) << 5], gd = g[px_g(pd) << 5];
  const uint32_t ba = g[Px<ORDER>::b(pa) << 5], bb = g[Px<ORDER>::b(pb) << 5];
  const uint32_t bc = g[Px<ORDER>::b(pc) << 5], bd = g[Px<ORDER>::b(pd) << 5];
  uint32_t lr, lg, lb;
  if (!ALPHA || al4 == 0xffffffffu || al4 == 0u) {
    lr = ra + rb + rc + rd;  // SUM4 (SUM2 with shift=1 for a duplicated column)
    lg = ga + gb + gc + gd;
    lb = ba + bb + bc + bd;
  } else {
    const uint32_t a0 = al4 & 0xff, a1 = (al4 >> 8) & 0xff, a2 = (al4 >> 16) & 0xff, a3 = al4 >> 24;
    const uint32_t asum = a0 + a1 + a2 + a3;
    lr = divide_by_alpha(a0 * ra + a1 * rb + a2 * rc + a3 * rd, asum, div_table, tab);
    lg = divide_by_alpha(a0 * ga + a1 * gb + a2 * gc + a3 * gd, asum, div_table, tab);
    lb = divide_by_alpha(a0 * ba + a1 * bb + a2 * bc + a3 * bd, asum, div_table, tab);
  }
  const int r = linear_to_gamma(t, lr, lane);
  const int gg = linear_to_gamma(t, lg, lane);
  const int b = linear_to_gamma(t, lb, lane);
  // conversion.go:51-69.  r,g,b <= 1020 so both results stay inside [16,240]:
  // ClipUV's clamp can never fire and is dropped (tests/test_oracle_properties.py).
  const int kRnd = (1 << 17) + (128 << 18);
  *u = (uint32_t)((-9719 * r - 19081 * gg + 28800 * b + kRnd) >> 18);
  *v = (uint32_t)((28800 * r - 24116 * gg - 4684 * b + kRnd) >> 18);
}

__device__ __forceinline__ void store_bytes4(uint8_t* p, uint32_t w, int n) {
  for (int i = 0; i < n; ++i) p[i] = (uint8_t)(w >> (8 * i));
}

// Processes 4 px x 2 rows held in p0[4], p1[4] (already clamped/duplicated at
// the image edges).  nv = valid pixels in this strip (1..4).
template <int ORDER, bool ALPHA>
__device__ __forceinline__ void process_strip(const Tables& t, int lane, const ImportJobDev& job,
                                              const uint32_t (&p0)[4], const uint32_t (&p1)[4],
                                              int x0, int y0, int nv, bool has_row1, bool div_table,
                                              const uint32_t* __restrict__ tab, bool* nonopaque) {
  const uint32_t yp0 = pack_byte2(luma_sum<ORDER>(p0[0]), luma_sum<ORDER>(p0[1]),
                                  luma_sum<ORDER>(p0[2]), luma_sum<ORDER>(p0[3]));
  const uint32_t yp1 = pack_byte2(luma_sum<ORDER>(p1[0]), luma_sum<ORDER>(p1[1]),
                                  luma_sum<ORDER>(p1[2]), luma_sum<ORDER>(p1[3]));
  uint32_t ap0 = 0xffffffffu, ap1 = 0xffffffffu;
  if (ALPHA) {
    ap0 = pack_byte3(p0[0], p0[1], p0[2], p0[3]);
    ap1 = pack_byte3(p1[0], p1[1], p1[2], p1[3]);
    *nonopaque = *nonopaque || ((ap0 & ap1) != 0xffffffffu);
  }
  uint32_t u0, v0, u1, v1;
  block_uv<ORDER, ALPHA>(t, lane, p0[0], p0[1], p1[0], p1[1], __byte_perm(ap0, ap1, 0x5410),
                         div_table, tab, &u0, &v0);
  block_uv<ORDER, ALPHA>(t, lane, p0[2], p0[3], p1[2], p1[3], __byte_perm(ap0, ap1, 0x7632),
                         div_table, tab, &u1, &v1);

  uint8_t* yrow = job.y + (ptrdiff_t)y0 * job.y_stride + x0;
  const ptrdiff_t uvo = (ptrdiff_t)(y0 >> 1) * job.uv_stride + (x0 >> 1);
  if (nv == 4) {
    if (job.flags & JOB_Y_ALIGNED) {
      b2_stg_stream_u32(yrow, yp0);
      if (has_row1) b2_stg_stream_u32(yrow + job.y_stride, yp1);
    } else {
      store_bytes4(yrow, yp0, 4);
      if (has_row1) store_bytes4(yrow + job.y_stride, yp1, 4);
    }
    if (job.a != nullptr) {
      uint8_t* arow = job.a + (ptrdiff_t)y0 * job.a_stride + x0;
      if (job.flags & JOB_A_ALIGNED) {
        b2_stg_stream_u32(arow, ap0);
        if (has_row1) b2_stg_stream_u32(arow + job.a_stride, ap1);
      } else {
        store_bytes4(arow, ap0, 4);
        if (has_row1) store_bytes4(arow + job.a_stride, ap1, 4);
      }
    }
    if (job.flags & JOB_UV_ALIGNED) {
      b2_stg_stream_u16(job.u + uvo, u0 | (u1 << 8));
      b2_stg_stream_u16(job.v + uvo, v0 | (v1 << 8));
    } else {
      job.u[uvo] = (uint8_t)u0; job.u[uvo + 1] = (uint8_t)u1;
      job.v[uvo] = (uint8_t)v0; job.v[uvo + 1] = (uint8_t)v1;
    }
  } else {
    store_bytes4(yrow, yp0, nv);
    if (has_row1) store_bytes4(yrow + job.y_stride, yp1, nv);
    if (job.a != nullptr) {
      uint8_t* arow = job.a + (ptrdiff_t)y0 * job.a_stride + x0;
      store_bytes4(arow, ap0, nv);
      if (has_row1) store_bytes4(arow + job.a_stride, ap1, nv);
    }
    job.u[uvo] = (uint8_t)u0; job.v[uvo] = (uint8_t)v0;
    if (nv >= 3) { job.u[uvo + 1] = (uint8_t)u1; job.v[uvo + 1] = (uint8_t)v1; }
  }
}

// Fast kernel: 4-byte pixels whose rows are 4-byte aligned.
template <int ORDER, bool ALPHA>
__global__ void __launch_bounds__(kThreads, 2)
import_kernel(const ImportJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix,
              uint32_t n_jobs, uint32_t uniform, uint32_t total_items,
              const uint32_t* __restrict__ tab, int div_table, int32_t* __restrict__ has_alpha) {
  __shared__ Tables t;
  load_tables(t, tab);
  const int lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5;
  for (uint32_t item = blockIdx.x * kWarps + warp; item < total_items; item += gridDim.x * kWarps) {
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, item, &local);
    const ImportJobDev job = jobs[img];
    const uint32_t rp = local / job.tiles_x;
    const uint32_t tx = local - rp * job.tiles_x;
    const int x0 = (int)tx * kItemW + lane * 4;
    const int y0 = 2 * (int)rp;
    const int nv = min(4, job.width - x0);
    bool nonopaque = false;
    if (nv > 0) {
      const bool has_row1 = (y0 + 1 < job.height);
      const uint8_t* r0 = job.src + (ptrdiff_t)y0 * job.src_stride;
      const uint8_t* r1 = has_row1 ? r0 + job.src_stride : r0;  // yuv.go:377 (rgb_stride=0)
      uint32_t p0[4], p1[4];
      if (nv == 4) {
        const uint8_t* q0 = r0 + 4 * x0;
        const uint8_t* q1 = r1 + 4 * x0;
        if ((((uintptr_t)q0 | (uintptr_t)q1) & 15) == 0) {
          const uint4 a = b2_ldg_stream_v4(q0);
          const uint4 b = b2_ldg_stream_v4(q1);
          p0[0] = a.x; p0[1] = a.y; p0[2] = a.z; p0[3] = a.w;
          p1[0] = b.x; p1[1] = b.y; p1[2] = b.z; p1[3] = b.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            p0[i] = __ldg((const uint32_t*)q0 + i);
            p1[i] = __ldg((const uint32_t*)q1 + i);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int x = min(x0 + i, job.width - 1);  // duplicated last column (SUM2)
          p0[i] = __ldg((const uint32_t*)(r0 + 4 * x));
          p1[i] = __ldg((const uint32_t*)(r1 + 4 * x));
        }
      }
      process_strip<ORDER, ALPHA>(t, lane, job, p0, p1, x0, y0, nv, has_row1, div_table != 0, tab,
                                  &nonopaque);
    }
    if (ALPHA && has_alpha != nullptr) {
      if (__any_sync(0xffffffffu, nonopaque) && lane == 0) atomicOr(has_alpha + img, 1);
    }
  }
}

// Generic kernel: any step (3/4), any channel offsets, any alignment.  Pixels
// are gathered byte-wise into canonical R,G,B,A words.
template <bool ALPHA>
__global__ void __launch_bounds__(kThreads, 2)
import_generic_kernel(const ImportJobDev* __restrict__ jobs, const uint32_t* __restrict__ prefix,
                      uint32_t n_jobs, uint32_t uniform, uint32_t total_items,
                      const uint32_t* __restrict__ tab, int div_table, int32_t* __restrict__ has_alpha,
                      int step, int r_off, int g_off, int b_off, int a_off) {
  __shared__ Tables t;
  load_tables(t, tab);
  const int lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5;
  for (uint32_t item = blockIdx.x * kWarps + warp; item < total_items; item += gridDim.x * kWarps) {
    uint32_t local;
    const uint32_t img = b2_find_image(prefix, n_jobs, uniform, item, &local);
    ImportJobDev job = jobs[img];
    job.flags = 0;  // byte stores everywhere
    const uint32_t rp = local / job.tiles_x;
    const uint32_t tx = local - rp * job.tiles_x;
    const int x0 = (int)tx * kItemW + lane * 4;
    const int y0 = 2 * (int)rp;
    const int nv = min(4, job.width - x0);
    bool nonopaque = false;
    if (nv > 0) {
      const bool has_row1 = (y0 + 1 < job.height);
      const uint8_t* r0 = job.src + (ptrdiff_t)y0 * job.src_stride;
      const uint8_t* r1 = has_row1 ? r0 + job.src_stride : r0;
      uint32_t p0[4], p1[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int x = min(x0 + i, job.width - 1);
        const uint8_t* q0 = r0 + (ptrdiff_t)step * x;
        const uint8_t* q1 = r1 + (ptrdiff_t)step * x;
        p0[i] = q0[r_off] | (q0[g_off] << 8) | (q0[b_off] << 16) | ((ALPHA ? q0[a_off] : 0xffu) << 24);
        p1[i] = q1[r_off] | (q1[g_off] << 8) | (q1[b_off] << 16) | ((ALPHA ? q1[a_off] : 0xffu) << 24);
      }
      process_strip<0, ALPHA>(t, lane, job, p0, p1, x0, y0, nv, has_row1, div_table != 0, tab,
                              &nonopaque);
    }
    if (ALPHA && has_alpha != nullptr) {
      if (__any_sync(0xffffffffu, nonopaque) && lane == 0) atomicOr(has_alpha + img, 1);
    }
  }
}

}  // namespace

static int import_impl(b2_ctx* ctx, int stream, const b2_import_job* jobs, int n_jobs,
                       int order, int flags, int32_t* has_alpha_dev, float dithering) {
  B2_CHECK_CTX(ctx, stream);
  if (n_jobs < 0) return B2_ERR_INVALID_CONFIG;
  if (n_jobs == 0) return B2_OK;
  if (jobs == NULL) return B2_ERR_NULL;
  if (order < B2_ORDER_RGBA || order > B2_ORDER_BGR) return B2_ERR_INVALID_CONFIG;
  const bool alpha = (flags & B2_IMPORT_ALPHA) != 0;
  const int step = (order >= B2_ORDER_RGB) ? 3 : 4;
  if (alpha && step != 4) return B2_ERR_INVALID_CONFIG;  // conversions.go:70 assert(step == 4)

  const size_t jb = sizeof(ImportJobDev) * (size_t)n_jobs;
  const size_t pb = (sizeof(uint32_t) * ((size_t)n_jobs + 1) + 15) & ~(size_t)15;
  if (jb + 2 * pb + 64 > B2_DESC_RING_BYTES) return B2_ERR_INVALID_CONFIG;
  uint8_t* host = (uint8_t*)malloc(jb + 2 * pb + 64);
  if (host == NULL) return B2_ERR_OOM;
  ImportJobDev* dj = (ImportJobDev*)host;
  uint32_t* pre = (uint32_t*)(host + ((jb + 15) & ~(size_t)15));
  uint32_t* upre = (uint32_t*)((uint8_t*)pre + pb);  // prefix in row pairs (fast kernel)
  bool fast = (step == 4);
  bool tuned = true;          // b2_import_fast.cu preconditions (3-byte pixels: rows of any alignment)
  bool unaligned = false;     // some source row is not 16-byte aligned
  int n_with_a = 0;
  uint64_t total = 0, utotal = 0;
  uint32_t uniform = 0, uuniform = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const b2_import_job& j = jobs[i];
    if (j.src == NULL || j.y == NULL || j.u == NULL || j.v == NULL) { free(host); return B2_ERR_NULL; }
    if (j.width <= 0 || j.height <= 0 || j.width > B2_MAX_DIMENSION || j.height > B2_MAX_DIMENSION) {
      free(host); return B2_ERR_BAD_DIM;
    }
    const int64_t as = j.src_stride < 0 ? -(int64_t)j.src_stride : j.src_stride;
    if (as < (int64_t)(alpha ? 4 : 3) * j.width || as < (int64_t)step * j.width) {
      free(host); return B2_ERR_BAD_DIM;  // picture_csp_enc.c.go:132
    }
    ImportJobDev& d = dj[i];
    d.src = j.src; d.y = j.y; d.u = j.u; d.v = j.v; d.a = j.a;
    d.src_stride = j.src_stride; d.y_stride = j.y_stride; d.uv_stride = j.uv_stride;
    d.a_stride = j.a_stride; d.width = j.width; d.height = j.height;
    d.tiles_x = (uint32_t)((j.width + kItemW - 1) / kItemW);
    d.flags = 0;
    if ((((uintptr_t)j.y | (uintptr_t)(uint32_t)j.y_stride) & 3) == 0) d.flags |= JOB_Y_ALIGNED;
    if (j.a && (((uintptr_t)j.a | (uintptr_t)(uint32_t)j.a_stride) & 3) == 0) d.flags |= JOB_A_ALIGNED;
    if ((((uintptr_t)j.u | (uintptr_t)j.v | (uintptr_t)(uint32_t)j.uv_stride) & 1) == 0) d.flags |= JOB_UV_ALIGNED;
    if ((((uintptr_t)j.src | (uintptr_t)(uint32_t)j.src_stride) & 3) != 0) fast = false;
    if ((((uintptr_t)j.src | (uintptr_t)(uint32_t)j.src_stride) & 15) != 0) unaligned = true;
    if ((step == 4 && (((uintptr_t)j.src | (uintptr_t)(uint32_t)j.src_stride) & 3) != 0) ||
        !(d.flags & JOB_Y_ALIGNED) || !(d.flags & JOB_UV_ALIGNED) || (j.a && !(d.flags & JOB_A_ALIGNED))) tuned = false;
    n_with_a += (j.a != NULL);
    const uint32_t units = (uint32_t)((j.height + 1) >> 1);
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
  if (n_with_a != 0 && n_with_a != n_jobs) tuned = false;  // WRITE_A is a kernel-wide switch
  void* dev = NULL;
  const size_t pre_off = (size_t)((uint8_t*)pre - host);
  int st = B2_OK;
  if (dithering > 0.f) {
    // the dithered import uploads twice (this table, then the jump polynomials) before the kernels that
    // read both are launched: reserve room for the pair so the second upload cannot wrap onto the first
    const size_t poly_bytes3 = 3 * sizeof(uint32_t) * B2_RNG_N * (size_t)n_jobs;
    st = b2_desc_reserve(ctx, stream, ((pre_off + 2 * pb + 255) & ~(size_t)255) + ((poly_bytes3 + 255) & ~(size_t)255));
  }
  if (st == B2_OK) st = b2_upload_desc(ctx, stream, host, pre_off + 2 * pb, &dev);
  free(host);
  if (st != B2_OK) return st;
  const ImportJobDev* djobs = (const ImportJobDev*)dev;
  const uint32_t* dpre = (const uint32_t*)((uint8_t*)dev + pre_off);
  const uint32_t* dupre = (const uint32_t*)((uint8_t*)dev + pre_off + pb);
  cudaStream_t s = ctx->streams[stream];
  if (alpha && has_alpha_dev != NULL) {
    B2_CUDA(ctx, cudaMemsetAsync(has_alpha_dev, 0, sizeof(int32_t) * (size_t)n_jobs, s));
  } else if (has_alpha_dev != NULL) {
    B2_CUDA(ctx, cudaMemsetAsync(has_alpha_dev, 0, sizeof(int32_t) * (size_t)n_jobs, s));
  }
  const int grid = b2_grid_for(ctx, 2, (total + kWarps - 1) / kWarps);
  const int divt = (flags & B2_IMPORT_DIV_TABLE) ? 1 : 0;
  const uint32_t tot = (uint32_t)total;
  ImportDither dither = {nullptr, nullptr, 0};
  const int amp = (dithering > 0.f) ? b2_rng_amp(dithering) : 0;   // amp 0: every draw is exactly YUV_HALF
  if (amp > 0) {
    // Dithered import (conversions.go:122-165): only on the tuned kernel's preconditions
    // (4-byte pixels, 4-byte aligned rows and planes), which PictureARGBToYUVA always meets.
    if (!tuned) return B2_ERR_INVALID_CONFIG;
    const size_t poly_bytes = sizeof(uint32_t) * B2_RNG_N * (size_t)n_jobs;
    if (3 * poly_bytes > B2_DESC_RING_BYTES) return B2_ERR_INVALID_CONFIG;
    uint32_t* polys = (uint32_t*)malloc(3 * poly_bytes);
    if (polys == NULL) return B2_ERR_OOM;
    int last_w = -1, last_i = 0;
    uint32_t max_units = 0;
    for (int i = 0; i < n_jobs; ++i) {
      uint32_t* pk = polys + (size_t)i * B2_RNG_N;
      uint32_t* pw = polys + (size_t)(n_jobs + i) * B2_RNG_N;
      uint32_t* ps = polys + (size_t)(2 * n_jobs + i) * B2_RNG_N;
      const uint32_t units_i = (uint32_t)((jobs[i].height + 1) >> 1);
      if (units_i > max_units) max_units = units_i;
      if (jobs[i].width == last_w) {   // batches are mostly of one size
        memcpy(pk, polys + (size_t)last_i * B2_RNG_N, sizeof(uint32_t) * B2_RNG_N);
        memcpy(pw, polys + (size_t)(n_jobs + last_i) * B2_RNG_N, sizeof(uint32_t) * B2_RNG_N);
        memcpy(ps, polys + (size_t)(2 * n_jobs + last_i) * B2_RNG_N, sizeof(uint32_t) * B2_RNG_N);
        continue;
      }
      const uint64_t w = (uint64_t)jobs[i].width;
      const uint64_t k = 2 * w + 2 * ((w + 1) >> 1);
      b2_rng_jump_poly(k, pk);                                       // draws of one row pair
      b2_rng_jump_poly(w, pw);                                       // draws of one luma row
      b2_rng_jump_poly(k * (uint64_t)b2_rng_segment_row_pairs(), ps);  // draws of one segment of row pairs
      last_w = jobs[i].width; last_i = i;
    }
    void* pdev = NULL;
    st = b2_upload_desc(ctx, stream, polys, 3 * poly_bytes, &pdev);
    free(polys);
    if (st != B2_OK) return st;
    void* states = NULL;
    st = b2_stream_scratch(ctx, stream, sizeof(uint32_t) * B2_RNG_N * (size_t)utotal, &states);
    if (st != B2_OK) return st;
    B2_CUDA(ctx, b2_rng_states_launch(s, djobs, dupre, (uint32_t)n_jobs, (const uint32_t*)pdev,
                                      (const uint32_t*)pdev + 2 * (size_t)n_jobs * B2_RNG_N, max_units, (uint32_t*)states));
    ctx->launches += 2;
    dither.rng_states = (const uint32_t*)states;
    dither.rng_poly_w = (const uint32_t*)pdev + (size_t)n_jobs * B2_RNG_N;
    dither.amp = amp;
  }
  // row pairs per ticket: about 16 strips of 128 pixels (a 1080p row pair has 15), at most 8 row pairs
  uint32_t ticket_batch = 1;
  if (utotal > 0) {
    const uint64_t strips_per_unit = (total + utotal - 1) / utotal;   // `total` counts strips (tiles_x per row pair)
    if (strips_per_unit < 16) ticket_batch = (uint32_t)(16 / strips_per_unit);
    if (ticket_batch > 8) ticket_batch = 8;
    if (ticket_batch < 1) ticket_batch = 1;
  }
  if (tuned) {
    B2_CUDA(ctx, b2_import_fast_launch(ctx, s, djobs, dupre, (uint32_t)n_jobs, uuniform, (uint32_t)utotal,
                                       order, alpha, n_with_a != 0, divt, has_alpha_dev, dither,
                                       b2_tickets(ctx, stream, 0), unaligned, ticket_batch));
  } else if (fast) {
    if (order == B2_ORDER_RGBA) {
      if (alpha) import_kernel<0, true><<<grid, kThreads, 0, s>>>(djobs, dpre, n_jobs, uniform, tot, ctx->tab_dev, divt, has_alpha_dev);
      else import_kernel<0, false><<<grid, kThreads, 0, s>>>(djobs, dpre, n_jobs, uniform, tot, ctx->tab_dev, divt, has_alpha_dev);
    } else {
      if (alpha) import_kernel<1, true><<<grid, kThreads, 0, s>>>(djobs, dpre, n_jobs, uniform, tot, ctx->tab_dev, divt, has_alpha_dev);
      else import_kernel<1, false><<<grid, kThreads, 0, s>>>(djobs, dpre, n_jobs, uniform, tot, ctx->tab_dev, divt, has_alpha_dev);
    }
  } else {
    const bool swap = (order == B2_ORDER_BGRA || order == B2_ORDER_BGR);
    const int r_off = swap ? 2 : 0, b_off = swap ? 0 : 2;
    if (alpha) import_generic_kernel<true><<<grid, kThreads, 0, s>>>(djobs, dpre, n_jobs, uniform, tot, ctx->tab_dev, divt, has_alpha_dev, step, r_off, 1, b_off, 3);
    else import_generic_kernel<false><<<grid, kThreads, 0, s>>>(djobs, dpre, n_jobs, uniform, tot, ctx->tab_dev, divt, has_alpha_dev, step, r_off, 1, b_off, 3);
  }
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}

extern "C" int b2_rgba_to_yuv420a(b2_ctx* ctx, int stream, const b2_import_job* jobs, int n_jobs,
                                  int order, int flags, int32_t* has_alpha_dev) {
  return import_impl(ctx, stream, jobs, n_jobs, order, flags, has_alpha_dev, 0.f);
}

extern "C" int b2_rgba_to_yuv420a_dithered(b2_ctx* ctx, int stream, const b2_import_job* jobs, int n_jobs,
                                           int order, int flags, int32_t* has_alpha_dev, float dithering) {
  // 732 B of descriptors + jump polynomials per image must fit the 4 MB ring together: big batches go in
  // slices (every image has its own generator, so the slicing changes nothing)
  const int kSlice = 4096;
  if (jobs == NULL || n_jobs <= kSlice) return import_impl(ctx, stream, jobs, n_jobs, order, flags, has_alpha_dev, dithering);
  for (int i = 0; i < n_jobs; i += kSlice) {
    const int m = (n_jobs - i < kSlice) ? n_jobs - i : kSlice;
    const int st = import_impl(ctx, stream, jobs + i, m, order, flags, has_alpha_dev ? has_alpha_dev + i : NULL, dithering);
    if (st != B2_OK) return st;
  }
  return B2_OK;
}
