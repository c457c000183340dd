// b2_common.cuh -- context, error plumbing and small device helpers shared by
// the sm_100a kernels of the go-webp pixel hot path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/b2webp.h"

#define B2_MAX_STREAMS 16
#define B2_TICKET_WORDS 32
#define B2_EVENT_SLOTS 32
// Staging ring for batch descriptors (pinned host + device mirror) per stream.
#define B2_DESC_RING_BYTES (4u << 20)
// Events of the host-staged entry points (b2_staged.cu): one per chunk in flight.
#define B2_STAGED_EVENTS 24

struct b2_ctx {
  int device;
  int sm_count;
  int n_streams;
  cudaStream_t streams[B2_MAX_STREAMS];
  cudaEvent_t events[B2_EVENT_SLOTS];
  // descriptor ring
  uint8_t* desc_host[B2_MAX_STREAMS];
  uint8_t* desc_dev[B2_MAX_STREAMS];
  size_t desc_off[B2_MAX_STREAMS];
  // constant tables living in global memory (gamma, inverse alpha)
  uint32_t* tab_dev;
  // grow-only device scratch per stream (dithered import: generator states per row pair)
  void* scratch[B2_MAX_STREAMS];
  size_t scratch_bytes[B2_MAX_STREAMS];
  uint64_t launches;
  // work counters of the kernels that hand their units out by ticket: B2_TICKET_WORDS words per stream,
  // zero between launches (the last warp of a launch clears the two it used)
  uint32_t* tickets_dev;
  // byte order of the 16-bit output modes (RGBA_4444, RGB_565), library-wide: 1 = the go-webp text
  // (libwebp built with WEBP_SWAP_16BIT_CSP: ba / gb byte first), 0 = the stock libwebp build (rg first)
  int swap16;
  // copies issued through the ABI (b2_copy_stats): calls and bytes per direction
  uint64_t h2d_calls, h2d_bytes, d2h_calls, d2h_bytes;
  // host-staged entry points: has-alpha flags (device + pinned mirror), per-chunk events
  int32_t* flags_dev;
  int32_t* flags_host;
  size_t flags_cap;
  cudaEvent_t staged_ev[B2_STAGED_EVENTS];
  char err[256];
};

// Layout of b2_ctx::tab_dev (uint32 words)
#define B2_TAB_G2L 0      // 256 entries: kGammaToLinearTab
#define B2_TAB_L2G 256    // 32 x {base', d} pairs for LinearToGamma
#define B2_TAB_INVA 320   // 1021 entries: kInvAlpha (binary-compat divide)
#define B2_TAB_MAGIC 1344 // 1021 entries: exact-division magic  M | L << 27
#define B2_TAB_WORDS (1344 + 1024)

int b2_fail_cuda(b2_ctx* ctx, cudaError_t e, const char* what);

#define B2_CUDA(ctx, call)                                    \
  do {                                                        \
    cudaError_t e_ = (call);                                  \
    if (e_ != cudaSuccess) return b2_fail_cuda(ctx, e_, #call); \
  } while (0)

#define B2_CHECK_CTX(ctx, stream)                                        \
  do {                                                                   \
    if ((ctx) == NULL) return B2_ERR_NULL;                               \
    if ((stream) < 0 || (stream) >= (ctx)->n_streams) return B2_ERR_INVALID_CONFIG; \
    B2_CUDA(ctx, cudaSetDevice((ctx)->device));                          \
  } while (0)

// Device scratch of at least 'bytes' for work enqueued on 'stream' (valid until the next
// call for the same stream; growing it waits for the stream).
int b2_stream_scratch(b2_ctx* ctx, int stream, size_t bytes, void** dev_out);

// Makes sure the NEXT uploads of this stream, 'bytes' in total (every part rounded up to 256), fit the
// ring without wrapping in between: a launch that reads two uploads (job table + generator
// polynomials) must not have the second one wrap onto the first.  Wraps now if it has to.
int b2_desc_reserve(b2_ctx* ctx, int stream, size_t bytes);
// Copies 'bytes' of host descriptors into the stream's ring; returns device ptr.
int b2_upload_desc(b2_ctx* ctx, int stream, const void* host, size_t bytes, void** dev_out);

// Persistent-grid sizing: ctas_per_sm resident CTAs on every SM, capped by work.
static inline uint32_t* b2_tickets(const b2_ctx* ctx, int stream, int slot) {
  return ctx->tickets_dev + (size_t)stream * B2_TICKET_WORDS + 2 * slot;
}
static inline int b2_grid_for(const b2_ctx* ctx, int ctas_per_sm, uint64_t units_of_work) {
  uint64_t g = (uint64_t)ctx->sm_count * (uint64_t)ctas_per_sm;
  if (units_of_work < g) g = units_of_work;
  if (g < 1) g = 1;
  return (int)g;
}

#ifdef __CUDACC__
// ---- device helpers --------------------------------------------------------
__device__ __forceinline__ uint32_t b2_lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
// streaming (read-once) 128/64/32-bit loads and stores: keep L1 for the halo data
__device__ __forceinline__ uint4 b2_ldg_stream_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// same without .nc: legal when the kernel also writes the buffer (in-place ops)
__device__ __forceinline__ uint4 b2_ld_stream_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ uint2 b2_ldg_stream_v2(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];"
               : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t b2_ldg_stream_u32(const void* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void b2_stg_stream_v4(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void b2_stg_stream_v2(void* p, uint2 v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};"
               :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void b2_stg_stream_u32(void* p, uint32_t v) {
  asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void b2_stg_stream_u16(void* p, uint32_t v) {
  asm volatile("st.global.L1::no_allocate.u16 [%0], %1;" :: "l"(p), "h"((unsigned short)v) : "memory");
}

// Finds the image a flat work item belongs to.  prefix[i] = first item of image
// i, prefix[n] = total.  uniform > 0: every image has 'uniform' items.
__device__ __forceinline__ uint32_t b2_find_image(const uint32_t* __restrict__ prefix,
                                                  uint32_t n, uint32_t uniform,
                                                  uint32_t item, uint32_t* local) {
  if (uniform) {
    const uint32_t img = item / uniform;
    *local = item - img * uniform;
    return img;
  }
  uint32_t lo = 0, hi = n;  // invariant: prefix[lo] <= item < prefix[hi]
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(prefix + mid) <= item) lo = mid; else hi = mid;
  }
  *local = item - __ldg(prefix + lo);
  return lo;
}
// The same for an item the whole warp looks up together (all lanes call it with the same
// item): a 32-way search, two dependent loads for up to 1024 images instead of ten.
__device__ __forceinline__ uint32_t b2_find_image_warp(const uint32_t* __restrict__ prefix,
                                                       uint32_t n, uint32_t uniform,
                                                       uint32_t item, uint32_t* local, uint32_t lane) {
  if (uniform) {
    const uint32_t img = item / uniform;
    *local = item - img * uniform;
    return img;
  }
  uint32_t lo = 0, span = n;  // invariant: prefix[lo] <= item < prefix[lo + span]
  while (span > 1) {
    const uint32_t step = (span + 31) >> 5;
    const uint32_t idx = lo + (lane + 1) * step;   // first image after this lane's share
    const bool le = (idx < lo + span) && (__ldg(prefix + idx) <= item);
    const uint32_t k = (uint32_t)__popc(__ballot_sync(0xffffffffu, le));
    const uint32_t nlo = lo + k * step;
    span = min(step, lo + span - nlo);
    lo = nlo;
  }
  *local = item - __ldg(prefix + lo);
  return lo;
}
#endif  // __CUDACC__
