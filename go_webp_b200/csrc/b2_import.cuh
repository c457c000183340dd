// b2_import.cuh -- descriptors shared by the general and the fast import kernels.
#pragma once
#include "b2_common.cuh"

struct ImportJobDev {
  const uint8_t* src;
  uint8_t* y;
  uint8_t* u;
  uint8_t* v;
  uint8_t* a;
  int32_t src_stride, y_stride, uv_stride, a_stride;
  int32_t width, height;
  uint32_t tiles_x;
  uint32_t flags;
};
enum { JOB_Y_ALIGNED = 1, JOB_A_ALIGNED = 2, JOB_UV_ALIGNED = 4 };

// Fast path (b2_import_fast.cu): every job is step 4, source rows 4-byte aligned
// (pointer and stride, any width), Y/A rows 4-byte and U/V rows 2-byte aligned.  unit_prefix counts
// row pairs.  tickets = the stream's two work-counter words (b2_tickets), zero between launches.
// unaligned = some job's source rows are not 16-byte aligned (pointer or stride).  ticket_batch = row pairs per ticket
// (>= 1; more than 1 for batches of narrow images).  Returns a cudaError_t
// from the launch.
// Dithered import (b2_dither.cu): rng_states = the VP8Random window (55 words) every row
// pair starts with, rng_poly_w = per image x^width mod P; amp = VP8Random.amp.  NULL
// rng_states = no dithering.
struct ImportDither {
  const uint32_t* rng_states;
  const uint32_t* rng_poly_w;
  int amp;
};
cudaError_t b2_import_fast_launch(const b2_ctx* ctx, cudaStream_t s, const ImportJobDev* jobs,
                                  const uint32_t* unit_prefix, uint32_t n_jobs,
                                  uint32_t uniform_units, uint32_t total_units, int order,
                                  bool alpha, bool write_a, int div_table, int32_t* has_alpha,
                                  ImportDither dither, uint32_t* tickets, bool unaligned, uint32_t ticket_batch);

// ---- VP8Random skip-ahead (b2_dither.cu) ---------------------------------------------
#define B2_RNG_N 55
// c = x^k mod (x^55 + x^31 - 1) over Z / 2^31 (host)
void b2_rng_jump_poly(uint64_t k, uint32_t* c);
int b2_rng_amp(float dithering);
// states[(unit_prefix[img] + r) * 55 ..] = generator window at the start of row pair r
// poly_k = x^k (k = draws of one row pair), poly_seg = x^(k * b2_rng_segment_row_pairs())
int b2_rng_segment_row_pairs();
cudaError_t b2_rng_states_launch(cudaStream_t s, const ImportJobDev* jobs, const uint32_t* unit_prefix, uint32_t n_jobs,
                                 const uint32_t* poly_k, const uint32_t* poly_seg, uint32_t max_units, uint32_t* states);
#ifdef __CUDACC__
// Warp-level skip-ahead.  ext[0..54] = the 55 values preceding the next draw (room for
// 110 words); c = x^k mod P.  On return ext[0..54] is the same window k draws later.
__device__ __forceinline__ void b2_rng_jump_warp(uint32_t* ext, const uint32_t* c, int lane) {
  // the next 54 draws, 24 at a time (the shortest lag): s_n = s_(n-55) - s_(n-24)
#pragma unroll
  for (int b = 0; b < 54; b += 24) {
    const int i = b + lane;
    if (lane < 24 && i < 54) ext[B2_RNG_N + i] = (ext[i] - ext[31 + i]) & 0x7fffffffu;
    __syncwarp();
  }
  uint32_t a0 = 0, a1 = 0;
  const int i1 = min(lane + 32, B2_RNG_N - 1);
  for (int j = 0; j < B2_RNG_N; ++j) {
    const uint32_t cj = c[j];
    a0 += cj * ext[lane + j];
    a1 += cj * ext[i1 + j];
  }
  __syncwarp();
  ext[lane] = a0 & 0x7fffffffu;
  if (lane + 32 < B2_RNG_N) ext[lane + 32] = a1 & 0x7fffffffu;
  __syncwarp();
}
#endif
