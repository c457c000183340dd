// b2_dither.cu -- VP8Random (pkg/vp8/random.go:13-95) for the dithered import
// (pkg/picture/conversions.go:122-165, pkg/libwebp/enc/picture_csp_enc.c.go:31-48).
//
// The reference runs ONE generator through the whole image: per row pair it draws `width`
// values for the first luma row, `width` for the second and 2 * ceil(width/2) for the
// chroma row (u, v per sample).  The generator is D. Knuth's subtractive lagged Fibonacci
// scheme: with s_n the value returned by call n and the table as s_-55 .. s_-1,
//     s_n = s_(n-55) - s_(n-24)   (mod 2^31),
// a LINEAR recurrence: the shift operator E satisfies E^55 + E^31 - 1 = 0, so for any k
//     s_(n+k) = sum_j c_j s_(n+j),  c = x^k mod (x^55 + x^31 - 1) over Z / 2^31.
// That is the skip-ahead: the host computes c for k = one row pair's draws (and for k =
// width, used by the import kernel to reach the second luma row and the chroma row), a
// pair of small kernels walks every image's row pairs (one 55 x 55 multiply-accumulate per
// jump, two levels) and leaves the generator state at the start of every row pair, and the import kernel
// regenerates its own draws from that state, 24 values per warp step (the shortest lag).
#include <stdlib.h>
#include <string.h>

#include "b2_import.cuh"

namespace {

const uint32_t kRandomTable[B2_RNG_N] = {   // random.go:17-32 (31-bit values)
    0x0de15230, 0x03b31886, 0x775faccb, 0x1c88626a, 0x68385c55, 0x14b3b828, 0x4a85fef8, 0x49ddb84b,
    0x64fcf397, 0x5c550289, 0x4a290000, 0x0d7ec1da, 0x5940b7ab, 0x5492577d, 0x4e19ca72, 0x38d38c69,
    0x0c01ee65, 0x32a1755f, 0x5437f652, 0x5abb2c32, 0x0faa57b1, 0x73f533e7, 0x685feeda, 0x7563cce2,
    0x6e990e83, 0x4730a7ed, 0x4fc0d9c6, 0x496b153c, 0x4f1403fa, 0x541afb0c, 0x73990b32, 0x26d7cb1c,
    0x6fcc3706, 0x2cbb77d8, 0x75762f2a, 0x6425ccdd, 0x24b35461, 0x0a7d8715, 0x220414a8, 0x141ebf67,
    0x56b41583, 0x73e502e3, 0x44cab16f, 0x28264d42, 0x73baaefb, 0x0a50ebed, 0x1d6ab6fb, 0x0d3ad40b,
    0x35db3b68, 0x2b081e83, 0x77ce6b95, 0x5181e5f0, 0x78853bbc, 0x009f9494, 0x27e5ed3c};

constexpr uint32_t kMask31 = 0x7fffffffu;

// a * b mod (x^55 + x^31 - 1), coefficients mod 2^31 (uint32 wrap-around keeps the low bits exact)
void poly_mulmod(const uint32_t* a, const uint32_t* b, uint32_t* out) {
  uint32_t t[2 * B2_RNG_N - 1];
  memset(t, 0, sizeof(t));
  for (int i = 0; i < B2_RNG_N; ++i) {
    if (a[i] == 0) continue;
    for (int j = 0; j < B2_RNG_N; ++j) t[i + j] += a[i] * b[j];
  }
  for (int d = 2 * B2_RNG_N - 2; d >= B2_RNG_N; --d) {   // x^55 = 1 - x^31
    const uint32_t c = t[d];
    t[d - B2_RNG_N] += c;
    t[d - B2_RNG_N + 31] -= c;
  }
  for (int i = 0; i < B2_RNG_N; ++i) out[i] = t[i] & kMask31;
}

}  // namespace

// c = x^k mod (x^55 + x^31 - 1) over Z / 2^31
void b2_rng_jump_poly(uint64_t k, uint32_t* c) {
  uint32_t result[B2_RNG_N] = {1}, base[B2_RNG_N] = {0, 1}, tmp[B2_RNG_N];
  while (k) {
    if (k & 1) { poly_mulmod(result, base, tmp); memcpy(result, tmp, sizeof(tmp)); }
    k >>= 1;
    if (k) { poly_mulmod(base, base, tmp); memcpy(base, tmp, sizeof(tmp)); }
  }
  memcpy(c, result, sizeof(result));
}

int b2_rng_amp(float dithering) {   // VP8InitRandom, random.go:49-63
  return (dithering < 0.f) ? 0 : (dithering > 1.f) ? (1 << 8) : (int)((1 << 8) * dithering);
}

namespace {

struct RngTable { uint32_t v[B2_RNG_N]; };

// States in two levels, so that an image's ~H/2 serial jumps become H/64 + 32:
//   level 1: one warp per image jumps kSeg row pairs at a time (poly_seg = x^(kSeg k)) and
//            leaves the state of every kSeg-th row pair;
//   level 2: one warp per (image, segment) walks the kSeg row pairs of its segment
//            (poly_k = x^k), starting from the level-1 state.
// A state = the 55 values that precede a row pair's first draw.
constexpr int kSeg = 32;

__global__ void __launch_bounds__(128)
rng_states_l1_kernel(const ImportJobDev* __restrict__ jobs, const uint32_t* __restrict__ unit_prefix, uint32_t n_jobs,
                     const uint32_t* __restrict__ poly_seg, uint32_t* __restrict__ states, const RngTable table) {
  __shared__ uint32_t ext_all[4][2 * B2_RNG_N], c_all[4][B2_RNG_N];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t img = blockIdx.x * 4 + warp;
  if (img >= n_jobs) return;
  uint32_t* ext = ext_all[warp];
  uint32_t* c = c_all[warp];
  for (int i = lane; i < B2_RNG_N; i += 32) {
    c[i] = poly_seg[(size_t)img * B2_RNG_N + i];
    ext[i] = table.v[i];
  }
  __syncwarp();
  const uint32_t units = (uint32_t)((jobs[img].height + 1) >> 1);
  uint32_t* out = states + (size_t)unit_prefix[img] * B2_RNG_N;
  for (uint32_t r = 0; r < units; r += kSeg) {
    for (int i = lane; i < B2_RNG_N; i += 32) out[(size_t)r * B2_RNG_N + i] = ext[i];
    if (r + kSeg >= units) break;
    b2_rng_jump_warp(ext, c, lane);
  }
}

// grid.x = segments of the tallest image, grid.y = image / 4
__global__ void __launch_bounds__(128)
rng_states_l2_kernel(const ImportJobDev* __restrict__ jobs, const uint32_t* __restrict__ unit_prefix, uint32_t n_jobs,
                     const uint32_t* __restrict__ poly_k, uint32_t* __restrict__ states) {
  __shared__ uint32_t ext_all[4][2 * B2_RNG_N], c_all[4][B2_RNG_N];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t img = blockIdx.y * 4 + warp;
  if (img >= n_jobs) return;
  const uint32_t units = (uint32_t)((jobs[img].height + 1) >> 1);
  const uint32_t r0 = blockIdx.x * kSeg;
  if (r0 >= units) return;
  uint32_t* ext = ext_all[warp];
  uint32_t* c = c_all[warp];
  uint32_t* out = states + (size_t)unit_prefix[img] * B2_RNG_N;
  for (int i = lane; i < B2_RNG_N; i += 32) {
    c[i] = poly_k[(size_t)img * B2_RNG_N + i];
    ext[i] = out[(size_t)r0 * B2_RNG_N + i];
  }
  __syncwarp();
  const uint32_t r1 = min(r0 + kSeg, units);
  for (uint32_t r = r0 + 1; r < r1; ++r) {
    b2_rng_jump_warp(ext, c, lane);
    for (int i = lane; i < B2_RNG_N; i += 32) out[(size_t)r * B2_RNG_N + i] = ext[i];
  }
}

}  // namespace

int b2_rng_segment_row_pairs() { return kSeg; }

cudaError_t b2_rng_states_launch(cudaStream_t s, const ImportJobDev* jobs, const uint32_t* unit_prefix, uint32_t n_jobs,
                                 const uint32_t* poly_k, const uint32_t* poly_seg, uint32_t max_units, uint32_t* states) {
  RngTable t;
  memcpy(t.v, kRandomTable, sizeof(t.v));
  rng_states_l1_kernel<<<(n_jobs + 3) / 4, 128, 0, s>>>(jobs, unit_prefix, n_jobs, poly_seg, states, t);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const dim3 grid((max_units + kSeg - 1) / kSeg, (n_jobs + 3) / 4);
  rng_states_l2_kernel<<<grid, 128, 0, s>>>(jobs, unit_prefix, n_jobs, poly_k, states);
  return cudaGetLastError();
}
