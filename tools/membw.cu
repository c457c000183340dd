// membw.cu -- HBM microbenchmark: achievable bandwidth for different read:write
// mixes with streaming 128-bit accesses (what the pixel kernels issue).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o membw membw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int R, int W>  // each thread-iteration reads R uint4 and writes W uint4
__global__ void mix(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n_iter) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_iter; i += stride) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      uint4 v;
      const uint4* p = src + (size_t)r * n_iter + i;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
      acc.x ^= v.x; acc.y += v.y; acc.z ^= v.z; acc.w += v.w;
    }
#pragma unroll
    for (int w = 0; w < W; ++w) {
      uint4* p = dst + (size_t)w * n_iter + i;
      uint4 o = make_uint4(acc.x + w, acc.y, acc.z, acc.w ^ (uint32_t)i);
      asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
    }
  }
  if (W == 0 && acc.x == 0x12345678u && acc.y == 0x9abcdef0u) dst[0] = acc;  // keep the reads alive
}

template <int R, int W>
void run(const char* name, uint4* a, uint4* b, size_t n_iter, int sms) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int ctas : {2, 4, 8}) {
    for (int i = 0; i < 3; ++i) mix<R, W><<<sms * ctas, 512 / (ctas > 4 ? 2 : 1)>>>(a, b, n_iter);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int i = 0; i < reps; ++i) mix<R, W><<<sms * ctas, 512 / (ctas > 4 ? 2 : 1)>>>(a, b, n_iter);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)(R + W) * 16.0 * n_iter * reps;
    printf("%-14s ctas/SM=%d  %8.1f GB/s total  (read %7.1f, write %7.1f)\n", name, ctas, bytes / ms / 1e6,
           (double)R * 16 * n_iter * reps / ms / 1e6, (double)W * 16 * n_iter * reps / ms / 1e6);
  }
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const size_t n_iter = (size_t)1 << 24;  // 256 MiB per stream
  uint4 *a, *b;
  cudaMalloc(&a, n_iter * 16 * 4); cudaMalloc(&b, n_iter * 16 * 4);
  cudaMemset(a, 1, n_iter * 16 * 4);
  run<1, 0>("read only", a, b, n_iter * 4, p.multiProcessorCount);
  run<0, 1>("write only", a, b, n_iter * 4, p.multiProcessorCount);
  run<1, 1>("copy 1:1", a, b, n_iter * 2, p.multiProcessorCount);
  run<2, 1>("2:1", a, b, n_iter, p.multiProcessorCount);
  run<4, 1>("4:1", a, b, n_iter, p.multiProcessorCount);
  run<1, 2>("1:2", a, b, n_iter, p.multiProcessorCount);
  run<1, 3>("1:3 (upsample)", a, b, n_iter, p.multiProcessorCount);
  run<1, 4>("1:4 (index 2b)", a, b, n_iter, p.multiProcessorCount);
  printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
