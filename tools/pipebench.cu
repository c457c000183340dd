// pipebench.cu -- per-SM issue throughput of the integer instructions the pixel
// kernels are built from (sm_100a).  Each test runs N_CHAIN independent dependency
// chains per thread, 1024 threads per SM, and reports warp-lanes retired per SM
// clock (128 = one instruction per scheduler per clock).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipebench pipebench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

constexpr int kChains = 8;
constexpr int kIters = 2048;

enum Op { IMAD, IMADHI, IMADWIDE, IADD3, LOP3, SHF, PRMT, I2IP, VIMNMX16, VIADDMNMX16, VIMNMX3, DP4A, DP2A, FFMA, FFMAIMM, HFMA2,
          FMNMX, I2F, F2I, LEAHI, MIX_IMAD_IADD, MIX_IMADHI_IADD, MIX_FFMA_IADD, MIX_FFMA_IMAD, MIX3, LDS32, LDS64, MIX_LDS_ALU,
          I2F8, IABS, POPC, SEL, N_OPS };
const char* kNames[N_OPS] = {"IMAD", "IMAD.HI", "IMAD.WIDE", "IADD3", "LOP3", "SHF", "PRMT", "I2IP(cvt.pack.sat)", "VIMNMX.S16x2.relu",
                             "VIADDMNMX.S16x2.relu", "VIMNMX3.s32", "IDP.4A", "IDP.2A", "FFMA reg", "FFMA imm", "HFMA2", "FMNMX", "I2F.S32",
                             "F2I", "LEA.HI ((a>>3)+b)", "mix IMAD+IADD3", "mix IMAD.HI+IADD3", "mix FFMA+IADD3", "mix FFMA+IMAD",
                             "mix FFMA+IMAD+IADD3+LOP3", "LDS.32", "LDS.64", "mix LDS.32+3 ALU", "I2F.U8 (byte sel)", "IABS", "POPC", "SEL"};

template <int OP>
__global__ void __launch_bounds__(1024, 1) bench(uint32_t* out, uint32_t seed, long long* cycles) {
  __shared__ uint32_t sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i * seed;
  __syncthreads();
  uint32_t x[kChains];
  float f[kChains];
#pragma unroll
  for (int k = 0; k < kChains; ++k) { x[k] = seed * (threadIdx.x + 1) + k * 77u; f[k] = (float)(x[k] & 1023) * 0.001f; }
  const uint32_t c1 = seed | 19077u, c2 = seed ^ 0x00ff00ffu;
  const float fc = (float)seed * 1e-9f + 0.999f;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int k = 0; k < kChains; ++k) {
      if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(c1), "r"(c2));
      if (OP == IMADHI) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(c1));
      if (OP == IMADWIDE) { uint64_t w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(x[k]), "r"(c1)); x[k] = (uint32_t)w ^ (uint32_t)(w >> 32); }
      if (OP == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(c1));
      if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(c1), "r"(c2));
      if (OP == SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(x[k]) : "r"(c1));
      if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x2714;" : "+r"(x[k]) : "r"(c1));
      if (OP == I2IP) asm volatile("cvt.pack.sat.u8.s32.b32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(c1), "r"(c2));
      if (OP == VIMNMX16) x[k] = __vimin_s16x2_relu(x[k], c1);
      if (OP == VIADDMNMX16) x[k] = __viaddmin_s16x2_relu(x[k], c1, c2);
      if (OP == VIMNMX3) x[k] = (uint32_t)__vimin3_s32((int)x[k], (int)c1, (int)c2);
      if (OP == DP4A) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(c1), "r"(c2));
      if (OP == DP2A) asm volatile("dp2a.lo.u32.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(c1), "r"(c2));
      if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[k]) : "f"(fc), "f"(fc));
      if (OP == FFMAIMM) asm volatile("fma.rn.f32 %0, %0, 0f3F7FF000, %1;" : "+f"(f[k]) : "f"(fc));
      if (OP == HFMA2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(c1), "r"(c2));
      if (OP == FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(f[k]) : "f"(fc));
      if (OP == I2F) { asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f[k]) : "r"(x[k])); x[k] = __float_as_uint(f[k]); }
      if (OP == I2F8) { asm volatile("{ .reg .b8 t0,t1,t2,t3; mov.b32 {t0,t1,t2,t3}, %1; cvt.rn.f32.u8 %0, t2; }" : "=f"(f[k]) : "r"(x[k])); x[k] = __float_as_uint(f[k]); }
      if (OP == F2I) { int t; asm volatile("cvt.rmi.s32.f32 %0, %1;" : "=r"(t) : "f"(f[k])); f[k] = __int_as_float(t | 0x3f000000); }
      if (OP == LEAHI) x[k] = (x[k] >> 3) + c1;
      if (OP == IABS) asm volatile("abs.s32 %0, %0;" : "+r"(x[k]));
      if (OP == POPC) asm volatile("popc.b32 %0, %0;" : "+r"(x[k]));
      if (OP == SEL) asm volatile("{ .reg .pred p; setp.lt.s32 p, %0, %1; selp.b32 %0, %1, %2, p; }" : "+r"(x[k]) : "r"(c1), "r"(c2));
      if (OP == MIX_IMAD_IADD) {
        if (k & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(c1), "r"(c2));
        else asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(c1));
      }
      if (OP == MIX_IMADHI_IADD) {
        if (k & 1) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(c1));
        else asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(c1));
      }
      if (OP == MIX_FFMA_IADD) {
        if (k & 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[k]) : "f"(fc), "f"(fc));
        else asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(c1));
      }
      if (OP == MIX_FFMA_IMAD) {
        if (k & 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[k]) : "f"(fc), "f"(fc));
        else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(c1), "r"(c2));
      }
      if (OP == MIX3) {
        if ((k & 3) == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[k]) : "f"(fc), "f"(fc));
        if ((k & 3) == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(c1), "r"(c2));
        if ((k & 3) == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(c1));
        if ((k & 3) == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(c1), "r"(c2));
      }
      if (OP == LDS32) { uint32_t a = (uint32_t)__cvta_generic_to_shared(sm) + ((x[k] * 4u) & 0x1ffcu & ~0x7cu) + (threadIdx.x & 31) * 4;
                         asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x[k]) : "r"(a)); }
      if (OP == LDS64) { uint32_t a = (uint32_t)__cvta_generic_to_shared(sm) + ((x[k] * 8u) & 0x1f00u) + (threadIdx.x & 31) * 8; uint32_t hi;
                         asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x[k]), "=r"(hi) : "r"(a)); x[k] ^= hi; }
      if (OP == MIX_LDS_ALU) {
        if ((k & 3) == 0) { uint32_t a = (uint32_t)__cvta_generic_to_shared(sm) + (x[k] & 0x1f80u) + (threadIdx.x & 31) * 4;
                            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x[k]) : "r"(a)); }
        else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(c1), "r"(c2));
      }
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int k = 0; k < kChains; ++k) acc ^= x[k] ^ __float_as_uint(f[k]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP> void run(uint32_t* out, long long* cyc, int sms) {
  bench<OP><<<sms, 1024>>>(out, 12345u, cyc);
  bench<OP><<<sms, 1024>>>(out, 12345u, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-28s error %s\n", kNames[OP], cudaGetErrorString(e)); return; }
  long long h[256];
  cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < sms; ++i) avg += (double)h[i];
  avg /= sms;
  const double lane_ops = 1024.0 * kChains * kIters;  // per SM, counting one "test step" per chain
  printf("%-28s %7.1f lane-steps/clk/SM   (%.0f cycles)\n", kNames[OP], lane_ops / avg, avg);
}

template <int OP> struct Runner { static void go(uint32_t* out, long long* cyc, int sms) { run<OP>(out, cyc, sms); Runner<OP + 1>::go(out, cyc, sms); } };
template <> struct Runner<N_OPS> { static void go(uint32_t*, long long*, int) {} };

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, sizeof(uint32_t) * 1024 * p.multiProcessorCount);
  cudaMalloc(&cyc, sizeof(long long) * p.multiProcessorCount);
  printf("%s, %d SMs; a 'step' is the listed instruction (mixes: one of the listed per step)\n", p.name, p.multiProcessorCount);
  Runner<0>::go(out, cyc, p.multiProcessorCount);
  return 0;
}
