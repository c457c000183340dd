// Timeline of the gradient unfilter's band pipeline on ONE 1920x1080 plane (clock64 per band and block).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DB2_GRAD_TRACE -o tools/grad_trace tools/grad_trace.cu go_webp_b200/csrc/build/b2_ctx.o
#include "../go_webp_b200/csrc/b2_filters.cu"
#include <vector>
int main() {
  const int w = 1920, h = 1080;
  uint8_t *src, *dst;
  cudaMalloc(&src, w * h); cudaMalloc(&dst, w * h);
  std::vector<uint8_t> host(w * h);
  for (size_t i = 0; i < host.size(); ++i) host[i] = (uint8_t)(i * 2654435761u >> 24);
  cudaMemcpy(src, host.data(), w * h, cudaMemcpyHostToDevice);
  FiltJobDev j{src, dst, w, w, w, h, 3, 15};
  FiltJobDev* dj; cudaMalloc(&dj, sizeof(j)); cudaMemcpy(dj, &j, sizeof(j), cudaMemcpyHostToDevice);
  const int smem = kGradWarps * kGWarpWords * 4;
  cudaFuncSetAttribute(alpha_unfilter_gradient_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int it = 0; it < 3; ++it) alpha_unfilter_gradient_kernel<<<1, 32 * kGradWarps, smem>>>(dj, 1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<long long> tr(64 * 128);
  cudaMemcpyFromSymbol(tr.data(), g_grad_trace, tr.size() * 8);
  const long long t0 = tr[0];
  const int nb = (h + 31) / 32, nblk = ((w + 3) / 4 + 31 + 7) / 8;
  printf("bands %d blocks %d\n", nb, nblk);
  for (int b = 0; b < nb; ++b) {
    printf("band %2d start %8lld first-block-end %8lld end %8lld  blocks:", b, tr[b * 128] - t0, tr[b * 128 + 1] - t0, tr[b * 128 + nblk] - t0);
    for (int k = 1; k < nblk; k += 6) printf(" %lld", tr[b * 128 + 1 + k] - tr[b * 128 + k]);
    printf("\n");
  }
  return 0;
}
