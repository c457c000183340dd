// b2_ctx.cu -- context, memory, copies, sync, timing and synthetic-input half of
// the C ABI (include/b2webp.h).  Kernels live in b2_import.cu, b2_upsample.cu and
// b2_vp8l.cu.
#include <math.h>
#include <stdlib.h>

#include <chrono>
#include <thread>

#include "b2_common.cuh"

int b2_fail_cuda(b2_ctx* ctx, cudaError_t e, const char* what) {
  if (ctx != NULL) {
    snprintf(ctx->err, sizeof(ctx->err), "%s: %s (%s)", what, cudaGetErrorName(e),
             cudaGetErrorString(e));
  }
  // keep the sticky-error state readable but do not clear real faults
  (void)cudaGetLastError();
  return (e == cudaErrorMemoryAllocation) ? B2_ERR_OOM : B2_ERR_CUDA;
}

extern "C" int b2_abi_version(void) { return B2_ABI_VERSION; }

extern "C" const char* b2_strerror(int status) {
  switch (status) {
    case B2_OK: return "ok";
    case B2_ERR_OOM: return "memory error allocating objects";
    case B2_ERR_BAD_DIM: return "picture has invalid width/height";
    case B2_ERR_NULL: return "a pointer parameter is nil";
    case B2_ERR_INVALID_CONFIG: return "configuration is invalid";
    case B2_ERR_CUDA: return "CUDA runtime or device error";
    default: return "unknown status";
  }
}

extern "C" int b2_device_count(int* count) {
  if (count == NULL) return B2_ERR_NULL;
  // A box that has just come up can answer cudaErrorSystemNotReady / NoDevice for a
  // moment while the driver's persistence daemon starts: ask again for up to 3 s
  // before reporting that there is no device.
  int n = 0;
  cudaError_t e = cudaSuccess;
  for (int attempt = 0; attempt < 10; ++attempt) {
    n = 0;
    e = cudaGetDeviceCount(&n);
    if (e == cudaSuccess && n > 0) break;
    if (e == cudaErrorInsufficientDriver) break;  // no driver at all: nothing to wait for
    (void)cudaGetLastError();
    std::this_thread::sleep_for(std::chrono::milliseconds(300));
  }
  if (e != cudaSuccess) { *count = 0; return B2_ERR_CUDA; }
  *count = n;
  return B2_OK;
}

// Gamma tables of pkg/libwebp/dsp/yuv.go:155-181 (C-intended float division,
// SURVEY Appendix B-1), generated on the host with the exact pow() expression,
// plus libwebp's kInvAlpha for the binary-compatible divide mode.
static void build_tables(uint32_t* tab) {
  const int kGammaFix = 12, kGammaTabFix = 7;
  const int kGammaScale = (1 << kGammaFix) - 1;
  const double kGamma = 0.80;
  const double scale = (double)(1 << kGammaTabFix) / kGammaScale;
  const double norm = 1. / 255.;
  int l2g[33];
  memset(tab, 0, sizeof(uint32_t) * B2_TAB_WORDS);
  for (int v = 0; v <= 255; ++v) {
    tab[B2_TAB_G2L + v] = (uint16_t)(pow(norm * v, kGamma) * kGammaScale + .5);
  }
  for (int v = 0; v <= 32; ++v) l2g[v] = (int)(255. * pow(scale * v, 1. / kGamma) + .5);
  for (int pos = 0; pos < 32; ++pos) {
    const int d = l2g[pos + 1] - l2g[pos];
    const int base = l2g[pos] * 512 + 64 - d * pos * 512;
    tab[B2_TAB_L2G + 2 * pos] = (uint32_t)base;
    tab[B2_TAB_L2G + 2 * pos + 1] = (uint32_t)d;
  }
  for (int a = 1; a <= 4 * 255; ++a) tab[B2_TAB_INVA + a] = (1u << 19) / (uint32_t)a;
  // Exact n / a for n < 2^24 (Granlund-Montgomery): L = ceil(log2 a),
  // M = floor(2^(24+L) / a) + 1  =>  n / a == umulhi(n << 8, M) >> L.
  // (n = 4 * sum of alpha-weighted linear values <= 4 * 1020 * 4095 < 2^24.)
  for (uint32_t a = 1; a <= 4 * 255; ++a) {
    uint32_t L = 0;
    while ((1u << L) < a) ++L;
    const uint64_t M = ((uint64_t)1 << (24 + L)) / a + 1;
    tab[B2_TAB_MAGIC + a] = (uint32_t)M | (L << 27);
  }
}

extern "C" int b2_ctx_create(int device, int n_streams, b2_ctx** out) {
  if (out == NULL) return B2_ERR_NULL;
  *out = NULL;
  if (n_streams < 1 || n_streams > B2_MAX_STREAMS) return B2_ERR_INVALID_CONFIG;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return B2_ERR_CUDA; }
  if (device < 0 || device >= n) return B2_ERR_INVALID_CONFIG;
  b2_ctx* ctx = (b2_ctx*)calloc(1, sizeof(b2_ctx));
  if (ctx == NULL) return B2_ERR_OOM;
  ctx->device = device;
  ctx->n_streams = n_streams;
  ctx->swap16 = 1;
  int st = B2_OK;
  do {
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) { st = b2_fail_cuda(ctx, e, "cudaSetDevice"); break; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { st = b2_fail_cuda(ctx, e, "props"); break; }
    ctx->sm_count = prop.multiProcessorCount;
    for (int i = 0; i < n_streams && st == B2_OK; ++i) {
      if ((e = cudaStreamCreateWithFlags(&ctx->streams[i], cudaStreamNonBlocking)) != cudaSuccess) st = b2_fail_cuda(ctx, e, "stream");
      if (st == B2_OK && (e = cudaHostAlloc((void**)&ctx->desc_host[i], B2_DESC_RING_BYTES, cudaHostAllocDefault)) != cudaSuccess) st = b2_fail_cuda(ctx, e, "desc host");
      if (st == B2_OK && (e = cudaMalloc((void**)&ctx->desc_dev[i], B2_DESC_RING_BYTES)) != cudaSuccess) st = b2_fail_cuda(ctx, e, "desc dev");
    }
    if (st != B2_OK) break;
    for (int i = 0; i < B2_EVENT_SLOTS && st == B2_OK; ++i) {
      if ((e = cudaEventCreate(&ctx->events[i])) != cudaSuccess) st = b2_fail_cuda(ctx, e, "event");
    }
    if (st != B2_OK) break;
    uint32_t* tab = (uint32_t*)malloc(sizeof(uint32_t) * B2_TAB_WORDS);
    if (tab == NULL) { st = B2_ERR_OOM; break; }
    build_tables(tab);
    if ((e = cudaMalloc((void**)&ctx->tab_dev, sizeof(uint32_t) * B2_TAB_WORDS)) != cudaSuccess) {
      free(tab); st = b2_fail_cuda(ctx, e, "tab"); break;
    }
    e = cudaMemcpy(ctx->tab_dev, tab, sizeof(uint32_t) * B2_TAB_WORDS, cudaMemcpyHostToDevice);
    free(tab);
    if (e != cudaSuccess) { st = b2_fail_cuda(ctx, e, "tab copy"); break; }
    const size_t tb = sizeof(uint32_t) * B2_TICKET_WORDS * B2_MAX_STREAMS;
    if ((e = cudaMalloc((void**)&ctx->tickets_dev, tb)) != cudaSuccess) { st = b2_fail_cuda(ctx, e, "tickets"); break; }
    if ((e = cudaMemset(ctx->tickets_dev, 0, tb)) != cudaSuccess) { st = b2_fail_cuda(ctx, e, "tickets clear"); break; }
  } while (0);
  if (st != B2_OK) { b2_ctx_destroy(ctx); return st; }
  *out = ctx;
  return B2_OK;
}

extern "C" void b2_ctx_destroy(b2_ctx* ctx) {
  if (ctx == NULL) return;
  (void)cudaSetDevice(ctx->device);
  for (int i = 0; i < ctx->n_streams; ++i) {
    if (ctx->streams[i]) { (void)cudaStreamSynchronize(ctx->streams[i]); (void)cudaStreamDestroy(ctx->streams[i]); }
    if (ctx->desc_host[i]) (void)cudaFreeHost(ctx->desc_host[i]);
    if (ctx->desc_dev[i]) (void)cudaFree(ctx->desc_dev[i]);
    if (ctx->scratch[i]) (void)cudaFree(ctx->scratch[i]);
  }
  for (int i = 0; i < B2_EVENT_SLOTS; ++i) if (ctx->events[i]) (void)cudaEventDestroy(ctx->events[i]);
  if (ctx->tab_dev) (void)cudaFree(ctx->tab_dev);
  if (ctx->tickets_dev) (void)cudaFree(ctx->tickets_dev);
  if (ctx->flags_dev) (void)cudaFree(ctx->flags_dev);
  if (ctx->flags_host) (void)cudaFreeHost(ctx->flags_host);
  for (int i = 0; i < B2_STAGED_EVENTS; ++i) if (ctx->staged_ev[i]) (void)cudaEventDestroy(ctx->staged_ev[i]);
  free(ctx);
}

extern "C" int b2_ctx_set_swap_16bit_csp(b2_ctx* ctx, int swap) {
  if (ctx == NULL) return B2_ERR_NULL;
  if (swap != 0 && swap != 1) return B2_ERR_INVALID_CONFIG;
  ctx->swap16 = swap;
  return B2_OK;
}
extern "C" int b2_ctx_swap_16bit_csp(const b2_ctx* ctx) { return ctx ? ctx->swap16 : -1; }
extern "C" int b2_ctx_device(const b2_ctx* ctx) { return ctx ? ctx->device : -1; }
extern "C" int b2_ctx_sm_count(const b2_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
extern "C" const char* b2_last_error(const b2_ctx* ctx) { return ctx ? ctx->err : "nil context"; }
extern "C" uint64_t b2_launch_count(const b2_ctx* ctx) { return ctx ? ctx->launches : 0; }

int b2_stream_scratch(b2_ctx* ctx, int stream, size_t bytes, void** dev_out) {
  if (bytes > ctx->scratch_bytes[stream]) {
    B2_CUDA(ctx, cudaStreamSynchronize(ctx->streams[stream]));
    if (ctx->scratch[stream]) { B2_CUDA(ctx, cudaFree(ctx->scratch[stream])); ctx->scratch[stream] = NULL; ctx->scratch_bytes[stream] = 0; }
    const size_t cap = bytes + bytes / 4 + (1u << 20);
    B2_CUDA(ctx, cudaMalloc(&ctx->scratch[stream], cap));
    ctx->scratch_bytes[stream] = cap;
  }
  *dev_out = ctx->scratch[stream];
  return B2_OK;
}

int b2_desc_reserve(b2_ctx* ctx, int stream, size_t bytes) {
  if (bytes > B2_DESC_RING_BYTES) return B2_ERR_INVALID_CONFIG;
  if (ctx->desc_off[stream] + bytes > B2_DESC_RING_BYTES) {
    B2_CUDA(ctx, cudaStreamSynchronize(ctx->streams[stream]));
    ctx->desc_off[stream] = 0;
  }
  return B2_OK;
}

int b2_upload_desc(b2_ctx* ctx, int stream, const void* host, size_t bytes, void** dev_out) {
  const size_t need = (bytes + 255) & ~(size_t)255;
  if (need > B2_DESC_RING_BYTES) return B2_ERR_INVALID_CONFIG;
  if (ctx->desc_off[stream] + need > B2_DESC_RING_BYTES) {
    // wrap: everything enqueued earlier on this stream must have consumed its descriptors
    B2_CUDA(ctx, cudaStreamSynchronize(ctx->streams[stream]));
    ctx->desc_off[stream] = 0;
  }
  const size_t off = ctx->desc_off[stream];
  memcpy(ctx->desc_host[stream] + off, host, bytes);
  B2_CUDA(ctx, cudaMemcpyAsync(ctx->desc_dev[stream] + off, ctx->desc_host[stream] + off, bytes,
                               cudaMemcpyHostToDevice, ctx->streams[stream]));
  ctx->desc_off[stream] = off + need;
  *dev_out = ctx->desc_dev[stream] + off;
  return B2_OK;
}

// ---- memory ------------------------------------------------------------------
extern "C" int b2_pinned_alloc(b2_ctx* ctx, size_t bytes, void** host_ptr) {
  if (ctx == NULL || host_ptr == NULL) return B2_ERR_NULL;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  *host_ptr = NULL;
  B2_CUDA(ctx, cudaHostAlloc(host_ptr, bytes ? bytes : 1, cudaHostAllocPortable));
  return B2_OK;
}
extern "C" int b2_pinned_free(b2_ctx* ctx, void* host_ptr) {
  if (ctx == NULL) return B2_ERR_NULL;
  if (host_ptr == NULL) return B2_OK;
  B2_CUDA(ctx, cudaFreeHost(host_ptr));
  return B2_OK;
}
extern "C" int b2_host_register(b2_ctx* ctx, void* host_ptr, size_t bytes) {
  if (ctx == NULL || host_ptr == NULL) return B2_ERR_NULL;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  B2_CUDA(ctx, cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable));
  return B2_OK;
}
extern "C" int b2_host_unregister(b2_ctx* ctx, void* host_ptr) {
  if (ctx == NULL || host_ptr == NULL) return B2_ERR_NULL;
  B2_CUDA(ctx, cudaHostUnregister(host_ptr));
  return B2_OK;
}
extern "C" int b2_dev_alloc(b2_ctx* ctx, size_t bytes, void** dev_ptr) {
  if (ctx == NULL || dev_ptr == NULL) return B2_ERR_NULL;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  *dev_ptr = NULL;
  B2_CUDA(ctx, cudaMalloc(dev_ptr, bytes ? bytes : 1));
  return B2_OK;
}
extern "C" int b2_dev_free(b2_ctx* ctx, void* dev_ptr) {
  if (ctx == NULL) return B2_ERR_NULL;
  if (dev_ptr == NULL) return B2_OK;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  B2_CUDA(ctx, cudaFree(dev_ptr));
  return B2_OK;
}
extern "C" int b2_h2d_async(b2_ctx* ctx, int stream, void* dst_dev, const void* src_host, size_t bytes) {
  B2_CHECK_CTX(ctx, stream);
  if (bytes == 0) return B2_OK;
  if (dst_dev == NULL || src_host == NULL) return B2_ERR_NULL;
  B2_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->streams[stream]));
  ctx->h2d_calls++; ctx->h2d_bytes += bytes;
  return B2_OK;
}
extern "C" int b2_d2h_async(b2_ctx* ctx, int stream, void* dst_host, const void* src_dev, size_t bytes) {
  B2_CHECK_CTX(ctx, stream);
  if (bytes == 0) return B2_OK;
  if (dst_host == NULL || src_dev == NULL) return B2_ERR_NULL;
  B2_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->streams[stream]));
  ctx->d2h_calls++; ctx->d2h_bytes += bytes;
  return B2_OK;
}
extern "C" int b2_h2d_2d_async(b2_ctx* ctx, int stream, void* dst_dev, size_t dst_pitch,
                               const void* src_host, size_t src_pitch, size_t width_bytes, size_t rows) {
  B2_CHECK_CTX(ctx, stream);
  if (width_bytes == 0 || rows == 0) return B2_OK;
  if (dst_dev == NULL || src_host == NULL) return B2_ERR_NULL;
  B2_CUDA(ctx, cudaMemcpy2DAsync(dst_dev, dst_pitch, src_host, src_pitch, width_bytes, rows,
                                 cudaMemcpyHostToDevice, ctx->streams[stream]));
  ctx->h2d_calls++; ctx->h2d_bytes += width_bytes * rows;
  return B2_OK;
}
extern "C" int b2_d2h_2d_async(b2_ctx* ctx, int stream, void* dst_host, size_t dst_pitch,
                               const void* src_dev, size_t src_pitch, size_t width_bytes, size_t rows) {
  B2_CHECK_CTX(ctx, stream);
  if (width_bytes == 0 || rows == 0) return B2_OK;
  if (dst_host == NULL || src_dev == NULL) return B2_ERR_NULL;
  B2_CUDA(ctx, cudaMemcpy2DAsync(dst_host, dst_pitch, src_dev, src_pitch, width_bytes, rows,
                                 cudaMemcpyDeviceToHost, ctx->streams[stream]));
  ctx->d2h_calls++; ctx->d2h_bytes += width_bytes * rows;
  return B2_OK;
}
extern "C" int b2_memset_async(b2_ctx* ctx, int stream, void* dst_dev, int value, size_t bytes) {
  B2_CHECK_CTX(ctx, stream);
  if (bytes == 0) return B2_OK;
  if (dst_dev == NULL) return B2_ERR_NULL;
  B2_CUDA(ctx, cudaMemsetAsync(dst_dev, value, bytes, ctx->streams[stream]));
  return B2_OK;
}
extern "C" int b2_sync(b2_ctx* ctx, int stream) {
  if (ctx == NULL) return B2_ERR_NULL;
  if (stream >= ctx->n_streams) return B2_ERR_INVALID_CONFIG;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  if (stream >= 0) {
    B2_CUDA(ctx, cudaStreamSynchronize(ctx->streams[stream]));
  } else {
    for (int i = 0; i < ctx->n_streams; ++i) B2_CUDA(ctx, cudaStreamSynchronize(ctx->streams[i]));
  }
  return B2_OK;
}
extern "C" int b2_event_record(b2_ctx* ctx, int stream, int slot) {
  B2_CHECK_CTX(ctx, stream);
  if (slot < 0 || slot >= B2_EVENT_SLOTS) return B2_ERR_INVALID_CONFIG;
  B2_CUDA(ctx, cudaEventRecord(ctx->events[slot], ctx->streams[stream]));
  return B2_OK;
}
extern "C" int b2_event_elapsed_ms(b2_ctx* ctx, int slot_start, int slot_stop, float* ms) {
  if (ctx == NULL || ms == NULL) return B2_ERR_NULL;
  if (slot_start < 0 || slot_start >= B2_EVENT_SLOTS || slot_stop < 0 || slot_stop >= B2_EVENT_SLOTS) return B2_ERR_INVALID_CONFIG;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  B2_CUDA(ctx, cudaEventSynchronize(ctx->events[slot_stop]));
  B2_CUDA(ctx, cudaEventElapsedTime(ms, ctx->events[slot_start], ctx->events[slot_stop]));
  return B2_OK;
}

// ---- synthetic inputs ----------------------------------------------------------
namespace {
__global__ void synth_kernel(uint32_t* dst, int words_per_row, int rows, int stride_bytes,
                             size_t image_bytes, int n_images, uint32_t seed, uint32_t first_img,
                             uint32_t plane, int alpha_mode) {
  const uint64_t per_img = (uint64_t)words_per_row * rows;
  const uint64_t total = per_img * n_images;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t img = (uint32_t)(i / per_img);
    const uint32_t r = (uint32_t)(i - img * per_img);
    const uint32_t y = r / (uint32_t)words_per_row;
    const uint32_t x = r - y * (uint32_t)words_per_row;
    uint32_t w = b2_lowbias32(seed ^ ((first_img + img) * 0x9E3779B1u) ^ (plane * 0x85EBCA77u) ^
                              (y * 0xC2B2AE3Du) ^ (x * 0x27D4EB2Fu));
    if (alpha_mode == B2_SYNTH_OPAQUE) {
      w |= 0xff000000u;
    } else if (alpha_mode != B2_SYNTH_RAW) {
      const uint32_t h2 = b2_lowbias32(w ^ 0x9E3779B9u);
      if (alpha_mode == B2_SYNTH_SPARSE) {
        w = ((h2 & 0xff) >= 77) ? (w | 0xff000000u) : ((w & 0x00ffffffu) | (h2 & 0xff000000u));
      } else {
        w = (w & 0x00ffffffu) | ((h2 & 1) ? 0xff000000u : 0u);
      }
    }
    uint32_t* row = (uint32_t*)((uint8_t*)dst + (size_t)img * image_bytes + (size_t)y * stride_bytes);
    row[x] = w;
  }
}
}  // namespace

extern "C" int b2_synth_fill_batch_u32(b2_ctx* ctx, int stream, void* dst_dev, int words_per_row,
                                       int rows, int stride_bytes, size_t image_bytes, int n_images,
                                       uint32_t seed, uint32_t first_img, uint32_t plane,
                                       int alpha_mode) {
  B2_CHECK_CTX(ctx, stream);
  if (dst_dev == NULL) return B2_ERR_NULL;
  if (words_per_row <= 0 || rows <= 0 || n_images <= 0 || stride_bytes < 4 * words_per_row ||
      (stride_bytes & 3) || (image_bytes & 3) || ((uintptr_t)dst_dev & 3)) return B2_ERR_INVALID_CONFIG;
  synth_kernel<<<ctx->sm_count * 8, 256, 0, ctx->streams[stream]>>>(
      (uint32_t*)dst_dev, words_per_row, rows, stride_bytes, image_bytes, n_images, seed, first_img,
      plane, alpha_mode);
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}
