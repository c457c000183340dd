// b2_staged.cu -- the host-buffer entry points of the C ABI: a whole batch goes
// host -> device -> kernel -> host in ONE call, cut into chunks that are spread over the
// context's streams so that the H2D of one chunk, the kernel of another and the D2H of a
// third overlap (PCIe is full duplex).  This is what the batched Go entry points
// (picture.ImportRGBABatch, picture.YUVAToRGBABatch, dsp.VP8LInverseTransformBatch) call
// through cgo; the device-side job arrays are the same ones the resident entry points
// take, so a picture whose planes already live on the device simply passes NULL host
// pointers for the sides that need no copy.
//
// Copies: one cudaMemcpyAsync per contiguous range.  Consecutive ranges that continue
// each other exactly on BOTH sides (the planes of a library-allocated picture, the
// pictures of a library-allocated batch) are merged, so a chunk of 16 frames is one
// copy, not 64; rows with padding become one pitched copy per plane.
#include <vector>

#include "b2_common.cuh"

namespace {

struct Copy { char* dst; const char* src; size_t dpitch, spitch, width, rows; };

struct CopyList {
  std::vector<Copy> v;
  // to_host: padding between rows on the host side belongs to the caller and must not
  // be written, so only tight rows collapse into one range; towards the device the
  // padding is read and lands in the library's own row padding.
  void add(void* dst, const void* src, size_t dpitch, size_t spitch, size_t width, size_t rows, bool to_host) {
    if (dst == NULL || src == NULL || width == 0 || rows == 0) return;
    Copy c = {(char*)dst, (const char*)src, dpitch, spitch, width, rows};
    const bool tight = (dpitch == width && spitch == width);
    if (rows == 1 || tight || (!to_host && dpitch == spitch)) {
      c.width = (rows - 1) * spitch + width;
      c.rows = 1;
      c.dpitch = c.spitch = c.width;
      if (!v.empty()) {
        Copy& p = v.back();
        if (p.rows == 1 && p.dst + p.width == c.dst && p.src + p.width == c.src) {
          p.width += c.width;
          p.dpitch = p.spitch = p.width;
          return;
        }
      }
    }
    v.push_back(c);
  }
  int issue(b2_ctx* ctx, int stream, cudaMemcpyKind kind) {
    for (const Copy& c : v) {
      if (c.rows == 1) {
        B2_CUDA(ctx, cudaMemcpyAsync(c.dst, c.src, c.width, kind, ctx->streams[stream]));
      } else {
        B2_CUDA(ctx, cudaMemcpy2DAsync(c.dst, c.dpitch, c.src, c.spitch, c.width, c.rows, kind, ctx->streams[stream]));
      }
      if (kind == cudaMemcpyHostToDevice) { ctx->h2d_calls++; ctx->h2d_bytes += c.width * c.rows; }
      else { ctx->d2h_calls++; ctx->d2h_bytes += c.width * c.rows; }
    }
    v.clear();
    return B2_OK;
  }
};

inline size_t uabs(int32_t s) { return (size_t)(s < 0 ? -(int64_t)s : (int64_t)s); }
// lowest address of a plane given the address of its row 0 and a signed stride
template <typename T> inline T* low(T* row0, int32_t stride, int32_t rows) {
  return (row0 == NULL || stride >= 0) ? row0 : (T*)((char*)row0 + (int64_t)stride * (rows - 1));
}

int ensure_flags(b2_ctx* ctx, int n) {
  if ((size_t)n <= ctx->flags_cap) return B2_OK;
  for (int i = 0; i < ctx->n_streams; ++i) B2_CUDA(ctx, cudaStreamSynchronize(ctx->streams[i]));
  if (ctx->flags_dev) { B2_CUDA(ctx, cudaFree(ctx->flags_dev)); ctx->flags_dev = NULL; }
  if (ctx->flags_host) { B2_CUDA(ctx, cudaFreeHost(ctx->flags_host)); ctx->flags_host = NULL; }
  ctx->flags_cap = 0;
  const size_t cap = (size_t)n + 1024;
  B2_CUDA(ctx, cudaMalloc((void**)&ctx->flags_dev, cap * sizeof(int32_t)));
  B2_CUDA(ctx, cudaHostAlloc((void**)&ctx->flags_host, cap * sizeof(int32_t), cudaHostAllocDefault));
  ctx->flags_cap = cap;
  return B2_OK;
}

int ensure_events(b2_ctx* ctx) {
  for (int i = 0; i < B2_STAGED_EVENTS; ++i) {
    if (ctx->staged_ev[i] == NULL) B2_CUDA(ctx, cudaEventCreateWithFlags(&ctx->staged_ev[i], cudaEventDisableTiming));
  }
  return B2_OK;
}

inline int bytes_per_pixel(int order) { return (order == B2_ORDER_RGB || order == B2_ORDER_BGR) ? 3 : 4; }

}  // namespace

extern "C" int b2_copy_stats(const b2_ctx* ctx, uint64_t out[4]) {
  if (ctx == NULL || out == NULL) return B2_ERR_NULL;
  out[0] = ctx->h2d_calls; out[1] = ctx->h2d_bytes; out[2] = ctx->d2h_calls; out[3] = ctx->d2h_bytes;
  return B2_OK;
}

extern "C" int b2_event_sync(b2_ctx* ctx, int slot) {
  if (ctx == NULL) return B2_ERR_NULL;
  if (slot < 0 || slot >= B2_EVENT_SLOTS) return B2_ERR_INVALID_CONFIG;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  B2_CUDA(ctx, cudaEventSynchronize(ctx->events[slot]));
  return B2_OK;
}

extern "C" int b2_import_batch_staged(b2_ctx* ctx, const b2_import_job* jobs, const b2_import_host* host,
                                      int n_jobs, int order, int flags, int chunk, float dithering,
                                      int32_t* has_alpha_host) {
  B2_CHECK_CTX(ctx, 0);
  if (n_jobs == 0) return B2_OK;
  if (jobs == NULL || host == NULL) return B2_ERR_NULL;
  if (n_jobs < 0 || order < B2_ORDER_RGBA || order > B2_ORDER_BGR) return B2_ERR_INVALID_CONFIG;
  if (chunk < 1) chunk = 16;
  int st;
  if ((st = ensure_flags(ctx, n_jobs)) != B2_OK) return st;
  if ((st = ensure_events(ctx)) != B2_OK) return st;
  const int ns = ctx->n_streams;
  const int bpp = bytes_per_pixel(order);
  const int n_chunks = (n_jobs + chunk - 1) / chunk;
  const int lag = ns < B2_STAGED_EVENTS ? ns : B2_STAGED_EVENTS;   // a multiple of ns keeps a chunk's waves on one stream
  CopyList cl;
  // second wave of one chunk: its flags are on the host -> copy back the alpha planes of
  // the pictures that turned out to have alpha (conversions.go:56-58 drops the others)
  auto finish = [&](int k) -> int {
    B2_CUDA(ctx, cudaEventSynchronize(ctx->staged_ev[k % B2_STAGED_EVENTS]));
    const int c0 = k * chunk, c1 = (c0 + chunk < n_jobs) ? c0 + chunk : n_jobs;
    for (int i = c0; i < c1; ++i) {
      const b2_import_job& j = jobs[i];
      const b2_import_host& h = host[i];
      if (ctx->flags_host[i] && h.a != NULL && j.a != NULL) {
        cl.add(h.a, j.a, (size_t)h.a_stride, (size_t)j.a_stride, (size_t)j.width, (size_t)j.height, true);
      }
    }
    return cl.issue(ctx, k % ns, cudaMemcpyDeviceToHost);
  };
  for (int k = 0; k < n_chunks; ++k) {
    const int s = k % ns;
    const int c0 = k * chunk, c1 = (c0 + chunk < n_jobs) ? c0 + chunk : n_jobs, m = c1 - c0;
    // the alpha planes of chunk k - lag go out BEFORE this chunk's work is queued behind them on the
    // same stream: the host stays 'lag' chunks ahead of the device, the D2H engine never idles
    if (k >= lag && (st = finish(k - lag)) != B2_OK) return st;
    for (int i = c0; i < c1; ++i) {
      const b2_import_job& j = jobs[i];
      if (host[i].src != NULL) {
        if (j.src == NULL || j.width <= 0 || j.height <= 0) return B2_ERR_NULL;
        cl.add((void*)low(j.src, j.src_stride, j.height), low(host[i].src, j.src_stride, j.height),
               uabs(j.src_stride), uabs(j.src_stride), (size_t)bpp * j.width, (size_t)j.height, false);
      }
    }
    if ((st = cl.issue(ctx, s, cudaMemcpyHostToDevice)) != B2_OK) return st;
    if (dithering > 0.f) st = b2_rgba_to_yuv420a_dithered(ctx, s, jobs + c0, m, order, flags, ctx->flags_dev + c0, dithering);
    else st = b2_rgba_to_yuv420a(ctx, s, jobs + c0, m, order, flags, ctx->flags_dev + c0);
    if (st != B2_OK) return st;
    B2_CUDA(ctx, cudaMemcpyAsync(ctx->flags_host + c0, ctx->flags_dev + c0, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, ctx->streams[s]));
    ctx->d2h_calls++; ctx->d2h_bytes += sizeof(int32_t) * m;
    B2_CUDA(ctx, cudaEventRecord(ctx->staged_ev[k % B2_STAGED_EVENTS], ctx->streams[s]));
    for (int i = c0; i < c1; ++i) {
      const b2_import_job& j = jobs[i];
      const b2_import_host& h = host[i];
      const size_t uvw = (size_t)(j.width + 1) >> 1, uvh = (size_t)(j.height + 1) >> 1;
      // host planes are top-down with positive strides: flipped device planes cannot be copied as they lie
      if (((h.y || h.a) && (j.y_stride < 0 || j.a_stride < 0)) || ((h.u || h.v) && j.uv_stride < 0)) return B2_ERR_INVALID_CONFIG;
      cl.add(h.y, j.y, (size_t)h.y_stride, (size_t)j.y_stride, (size_t)j.width, (size_t)j.height, true);
      cl.add(h.u, j.u, (size_t)h.uv_stride, (size_t)j.uv_stride, uvw, uvh, true);
      cl.add(h.v, j.v, (size_t)h.uv_stride, (size_t)j.uv_stride, uvw, uvh, true);
    }
    if ((st = cl.issue(ctx, s, cudaMemcpyDeviceToHost)) != B2_OK) return st;
  }
  for (int k = (n_chunks > lag ? n_chunks - lag : 0); k < n_chunks; ++k) {
    if ((st = finish(k)) != B2_OK) return st;
  }
  for (int i = 0; i < ns; ++i) B2_CUDA(ctx, cudaStreamSynchronize(ctx->streams[i]));
  if (has_alpha_host != NULL) memcpy(has_alpha_host, ctx->flags_host, sizeof(int32_t) * n_jobs);
  return B2_OK;
}

extern "C" int b2_upsample_batch_staged(b2_ctx* ctx, const b2_upsample_job* jobs, const b2_upsample_host* host,
                                        int n_jobs, int mode, int chunk) {
  B2_CHECK_CTX(ctx, 0);
  if (n_jobs == 0) return B2_OK;
  if (jobs == NULL || host == NULL) return B2_ERR_NULL;
  if (n_jobs < 0 || mode < B2_MODE_RGB || mode > B2_MODE_rgbA_4444) return B2_ERR_INVALID_CONFIG;
  if (chunk < 1) chunk = 8;
  static const int kBpp[11] = {3, 4, 3, 4, 4, 2, 2, 4, 4, 4, 2};
  const int ns = ctx->n_streams;
  CopyList cl;
  int st;
  for (int k = 0, c0 = 0; c0 < n_jobs; ++k, c0 += chunk) {
    const int s = k % ns;
    const int c1 = (c0 + chunk < n_jobs) ? c0 + chunk : n_jobs;
    for (int i = c0; i < c1; ++i) {
      const b2_upsample_job& j = jobs[i];
      const b2_upsample_host& h = host[i];
      if (j.width <= 0 || j.height <= 0) return B2_ERR_BAD_DIM;
      const size_t uvw = (size_t)(j.width + 1) >> 1, uvh = (size_t)(j.height + 1) >> 1;
      if (((h.y || h.a) && (j.y_stride < 0 || j.a_stride < 0)) || ((h.u || h.v) && j.uv_stride < 0) ||
          (h.dst && j.dst_stride < 0)) return B2_ERR_INVALID_CONFIG;   // host buffers are top-down
      cl.add((void*)j.y, h.y, (size_t)j.y_stride, (size_t)h.y_stride, (size_t)j.width, (size_t)j.height, false);
      cl.add((void*)j.u, h.u, (size_t)j.uv_stride, (size_t)h.uv_stride, uvw, uvh, false);
      cl.add((void*)j.v, h.v, (size_t)j.uv_stride, (size_t)h.uv_stride, uvw, uvh, false);
      if (j.a != NULL) cl.add((void*)j.a, h.a, (size_t)j.a_stride, (size_t)h.a_stride, (size_t)j.width, (size_t)j.height, false);
    }
    if ((st = cl.issue(ctx, s, cudaMemcpyHostToDevice)) != B2_OK) return st;
    if ((st = b2_yuv420a_to_rgba_fancy(ctx, s, jobs + c0, c1 - c0, mode)) != B2_OK) return st;
    for (int i = c0; i < c1; ++i) {
      const b2_upsample_job& j = jobs[i];
      const b2_upsample_host& h = host[i];
      // only the rows the job produces (EmitFancyRGB's batches) come back
      const int r0 = (j.row_begin == 0 && j.row_end == 0) ? 0 : j.row_begin;
      const int r1 = (j.row_begin == 0 && j.row_end == 0) ? j.height : j.row_end;
      if (h.dst == NULL || r1 <= r0) continue;
      cl.add((char*)h.dst + (size_t)h.dst_stride * r0, (const char*)j.dst + (size_t)j.dst_stride * r0, (size_t)h.dst_stride, (size_t)j.dst_stride,
             (size_t)kBpp[mode] * j.width, (size_t)(r1 - r0), true);
    }
    if ((st = cl.issue(ctx, s, cudaMemcpyDeviceToHost)) != B2_OK) return st;
  }
  for (int i = 0; i < ns; ++i) B2_CUDA(ctx, cudaStreamSynchronize(ctx->streams[i]));
  return B2_OK;
}

extern "C" int b2_vp8l_batch_staged(b2_ctx* ctx, int op, const b2_argb_job* jobs, const b2_argb_host* host,
                                    int n_jobs, int chunk) {
  B2_CHECK_CTX(ctx, 0);
  if (n_jobs == 0) return B2_OK;
  if (jobs == NULL || host == NULL) return B2_ERR_NULL;
  typedef int (*op_fn)(b2_ctx*, int, const b2_argb_job*, int);
  static const op_fn kOps[] = {b2_vp8l_add_green, b2_vp8l_sub_green, b2_vp8l_color_inverse, b2_vp8l_color_forward,
                               b2_vp8l_inverse_chain, b2_vp8l_forward_chain, b2_vp8l_color_index_inverse,
                               b2_vp8l_bundle_color_map, b2_vp8l_predictor_inverse, b2_vp8l_predictor_forward};
  if (n_jobs < 0 || op < 0 || op >= (int)(sizeof(kOps) / sizeof(kOps[0]))) return B2_ERR_INVALID_CONFIG;
  if (chunk < 1) chunk = 4;
  const int ns = ctx->n_streams;
  CopyList cl;
  int st;
  for (int k = 0, c0 = 0; c0 < n_jobs; ++k, c0 += chunk) {
    const int s = k % ns;
    const int c1 = (c0 + chunk < n_jobs) ? c0 + chunk : n_jobs;
    for (int i = c0; i < c1; ++i) {
      cl.add((void*)jobs[i].src, host[i].src, host[i].src_bytes, host[i].src_bytes, host[i].src_bytes, 1, false);
      cl.add((void*)jobs[i].data, host[i].data, host[i].data_bytes, host[i].data_bytes, host[i].data_bytes, 1, false);
    }
    if ((st = cl.issue(ctx, s, cudaMemcpyHostToDevice)) != B2_OK) return st;
    if ((st = kOps[op](ctx, s, jobs + c0, c1 - c0)) != B2_OK) return st;
    for (int i = c0; i < c1; ++i) cl.add(host[i].dst, jobs[i].dst, host[i].dst_bytes, host[i].dst_bytes, host[i].dst_bytes, 1, true);
    if ((st = cl.issue(ctx, s, cudaMemcpyDeviceToHost)) != B2_OK) return st;
  }
  for (int i = 0; i < ns; ++i) B2_CUDA(ctx, cudaStreamSynchronize(ctx->streams[i]));
  return B2_OK;
}

extern "C" int b2_vp8l_chain_batch_staged(b2_ctx* ctx, const b2_vp8l_chain_job* jobs, const b2_vp8l_chain_host* host,
                                          int n_jobs, int out_mode, int chunk) {
  B2_CHECK_CTX(ctx, 0);
  if (n_jobs == 0) return B2_OK;
  if (jobs == NULL || host == NULL) return B2_ERR_NULL;
  if (n_jobs < 0) return B2_ERR_INVALID_CONFIG;
  if (chunk < 1) chunk = 4;
  const int ns = ctx->n_streams;
  CopyList cl;
  int st;
  for (int k = 0, c0 = 0; c0 < n_jobs; ++k, c0 += chunk) {
    const int s = k % ns;
    const int c1 = (c0 + chunk < n_jobs) ? c0 + chunk : n_jobs;
    for (int i = c0; i < c1; ++i) {
      const size_t px = (size_t)4 * (size_t)jobs[i].width * (size_t)jobs[i].height;
      cl.add((void*)jobs[i].src, host[i].src, px, px, px, 1, false);
      cl.add((void*)jobs[i].predictor_data, host[i].predictor_data, host[i].predictor_bytes, host[i].predictor_bytes, host[i].predictor_bytes, 1, false);
      cl.add((void*)jobs[i].color_data, host[i].color_data, host[i].color_bytes, host[i].color_bytes, host[i].color_bytes, 1, false);
    }
    if ((st = cl.issue(ctx, s, cudaMemcpyHostToDevice)) != B2_OK) return st;
    if ((st = b2_vp8l_predictor_inverse_chain(ctx, s, jobs + c0, c1 - c0, out_mode)) != B2_OK) return st;
    for (int i = c0; i < c1; ++i) {
      const size_t px = (size_t)4 * (size_t)jobs[i].width * (size_t)jobs[i].height;
      cl.add(host[i].dst, jobs[i].dst, px, px, px, 1, true);
    }
    if ((st = cl.issue(ctx, s, cudaMemcpyDeviceToHost)) != B2_OK) return st;
  }
  for (int i = 0; i < ns; ++i) B2_CUDA(ctx, cudaStreamSynchronize(ctx->streams[i]));
  return B2_OK;
}
