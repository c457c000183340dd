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
// row pairs.  Returns a cudaError_t from the launch.
cudaError_t b2_import_fast_launch(const b2_ctx* ctx, cudaStream_t s, const ImportJobDev* jobs,
                                  const uint32_t* unit_prefix, uint32_t n_jobs,
                                  uint32_t uniform_units, uint32_t total_units, int order,
                                  bool alpha, bool write_a, int div_table, int32_t* has_alpha);
