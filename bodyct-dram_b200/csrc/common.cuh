// Shared helpers for libdram_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/dram_b200.h"

namespace dram {

void set_error(const char* fmt, ...);

#define DRAM_REQUIRE(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      ::dram::set_error(__VA_ARGS__);           \
      return DRAM_E_INVALID;                    \
    }                                           \
  } while (0)

#define DRAM_CUDA(expr)                                                                  \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::dram::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DRAM_E_CUDA;                                                                \
    }                                                                                    \
  } while (0)

#define DRAM_LAUNCH_CHECK()                                                              \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      ::dram::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DRAM_E_CUDA;                                                                \
    }                                                                                    \
  } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

static inline int grid_for(long long work_items, int block, int max_waves = 32) {
  long long g = (work_items + block - 1) / block;
  long long cap = (long long)kNumSMs * max_waves;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// Grid-stride walk over (row, group) pairs of a [rows][groups] problem without a 64-bit division per element: the
// per-thread start is divided once, every further step is an add and a compare.
struct RowGroupIter {
  long long r, dr;
  int g, dg, groups;
  __device__ __forceinline__ RowGroupIter(int groups_) : groups(groups_) {
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long T = (long long)gridDim.x * blockDim.x;
    r = tid / groups; g = (int)(tid - r * groups);
    dr = T / groups; dg = (int)(T - dr * groups);
  }
  __device__ __forceinline__ void next() {
    r += dr; g += dg;
    if (g >= groups) { g -= groups; ++r; }
  }
};

// Grid whose total thread count is a multiple of `groups`, so that a thread of a [rows][groups] grid-stride loop keeps
// ONE group (its per-channel constants are loaded once, outside the loop) and only walks rows.
static inline int grid_fixed_group(long long rows, int groups, int block, int max_waves = 32) {
  int a = groups, b = block;
  while (b) { int t = a % b; a = b; b = t; }
  const int m = groups / a;                       // grid must be a multiple of groups / gcd(groups, block)
  long long g = grid_for(rows * groups, block, max_waves);
  g = (g + m - 1) / m * m;
  return (int)g;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// align_corners=True source index (ATen area_pixel_compute_source_index + guard_index_and_lambda)
struct Lerp {
  int i0, i1;
  float w0, w1;
};
__device__ __forceinline__ Lerp lerp_setup(int o, float scale, int in_size) {
  float real = scale * (float)o;
  int i0 = (int)real;
  if (i0 > in_size - 1) i0 = in_size - 1;
  float l1 = real - (float)i0;
  l1 = fminf(fmaxf(l1, 0.f), 1.f);
  Lerp r;
  r.i0 = i0;
  r.i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  r.w1 = l1;
  r.w0 = 1.f - l1;
  return r;
}
static __host__ __device__ __forceinline__ float ac_scale(int in_size, int out_size) {
  return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
}

}  // namespace dram
