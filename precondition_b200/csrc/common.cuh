// Shared helpers for the precond_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/precond_b200.h"

namespace pc {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// GEMM-phase timing hooks (no-ops unless pc_stats_reset(1) was called)
bool gemm_timing_enabled();
void gemm_timing_record(cudaEvent_t start, cudaEvent_t stop);  // takes ownership
void gemm_count(int launches);
void gemm_add_flops(double flops);

#define PC_CUDA_CHECK(expr)                                                    \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) {                                                   \
      ::pc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                      __FILE__, __LINE__);                                     \
      return PC_ERR_CUDA;                                                      \
    }                                                                          \
  } while (0)

#define PC_REQUIRE(cond, ...)      \
  do {                             \
    if (!(cond)) {                 \
      ::pc::set_error(__VA_ARGS__); \
      return PC_ERR_INVALID;       \
    }                              \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// max of |x| that PROPAGATES NaN the way jnp.max does: the int view of a
// non-negative float orders like the float and a quiet NaN (0x7fc00000) sorts
// above +inf, so an integer max keeps it.
__device__ __forceinline__ uint32_t absbits(float v) {
  return __float_as_uint(v) & 0x7fffffffu;
}
__device__ __forceinline__ uint32_t warp_max_u32(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint32_t t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t > v ? t : v;
  }
  return v;
}

// Block-wide sum with a fixed (deterministic) order. `scratch` >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarp = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = lane < nwarp ? scratch[lane] : 0.f;
    r = warp_sum(r);
    if (lane == 0) scratch[0] = r;
  }
  __syncthreads();
  r = scratch[0];
  return r;
}

__device__ __forceinline__ uint32_t block_max_u32(uint32_t v, uint32_t* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarp = (blockDim.x + 31) >> 5;
  v = warp_max_u32(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  uint32_t r = 0;
  if (warp == 0) {
    r = lane < nwarp ? scratch[lane] : 0u;
    r = warp_max_u32(r);
    if (lane == 0) scratch[0] = r;
  }
  __syncthreads();
  return scratch[0];
}

}  // namespace pc
