// Shared helpers for the ExpertSim sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/expertsim_b200.h"

namespace es {

void set_error(const std::string& msg);
int* pipeline_err_flag();   // per-device, host-mapped (api.cu); nullptr if it cannot be allocated

#define ES_REQUIRE(cond, msg)                                               \
  do {                                                                      \
    if (!(cond)) {                                                          \
      ::es::set_error(std::string(__func__) + ": " + (msg));                \
      return ES_ERR_INVALID;                                                \
    }                                                                       \
  } while (0)

#define ES_LAUNCH_CHECK()                                                   \
  do {                                                                      \
    cudaError_t _e = cudaGetLastError();                                    \
    if (_e != cudaSuccess) {                                                \
      ::es::set_error(std::string(__func__) + ": " + cudaGetErrorString(_e)); \
      return ES_ERR_CUDA;                                                   \
    }                                                                       \
  } while (0)

#define ES_CUDA(call)                                                       \
  do {                                                                      \
    cudaError_t _e = (call);                                                \
    if (_e != cudaSuccess) {                                                \
      ::es::set_error(std::string(__func__) + ": " + cudaGetErrorString(_e)); \
      return ES_ERR_CUDA;                                                   \
    }                                                                       \
  } while (0)

constexpr int kMaxGroups = 64;
constexpr float kLReLU = 0.1f;
constexpr float kNormEps = 1e-5f;

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long ceil_div_l(long a, long b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum; `red` is a shared array of >= 32 floats.  All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

// Find the group that owns `row` (groups are sorted by row_start; inactive groups have rows == 0).
__device__ __forceinline__ int find_group(const es_group* __restrict__ grp, int n_groups, int row) {
  for (int g = 0; g < n_groups; ++g) {
    const int s = grp[g].row_start, n = grp[g].rows;
    if (row >= s && row < s + n) return g;
  }
  return -1;
}

__device__ __forceinline__ float lrelu(float x) { return x > 0.f ? x : kLReLU * x; }
__device__ __forceinline__ float act_fwd(float x, int act) {
  return act == 1 ? fmaxf(x, 0.f) : (act == 2 ? lrelu(x) : x);
}
// derivative expressed on the pre-activation value
__device__ __forceinline__ float act_grad(float pre, int act) {
  return act == 1 ? (pre > 0.f ? 1.f : 0.f) : (act == 2 ? (pre > 0.f ? 1.f : kLReLU) : 1.f);
}

__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ __nv_bfloat16 f2bf(float v) { return __float2bfloat16_rn(v); }

// unpack 8 bf16 held in a uint4 to floats
__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 q;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return q;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace es
