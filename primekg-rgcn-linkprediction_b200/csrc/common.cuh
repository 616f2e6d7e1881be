// Shared helpers for the RGCN B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/rgcn_b200.h"

namespace rgcn {

// thread-local last-error message, read through rgcn_last_error()
char* err_buf();
void set_error(const char* fmt, ...);

#define RGCN_CHECK_ARG(cond, ...)                                   \
  do {                                                              \
    if (!(cond)) {                                                  \
      ::rgcn::set_error(__VA_ARGS__);                               \
      return RGCN_EINVAL;                                           \
    }                                                               \
  } while (0)

#define RGCN_CUDA(call)                                                              \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      ::rgcn::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,         \
                        cudaGetErrorString(e__));                                    \
      return RGCN_ECUDA;                                                             \
    }                                                                                \
  } while (0)

// every launch of one of OUR kernels goes through this macro, which also counts it
#define RGCN_LAUNCH_CHECK()          \
  do {                               \
    ::rgcn::count_launch(1);         \
    RGCN_CUDA(cudaGetLastError());   \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

void count_launch(int n);

// Number of SMs of the current device (148 on B200), cached.
int sm_count();

// Hub segments: a (row, relation) segment with more than kHubThreshold edges is cut into
// chunks of kHubChunk edges that whole thread blocks reduce in a fixed order.
constexpr int kHubThreshold = 128;
constexpr int kHubChunk = 128;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void add4(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void fma4(float4& a, float w, const float4& b) {
  a.x = fmaf(w, b.x, a.x); a.y = fmaf(w, b.y, a.y); a.z = fmaf(w, b.z, a.z); a.w = fmaf(w, b.w, a.w);
}
__device__ __forceinline__ float4 scale4(const float4& a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 div4(const float4& a, float s) { return make_float4(a.x / s, a.y / s, a.z / s, a.w / s); }

}  // namespace rgcn
