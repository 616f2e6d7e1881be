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

// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl() may be SCHEDULED while its predecessor
// on the stream is still running; it must execute pdl_wait() before it touches anything the predecessor reads or
// writes (every hot-path kernel does so first thing, or right after a data-independent prologue) and signals with
// pdl_launch_dependents() that its own successor may be scheduled.  This hides the launch / scheduling latency between
// the ~40 dependent kernels of a training step (2-3 us each), eagerly and inside a captured CUDA graph.
// RGCN_PDL=0 in the environment turns it off (plain stream-ordered launches).
bool pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

void count_launch(int n);

// Number of SMs of the current device (148 on B200), cached.
int sm_count();

// Hub segments: a (row, relation) segment with more than kHubThreshold edges is cut into
// chunks of kHubChunk edges that whole thread blocks reduce in a fixed order.
constexpr int kHubThreshold = 128;
constexpr int kHubChunk = 128;

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// first statement of a hot-path kernel: let the successor be scheduled, then wait for the predecessor's results
__device__ __forceinline__ void pdl_enter() { pdl_launch_dependents(); pdl_wait(); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// Packed fp32 arithmetic of sm_100 (add.f32x2 / fma.f32x2 -> FADD2 / FFMA2: two IEEE round-to-nearest results per
// instruction, bit-identical to the scalar forms).  The gather loops are instruction-issue sensitive (16 B per lane per
// edge): the accumulate of one 128-bit vector is 2 instructions instead of 4.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void add4(float4& a, const float4& b) {
  unsigned long long a0 = pack2(a.x, a.y), a1 = pack2(a.z, a.w);
  const unsigned long long b0 = pack2(b.x, b.y), b1 = pack2(b.z, b.w);
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a0) : "l"(b0));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a1) : "l"(b1));
  unpack2(a0, a.x, a.y);
  unpack2(a1, a.z, a.w);
}
__device__ __forceinline__ void fma4(float4& a, float w, const float4& b) {
  unsigned long long a0 = pack2(a.x, a.y), a1 = pack2(a.z, a.w);
  const unsigned long long b0 = pack2(b.x, b.y), b1 = pack2(b.z, b.w), ww = pack2(w, w);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a0) : "l"(ww), "l"(b0));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a1) : "l"(ww), "l"(b1));
  unpack2(a0, a.x, a.y);
  unpack2(a1, a.z, a.w);
}
__device__ __forceinline__ float4 scale4(const float4& a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
// a / s with a TRUE (correctly rounded) division, like the reference's `s / cnt.clamp(min=1)` — but with ONE reciprocal
// for all components: r = RN(1 / s), q = RN(a r), then two exact-remainder corrections q += (a - s q) r (Markstein: a
// faithful quotient corrected through the fma remainder with the correctly rounded reciprocal is the correctly rounded
// quotient).  Same bits as the IEEE division for a = 0 and for |a| >= 2^-100 (no underflow in the remainder), s a small
// positive integer — checked exhaustively against exact rational arithmetic for s in 2 .. 129 on random a.  The
// compiler's division is ~25 instructions per component; this is 5, plus the shared reciprocal: the walks finish
// 93 k (row, relation) segments per cfg2 layer and spent a quarter (d = 64) of their issue slots dividing.
__device__ __forceinline__ float div_by(float a, float s, float r) {
  float q = a * r;
  q = fmaf(fmaf(-s, q, a), r, q);
  return fmaf(fmaf(-s, q, a), r, q);
}
__device__ __forceinline__ float4 div4(const float4& a, float s) {
  const float r = __frcp_rn(s);
  return make_float4(div_by(a.x, s, r), div_by(a.y, s, r), div_by(a.z, s, r), div_by(a.w, s, r));
}

}  // namespace rgcn
