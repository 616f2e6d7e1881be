// Basis decomposition of the relation weights (RGCNConv(num_bases = B), plumbed at reference src/models/rgcn.py:58, :76,
// :84): W_r = sum_b comp[r, b] * V_b, i.e. PyG's `(comp @ weight.view(B, -1)).view(R, in, out)`, and its backward
//   g_V[b]       = sum_r comp[r, b] * g_W[r]
//   g_comp[r, b] = <g_W[r], V_b>
// B and R are tens, in * out is 10^4 .. 10^5: three bandwidth-bound streaming kernels over the [*, in * out] matrices
// (a library SGEMM sees a 30 x 8 product with K = 65,536 and spends 150 us on it; these take a few us).
// Fixed summation orders: deterministic.
#include "common.cuh"

namespace rgcn {

constexpr int kMaxCombine = 64;      // rows of comp held per thread loop (R and B up to this)

// out[i, :] = sum_j coef(i, j) * in[j, :]   over float4 columns; coef(i, j) = c[i * ldc + j] or c[j * ldc + i]
template <bool TRANSPOSED>
__global__ void __launch_bounds__(256) combine_rows_kernel(const float* __restrict__ c, int ldc, int n_out, int n_in,
                                                           const float* __restrict__ in, float* __restrict__ out,
                                                           int64_t cols4) {
  pdl_enter();
  extern __shared__ float s_c[];                       // [n_out][n_in]
  for (int t = threadIdx.x; t < n_out * n_in; t += 256) {
    const int i = t / n_in, j = t % n_in;
    s_c[t] = TRANSPOSED ? c[j * ldc + i] : c[i * ldc + j];
  }
  __syncthreads();
  const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (k >= cols4) return;
  const float4* __restrict__ in4 = reinterpret_cast<const float4*>(in);
  float4* __restrict__ out4 = reinterpret_cast<float4*>(out);
  // inputs in chunks of 8 rows held in registers, every output row accumulated in input order
  for (int i0 = 0; i0 < n_out; i0 += 8) {
    float4 acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < n_in; ++j) {
      const float4 v = __ldg(in4 + (int64_t)j * cols4 + k);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i0 + u < n_out) fma4(acc[u], s_c[(i0 + u) * n_in + j], v);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (i0 + u < n_out) out4[(int64_t)(i0 + u) * cols4 + k] = acc[u];
  }
}

// g_comp[r, b] = <gW[r, :], V[b, :]>: one block per r, B accumulators per thread (B <= 16), fixed-order block reduce
__global__ void __launch_bounds__(256) comp_grad_kernel(const float* __restrict__ gW, const float* __restrict__ V, int B,
                                                        int64_t cols4, float* __restrict__ g_comp, int ldg) {
  pdl_enter();
  __shared__ float red[8][16];
  const int r = blockIdx.x;
  const float4* __restrict__ g4 = reinterpret_cast<const float4*>(gW) + (int64_t)r * cols4;
  const float4* __restrict__ v4 = reinterpret_cast<const float4*>(V);
  float acc[16];
#pragma unroll
  for (int b = 0; b < 16; ++b) acc[b] = 0.f;
  for (int64_t k = threadIdx.x; k < cols4; k += 256) {
    const float4 g = __ldg(g4 + k);
#pragma unroll
    for (int b = 0; b < 16; ++b) {
      if (b < B) {
        const float4 v = __ldg(v4 + (int64_t)b * cols4 + k);
        acc[b] = fmaf(g.x, v.x, acc[b]); acc[b] = fmaf(g.y, v.y, acc[b]);
        acc[b] = fmaf(g.z, v.z, acc[b]); acc[b] = fmaf(g.w, v.w, acc[b]);
      }
    }
  }
#pragma unroll
  for (int b = 0; b < 16; ++b) {
    float s = acc[b];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][b] = s;
  }
  __syncthreads();
  if (threadIdx.x < B) {
    float s = red[0][threadIdx.x];
    for (int w = 1; w < 8; ++w) s += red[w][threadIdx.x];
    g_comp[r * ldg + threadIdx.x] = s;
  }
}

}  // namespace rgcn

using namespace rgcn;

static int check_combine(const float* comp, const float* a, const float* b, int32_t R, int32_t B, int64_t io) {
  RGCN_CHECK_ARG(comp && a && b && R >= 1 && B >= 1 && R <= kMaxCombine && B <= kMaxCombine && io >= 4 && io % 4 == 0,
                 "basis_combine: need 1 <= R, B <= %d and in * out a positive multiple of 4", kMaxCombine);
  RGCN_CHECK_ARG((((uintptr_t)a | (uintptr_t)b) & 15) == 0, "basis_combine: matrices must be 16-byte aligned");
  return RGCN_OK;
}

extern "C" int rgcn_basis_combine(const float* comp, const float* V, int32_t R, int32_t B, int64_t in_out, float* W,
                                  rgcn_stream_t stream) {
  int rc = check_combine(comp, V, W, R, B, in_out);
  if (rc) return rc;
  const int64_t cols4 = in_out / 4;
  RGCN_CUDA(launch_pdl(combine_rows_kernel<false>, dim3((unsigned)((cols4 + 255) / 256)), dim3(256), (size_t)R * B * sizeof(float),
                       (cudaStream_t)stream, comp, (int)B, (int)R, (int)B, V, W, cols4));
  RGCN_LAUNCH_CHECK();
  return RGCN_OK;
}

extern "C" int rgcn_basis_combine_bwd(const float* comp, const float* V, const float* gW, int32_t R, int32_t B,
                                      int64_t in_out, float* g_V, float* g_comp, rgcn_stream_t stream) {
  int rc = check_combine(comp, V, gW, R, B, in_out);
  if (rc) return rc;
  RGCN_CHECK_ARG(!g_comp || B <= 16, "basis_combine_bwd: the coefficient gradient handles at most 16 bases");
  RGCN_CHECK_ARG(!g_V || ((uintptr_t)g_V & 15) == 0, "basis_combine_bwd: g_V must be 16-byte aligned");
  const int64_t cols4 = in_out / 4;
  if (g_V) {
    RGCN_CUDA(launch_pdl(combine_rows_kernel<true>, dim3((unsigned)((cols4 + 255) / 256)), dim3(256), (size_t)R * B * sizeof(float),
                         (cudaStream_t)stream, comp, (int)B, (int)B, (int)R, gW, g_V, cols4));
    RGCN_LAUNCH_CHECK();
  }
  if (g_comp) {
    RGCN_CUDA(launch_pdl(comp_grad_kernel, dim3((unsigned)R), dim3(256), 0, (cudaStream_t)stream, gW, V, (int)B, cols4, g_comp, (int)B));
    RGCN_LAUNCH_CHECK();
  }
  return RGCN_OK;
}
