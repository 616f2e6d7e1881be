// Row-sparse output gradient of the LAST encoder layer.
//
// The training step of the reference (src/train.py:291-306) scores 2 * batch (head, tail) rows of the encoder output
// (src/models/rgcn.py:325-326), so autograd hands the last RGCNConv an [N, d_out] output gradient that is zero outside
// the <= 2 * batch listed rows (cfg2: <= 4,096 of 30,926).  PyTorch / PyG propagate that matrix densely.  Here the
// listed rows are compacted and the layer's backward runs on the compact matrices:
//
//   slot[i]   = first position c of node i in the row list, or m_c (the zero row) when i is not listed
//   Gc[c]     = g_out[rows[c]] as bf16 planes when slot[rows[c]] == c (duplicates keep ONE copy), else zeros
//   Ac[c]     = saved operand planes A[rows[c]] for the same rows, else zeros
//   gAc       = Gc @ [W; root]^T        (dgrad over m_c rows instead of N)
//   g_x[j]    = gAc[slot[j], root block] + sum_r sum_{(j -> i, r)} w_e * gAc[slot[i], r block]      (absent edges skipped)
//   gW, groot = Ac^T @ Gc, g_bias = column sums of Gc                                        (wgrad over m_c rows)
//
// Rows that are absent contribute exact zeros to every sum of the dense formulation, so the results are the same
// numbers; only the zero work is gone.
#include "common.cuh"

namespace rgcn {

// slot[i] = m_c for every node; the zero row (row m_c) of gAc is cleared in the same launch
__global__ void __launch_bounds__(256) rows_slot_fill_kernel(int32_t* __restrict__ slot, int64_t n, int32_t m_c,
                                                             float* __restrict__ zero_row, int32_t zero_cols) {
  pdl_enter();
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) slot[i] = m_c;
  if (zero_row && i < zero_cols) zero_row[i] = 0.f;
}

// slot[rows[c]] = min over the positions c that list the node (deterministic winner among duplicates)
__global__ void __launch_bounds__(256) rows_slot_min_kernel(const int64_t* __restrict__ rows, int64_t n_list,
                                                            int32_t* __restrict__ slot) {
  pdl_enter();
  const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (c < n_list) atomicMin(slot + rows[c], (int32_t)c);
}

// rows[c] = head[c] (c < n) or tail[c - n], an out-of-range index parked on row 0 (which it never owns);
// slot[rows[c]] = min over the valid positions that list the node
__global__ void __launch_bounds__(256) rows_list_kernel(const int64_t* __restrict__ head, const int64_t* __restrict__ tail,
                                                        int64_t n, int64_t n_nodes, int64_t* __restrict__ rows,
                                                        int32_t* __restrict__ slot) {
  pdl_enter();
  const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (c >= 2 * n) return;
  const int64_t v = c < n ? head[c] : tail[c - n];
  const bool ok = v >= 0 && v < n_nodes;
  rows[c] = ok ? v : 0;
  if (ok) atomicMin(slot + v, (int32_t)c);
}

// 8 warps per block, one compact row per warp: G planes from g_out, operand planes copied, column sums of G per block
constexpr int kRowsPerBlock = 8;

__device__ __forceinline__ void split4_rs(const float4& v, uint2& hi, uint2& lo) {
  __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
  hi.x = *reinterpret_cast<uint32_t*>(&h01);
  hi.y = *reinterpret_cast<uint32_t*>(&h23);
  const float hx = __uint_as_float(hi.x << 16), hy = __uint_as_float(hi.x & 0xffff0000u);
  const float hz = __uint_as_float(hi.y << 16), hw = __uint_as_float(hi.y & 0xffff0000u);
  __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - hx, v.y - hy), l23 = __floats2bfloat162_rn(v.z - hz, v.w - hw);
  lo.x = *reinterpret_cast<uint32_t*>(&l01);
  lo.y = *reinterpret_cast<uint32_t*>(&l23);
}

struct CompactParams {
  const int64_t* rows; int64_t n_list; const int32_t* slot; int32_t m_c;
  const float* g_out; int64_t ld_g_out; int32_t d_out;
  __nv_bfloat16* G_hi; __nv_bfloat16* G_lo; int64_t ldg;
  const __nv_bfloat16* A_hi; const __nv_bfloat16* A_lo; int64_t lda; int32_t K;      // K % 4 == 0
  __nv_bfloat16* Ac_hi; __nv_bfloat16* Ac_lo; int64_t ldac;                          // nullable: no weight gradient
  float* colsum_partial;                                                              // nullable, [gridDim.x, d_out]
  float* zero_row; int32_t zero_cols;                                                 // nullable: cleared by block 0
};

__global__ void __launch_bounds__(256) rows_compact_kernel(const CompactParams p) {
  pdl_enter();
  extern __shared__ float4 s_cs[];                // [8][d_out / 4]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (p.zero_row && blockIdx.x == 0)
    for (int i = threadIdx.x; i < p.zero_cols; i += 256) p.zero_row[i] = 0.f;
  const int nv = p.d_out >> 2, nk = p.K >> 2;     // float4 of G per row, 8-byte groups of 4 bf16 per plane row
  const int c = blockIdx.x * kRowsPerBlock + warp;
  int64_t node = -1;
  if (c < p.n_list) {
    node = p.rows[c];
    if (__ldg(p.slot + node) != c) node = -1;     // a duplicate of an earlier position: its gradient is in that row
  }
  if (c < p.m_c) {
    if (node >= 0) {
      const float* __restrict__ g = p.g_out + node * p.ld_g_out;
#pragma unroll 2
      for (int vi = lane; vi < nv; vi += 32) {
        const float4 v = ldg4(g + vi * 4);
        uint2 h, l;
        split4_rs(v, h, l);
        *reinterpret_cast<uint2*>(p.G_hi + (int64_t)c * p.ldg + vi * 4) = h;
        if (p.G_lo) *reinterpret_cast<uint2*>(p.G_lo + (int64_t)c * p.ldg + vi * 4) = l;
        if (p.colsum_partial) s_cs[warp * nv + vi] = v;
      }
      if (p.Ac_hi) {
        const uint2* __restrict__ ah = reinterpret_cast<const uint2*>(p.A_hi + node * p.lda);
        const uint2* __restrict__ al = p.A_lo ? reinterpret_cast<const uint2*>(p.A_lo + node * p.lda) : nullptr;
        uint2* __restrict__ ch = reinterpret_cast<uint2*>(p.Ac_hi + (int64_t)c * p.ldac);
        uint2* __restrict__ cl = p.Ac_lo ? reinterpret_cast<uint2*>(p.Ac_lo + (int64_t)c * p.ldac) : nullptr;
#pragma unroll 4
        for (int ki = lane; ki < nk; ki += 32) {
          const uint2 h = __ldg(ah + ki);
          uint2 l = make_uint2(0u, 0u);
          if (al) l = __ldg(al + ki);
          ch[ki] = h;
          if (cl) cl[ki] = l;
        }
      }
    } else {
      // padding row or duplicate: zeros (stale planes could hold NaN patterns; 0 * NaN would poison the products)
      const uint2 z = make_uint2(0u, 0u);
      for (int vi = lane; vi < nv; vi += 32) {
        *reinterpret_cast<uint2*>(p.G_hi + (int64_t)c * p.ldg + vi * 4) = z;
        if (p.G_lo) *reinterpret_cast<uint2*>(p.G_lo + (int64_t)c * p.ldg + vi * 4) = z;
        if (p.colsum_partial) s_cs[warp * nv + vi] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (p.Ac_hi)
        for (int ki = lane; ki < nk; ki += 32) {
          *reinterpret_cast<uint2*>(p.Ac_hi + (int64_t)c * p.ldac + ki * 4) = z;
          if (p.Ac_lo) *reinterpret_cast<uint2*>(p.Ac_lo + (int64_t)c * p.ldac + ki * 4) = z;
        }
    }
  } else if (p.colsum_partial) {
    for (int vi = lane; vi < nv; vi += 32) s_cs[warp * nv + vi] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (p.colsum_partial) {
    __syncthreads();
    for (int vi = threadIdx.x; vi < nv; vi += 256) {
      float4 s = s_cs[vi];
      for (int w = 1; w < 8; ++w) add4(s, s_cs[w * nv + vi]);                        // fixed order: deterministic
      *reinterpret_cast<float4*>(p.colsum_partial + (size_t)blockIdx.x * p.d_out + vi * 4) = s;
    }
  }
}

}  // namespace rgcn

using namespace rgcn;

extern "C" int64_t rgcn_rows_compact_size(int64_t n_list) { return n_list <= 0 ? 0 : (n_list + 127) / 128 * 128; }
extern "C" int64_t rgcn_rows_compact_blocks(int64_t n_list) {
  return (rgcn_rows_compact_size(n_list) + kRowsPerBlock - 1) / kRowsPerBlock;
}

extern "C" int rgcn_rows_compact(const int64_t* rows, int64_t n_list, int64_t n_nodes, int32_t* slot,
                                 const float* g_out, int64_t ld_g_out, int32_t d_out, void* G_hi, void* G_lo, int64_t ldg,
                                 const void* A_hi, const void* A_lo, int64_t lda, int32_t K, void* Ac_hi, void* Ac_lo,
                                 int64_t ldac, float* colsum_partial, float* zero_row, int32_t zero_cols,
                                 int32_t slot_ready, rgcn_stream_t stream) {
  const int64_t m_c = rgcn_rows_compact_size(n_list);
  RGCN_CHECK_ARG(rows && n_list > 0 && n_nodes > 0 && slot && m_c < (1ll << 30), "rows_compact: bad row list");
  RGCN_CHECK_ARG(g_out && ((uintptr_t)g_out & 15) == 0 && ld_g_out % 4 == 0 && d_out >= 4 && d_out % 4 == 0 && d_out <= 1024,
                 "rows_compact: g_out must be 16-byte aligned, d_out a multiple of 4 up to 1024");
  RGCN_CHECK_ARG(G_hi && ((uintptr_t)G_hi & 7) == 0 && (!G_lo || ((uintptr_t)G_lo & 7) == 0) && ldg % 4 == 0,
                 "rows_compact: G planes must be 8-byte aligned, ld %% 4 == 0");
  RGCN_CHECK_ARG(!Ac_hi || (A_hi && K % 4 == 0 && lda % 4 == 0 && ldac % 4 == 0 && (((uintptr_t)A_hi | (uintptr_t)Ac_hi) & 7) == 0 &&
                            (!Ac_lo || (A_lo && (((uintptr_t)A_lo | (uintptr_t)Ac_lo) & 7) == 0))),
                 "rows_compact: operand planes must be 8-byte aligned with K, ld %% 4 == 0");
  cudaStream_t st = (cudaStream_t)stream;
  if (!slot_ready) {
    // (slot_ready: the decoder's deterministic backward, rgcn_link_loss_bwd_rows, has built slot for this very list)
    RGCN_CUDA(launch_pdl(rows_slot_fill_kernel, dim3((unsigned)((n_nodes + 255) / 256)), dim3(256), 0, st, slot, n_nodes,
                         (int32_t)m_c, (float*)nullptr, 0));
    RGCN_LAUNCH_CHECK();
    RGCN_CUDA(launch_pdl(rows_slot_min_kernel, dim3((unsigned)((n_list + 255) / 256)), dim3(256), 0, st, rows, n_list, slot));
    RGCN_LAUNCH_CHECK();
  }
  CompactParams p{};
  p.rows = rows; p.n_list = n_list; p.slot = slot; p.m_c = (int32_t)m_c;
  p.g_out = g_out; p.ld_g_out = ld_g_out; p.d_out = d_out;
  p.G_hi = (__nv_bfloat16*)G_hi; p.G_lo = (__nv_bfloat16*)G_lo; p.ldg = ldg;
  p.A_hi = (const __nv_bfloat16*)A_hi; p.A_lo = (const __nv_bfloat16*)A_lo; p.lda = lda; p.K = K;
  p.Ac_hi = (__nv_bfloat16*)Ac_hi; p.Ac_lo = (__nv_bfloat16*)Ac_lo; p.ldac = ldac;
  p.colsum_partial = colsum_partial;
  p.zero_row = zero_row; p.zero_cols = zero_cols;
  const size_t smem = (size_t)8 * (d_out / 4) * sizeof(float4);
  RGCN_CUDA(launch_pdl(rows_compact_kernel, dim3((unsigned)rgcn_rows_compact_blocks(n_list)), dim3(256), smem, st, p));
  RGCN_LAUNCH_CHECK();
  return RGCN_OK;
}

// Row list of a link-prediction batch for the listed-rows FORWARD of the last layer (rgcn_layer_fwd, rows != NULL):
// rows [2 n] = heads then tails (the order rgcn_link_loss_bwd_rows uses), slot [n_nodes] = node -> first position or m_c.
extern "C" int rgcn_rows_list_build(const int64_t* head, const int64_t* tail, int64_t n_pairs, int64_t n_nodes,
                                    int64_t* rows, int32_t* slot, rgcn_stream_t stream) {
  RGCN_CHECK_ARG(head && tail && rows && slot && n_pairs > 0 && n_nodes > 0 && 2 * n_pairs < (1ll << 30), "rows_list_build: bad arguments");
  const int64_t m_c = rgcn_rows_compact_size(2 * n_pairs);
  cudaStream_t st = (cudaStream_t)stream;
  RGCN_CUDA(launch_pdl(rows_slot_fill_kernel, dim3((unsigned)((n_nodes + 255) / 256)), dim3(256), 0, st, slot, n_nodes,
                       (int32_t)m_c, (float*)nullptr, 0));
  RGCN_LAUNCH_CHECK();
  RGCN_CUDA(launch_pdl(rows_list_kernel, dim3((unsigned)((2 * n_pairs + 255) / 256)), dim3(256), 0, st, head, tail, n_pairs, n_nodes,
                       rows, slot));
  RGCN_LAUNCH_CHECK();
  return RGCN_OK;
}
