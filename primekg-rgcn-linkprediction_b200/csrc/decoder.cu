// Fused DistMult decoder: row gathers node_embeddings[head], [tail] (reference
// src/models/rgcn.py:325-326) + LinkPredictor.forward (src/models/rgcn.py:207-211) in one kernel,
// and its backward.  One warp per (head, relation, tail) pair, 128-bit loads, shuffle reduction.
#include "common.cuh"

namespace rgcn {

__global__ void __launch_bounds__(256) distmult_fwd_kernel(const float* __restrict__ emb_h, int64_t ld_h,
                                                           const float* __restrict__ emb_t, int64_t ld_t,
                                                           const int64_t* __restrict__ head,
                                                           const int64_t* __restrict__ tail,
                                                           const int64_t* __restrict__ rel,
                                                           const float* __restrict__ rel_table,
                                                           const float* __restrict__ rel_rows,
                                                           const float* __restrict__ rel_scale, int64_t n_pairs,
                                                           int32_t d, float* __restrict__ score) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int64_t p = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (p >= n_pairs) return;
  const float* h = emb_h + (head ? head[p] : p) * ld_h;
  const float* t = emb_t + (tail ? tail[p] : p) * ld_t;
  const float* r = rel_rows ? rel_rows + p * d : rel_table + rel[p] * d;
  float s = 0.f;
  for (int vi = lane; vi < (d >> 2); vi += 32) {
    const float4 a = ldg4(h + vi * 4), c = ldg4(t + vi * 4);
    float4 b = ldg4(r + vi * 4);
    if (rel_scale) {   // dropout on the relation row: mask / (1 - p), drawn by the caller
      const float4 m = ldg4(rel_scale + p * d + vi * 4);
      b.x *= m.x; b.y *= m.y; b.z *= m.z; b.w *= m.w;
    }
    // (h * r) * t, summed in element order inside the lane like torch.sum's pairwise tree is not
    // reproduced bit for bit; the tolerance of the parity tests covers the reduction order
    s += a.x * b.x * c.x;
    s += a.y * b.y * c.y;
    s += a.z * b.z * c.z;
    s += a.w * b.w * c.w;
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) score[p] = s;
}

__global__ void __launch_bounds__(256) distmult_bwd_kernel(const float* __restrict__ emb_h, int64_t ld_h,
                                                           const float* __restrict__ emb_t, int64_t ld_t,
                                                           const int64_t* __restrict__ head,
                                                           const int64_t* __restrict__ tail,
                                                           const int64_t* __restrict__ rel,
                                                           const float* __restrict__ rel_table,
                                                           const float* __restrict__ rel_rows,
                                                           const float* __restrict__ rel_scale,
                                                           const float* __restrict__ g_score, int64_t n_pairs,
                                                           int32_t d, float* __restrict__ g_h, int64_t ld_gh,
                                                           float* __restrict__ g_t, int64_t ld_gt,
                                                           float* __restrict__ g_rel_table,
                                                           float* __restrict__ g_rel_rows) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int64_t p = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (p >= n_pairs) return;
  const int64_t hi = head ? head[p] : p, ti = tail ? tail[p] : p;
  const float* h = emb_h + hi * ld_h;
  const float* t = emb_t + ti * ld_t;
  const float* r = rel_rows ? rel_rows + p * d : rel_table + rel[p] * d;
  const float g = g_score[p];
  for (int vi = lane; vi < (d >> 2); vi += 32) {
    const float4 a = ldg4(h + vi * 4), c = ldg4(t + vi * 4);
    float4 b = ldg4(r + vi * 4);
    float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
    if (rel_scale) {
      m = ldg4(rel_scale + p * d + vi * 4);
      b.x *= m.x; b.y *= m.y; b.z *= m.z; b.w *= m.w;
    }
    const float4 gh = make_float4(g * b.x * c.x, g * b.y * c.y, g * b.z * c.z, g * b.w * c.w);
    const float4 gt = make_float4(g * a.x * b.x, g * a.y * b.y, g * a.z * b.z, g * a.w * b.w);
    const float4 gr = make_float4(g * a.x * c.x * m.x, g * a.y * c.y * m.y, g * a.z * c.z * m.z, g * a.w * c.w * m.w);
    // gathered rows may repeat => fp32 atomics; identity rows are written exactly once => plain stores
    if (head) atomicAdd(reinterpret_cast<float4*>(g_h + hi * ld_gh + vi * 4), gh);
    else *reinterpret_cast<float4*>(g_h + hi * ld_gh + vi * 4) = gh;
    if (tail) atomicAdd(reinterpret_cast<float4*>(g_t + ti * ld_gt + vi * 4), gt);
    else *reinterpret_cast<float4*>(g_t + ti * ld_gt + vi * 4) = gt;
    if (g_rel_rows) *reinterpret_cast<float4*>(g_rel_rows + p * d + vi * 4) = gr;
    if (g_rel_table) atomicAdd(reinterpret_cast<float4*>(g_rel_table + rel[p] * d + vi * 4), gr);
  }
}

// loss = mean_i [ max(x,0) - x*y + log1p(exp(-|x|)) ]   (= torch BCEWithLogitsLoss, reference src/train.py:139, :300)
// one block, fixed-order tree => deterministic; also counts sigmoid(x) > 0.5 == y (the accuracy of src/train.py:321-322)
__global__ void __launch_bounds__(1024) bce_logits_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                              int64_t n, float* __restrict__ loss,
                                                              int32_t* __restrict__ n_correct) {
  pdl_enter();
  __shared__ float sl[32];
  __shared__ int sc[32];
  float acc = 0.f;
  int cor = 0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) {
    const float v = x[i], t = y[i];
    acc += fmaxf(v, 0.f) - v * t + log1pf(expf(-fabsf(v)));
    cor += ((v > 0.f) ? 1.f : 0.f) == t;
  }
  for (int o = 16; o; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    cor += __shfl_xor_sync(0xffffffffu, cor, o);
  }
  if ((threadIdx.x & 31) == 0) { sl[threadIdx.x >> 5] = acc; sc[threadIdx.x >> 5] = cor; }
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = sl[threadIdx.x]; cor = sc[threadIdx.x];
    for (int o = 16; o; o >>= 1) {
      acc += __shfl_xor_sync(0xffffffffu, acc, o);
      cor += __shfl_xor_sync(0xffffffffu, cor, o);
    }
    if (threadIdx.x == 0) {
      *loss = acc / (float)n;
      if (n_correct) *n_correct = cor;
    }
  }
}

// g_x[i] = g_loss * (sigmoid(x_i) - y_i) / n
__global__ void bce_logits_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t n,
                                      const float* __restrict__ g_loss, float* __restrict__ g_x) {
  pdl_enter();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = 1.f / (1.f + expf(-x[i]));
  g_x[i] = (*g_loss) * (s - y[i]) / (float)n;
}

// 1 when any index is out of range
__global__ void check_pairs_kernel(const int64_t* head, const int64_t* tail, const int64_t* rel, int64_t n_pairs,
                                   int64_t n_nodes, int32_t n_rel, int32_t* flag) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  if (head[p] < 0 || head[p] >= n_nodes || tail[p] < 0 || tail[p] >= n_nodes || (rel && (rel[p] < 0 || rel[p] >= n_rel)))
    atomicOr(flag, 1);
}

}  // namespace rgcn

using namespace rgcn;

static int check_dm(const float* emb_h, int64_t ld_h, const float* emb_t, int64_t ld_t, const int64_t* rel,
                    const float* rel_table, const float* rel_rows, int64_t n_pairs, int32_t d) {
  RGCN_CHECK_ARG(n_pairs >= 0 && d >= 4 && d % 4 == 0, "distmult: d=%d must be a positive multiple of 4", d);
  RGCN_CHECK_ARG(n_pairs == 0 || (emb_h && emb_t), "distmult: null embeddings");
  RGCN_CHECK_ARG(ld_h % 4 == 0 && ld_t % 4 == 0 && (((uintptr_t)emb_h | (uintptr_t)emb_t) & 15) == 0,
                 "distmult: embeddings must be 16-byte aligned rows");
  RGCN_CHECK_ARG(rel_rows || (rel_table && rel), "distmult: need rel_rows or (rel_table, rel)");
  RGCN_CHECK_ARG((((uintptr_t)rel_rows | (uintptr_t)rel_table) & 15) == 0, "distmult: relation rows must be 16-byte aligned");
  return RGCN_OK;
}
#define CHECK_SCALE(s) RGCN_CHECK_ARG((((uintptr_t)(s)) & 15) == 0, "distmult: rel_scale must be 16-byte aligned")

extern "C" int rgcn_distmult_fwd(const float* emb_h, int64_t ld_h, const float* emb_t, int64_t ld_t,
                                 const int64_t* head, const int64_t* tail, const int64_t* rel,
                                 const float* rel_table, const float* rel_rows, const float* rel_scale,
                                 int64_t n_pairs, int32_t d, float* score, rgcn_stream_t stream) {
  int rc = check_dm(emb_h, ld_h, emb_t, ld_t, rel, rel_table, rel_rows, n_pairs, d);
  if (rc) return rc;
  if (n_pairs == 0) return RGCN_OK;
  CHECK_SCALE(rel_scale);
  RGCN_CHECK_ARG(score, "distmult_fwd: null output");
  RGCN_CUDA(launch_pdl(distmult_fwd_kernel, dim3((unsigned)((n_pairs + 7) / 8)), dim3(256), 0, (cudaStream_t)stream, 
      emb_h, ld_h, emb_t, ld_t, head, tail, rel, rel_table, rel_rows, rel_scale, n_pairs, d, score));
  RGCN_LAUNCH_CHECK();
  return RGCN_OK;
}

extern "C" int rgcn_distmult_bwd(const float* emb_h, int64_t ld_h, const float* emb_t, int64_t ld_t,
                                 const int64_t* head, const int64_t* tail, const int64_t* rel,
                                 const float* rel_table, const float* rel_rows, const float* rel_scale,
                                 const float* g_score, int64_t n_pairs, int32_t d, float* g_h, int64_t ld_gh, float* g_t, int64_t ld_gt,
                                 float* g_rel_table, float* g_rel_rows, rgcn_stream_t stream) {
  int rc = check_dm(emb_h, ld_h, emb_t, ld_t, rel, rel_table, rel_rows, n_pairs, d);
  if (rc) return rc;
  if (n_pairs == 0) return RGCN_OK;
  RGCN_CHECK_ARG(g_score && g_h && g_t && ld_gh % 4 == 0 && ld_gt % 4 == 0 &&
                 (((uintptr_t)g_h | (uintptr_t)g_t) & 15) == 0, "distmult_bwd: bad gradient buffers");
  CHECK_SCALE(rel_scale);
  RGCN_CHECK_ARG(!g_rel_table || rel, "distmult_bwd: g_rel_table needs rel");
  RGCN_CUDA(launch_pdl(distmult_bwd_kernel, dim3((unsigned)((n_pairs + 7) / 8)), dim3(256), 0, (cudaStream_t)stream, 
      emb_h, ld_h, emb_t, ld_t, head, tail, rel, rel_table, rel_rows, rel_scale, g_score, n_pairs, d, g_h, ld_gh, g_t, ld_gt,
      g_rel_table, g_rel_rows));
  RGCN_LAUNCH_CHECK();
  return RGCN_OK;
}
extern "C" int rgcn_check_pairs(const int64_t* head, const int64_t* tail, const int64_t* rel, int64_t n_pairs,
                                int64_t n_nodes, int32_t n_rel, int32_t* flag, rgcn_stream_t stream) {
  RGCN_CHECK_ARG(flag && (n_pairs == 0 || (head && tail)), "check_pairs: null argument");
  RGCN_CUDA(cudaMemsetAsync(flag, 0, sizeof(int32_t), (cudaStream_t)stream));
  if (n_pairs == 0) return RGCN_OK;
  check_pairs_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(head, tail, rel, n_pairs,
                                                                                          n_nodes, n_rel, flag);
  RGCN_LAUNCH_CHECK();
  return RGCN_OK;
}

extern "C" int rgcn_bce_logits_fwd(const float* logits, const float* labels, int64_t n, float* loss, int32_t* n_correct,
                                   rgcn_stream_t stream) {
  RGCN_CHECK_ARG(n > 0 && logits && labels && loss, "bce_logits_fwd: bad arguments");
  RGCN_CUDA(launch_pdl(bce_logits_fwd_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, logits, labels, n, loss, n_correct));
  RGCN_LAUNCH_CHECK();
  return RGCN_OK;
}

extern "C" int rgcn_bce_logits_bwd(const float* logits, const float* labels, int64_t n, const float* g_loss,
                                   float* g_logits, rgcn_stream_t stream) {
  RGCN_CHECK_ARG(n > 0 && logits && labels && g_loss && g_logits, "bce_logits_bwd: bad arguments");
  RGCN_CUDA(launch_pdl(bce_logits_bwd_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, logits, labels, n, g_loss, g_logits));
  RGCN_LAUNCH_CHECK();
  return RGCN_OK;
}
