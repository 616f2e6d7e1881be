// Peer-memory kernels of the destination-range partitioned path (BASELINE.json cfg5): the exchange
// steps between GPUs of one NVSwitch domain, done by loads / stores on peer-mapped buffers instead of collectives.
//
//   rgcn_p2p_push_rows     all-gather by push: this GPU's feature shard -> the same rows of every GPU's full matrix
//                          (layer-0 input, i.e. the embedding-table shard; later layers are pushed by the transform
//                          epilogue itself, see rgcn_transform_fwd)
//   rgcn_p2p_reduce_split  reduce-scatter by pull, fused with what consumes it: this GPU's rows of every GPU's
//                          full-length partial gradient are loaded over NVLink, summed in RANK ORDER (deterministic),
//                          added to the local root-term gradient, masked by ReLU / dropout and written as the bf16
//                          planes the tensor-core kernels read (+ column-sum partials), or as fp32.
// The reference has no distributed code; this is the scale-out of src/models/rgcn.py:123-128 and of its autograd
// backward (src/train.py:306).
#include "common.cuh"

namespace rgcn {

constexpr int kMaxPeers = 8;

struct PeerPtrs {
  float* p[kMaxPeers];
};
struct ConstPeerPtrs {
  const float* p[kMaxPeers];
};

__global__ void __launch_bounds__(256) push_rows_kernel(const float* __restrict__ src, int64_t ld_src, int64_t rows,
                                                        int cols, PeerPtrs dst, int n_dst, int64_t row0,
                                                        int64_t ld_dst) {
  pdl_enter();
  const int c4n = cols >> 2;
  const int64_t total = rows * c4n;
  constexpr int U = 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += stride * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < total) v[u] = ldg4(src + (i / c4n) * ld_src + (i % c4n) * 4);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < total) {
        const int64_t off = (row0 + i / c4n) * ld_dst + (i % c4n) * 4;
        for (int q = 0; q < n_dst; ++q) *reinterpret_cast<float4*>(dst.p[q] + off) = v[u];
      }
    }
  }
}

__device__ __forceinline__ void split4p(const float4& v, uint2& hi, uint2& lo) {
  __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
  hi.x = *reinterpret_cast<uint32_t*>(&h01);
  hi.y = *reinterpret_cast<uint32_t*>(&h23);
  const float hx = __uint_as_float(hi.x << 16), hy = __uint_as_float(hi.x & 0xffff0000u);
  const float hz = __uint_as_float(hi.y << 16), hw = __uint_as_float(hi.y & 0xffff0000u);
  __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - hx, v.y - hy), l23 = __floats2bfloat162_rn(v.z - hz, v.w - hw);
  lo.x = *reinterpret_cast<uint32_t*>(&l01);
  lo.y = *reinterpret_cast<uint32_t*>(&l23);
}

// value[r, c] = (extra[r, c] + part_0[row0 + r, c]) + part_1[row0 + r, c] + ...   (rank order, left to right)
// then the optional mask (zero where mask <= 0, times scale) and the outputs.  Same thread mapping and column-sum
// partial layout as split_planes_kernel (transform.cu), so the bias-gradient reduction is shared.
__global__ void __launch_bounds__(256) reduce_split_kernel(ConstPeerPtrs part, int n_part, int64_t row0, int64_t ld_part,
                                                           const float* __restrict__ extra, int64_t ld_extra,
                                                           const float* __restrict__ mask, int64_t ldm, float scale,
                                                           int64_t rows, int cols, float* __restrict__ out, int64_t ldo,
                                                           __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                           int64_t ldp, float* __restrict__ colsum_partial,
                                                           int64_t rows_per_block) {
  pdl_enter();
  __shared__ float4 red[256];
  const int tpr = cols >> 2;                       // threads per row (<= 256)
  const int rpp = 256 / tpr;                       // rows per pass
  const int c4 = threadIdx.x % tpr, rsub = threadIdx.x / tpr;
  const int64_t r_beg = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r_end = min(r_beg + rows_per_block, rows);
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rsub < rpp) {
    constexpr int U = 2;                           // rows in flight per thread (x n_part peer loads each)
    for (int64_t r0 = r_beg + rsub; r0 < r_end; r0 += (int64_t)U * rpp) {
      float4 v[U], m[U], pv[U][kMaxPeers];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t r = r0 + (int64_t)u * rpp;
        const bool ok = r < r_end;
        v[u] = (ok && extra) ? ldg4(extra + r * ld_extra + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (mask) m[u] = ok ? ldg4(mask + r * ldm + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q)
          if (q < n_part)
            pv[u][q] = ok ? __ldcg(reinterpret_cast<const float4*>(part.p[q] + (row0 + r) * ld_part + c4 * 4))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t r = r0 + (int64_t)u * rpp;
        if (r < r_end) {
#pragma unroll
          for (int q = 0; q < kMaxPeers; ++q)
            if (q < n_part) add4(v[u], pv[u][q]);
          if (mask) {
            if (!(m[u].x > 0.f)) v[u].x = 0.f;
            if (!(m[u].y > 0.f)) v[u].y = 0.f;
            if (!(m[u].z > 0.f)) v[u].z = 0.f;
            if (!(m[u].w > 0.f)) v[u].w = 0.f;
            v[u] = scale4(v[u], scale);
          }
          add4(cs, v[u]);
          if (out) *reinterpret_cast<float4*>(out + r * ldo + c4 * 4) = v[u];
          if (hi) {
            uint2 h, l;
            split4p(v[u], h, l);
            *reinterpret_cast<uint2*>(hi + r * ldp + c4 * 4) = h;
            if (lo) *reinterpret_cast<uint2*>(lo + r * ldp + c4 * 4) = l;
          }
        }
      }
    }
  }
  if (colsum_partial) {
    red[threadIdx.x] = cs;
    __syncthreads();
    if (rsub == 0) {
      float4 s = red[c4];
      for (int j = 1; j < rpp; ++j) add4(s, red[j * tpr + c4]);
      *reinterpret_cast<float4*>(colsum_partial + (size_t)blockIdx.x * cols + c4 * 4) = s;
    }
  }
}

// out[rows[c], :] = sum_q part_q[row0 + rows[c], :]  (rank order) for the listed rows only — the reduce-scatter of a gradient
// that is zero outside a short row list (the last layer of the partitioned encoder: the loss reads 2 * batch rows), so the
// exchange moves kilobytes instead of this rank's whole shard of every rank's buffer.  One warp per list entry; duplicates
// write the same value twice.
__global__ void __launch_bounds__(256) pull_rows_kernel(ConstPeerPtrs part, int n_part, int64_t row0, int64_t ld_part,
                                                        const int64_t* __restrict__ rows, int64_t n_list, int cols,
                                                        float* __restrict__ out, int64_t ldo) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int64_t c = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= n_list) return;
  const int64_t r = rows[c];
  for (int vi = lane; vi < (cols >> 2); vi += 32) {
    float4 v = __ldcg(reinterpret_cast<const float4*>(part.p[0] + (row0 + r) * ld_part + vi * 4));
    for (int q = 1; q < n_part; ++q) add4(v, __ldcg(reinterpret_cast<const float4*>(part.p[q] + (row0 + r) * ld_part + vi * 4)));
    *reinterpret_cast<float4*>(out + r * ldo + vi * 4) = v;
  }
}

// ---- all-reduce (average) of one flat fp32 buffer over the GPUs of one NVSwitch domain ----------------------------
// The data-parallel gradient exchange (every replica holds the whole cfg1-4 graph; the reference is single-device, its
// README lists multi-GPU as future work).  Two-shot over peer-mapped memory, no collective library, CUDA-graph capturable:
//   signal : after the producers of `in` (stream order) -> arrive flag (epoch) in every peer's flag block
//   reduce : every block waits for all arrive flags, then this rank's 1/P slice is summed over all ranks' `in` buffers in
//            RANK ORDER (every rank computes the same bits), scaled, and stored into the slice of EVERY rank's `out`
//   done   : done flag to every peer, then wait for everybody's: all slices of this rank's `out` have landed
// Flags are monotone epochs (device-side counter, advanced by the signal kernel, so a graph replay needs no host help).
// Waits are bounded (~seconds): a missing peer sets bit 1 of *status instead of hanging the GPU.
struct FlagPtrs {
  unsigned int* p[kMaxPeers];                    // every rank's flag block: arrive[kMaxPeers], done[kMaxPeers]
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// true when the flag reached `epoch` (wrap-safe comparison); false after ~4 s of polling
__device__ __forceinline__ bool wait_flag(const unsigned int* p, unsigned int epoch) {
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys(p) - epoch) < 0) {
    if (clock64() - t0 > 8000000000ll) return false;
    __nanosleep(64);
  }
  return true;
}

// which = 0: advance the epoch and raise the arrive flags; which = 1: raise the done flags, then wait for all of them
__global__ void __launch_bounds__(32) allreduce_flag_kernel(FlagPtrs flags, int n, int rank, unsigned int* epoch_ctr,
                                                            int which, int* status) {
  pdl_enter();                                   // everything before this launch in the stream is complete and flushed
  unsigned int e = 0;
  if (threadIdx.x == 0) {
    e = *epoch_ctr + (which == 0 ? 1u : 0u);
    if (which == 0) *epoch_ctr = e;
  }
  e = __shfl_sync(0xffffffffu, e, 0);
  __threadfence_system();
  const int q = threadIdx.x;
  if (q < n) st_release_sys(flags.p[q] + which * kMaxPeers + rank, e);
  if (which == 1 && q < n) {
    if (!wait_flag(flags.p[rank] + kMaxPeers + q, e) && status) atomicOr(status, 2);
  }
}

__global__ void __launch_bounds__(256) allreduce_reduce_kernel(ConstPeerPtrs in, PeerPtrs out, int n, int rank,
                                                               int64_t n_vec, float scale, const unsigned int* my_flags,
                                                               const unsigned int* epoch_ctr, int* status) {
  pdl_enter();
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (threadIdx.x < n) {
    const unsigned int e = *epoch_ctr;
    if (!wait_flag(my_flags + threadIdx.x, e)) { s_ok = 0; if (status) atomicOr(status, 2); }
  }
  __syncthreads();
  if (!s_ok) return;
  // this rank's slice, in units of float4
  const int64_t per = (n_vec + n - 1) / n;
  const int64_t v0 = per * rank, v1 = min(v0 + per, n_vec);
  constexpr int U = 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = v0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < v1; i0 += stride * U) {
    float4 pv[U][kMaxPeers];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
#pragma unroll
      for (int q = 0; q < kMaxPeers; ++q)
        if (q < n) pv[u][q] = i < v1 ? __ldcg(reinterpret_cast<const float4*>(in.p[q]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < v1) {
        float4 v = pv[u][0];
#pragma unroll
        for (int q = 1; q < kMaxPeers; ++q)
          if (q < n) add4(v, pv[u][q]);
        v = scale4(v, scale);
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q)
          if (q < n) *(reinterpret_cast<float4*>(out.p[q]) + i) = v;
      }
    }
  }
}

}  // namespace rgcn

using namespace rgcn;

extern "C" size_t rgcn_p2p_allreduce_flag_bytes(void) { return 2 * kMaxPeers * sizeof(unsigned int); }

extern "C" int rgcn_p2p_allreduce(const float* const* in_host, float* const* out_host, unsigned int* const* flags_host,
                                  int32_t n_ranks, int32_t rank, int64_t n_floats, float scale, unsigned int* epoch_counter,
                                  int32_t* status, rgcn_stream_t stream) {
  RGCN_CHECK_ARG(n_ranks >= 1 && n_ranks <= kMaxPeers && rank >= 0 && rank < n_ranks, "p2p_allreduce: between 1 and %d ranks", kMaxPeers);
  RGCN_CHECK_ARG(in_host && out_host && flags_host && epoch_counter, "p2p_allreduce: null argument");
  RGCN_CHECK_ARG(n_floats >= 0 && n_floats % 4 == 0, "p2p_allreduce: the buffer length must be a multiple of 4 floats");
  ConstPeerPtrs in{};
  PeerPtrs out{};
  FlagPtrs fl{};
  for (int q = 0; q < n_ranks; ++q) {
    RGCN_CHECK_ARG(in_host[q] && out_host[q] && flags_host[q] && (((uintptr_t)in_host[q] | (uintptr_t)out_host[q]) & 15) == 0 &&
                   ((uintptr_t)flags_host[q] & 3) == 0, "p2p_allreduce: buffer %d is null or misaligned", q);
    in.p[q] = in_host[q]; out.p[q] = out_host[q]; fl.p[q] = flags_host[q];
  }
  if (n_floats == 0) return RGCN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  RGCN_CUDA(launch_pdl(allreduce_flag_kernel, dim3(1), dim3(32), 0, st, fl, (int)n_ranks, (int)rank, epoch_counter, 0, (int*)status));
  RGCN_LAUNCH_CHECK();
  const int64_t n_vec = n_floats / 4;
  const int64_t slice = (n_vec + n_ranks - 1) / n_ranks;
  int64_t blocks = (slice + 256 * 2 - 1) / (256 * 2);
  const int64_t cap = (int64_t)sm_count() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  RGCN_CUDA(launch_pdl(allreduce_reduce_kernel, dim3((unsigned)blocks), dim3(256), 0, st, in, out, (int)n_ranks, (int)rank, n_vec,
                       scale, (const unsigned int*)fl.p[rank], (const unsigned int*)epoch_counter, (int*)status));
  RGCN_LAUNCH_CHECK();
  RGCN_CUDA(launch_pdl(allreduce_flag_kernel, dim3(1), dim3(32), 0, st, fl, (int)n_ranks, (int)rank, epoch_counter, 1, (int*)status));
  RGCN_LAUNCH_CHECK();
  return RGCN_OK;
}

extern "C" int rgcn_p2p_pull_rows(const float* const* part_host, int32_t n_part, int64_t row0, int64_t ld_part,
                                  const int64_t* rows, int64_t n_list, int64_t n_rows_out, int32_t cols, float* out, int64_t ldo,
                                  rgcn_stream_t stream) {
  RGCN_CHECK_ARG(cols >= 4 && cols % 4 == 0 && n_list >= 0 && n_rows_out >= 0, "p2p_pull_rows: cols=%d must be a positive multiple of 4", cols);
  RGCN_CHECK_ARG(n_part >= 1 && n_part <= kMaxPeers && part_host, "p2p_pull_rows: between 1 and %d partial buffers", kMaxPeers);
  RGCN_CHECK_ARG(ld_part % 4 == 0 && row0 >= 0 && rows && out && ((uintptr_t)out & 15) == 0 && ldo % 4 == 0, "p2p_pull_rows: bad buffers");
  ConstPeerPtrs pp{};
  for (int q = 0; q < n_part; ++q) {
    RGCN_CHECK_ARG(part_host[q] && ((uintptr_t)part_host[q] & 15) == 0, "p2p_pull_rows: partial %d is null or misaligned", q);
    pp.p[q] = part_host[q];
  }
  if (n_list == 0) return RGCN_OK;
  RGCN_CUDA(launch_pdl(pull_rows_kernel, dim3((unsigned)((n_list + 7) / 8)), dim3(256), 0, (cudaStream_t)stream, pp, (int)n_part, row0,
                       ld_part, rows, n_list, (int)cols, out, ldo));
  RGCN_LAUNCH_CHECK();
  return RGCN_OK;
}

extern "C" int rgcn_p2p_push_rows(const float* src, int64_t ld_src, int64_t rows, int32_t cols,
                                  float* const* dst_host, int32_t n_dst, int64_t row0, int64_t ld_dst,
                                  rgcn_stream_t stream) {
  RGCN_CHECK_ARG(rows >= 0 && cols >= 4 && cols % 4 == 0, "p2p_push_rows: cols=%d must be a positive multiple of 4", cols);
  RGCN_CHECK_ARG(n_dst >= 1 && n_dst <= kMaxPeers && dst_host, "p2p_push_rows: between 1 and %d destinations", kMaxPeers);
  RGCN_CHECK_ARG(src && ((uintptr_t)src & 15) == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0 && row0 >= 0,
                 "p2p_push_rows: buffers must be 16-byte aligned with ld %% 4 == 0");
  PeerPtrs d{};
  for (int q = 0; q < n_dst; ++q) {
    RGCN_CHECK_ARG(dst_host[q] && ((uintptr_t)dst_host[q] & 15) == 0, "p2p_push_rows: destination %d is null or misaligned", q);
    d.p[q] = dst_host[q];
  }
  if (rows == 0) return RGCN_OK;
  const int64_t total = rows * (cols / 4);
  int64_t blocks = (total + 256 * 4 - 1) / (256 * 4);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  RGCN_CUDA(launch_pdl(push_rows_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, src, ld_src, rows, cols, d, n_dst, row0, ld_dst));
  RGCN_LAUNCH_CHECK();
  return RGCN_OK;
}

extern "C" int rgcn_p2p_reduce_split(const float* const* part_host, int32_t n_part, int64_t row0, int64_t ld_part,
                                     const float* extra, int64_t ld_extra, const float* relu_mask, int64_t ldm,
                                     float mask_scale, int64_t rows, int32_t cols, float* out, int64_t ldo, void* hi,
                                     void* lo, int64_t ldp, float* colsum_partial, rgcn_stream_t stream) {
  RGCN_CHECK_ARG(rows >= 0 && cols >= 4 && cols % 4 == 0 && cols <= 1024, "p2p_reduce_split: cols=%d must be a multiple of 4 in [4, 1024]", cols);
  RGCN_CHECK_ARG(n_part >= 1 && n_part <= kMaxPeers && part_host, "p2p_reduce_split: between 1 and %d partial buffers", kMaxPeers);
  RGCN_CHECK_ARG(ld_part % 4 == 0 && row0 >= 0, "p2p_reduce_split: ld_part %% 4 == 0");
  RGCN_CHECK_ARG(!extra || (((uintptr_t)extra & 15) == 0 && ld_extra % 4 == 0), "p2p_reduce_split: extra misaligned");
  RGCN_CHECK_ARG(!relu_mask || (((uintptr_t)relu_mask & 15) == 0 && ldm % 4 == 0), "p2p_reduce_split: mask misaligned");
  RGCN_CHECK_ARG(out || hi, "p2p_reduce_split: no output");
  RGCN_CHECK_ARG(!out || (((uintptr_t)out & 15) == 0 && ldo % 4 == 0), "p2p_reduce_split: out misaligned");
  RGCN_CHECK_ARG(!hi || (((uintptr_t)hi & 7) == 0 && (!lo || ((uintptr_t)lo & 7) == 0) && ldp % 4 == 0),
                 "p2p_reduce_split: planes must be 8-byte aligned, ld %% 4 == 0");
  ConstPeerPtrs pp{};
  for (int q = 0; q < n_part; ++q) {
    RGCN_CHECK_ARG(part_host[q] && ((uintptr_t)part_host[q] & 15) == 0, "p2p_reduce_split: partial %d is null or misaligned", q);
    pp.p[q] = part_host[q];
  }
  if (rows == 0) return RGCN_OK;
  // same block decomposition as rgcn_split_planes, so rgcn_split_planes_blocks() sizes the column-sum partials
  const int64_t nb = rgcn_split_planes_blocks(rows, cols);
  const int rpp = 256 / (cols / 4) > 0 ? 256 / (cols / 4) : 1;
  const int64_t rows_per_block = ((rows + nb - 1) / nb + rpp - 1) / rpp * rpp;
  RGCN_CUDA(launch_pdl(reduce_split_kernel, dim3((unsigned)nb), dim3(256), 0, (cudaStream_t)stream, 
      pp, n_part, row0, ld_part, extra, ld_extra, relu_mask, ldm, mask_scale, rows, cols, out, ldo, (__nv_bfloat16*)hi,
      (__nv_bfloat16*)lo, ldp, colsum_partial, rows_per_block));
  RGCN_LAUNCH_CHECK();
  return RGCN_OK;
}
