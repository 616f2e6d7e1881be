// Graph preprocessing: (dst, relation)-keyed CSR + (src, relation)-keyed transposed CSR.
// Replaces the per-call boolean masks `edge_index[:, edge_type == r]` of RGCNConv's loop path
// (reference call sites src/models/rgcn.py:123, :128; input format src/preprocess.py:240-261).
// Stable LSD radix sort (cub) of key = major * R + rel with the edge id as payload, so the order
// inside a key is the original edge order — bit-identical to a stable sort (oracle/rgcn_ref.py).
#include <cub/cub.cuh>

#include "common.cuh"

namespace rgcn {

__global__ void make_keys_kernel(const int64_t* __restrict__ major, const int64_t* __restrict__ minor,
                                 const int64_t* __restrict__ rel, int64_t E, int64_t n_major, int64_t n_minor,
                                 int32_t R, int major_bit, int minor_bit,
                                 uint32_t* __restrict__ keys, int32_t* __restrict__ vals,
                                 int32_t* __restrict__ hist, int32_t* __restrict__ status) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
    const int64_t a = major[e], b = minor[e], r = rel[e];
    int bad = 0;
    if (a < 0 || a >= n_major) bad |= major_bit;
    if (b < 0 || b >= n_minor) bad |= minor_bit;
    if (r < 0 || r >= R) bad |= 4;
    uint32_t k = 0;
    if (bad) {
      atomicOr(status, bad);
    } else {
      k = (uint32_t)(a * R + r);
      atomicAdd(hist + k, 1);
    }
    keys[e] = k;
    vals[e] = (int32_t)e;
  }
}

__global__ void gather_minor_kernel(const int64_t* __restrict__ minor, const int32_t* __restrict__ perm, int64_t E,
                                    int32_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < E; p += stride)
    out[p] = (int32_t)minor[perm[p]];
}

// inv_cnt[k] = 1 / max(cnt, 1); also segment statistics
__global__ void seg_stats_kernel(const int32_t* __restrict__ rowptr, int64_t n_keys, float* __restrict__ inv_cnt,
                                 int32_t* __restrict__ max_len, int32_t* __restrict__ n_hub) {
  int local_max = 0, local_hub = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_keys; k += stride) {
    const int c = rowptr[k + 1] - rowptr[k];
    if (inv_cnt) inv_cnt[k] = 1.0f / (float)max(c, 1);
    local_max = max(local_max, c);
    local_hub += (c > kHubThreshold);
  }
  for (int o = 16; o; o >>= 1) {
    local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    local_hub += __shfl_xor_sync(0xffffffffu, local_hub, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (local_max) atomicMax(max_len, local_max);
    if (local_hub && n_hub) atomicAdd(n_hub, local_hub);
  }
}

__global__ void edge_weight_t_kernel(const uint32_t* __restrict__ keys_t_sorted, const int32_t* __restrict__ row_t,
                                     const float* __restrict__ inv_cnt, int64_t E, int32_t R, float* __restrict__ w_t) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < E; p += stride) {
    const int r = (int)(keys_t_sorted[p] % (uint32_t)R);
    w_t[p] = inv_cnt[(int64_t)row_t[p] * R + r];
  }
}

struct WsLayout {
  size_t keys_in, keys_out, vals_in, cub_tmp, total;
  size_t cub_bytes;
};

static WsLayout ws_layout(int64_t E, int64_t n_dst, int64_t n_src, int32_t R) {
  WsLayout w{};
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)E, 0, 32);
  const int64_t nk = (n_dst > n_src ? n_dst : n_src) * R + 1;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (int32_t*)nullptr, (int32_t*)nullptr, (int)nk);
  w.cub_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
  size_t off = 0;
  w.keys_in = off;  off += align_up((size_t)E * 4, 256);
  w.keys_out = off; off += align_up((size_t)E * 4, 256);
  w.vals_in = off;  off += align_up((size_t)E * 4, 256);
  w.cub_tmp = off;  off += align_up(w.cub_bytes, 256);
  w.total = off + 256;
  return w;
}

static int grid_for(int64_t blocks) {
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

static int bits_for(uint64_t n) {  // bits needed to represent values in [0, n)
  int b = 1;
  while (b < 32 && (1ull << b) < n) ++b;
  return b;
}

static int build_one(const int64_t* major, const int64_t* minor, const int64_t* rel, int64_t E, int64_t n_major,
                     int64_t n_minor, int32_t R, int major_bit, int minor_bit, int32_t* rowptr, int32_t* idx,
                     int32_t* perm, int32_t* status, char* ws, const WsLayout& L, cudaStream_t st) {
  const int64_t nk = n_major * R;
  uint32_t* keys_in = (uint32_t*)(ws + L.keys_in);
  uint32_t* keys_out = (uint32_t*)(ws + L.keys_out);
  int32_t* vals_in = (int32_t*)(ws + L.vals_in);
  RGCN_CUDA(cudaMemsetAsync(rowptr, 0, (size_t)(nk + 1) * sizeof(int32_t), st));
  if (E > 0) {
    const int grid = grid_for((E + 255) / 256);
    make_keys_kernel<<<grid, 256, 0, st>>>(major, minor, rel, E, n_major, n_minor, R, major_bit, minor_bit, keys_in,
                                           vals_in, rowptr, status);
    RGCN_LAUNCH_CHECK();
    size_t tmp = L.cub_bytes;
    RGCN_CUDA(cub::DeviceRadixSort::SortPairs(ws + L.cub_tmp, tmp, keys_in, keys_out, vals_in, perm, (int)E, 0,
                                              bits_for((uint64_t)nk), st));
    gather_minor_kernel<<<grid, 256, 0, st>>>(minor, perm, E, idx);
    RGCN_LAUNCH_CHECK();
  }
  size_t tmp = L.cub_bytes;
  RGCN_CUDA(cub::DeviceScan::ExclusiveSum(ws + L.cub_tmp, tmp, rowptr, rowptr, (int)(nk + 1), st));
  return RGCN_OK;
}

// ---- hub plan ---------------------------------------------------------------------------------
struct IsHub {
  const int32_t* rowptr;
  int threshold;
  __device__ bool operator()(int32_t k) const { return rowptr[k + 1] - rowptr[k] > threshold; }
};

__global__ void hub_chunks_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ hub_keys,
                                  const int32_t* __restrict__ n_hubs, int64_t cap, int32_t* __restrict__ chunk_cnt) {
  const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h > cap) return;
  int c = 0;
  if (h < *n_hubs) {
    const int k = hub_keys[h];
    c = (rowptr[k + 1] - rowptr[k] + kHubChunk - 1) / kHubChunk;
  }
  chunk_cnt[h] = c;
}

// one thread per hub segment: fills the table entries of its chunks
__global__ void hub_chunk_table_kernel(const int32_t* __restrict__ hub_keys, const int32_t* __restrict__ chunk_ptr,
                                       int n_hubs, int R, int4* __restrict__ table) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n_hubs) return;
  const int key = hub_keys[h], row = key / R;
  int first = h, end = h + 1;
  while (first > 0 && hub_keys[first - 1] / R == row) --first;
  while (end < n_hubs && hub_keys[end] / R == row) ++end;
  const int need = chunk_ptr[end] - chunk_ptr[first];
  const int c0 = chunk_ptr[h], c1 = chunk_ptr[h + 1];
  for (int c = c0; c < c1; ++c) table[c] = make_int4(key, c0, first, need);
}

}  // namespace rgcn

using namespace rgcn;

extern "C" int rgcn_hub_chunk_table(const int32_t* hub_keys, const int32_t* hub_chunk_ptr, int32_t n_hubs, int32_t R,
                                    int32_t* chunk_table, rgcn_stream_t stream) {
  RGCN_CHECK_ARG(n_hubs >= 0 && R >= 1, "hub_chunk_table: bad sizes");
  if (n_hubs == 0) return RGCN_OK;
  RGCN_CHECK_ARG(hub_keys && hub_chunk_ptr && chunk_table && ((uintptr_t)chunk_table & 15) == 0,
                 "hub_chunk_table: null or misaligned argument");
  hub_chunk_table_kernel<<<(unsigned)((n_hubs + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      hub_keys, hub_chunk_ptr, n_hubs, R, reinterpret_cast<int4*>(chunk_table));
  RGCN_LAUNCH_CHECK();
  return RGCN_OK;
}

extern "C" size_t rgcn_csr_build_workspace_bytes(int64_t E, int64_t n_dst, int64_t n_src, int32_t R) {
  if (E < 0 || n_dst < 0 || n_src < 0 || R < 1) return 0;
  return ws_layout(E, n_dst, n_src, R).total;
}

extern "C" int rgcn_csr_build(const int64_t* src, const int64_t* dst, const int64_t* rel, int64_t E, int64_t n_dst,
                              int64_t n_src, int32_t R, int32_t* rowptr, int32_t* col, int32_t* perm,
                              int32_t* rowptr_t, int32_t* row_t, int32_t* perm_t, float* inv_cnt, float* w_t,
                              int32_t* status, void* workspace, size_t workspace_bytes, rgcn_stream_t stream) {
  RGCN_CHECK_ARG(E >= 0 && n_dst >= 0 && n_src >= 0 && R >= 1, "csr_build: negative size");
  RGCN_CHECK_ARG(E < (1ll << 31) - 1, "csr_build: E=%lld does not fit int32 offsets", (long long)E);
  RGCN_CHECK_ARG(n_dst * R < (1ll << 31) - 1 && n_src * R < (1ll << 31) - 1,
                 "csr_build: n * R must fit int32 keys");
  RGCN_CHECK_ARG(E == 0 || (src && dst && rel), "csr_build: null edge arrays");
  RGCN_CHECK_ARG(rowptr && rowptr_t && status && inv_cnt, "csr_build: null outputs");
  RGCN_CHECK_ARG(E == 0 || (col && perm && row_t && perm_t && w_t), "csr_build: null outputs");
  const WsLayout L = ws_layout(E, n_dst, n_src, R);
  if (!workspace || workspace_bytes < L.total) {
    set_error("csr_build: workspace too small (%zu < %zu)", workspace_bytes, L.total);
    return RGCN_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)align_up((size_t)workspace, 256);
  RGCN_CUDA(cudaMemsetAsync(status, 0, 4 * sizeof(int32_t), st));
  // (dst, rel) CSR: major = dst (status bit 1), minor = src (status bit 0)
  int rc = build_one(dst, src, rel, E, n_dst, n_src, R, 2, 1, rowptr, col, perm, status, ws, L, st);
  if (rc) return rc;
  const int grid_k = grid_for((n_dst * R + 255) / 256 + 1);
  seg_stats_kernel<<<grid_k, 256, 0, st>>>(rowptr, n_dst * R, inv_cnt, status + 2, status + 1);
  RGCN_LAUNCH_CHECK();
  // (src, rel) transposed CSR
  rc = build_one(src, dst, rel, E, n_src, n_dst, R, 1, 2, rowptr_t, row_t, perm_t, status, ws, L, st);
  if (rc) return rc;
  const int grid_t = grid_for((n_src * R + 255) / 256 + 1);
  seg_stats_kernel<<<grid_t, 256, 0, st>>>(rowptr_t, n_src * R, nullptr, status + 3, nullptr);
  RGCN_LAUNCH_CHECK();
  if (E > 0) {
    const int grid = grid_for((E + 255) / 256);
    // keys_out still holds the sorted transposed keys
    edge_weight_t_kernel<<<grid, 256, 0, st>>>((const uint32_t*)(ws + L.keys_out), row_t, inv_cnt, E, R, w_t);
    RGCN_LAUNCH_CHECK();
  }
  return RGCN_OK;
}

extern "C" size_t rgcn_hub_plan_workspace_bytes(int64_t n_keys, int64_t cap_hubs) {
  size_t sel = 0, scan = 0;
  cub::CountingInputIterator<int32_t> it(0);
  IsHub pred{nullptr, kHubThreshold};
  cub::DeviceSelect::If(nullptr, sel, it, (int32_t*)nullptr, (int32_t*)nullptr, (int)n_keys, pred);
  cub::DeviceScan::ExclusiveSum(nullptr, scan, (int32_t*)nullptr, (int32_t*)nullptr, (int)(cap_hubs + 1));
  return align_up(sel > scan ? sel : scan, 256) + 512;
}

extern "C" int rgcn_hub_plan(const int32_t* rowptr, int64_t n_keys, int32_t hub_threshold, int32_t* hub_keys,
                             int32_t* hub_chunk_ptr, int64_t cap_hubs, int32_t* n_hubs_host, int32_t* n_chunks_host,
                             void* workspace, size_t workspace_bytes, rgcn_stream_t stream) {
  RGCN_CHECK_ARG(hub_threshold >= 1, "hub_plan: hub_threshold must be positive");
  RGCN_CHECK_ARG(rowptr && hub_keys && hub_chunk_ptr && n_hubs_host && n_chunks_host, "hub_plan: null argument");
  RGCN_CHECK_ARG(n_keys >= 0 && n_keys < (1ll << 31) - 1 && cap_hubs >= 0, "hub_plan: bad sizes");
  const size_t need = rgcn_hub_plan_workspace_bytes(n_keys, cap_hubs);
  if (!workspace || workspace_bytes < need) {
    set_error("hub_plan: workspace too small (%zu < %zu)", workspace_bytes, need);
    return RGCN_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)align_up((size_t)workspace, 256);
  int32_t* d_n = (int32_t*)ws;           // number of hubs (device)
  char* tmp = ws + 256;
  size_t tmp_bytes = need - 512;
  cub::CountingInputIterator<int32_t> it(0);
  IsHub pred{rowptr, hub_threshold};
  RGCN_CUDA(cudaMemsetAsync(d_n, 0, sizeof(int32_t), st));
  if (n_keys > 0 && cap_hubs > 0) {
    // keys come out in increasing order (DeviceSelect keeps the input order)
    RGCN_CUDA(cub::DeviceSelect::If(tmp, tmp_bytes, it, hub_keys, d_n, (int)n_keys, pred, st));
  }
  hub_chunks_kernel<<<(unsigned)((cap_hubs + 1 + 255) / 256), 256, 0, st>>>(rowptr, hub_keys, d_n, cap_hubs,
                                                                         hub_chunk_ptr);
  RGCN_LAUNCH_CHECK();
  tmp_bytes = need - 512;
  RGCN_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, hub_chunk_ptr, hub_chunk_ptr, (int)(cap_hubs + 1), st));
  RGCN_CUDA(cudaMemcpyAsync(n_hubs_host, d_n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  RGCN_CUDA(cudaMemcpyAsync(n_chunks_host, hub_chunk_ptr + cap_hubs, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  RGCN_CUDA(cudaStreamSynchronize(st));
  if (*n_hubs_host > cap_hubs) {
    set_error("hub_plan: %d hub segments exceed the capacity %lld", *n_hubs_host, (long long)cap_hubs);
    return RGCN_EINVAL;
  }
  return RGCN_OK;
}
