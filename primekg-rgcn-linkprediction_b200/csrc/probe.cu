// Bandwidth probe: the row-gather access pattern of the neighbourhood aggregation (csrc/aggregate.cu) with everything
// else removed — no CSR walk, no mean, no operand-plane stores.  bench.py runs it ON THE BOX THE BENCH RUNS ON to measure
// the ceilings the aggregation is judged against:
//   * table smaller than the L2 (cfg2: 30,926 x 256 fp32 = 31.7 MB), real or uniform source indices -> what the L2 can
//     deliver to the SMs for this gather (the "L2 peak" of the roofline block, instead of a figure quoted from a guide);
//   * idx == NULL: rows taken in order (a streaming read) -> L2 read bandwidth for an L2-sized table, HBM for a large one.
// A group of G = min(32, d / 4) lanes sums one row per step with U independent 128-bit loads in flight per lane, like the
// hot loop of aggregate_rows_kernel; each group finally stores one vector so nothing is optimised away.
#include "common.cuh"

namespace rgcn {

template <int G, int VPL>
__global__ void __launch_bounds__(256) probe_gather_kernel(const float* __restrict__ table, int64_t ld, int64_t n_rows,
                                                           const int32_t* __restrict__ idx, int64_t n_idx, int32_t d,
                                                           float* __restrict__ sink) {
  constexpr int GROUPS = 256 / G;
  constexpr int U = (VPL >= 4) ? 2 : (VPL == 2 ? 4 : 8);
  const int lane = threadIdx.x % G, grp = threadIdx.x / G;
  const int64_t n_groups = (int64_t)gridDim.x * GROUPS;
  const int64_t gid = (int64_t)blockIdx.x * GROUPS + grp;
  // contiguous slice of the index list per group (neighbouring edges of the CSR belong to one destination row)
  const int64_t per = (n_idx + n_groups - 1) / n_groups;
  const int64_t beg = gid * per, end = beg + per < n_idx ? beg + per : n_idx;
  const int nvec = d >> 2;
  int vcol[VPL];
  float4 acc[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int vi = k * G + lane;
    vcol[k] = vi < nvec ? vi * 4 : 0;
    acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int64_t e = beg;
  for (; e + U <= end; e += U) {
    float4 v[U][VPL];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t j = idx ? (int64_t)__ldg(idx + e + u) : (e + u) % n_rows;
      const float* __restrict__ rp = table + j * ld;
#pragma unroll
      for (int k = 0; k < VPL; ++k) v[u][k] = ldg4(rp + vcol[k]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < VPL; ++k) add4(acc[k], v[u][k]);
  }
  for (; e < end; ++e) {
    const int64_t j = idx ? (int64_t)__ldg(idx + e) : e % n_rows;
#pragma unroll
    for (int k = 0; k < VPL; ++k) add4(acc[k], ldg4(table + j * ld + vcol[k]));
  }
  float4 s = acc[0];
#pragma unroll
  for (int k = 1; k < VPL; ++k) add4(s, acc[k]);
  reinterpret_cast<float4*>(sink)[gid * G + lane] = s;
}

template <int G, int VPL>
static int launch_probe(const float* table, int64_t ld, int64_t n_rows, const int32_t* idx, int64_t n_idx, int32_t d,
                        float* sink, int blocks, cudaStream_t st) {
  probe_gather_kernel<G, VPL><<<blocks, 256, 0, st>>>(table, ld, n_rows, idx, n_idx, d, sink);
  RGCN_LAUNCH_CHECK();
  return RGCN_OK;
}

}  // namespace rgcn

using namespace rgcn;

extern "C" int64_t rgcn_probe_gather_sink_floats(int32_t blocks_per_sm) {
  return (int64_t)sm_count() * (blocks_per_sm > 0 ? blocks_per_sm : 1) * 256 * 4;
}

extern "C" int rgcn_probe_gather(const float* table, int64_t ld, int64_t n_rows, int32_t d, const int32_t* idx,
                                 int64_t n_idx, int32_t blocks_per_sm, float* sink, rgcn_stream_t stream) {
  RGCN_CHECK_ARG(table && sink && n_rows > 0 && n_idx >= 0 && ld % 4 == 0 && ((uintptr_t)table & 15) == 0 &&
                 ((uintptr_t)sink & 15) == 0, "probe_gather: bad arguments");
  RGCN_CHECK_ARG(d >= 16 && d % 4 == 0 && d <= 1024 && blocks_per_sm >= 1 && blocks_per_sm <= 8,
                 "probe_gather: d must be a multiple of 4 in [16, 1024], blocks_per_sm in [1, 8]");
  if (n_idx == 0) return RGCN_OK;
  const int blocks = sm_count() * blocks_per_sm;
  cudaStream_t st = (cudaStream_t)stream;
  const int nvec = d >> 2;
  if (nvec <= 4) return launch_probe<4, 1>(table, ld, n_rows, idx, n_idx, d, sink, blocks, st);
  if (nvec <= 8) return launch_probe<8, 1>(table, ld, n_rows, idx, n_idx, d, sink, blocks, st);
  if (nvec <= 16) return launch_probe<16, 1>(table, ld, n_rows, idx, n_idx, d, sink, blocks, st);
  if (nvec <= 32) return launch_probe<32, 1>(table, ld, n_rows, idx, n_idx, d, sink, blocks, st);
  if (nvec <= 64) return launch_probe<32, 2>(table, ld, n_rows, idx, n_idx, d, sink, blocks, st);
  if (nvec <= 128) return launch_probe<32, 4>(table, ld, n_rows, idx, n_idx, d, sink, blocks, st);
  return launch_probe<32, 8>(table, ld, n_rows, idx, n_idx, d, sink, blocks, st);
}
