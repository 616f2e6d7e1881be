// Relational transform of one RGCN layer on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
//   forward  O  = [H | X] @ [Wf ; root] + bias (, ReLU)          M = nodes, N = d_out, K = (R+1) d_in
//   dgrad    gA = (gO * relu') @ [Wf ; root]^T                   M = nodes, N = (R+1) d_in, K = d_out
//   wgrad    [gWf ; g_root] = [H | X]^T @ (gO * relu'),  g_bias = column sums      M = (R+1) d_in, N = d_out, K = nodes
//
// These are the R+1 per-relation `h_r @ W_r` / `x @ root` products of RGCNConv (reference call sites
// src/models/rgcn.py:123, :128) and their autograd transposes (src/train.py:306), concatenated along K so
// that one tile pass serves all relations.
//
// Precision modes (fp32 accumulation in TMEM in both):
//   mode 0 "fp32": every fp32 operand x is split on the fly into bf16 hi + lo (x = hi + lo to 2^-17) and the
//                  product is formed as hi*hi + hi*lo + lo*hi  -> ~1e-5 relative, inside the 1e-4 tolerance.
//   mode 1 "bf16": operands rounded to bf16, one product (the "bf16-transform" mode, tolerance 2e-2).
//
// Kernel anatomy (one CTA = one 128 x BN output tile, BN <= 256 TMEM columns):
//   warps 0..3 (fwd/dgrad) or 0..7 (wgrad): operand loaders: 128-bit coalesced global loads of the fp32 activations,
//              convert / split to bf16 in registers, store into the 128B-swizzled UMMA layout, fence.proxy.async,
//              mbarrier arrive.  The same warps run the epilogue (tcgen05.ld -> bias / ReLU -> global) afterwards.
//   TMA warp : (fwd/dgrad) one thread streams the pre-split bf16 weight tiles with cp.async.bulk.tensor (SWIZZLE_128B).
//   MMA warp : one thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) and tcgen05.commit.
#include <cuda.h>

#include "common.cuh"
#include "tc05.cuh"

namespace rgcn {
using namespace tc05;

constexpr int BM = 128;         // UMMA M
constexpr int BK = 64;          // K per stage for K-major operands (= one 128-byte swizzle row of bf16)
constexpr int BNMAX = 256;      // UMMA N max = TMEM columns per accumulator
constexpr int WG_BK = 32;       // nodes per stage in the weight-gradient kernel

// ------------------------------------------------------------------------------------------------
// weight preparation: fp32 [K1 + K2, N] (two row blocks) -> bf16 hi / lo, optionally transposed, zero padded
// ------------------------------------------------------------------------------------------------
__global__ void split_weights_kernel(const float* __restrict__ w1, int K1, const float* __restrict__ w2, int K2, int N,
                                     int transpose, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                     int rows_pad, int cols_pad) {
  // output [rows_pad, cols_pad]; transpose: out[n][k] = W[k][n], else out[k][n] = W[k][n]
  const int64_t total = (int64_t)rows_pad * cols_pad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols_pad), c = (int)(i % cols_pad);
    const int k = transpose ? c : r, n = transpose ? r : c;
    float v = 0.f;
    if (k < K1 + K2 && n < N) v = (k < K1) ? w1[(int64_t)k * N + n] : w2[(int64_t)(k - K1) * N + n];
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}

// ------------------------------------------------------------------------------------------------
// K-major GEMM: out[M, N] = A[M, K] @ B[N, K]^T   (forward and dgrad)
// ------------------------------------------------------------------------------------------------
struct GemmKParams {
  const float* a1; int64_t lda1; int K1;      // A columns [0, K1)
  const float* a2; int64_t lda2; int K2;      // A columns [K1, K1 + K2)
  const float* mask; int64_t ldmask;          // optional: A1 element is zeroed where mask <= 0 (ReLU backward)
  int64_t M;
  int N;                                      // valid output columns
  int BN;                                     // tile width: multiple of 32, <= 256
  int num_kb;                                 // K blocks of 64 (weights are zero padded to this)
  const float* bias; int relu;
  float* out; int64_t ldo;
};

template <bool SPLIT>
struct KStage {
  static constexpr int A_BYTES = BM * 128;              // 128 rows x 64 bf16
  static constexpr int B_BYTES = BNMAX * 128;
  static constexpr int BYTES = (SPLIT ? 2 : 1) * (A_BYTES + B_BYTES);
  static constexpr int STAGES = SPLIT ? 2 : 4;
  static constexpr int A_HI = 0, A_LO = A_BYTES;
  static constexpr int B_HI = (SPLIT ? 2 : 1) * A_BYTES, B_LO = B_HI + B_BYTES;
};

constexpr int K_LOADERS = 256;   // 8 loader / epilogue warps

// 4 fp32 -> 4 bf16 hi (+ 4 bf16 lo) packed into 8-byte words.  Packed conversions only
// (cvt.rn.bf16x2.f32 = F2FP on the ALU pipe); the scalar F2F.BF16.F32 runs on the slow conversion pipe and
// made the loaders the bottleneck of every GEMM.
template <bool SPLIT>
__device__ __forceinline__ void split_store4(const float4& v, uint8_t* hi_ptr, uint8_t* lo_ptr) {
  __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&h01);
  pk.y = *reinterpret_cast<uint32_t*>(&h23);
  *reinterpret_cast<uint2*>(hi_ptr) = pk;
  if (SPLIT) {
    // bf16 -> fp32 is a shift / mask of the bit pattern
    const float hx = __uint_as_float(pk.x << 16), hy = __uint_as_float(pk.x & 0xffff0000u);
    const float hz = __uint_as_float(pk.y << 16), hw = __uint_as_float(pk.y & 0xffff0000u);
    __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - hx, v.y - hy), l23 = __floats2bfloat162_rn(v.z - hz, v.w - hw);
    uint2 pl;
    pl.x = *reinterpret_cast<uint32_t*>(&l01);
    pl.y = *reinterpret_cast<uint32_t*>(&l23);
    *reinterpret_cast<uint2*>(lo_ptr) = pl;
  }
}
__device__ __forceinline__ void relu_mask4(float4& v, const float4& m) {
  if (!(m.x > 0.f)) v.x = 0.f;
  if (!(m.y > 0.f)) v.y = 0.f;
  if (!(m.z > 0.f)) v.z = 0.f;
  if (!(m.w > 0.f)) v.w = 0.f;
}

template <bool SPLIT>
__global__ void __launch_bounds__(K_LOADERS + 64, 1)
gemm_kmajor_kernel(const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                   const GemmKParams p) {
  using S = KStage<SPLIT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[S::STAGES], empty_bar[S::STAGES], accum_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_bias[BNMAX];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x, n_tile = blockIdx.y;
  const int64_t m0 = (int64_t)m_tile * BM;
  const int n0 = n_tile * p.BN;
  const uint32_t tmem_cols = p.BN <= 32 ? 32 : p.BN <= 64 ? 64 : p.BN <= 128 ? 128 : 256;

  if (threadIdx.x < BNMAX) {
    const int n = n0 + threadIdx.x;
    s_bias[threadIdx.x] = (p.bias && threadIdx.x < p.BN && n < p.N) ? p.bias[n] : 0.f;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < S::STAGES; ++s) {
      mbar_init(&full_bar[s], K_LOADERS + 1);   // loader threads + the TMA thread's expect_tx arrive
      mbar_init(&empty_bar[s], 1);              // one tcgen05.commit
    }
    mbar_init(&accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_b_hi);
    if (SPLIT) tma_prefetch_desc(&tm_b_lo);
  }
  if (warp == 0) tmem_alloc(&tmem_base_smem, tmem_cols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp < 8) {
    // ===================== operand-A loaders (register double buffering: K block kb+1 is in flight
    // while kb is converted and stored) =====================
    const int c = threadIdx.x & 15;          // float4 index inside the 64-wide K block
    const int rsub = threadIdx.x >> 4;       // 0..15
    float4 cur[8], nxt[8], mcur[8], mnxt[8];
    bool cur_masked = false, nxt_masked = false;
    auto issue = [&](int kb, float4 (&v)[8], float4 (&mk)[8], bool& masked) {
      const int k = kb * BK + c * 4;
      const float* src = nullptr;
      const float* msk = nullptr;
      int64_t ld = 0;
      if (k < p.K1) { src = p.a1 + k; ld = p.lda1; if (p.mask) msk = p.mask + k; }
      else if (k < p.K1 + p.K2) { src = p.a2 + (k - p.K1); ld = p.lda2; }
      masked = msk != nullptr;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int64_t row = m0 + it * 16 + rsub;
        const bool ok = src && row < p.M;
        v[it] = ok ? ldg4(src + row * ld) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (msk) mk[it] = ok ? ldg4(msk + row * p.ldmask) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    issue(0, cur, mcur, cur_masked);
    for (int kb = 0; kb < p.num_kb; ++kb) {
      const int s = kb % S::STAGES;
      const uint32_t ph = (kb / S::STAGES) & 1;
      if (kb + 1 < p.num_kb) issue(kb + 1, nxt, mnxt, nxt_masked);
      mbar_wait(&empty_bar[s], ph ^ 1);      // slot free (passes immediately on the first round)
      uint8_t* st = smem + (size_t)s * S::BYTES;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const uint32_t r = it * 16 + rsub;
        const uint32_t off = sw128_offset(r, c >> 1) + (c & 1) * 8;
        if (cur_masked) relu_mask4(cur[it], mcur[it]);
        split_store4<SPLIT>(cur[it], st + S::A_HI + off, st + S::A_LO + off);
      }
      fence_proxy_async();
      mbar_arrive(&full_bar[s]);
#pragma unroll
      for (int it = 0; it < 8; ++it) { cur[it] = nxt[it]; mcur[it] = mnxt[it]; }
      cur_masked = nxt_masked;
    }
    // ===================== epilogue =====================
    // TMEM -> registers (+ bias, ReLU) -> staging tile in the (now idle) operand smem -> coalesced row stores.
    // Warps w and w+4 share TMEM lane quadrant w and alternate 32-column chunks.
    mbar_wait(&accum_bar, 0);
    fence_after_sync();
    const int q = warp & 3, half = warp >> 2;
    const int pitch = p.BN + 4;                      // floats; +4 keeps the float4 row stores bank-conflict free
    float* stile = reinterpret_cast<float*>(smem);
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    float* srow = stile + (size_t)(q * 32 + lane) * pitch;
    for (int cc = half * 32; cc < p.BN; cc += 64) {
      uint32_t r[32];
      tmem_ld_32x32(t_lane + cc, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = *reinterpret_cast<const float4*>(s_bias + cc + j);
        float4 qv = make_float4(__uint_as_float(r[j]) + b.x, __uint_as_float(r[j + 1]) + b.y,
                                __uint_as_float(r[j + 2]) + b.z, __uint_as_float(r[j + 3]) + b.w);
        if (p.relu) { qv.x = fmaxf(qv.x, 0.f); qv.y = fmaxf(qv.y, 0.f); qv.z = fmaxf(qv.z, 0.f); qv.w = fmaxf(qv.w, 0.f); }
        *reinterpret_cast<float4*>(srow + cc + j) = qv;
      }
    }
    fence_before_sync();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int ncol4 = min(p.BN, p.N - n0) >> 2;      // valid float4 columns of this tile
    for (int rl = warp; rl < BM; rl += 8) {
      const int64_t row = m0 + rl;
      if (row >= p.M) break;
      const float* src = stile + (size_t)rl * pitch;
      float* dst = p.out + row * p.ldo + n0;
      for (int c4 = lane; c4 < ncol4; c4 += 32)
        *reinterpret_cast<float4*>(dst + c4 * 4) = *reinterpret_cast<const float4*>(src + c4 * 4);
    }
  } else if (warp == 8) {
    // ===================== TMA producer for the weight tiles =====================
    if (lane == 0) {
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % S::STAGES;
        const uint32_t ph = (kb / S::STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* st = smem + (size_t)s * S::BYTES;
        const uint32_t bytes = (uint32_t)p.BN * 128u * (SPLIT ? 2u : 1u);
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        tma_load_2d(st + S::B_HI, &tm_b_hi, &full_bar[s], kb * BK, n0);
        if (SPLIT) tma_load_2d(st + S::B_LO, &tm_b_lo, &full_bar[s], kb * BK, n0);
      }
    }
  } else {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16(BM, p.BN, 0, 0);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % S::STAGES;
        const uint32_t ph = (kb / S::STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        fence_after_sync();
        const uint32_t base = smem_u32(smem + (size_t)s * S::BYTES);
        const uint64_t da_hi = smem_desc_sw128(base + S::A_HI, 16, 1024);
        const uint64_t db_hi = smem_desc_sw128(base + S::B_HI, 16, 1024);
        const uint64_t da_lo = smem_desc_sw128(base + S::A_LO, 16, 1024);
        const uint64_t db_lo = smem_desc_sw128(base + S::B_LO, 16, 1024);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t adv = (uint64_t)(k * 32 >> 4);      // 16 bf16 = 32 bytes along K inside the swizzle row
          mma_bf16_ss(tmem_base, da_hi + adv, db_hi + adv, idesc, (kb | k) ? 1u : 0u);
          if (SPLIT) {
            mma_bf16_ss(tmem_base, da_hi + adv, db_lo + adv, idesc, 1u);
            mma_bf16_ss(tmem_base, da_lo + adv, db_hi + adv, idesc, 1u);
          }
        }
        mma_commit(&empty_bar[s]);            // frees the stage once these MMAs have read it
      }
      mma_commit(&accum_bar);                 // accumulator complete
    }
  }
  __syncthreads();
  if (warp == 0) {
    fence_after_sync();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient: P[split][m, n] = sum_{node in split} A[node, m] * G[node, n]   (both operands MN-major)
// ------------------------------------------------------------------------------------------------
struct WgradParams {
  const float* a1; int64_t lda1; int K1;      // A columns (= output rows m) [0, K1)
  const float* a2; int64_t lda2; int K2;      // [K1, K1 + K2)
  const float* g; int64_t ldg;                // upstream gradient [nodes, N]
  const float* mask; int64_t ldmask;          // optional ReLU mask source (same shape as g)
  int64_t nodes;
  int N, BN;                                  // valid / tile columns (BN multiple of 32 <= 256)
  int splits;                                 // node range is cut into `splits` contiguous slices
  int64_t nodes_per_split;                    // multiple of WG_BK
  float* partial;                             // [splits, m_tiles * 128, BN_total]
  float* partial_bias;                        // [splits, BN_total]
  int ldp;                                    // = n_tiles * BN
};

template <bool SPLIT>
struct WStage {
  static constexpr int A_BYTES = 2 * WG_BK * 128;        // 2 chunks of 64 m  x 32 nodes x 128 B
  static constexpr int B_BYTES = 4 * WG_BK * 128;        // 4 chunks of 64 n
  static constexpr int BYTES = (SPLIT ? 2 : 1) * (A_BYTES + B_BYTES);
  static constexpr int STAGES = 4;
  static constexpr int A_HI = 0, A_LO = A_BYTES;
  static constexpr int B_HI = (SPLIT ? 2 : 1) * A_BYTES, B_LO = B_HI + B_BYTES;
};

template <bool SPLIT>
__global__ void __launch_bounds__(288, 1) gemm_wgrad_kernel(const WgradParams p) {
  using S = WStage<SPLIT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[S::STAGES], empty_bar[S::STAGES], accum_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float4 bias_red[4][64];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x, n_tile = blockIdx.y, split = blockIdx.z;
  const int m0 = m_tile * BM, n0 = n_tile * p.BN;
  const uint32_t tmem_cols = p.BN <= 32 ? 32 : p.BN <= 64 ? 64 : p.BN <= 128 ? 128 : 256;
  const int64_t node_beg = (int64_t)split * p.nodes_per_split;
  const int64_t node_end = min(node_beg + p.nodes_per_split, p.nodes);
  const int num_kb = node_end > node_beg ? (int)((node_end - node_beg + WG_BK - 1) / WG_BK) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S::STAGES; ++s) {
      mbar_init(&full_bar[s], 256);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_smem, tmem_cols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp < 8) {
    // ===================== loaders =====================
    const int ca = threadIdx.x & 31, ra = threadIdx.x >> 5;     // A: float4 column 0..31 (128 m), node sub-row 0..7
    const int cb = threadIdx.x & 63, rb = threadIdx.x >> 6;     // B: float4 column 0..63 (256 n), node sub-row 0..3
    const int ma = m0 + ca * 4;
    const float* srca = nullptr; int64_t lda = 0;
    if (ma < p.K1) { srca = p.a1 + ma; lda = p.lda1; }
    else if (ma < p.K1 + p.K2) { srca = p.a2 + (ma - p.K1); lda = p.lda2; }
    const int nb = n0 + cb * 4;
    const bool b_on = (cb * 4 < p.BN);
    const bool b_valid = b_on && nb < p.N;
    float4 colsum = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 va[4], vb[8], vm[8], na[4], nbv[8], nm[8];
    auto issue = [&](int kb, float4 (&xa)[4], float4 (&xb)[8], float4 (&xm)[8]) {
      const int64_t nd0 = node_beg + (int64_t)kb * WG_BK;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int64_t nd = nd0 + it * 8 + ra;
        xa[it] = (srca && nd < node_end) ? ldg4(srca + nd * lda) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int64_t nd = nd0 + it * 4 + rb;
        const bool ok = b_valid && nd < node_end;
        xb[it] = ok ? ldg4(p.g + nd * p.ldg + nb) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.mask) xm[it] = ok ? ldg4(p.mask + nd * p.ldmask + nb) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    if (num_kb > 0) issue(0, va, vb, vm);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % S::STAGES;
      const uint32_t ph = (kb / S::STAGES) & 1;
      if (kb + 1 < num_kb) issue(kb + 1, na, nbv, nm);
      if (p.mask) {
#pragma unroll
        for (int it = 0; it < 8; ++it) relu_mask4(vb[it], vm[it]);
      }
      if (m_tile == 0) {
#pragma unroll
        for (int it = 0; it < 8; ++it) add4(colsum, vb[it]);
      }
      mbar_wait(&empty_bar[s], ph ^ 1);
      uint8_t* st = smem + (size_t)s * S::BYTES;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const uint32_t kr = it * 8 + ra;                       // node row inside the stage
        const uint32_t off = (uint32_t)(ca >> 4) * (WG_BK * 128) + sw128_offset(kr, (ca & 15) >> 1) + (ca & 1) * 8;
        split_store4<SPLIT>(va[it], st + S::A_HI + off, st + S::A_LO + off);
      }
      if (b_on) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const uint32_t kr = it * 4 + rb;
          const uint32_t off = (uint32_t)(cb >> 4) * (WG_BK * 128) + sw128_offset(kr, (cb & 15) >> 1) + (cb & 1) * 8;
          split_store4<SPLIT>(vb[it], st + S::B_HI + off, st + S::B_LO + off);
        }
      }
      fence_proxy_async();
      mbar_arrive(&full_bar[s]);
#pragma unroll
      for (int it = 0; it < 4; ++it) va[it] = na[it];
#pragma unroll
      for (int it = 0; it < 8; ++it) { vb[it] = nbv[it]; vm[it] = nm[it]; }
    }
    // bias gradient partial: fixed-order sum of the 4 node sub-rows sharing a column group
    if (m_tile == 0) bias_red[rb][cb] = colsum;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (m_tile == 0 && threadIdx.x < 64 && b_valid) {
      float4 s4 = bias_red[0][cb];
      add4(s4, bias_red[1][cb]); add4(s4, bias_red[2][cb]); add4(s4, bias_red[3][cb]);
      *reinterpret_cast<float4*>(p.partial_bias + (size_t)split * p.ldp + nb) = s4;
    }
    // ===================== epilogue: TMEM -> partial buffer (warps 0..3 own TMEM lanes 32w..32w+31) ==========
    if (warp < 4) {
      float* prow = p.partial + ((size_t)split * gridDim.x * BM + (size_t)(m0 + warp * 32 + lane)) * p.ldp + n0;
      if (num_kb > 0) {
        mbar_wait(&accum_bar, 0);
        fence_after_sync();
        const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int cc = 0; cc < p.BN; cc += 32) {
          uint32_t r[32];
          tmem_ld_32x32(t_lane + cc, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(prow + cc + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                    __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
        fence_before_sync();
      } else {
        for (int cc = 0; cc < p.BN; cc += 4) *reinterpret_cast<float4*>(prow + cc) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  } else {
    // ===================== MMA issuer (warp 8) =====================
    if (lane == 0 && num_kb > 0) {
      const uint32_t idesc = idesc_bf16(BM, p.BN, 1, 1);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % S::STAGES;
        const uint32_t ph = (kb / S::STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        fence_after_sync();
        const uint32_t base = smem_u32(smem + (size_t)s * S::BYTES);
        // MN-major: LBO = stride between 64-element M/N chunks, SBO = stride between 8-node groups
        const uint64_t da_hi = smem_desc_sw128(base + S::A_HI, WG_BK * 128, 1024);
        const uint64_t db_hi = smem_desc_sw128(base + S::B_HI, WG_BK * 128, 1024);
        const uint64_t da_lo = smem_desc_sw128(base + S::A_LO, WG_BK * 128, 1024);
        const uint64_t db_lo = smem_desc_sw128(base + S::B_LO, WG_BK * 128, 1024);
#pragma unroll
        for (int k = 0; k < WG_BK / 16; ++k) {
          const uint64_t adv = (uint64_t)(k * 2048 >> 4);      // 16 nodes = two 8-node atoms of 1024 B
          mma_bf16_ss(tmem_base, da_hi + adv, db_hi + adv, idesc, (kb | k) ? 1u : 0u);
          if (SPLIT) {
            mma_bf16_ss(tmem_base, da_hi + adv, db_lo + adv, idesc, 1u);
            mma_bf16_ss(tmem_base, da_lo + adv, db_hi + adv, idesc, 1u);
          }
        }
        mma_commit(&empty_bar[s]);
      }
      mma_commit(&accum_bar);
    }
  }
  __syncthreads();
  if (warp == 0) {
    fence_after_sync();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// fixed-order reduction of the split partials into gW1 [K1, N], gW2 [K2, N], g_bias [N]
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ partial_bias, int splits,
                                    int m_pad, int ldp, int K1, int K2, int N, float* __restrict__ gw1,
                                    float* __restrict__ gw2, float* __restrict__ gbias) {
  const int64_t total = (int64_t)(K1 + K2 + 1) * (N >> 2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / (N >> 2)), n = (int)(i % (N >> 2)) * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < K1 + K2) {
      for (int sp = 0; sp < splits; ++sp)
        add4(s, *reinterpret_cast<const float4*>(partial + ((size_t)sp * m_pad + m) * ldp + n));
      float* dst = (m < K1) ? gw1 + (size_t)m * N + n : gw2 + (size_t)(m - K1) * N + n;
      if (dst) *reinterpret_cast<float4*>(dst) = s;
    } else if (gbias) {
      for (int sp = 0; sp < splits; ++sp) add4(s, *reinterpret_cast<const float4*>(partial_bias + (size_t)sp * ldp + n));
      *reinterpret_cast<float4*>(gbias + n) = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// bf16 matrix [rows, cols] row-major, box = 64 cols x box_rows rows, 128-byte swizzle
static int make_map(CUtensorMap* m, const void* base, int rows, int cols, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return RGCN_EUNSUPPORTED; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with code %d", (int)r); return RGCN_ECUDA; }
  return RGCN_OK;
}

static unsigned grid_cap(int64_t blocks, int64_t cap) { return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks)); }

static int round_up(int x, int a) { return (x + a - 1) / a * a; }

struct Tiling { int n_tiles, BN, n_pad; };
static Tiling tile_n(int N) {
  Tiling t;
  const int np = round_up(N, 32);
  t.n_tiles = (np + BNMAX - 1) / BNMAX;
  t.BN = round_up((np + t.n_tiles - 1) / t.n_tiles, 32);
  t.n_pad = t.n_tiles * t.BN;
  return t;
}

// MN-major operands come in 64-element swizzle atoms: the weight-gradient tile width is a multiple of 64
static Tiling tile_n64(int N) {
  Tiling t;
  const int np = round_up(N, 64);
  t.n_tiles = (np + BNMAX - 1) / BNMAX;
  t.BN = round_up((np + t.n_tiles - 1) / t.n_tiles, 64);
  t.n_pad = t.n_tiles * t.BN;
  return t;
}

template <typename K>
static int set_smem(K kernel, int bytes) {
  RGCN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return RGCN_OK;
}

// A[M, K1+K2] @ W^T with W given as bf16 hi/lo [n_pad, k_pad] (K-major rows)
static int launch_kmajor(const GemmKParams& p, const __nv_bfloat16* bhi, const __nv_bfloat16* blo, int n_pad, int k_pad,
                         int n_tiles, bool split, cudaStream_t st) {
  CUtensorMap mhi, mlo;
  int rc = make_map(&mhi, bhi, n_pad, k_pad, p.BN);
  if (rc) return rc;
  rc = make_map(&mlo, split ? blo : bhi, n_pad, k_pad, p.BN);
  if (rc) return rc;
  dim3 grid((unsigned)((p.M + BM - 1) / BM), (unsigned)n_tiles);
  if (split) {
    const int smem = KStage<true>::STAGES * KStage<true>::BYTES + 1024;
    rc = set_smem(gemm_kmajor_kernel<true>, smem);
    if (rc) return rc;
    gemm_kmajor_kernel<true><<<grid, K_LOADERS + 64, smem, st>>>(mhi, mlo, p);
  } else {
    const int smem = KStage<false>::STAGES * KStage<false>::BYTES + 1024;
    rc = set_smem(gemm_kmajor_kernel<false>, smem);
    if (rc) return rc;
    gemm_kmajor_kernel<false><<<grid, K_LOADERS + 64, smem, st>>>(mhi, mlo, p);
  }
  RGCN_LAUNCH_CHECK();
  return RGCN_OK;
}

static int wgrad_splits(int64_t nodes, int tiles) {
  int s = sm_count() / tiles;                                     // one wave: tiles * splits <= #SMs
  const int64_t max_s = (nodes + 4 * WG_BK - 1) / (4 * WG_BK);      // at least 4 stages of work per split
  if (s > max_s) s = (int)max_s;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  return s;
}

}  // namespace rgcn

using namespace rgcn;

static int check_mat(const float* p, int64_t ld, const char* what) {
  RGCN_CHECK_ARG(p && ((uintptr_t)p & 15) == 0 && ld % 4 == 0, "transform: %s must be non-null, 16-byte aligned, ld %% 4 == 0", what);
  return RGCN_OK;
}

extern "C" size_t rgcn_transform_workspace_bytes(int64_t n_rows, int32_t K1, int32_t K2, int32_t d_out) {
  if (n_rows < 0 || K1 < 0 || K2 < 0 || d_out <= 0) return 0;
  const int K = K1 + K2;
  // forward: weights^T [n_pad(d_out), k_pad(K)];  dgrad: weights [n_pad(K), k_pad(d_out)];  2 bf16 planes each
  const size_t fwd = (size_t)tile_n(d_out).n_pad * round_up(K, BK) * 2 * 2;
  const size_t dgr = (size_t)tile_n(K).n_pad * round_up(d_out, BK) * 2 * 2;
  const Tiling t = tile_n64(d_out);
  const int m_tiles = (K + BM - 1) / BM;
  const int splits = wgrad_splits(n_rows, m_tiles * t.n_tiles);
  const size_t wg = ((size_t)splits * m_tiles * BM + splits) * t.n_pad * 4;
  size_t need = fwd > dgr ? fwd : dgr;
  if (wg > need) need = wg;
  return align_up(need, 256) + 1024;
}

extern "C" int rgcn_transform_fwd(const float* A1, int64_t lda1, int32_t K1, const float* A2, int64_t lda2, int32_t K2,
                                  const float* W1, const float* W2, const float* bias, int32_t relu, int64_t n_rows,
                                  int32_t d_out, float* out, int64_t ldo, int32_t mode, void* workspace,
                                  size_t workspace_bytes, rgcn_stream_t stream) {
  RGCN_CHECK_ARG(n_rows >= 0 && K1 > 0 && K2 >= 0 && d_out > 0, "transform_fwd: bad sizes");
  RGCN_CHECK_ARG(K1 % 4 == 0 && K2 % 4 == 0 && d_out % 4 == 0, "transform_fwd: K1, K2, d_out must be multiples of 4");
  RGCN_CHECK_ARG(mode == 0 || mode == 1, "transform_fwd: mode must be 0 (fp32) or 1 (bf16)");
  int rc = check_mat(A1, lda1, "A1"); if (rc) return rc;
  if (K2) { rc = check_mat(A2, lda2, "A2"); if (rc) return rc; RGCN_CHECK_ARG(W2, "transform_fwd: W2 is null"); }
  rc = check_mat(out, ldo, "out"); if (rc) return rc;
  RGCN_CHECK_ARG(W1 && (!bias || ((uintptr_t)bias & 15) == 0), "transform_fwd: W1 null or bias misaligned");
  if (n_rows == 0) return RGCN_OK;
  const int K = K1 + K2, k_pad = round_up(K, BK);
  const Tiling t = tile_n(d_out);
  const size_t plane = (size_t)t.n_pad * k_pad * 2;
  if (!workspace || workspace_bytes < align_up(2 * plane, 256) + 1024) {
    set_error("transform_fwd: workspace too small"); return RGCN_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)align_up((size_t)workspace, 1024);
  __nv_bfloat16* bhi = (__nv_bfloat16*)ws;
  __nv_bfloat16* blo = (__nv_bfloat16*)(ws + plane);
  const bool split = mode == 0;
  {
    const int64_t total = (int64_t)t.n_pad * k_pad;
    split_weights_kernel<<<grid_cap((total + 255) / 256, 1184), 256, 0, st>>>(
        W1, K1, W2, K2, d_out, 1, bhi, split ? blo : nullptr, t.n_pad, k_pad);
    RGCN_LAUNCH_CHECK();
  }
  GemmKParams p{};
  p.a1 = A1; p.lda1 = lda1; p.K1 = K1; p.a2 = A2; p.lda2 = lda2; p.K2 = K2;
  p.M = n_rows; p.N = d_out; p.BN = t.BN; p.num_kb = k_pad / BK;
  p.bias = bias; p.relu = relu; p.out = out; p.ldo = ldo;
  return launch_kmajor(p, bhi, blo, t.n_pad, k_pad, t.n_tiles, split, st);
}

extern "C" int rgcn_transform_dgrad(const float* gO, int64_t ldg, const float* relu_out, int64_t ld_ro, int32_t d_out,
                                    const float* W1, int32_t K1, const float* W2, int32_t K2, int64_t n_rows, float* gA,
                                    int64_t ldga, int32_t mode, void* workspace, size_t workspace_bytes,
                                    rgcn_stream_t stream) {
  RGCN_CHECK_ARG(n_rows >= 0 && K1 > 0 && K2 >= 0 && d_out > 0, "transform_dgrad: bad sizes");
  RGCN_CHECK_ARG(K1 % 4 == 0 && K2 % 4 == 0 && d_out % 4 == 0, "transform_dgrad: K1, K2, d_out must be multiples of 4");
  RGCN_CHECK_ARG(mode == 0 || mode == 1, "transform_dgrad: mode must be 0 (fp32) or 1 (bf16)");
  int rc = check_mat(gO, ldg, "gO"); if (rc) return rc;
  rc = check_mat(gA, ldga, "gA"); if (rc) return rc;
  if (relu_out) { rc = check_mat(relu_out, ld_ro, "relu_out"); if (rc) return rc; }
  RGCN_CHECK_ARG(W1 && (K2 == 0 || W2), "transform_dgrad: null weights");
  if (n_rows == 0) return RGCN_OK;
  const int K = K1 + K2;                      // = N of this GEMM
  const int k_pad = round_up(d_out, BK);      // = K of this GEMM
  const Tiling t = tile_n(K);
  const size_t plane = (size_t)t.n_pad * k_pad * 2;
  if (!workspace || workspace_bytes < align_up(2 * plane, 256) + 1024) {
    set_error("transform_dgrad: workspace too small"); return RGCN_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)align_up((size_t)workspace, 1024);
  __nv_bfloat16* bhi = (__nv_bfloat16*)ws;
  __nv_bfloat16* blo = (__nv_bfloat16*)(ws + plane);
  const bool split = mode == 0;
  {
    const int64_t total = (int64_t)t.n_pad * k_pad;
    split_weights_kernel<<<grid_cap((total + 255) / 256, 1184), 256, 0, st>>>(
        W1, K1, W2, K2, d_out, 0, bhi, split ? blo : nullptr, t.n_pad, k_pad);
    RGCN_LAUNCH_CHECK();
  }
  GemmKParams p{};
  p.a1 = gO; p.lda1 = ldg; p.K1 = d_out; p.a2 = nullptr; p.K2 = 0;
  p.mask = relu_out; p.ldmask = ld_ro;
  p.M = n_rows; p.N = K; p.BN = t.BN; p.num_kb = k_pad / BK;
  p.bias = nullptr; p.relu = 0; p.out = gA; p.ldo = ldga;
  return launch_kmajor(p, bhi, blo, t.n_pad, k_pad, t.n_tiles, split, st);
}

extern "C" int rgcn_transform_wgrad(const float* A1, int64_t lda1, int32_t K1, const float* A2, int64_t lda2, int32_t K2,
                                    const float* gO, int64_t ldg, const float* relu_out, int64_t ld_ro, int32_t d_out,
                                    int64_t n_rows, float* gW1, float* gW2, float* gbias, int32_t mode, void* workspace,
                                    size_t workspace_bytes, rgcn_stream_t stream) {
  RGCN_CHECK_ARG(n_rows >= 0 && K1 > 0 && K2 >= 0 && d_out > 0, "transform_wgrad: bad sizes");
  RGCN_CHECK_ARG(K1 % 4 == 0 && K2 % 4 == 0 && d_out % 4 == 0, "transform_wgrad: K1, K2, d_out must be multiples of 4");
  RGCN_CHECK_ARG(mode == 0 || mode == 1, "transform_wgrad: mode must be 0 (fp32) or 1 (bf16)");
  int rc = check_mat(A1, lda1, "A1"); if (rc) return rc;
  if (K2) { rc = check_mat(A2, lda2, "A2"); if (rc) return rc; }
  rc = check_mat(gO, ldg, "gO"); if (rc) return rc;
  if (relu_out) { rc = check_mat(relu_out, ld_ro, "relu_out"); if (rc) return rc; }
  RGCN_CHECK_ARG(gW1 && (K2 == 0 || gW2), "transform_wgrad: null outputs");
  const int K = K1 + K2;
  const Tiling t = tile_n64(d_out);
  const int m_tiles = (K + BM - 1) / BM;
  const int splits = wgrad_splits(n_rows, m_tiles * t.n_tiles);
  const size_t part = (size_t)splits * m_tiles * BM * t.n_pad * 4, pbias = (size_t)splits * t.n_pad * 4;
  if (!workspace || workspace_bytes < align_up(part + pbias, 256) + 1024) {
    set_error("transform_wgrad: workspace too small"); return RGCN_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)align_up((size_t)workspace, 1024);
  WgradParams p{};
  p.a1 = A1; p.lda1 = lda1; p.K1 = K1; p.a2 = A2; p.lda2 = lda2; p.K2 = K2;
  p.g = gO; p.ldg = ldg; p.mask = relu_out; p.ldmask = ld_ro;
  p.nodes = n_rows; p.N = d_out; p.BN = t.BN; p.splits = splits;
  p.nodes_per_split = ((n_rows + splits - 1) / splits + WG_BK - 1) / WG_BK * WG_BK;
  if (p.nodes_per_split == 0) p.nodes_per_split = WG_BK;
  p.partial = (float*)ws; p.partial_bias = (float*)(ws + part); p.ldp = t.n_pad;
  dim3 grid((unsigned)m_tiles, (unsigned)t.n_tiles, (unsigned)splits);
  if (mode == 0) {
    const int smem = WStage<true>::STAGES * WStage<true>::BYTES + 1024;
    rc = set_smem(gemm_wgrad_kernel<true>, smem); if (rc) return rc;
    gemm_wgrad_kernel<true><<<grid, 288, smem, st>>>(p);
  } else {
    const int smem = WStage<false>::STAGES * WStage<false>::BYTES + 1024;
    rc = set_smem(gemm_wgrad_kernel<false>, smem); if (rc) return rc;
    gemm_wgrad_kernel<false><<<grid, 288, smem, st>>>(p);
  }
  RGCN_LAUNCH_CHECK();
  const int64_t total = (int64_t)(K + 1) * (d_out / 4);
  wgrad_reduce_kernel<<<grid_cap((total + 255) / 256, 2368), 256, 0, st>>>(
      p.partial, p.partial_bias, splits, m_tiles * BM, t.n_pad, K1, K2, d_out, gW1, gW2, gbias);
  RGCN_LAUNCH_CHECK();
  return RGCN_OK;
}
