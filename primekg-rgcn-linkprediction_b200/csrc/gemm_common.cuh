// Helpers shared by the tensor-core translation units (transform.cu, fused_layer.cu): operand-plane conversion, the
// counter-based dropout hash, TMA tensor maps (cached driver calls), dynamic shared-memory opt-in, weight-plane layout.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc05.cuh"

namespace rgcn {

// fp32 -> bf16 hi / lo pairs (hi + lo = x to 2^-17)
__device__ __forceinline__ void split4(const float4& v, uint2& hi, uint2& lo) {
  // packed conversions only (cvt.rn.bf16x2.f32 = F2FP on the ALU pipe)
  __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
  hi.x = *reinterpret_cast<uint32_t*>(&h01);
  hi.y = *reinterpret_cast<uint32_t*>(&h23);
  const float hx = __uint_as_float(hi.x << 16), hy = __uint_as_float(hi.x & 0xffff0000u);
  const float hz = __uint_as_float(hi.y << 16), hw = __uint_as_float(hi.y & 0xffff0000u);
  __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - hx, v.y - hy), l23 = __floats2bfloat162_rn(v.z - hz, v.w - hw);
  lo.x = *reinterpret_cast<uint32_t*>(&l01);
  lo.y = *reinterpret_cast<uint32_t*>(&l23);
}

// Counter-based dropout: element (row, col) of step `ctr` is kept iff 16 bits of hash(seed, ctr, row * N + col) >= thresh.
// The same function regenerates the mask anywhere; the backward pass does not even need it (out == 0 where dropped).
__device__ __forceinline__ uint32_t pcg_hash(uint32_t x) {
  uint32_t state = x * 747796405u + 2891336453u;
  uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
  return (word >> 22u) ^ word;
}
// key of the 2^32-element block that `elem` lies in; the bits of element e are pcg_hash((uint32_t)e ^ block_key)
__device__ __forceinline__ uint32_t drop_block_key(uint32_t key, uint64_t elem) {
  return pcg_hash(key + (uint32_t)(elem >> 32) * 0x9E3779B9u);
}

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

// bf16 matrix [rows, cols] row-major with leading dimension ld (elements); box = 64 cols x box_rows rows, 128B swizzle;
// out-of-range elements read as zero
struct MapKey { const void* base; int64_t rows, cols, ld; int box_rows; };
static int make_map(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  // the same planes / weight buffers come back every step: a small thread-local cache skips the driver call
  constexpr int NC = 128;
  static thread_local MapKey keys[NC];
  static thread_local CUtensorMap vals[NC];
  static thread_local int n_cached = 0, next = 0;
  for (int i = 0; i < n_cached; ++i) {
    const MapKey& k = keys[i];
    if (k.base == base && k.rows == rows && k.cols == cols && k.ld == ld && k.box_rows == box_rows) {
      *m = vals[i];
      return RGCN_OK;
    }
  }
  EncodeTiledFn enc = encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return RGCN_EUNSUPPORTED; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with code %d (rows %lld cols %lld ld %lld box_rows %d)", (int)r,
              (long long)rows, (long long)cols, (long long)ld, box_rows);
    return RGCN_ECUDA;
  }
  keys[next] = MapKey{base, rows, cols, ld, box_rows};
  vals[next] = *m;
  next = (next + 1) % NC;
  if (n_cached < NC) ++n_cached;
  return RGCN_OK;
}

// fp32 matrix [rows, cols] row-major with leading dimension ld (elements); box = 16 cols x 32 rows, 64-byte swizzle: the
// per-warp output patch of the transform's storing epilogue (TMA store; partial boxes are clipped)
static int make_out_map(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld) {
  struct Key { const void* base; int64_t rows, cols, ld; };
  constexpr int NC = 32;
  static thread_local Key keys[NC];
  static thread_local CUtensorMap vals[NC];
  static thread_local int n_cached = 0, next = 0;
  for (int i = 0; i < n_cached; ++i)
    if (keys[i].base == base && keys[i].rows == rows && keys[i].cols == cols && keys[i].ld == ld) { *m = vals[i]; return RGCN_OK; }
  EncodeTiledFn enc = encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return RGCN_EUNSUPPORTED; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {16u, 32u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (output) failed with code %d (rows %lld cols %lld ld %lld)", (int)r, (long long)rows,
              (long long)cols, (long long)ld);
    return RGCN_ECUDA;
  }
  keys[next] = Key{base, rows, cols, ld};
  vals[next] = *m;
  next = (next + 1) % NC;
  if (n_cached < NC) ++n_cached;
  return RGCN_OK;
}

static unsigned grid_cap(int64_t blocks, int64_t cap) { return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks)); }

static int round_up(int x, int a) { return (x + a - 1) / a * a; }

// opt in to > 48 KB dynamic shared memory, once per kernel (function pointer) and device
template <typename K>
static int set_smem(K kernel, int bytes) {
  struct Seen { const void* fn; int dev; };
  static Seen seen[64];
  static int n_seen = 0;
  int dev = 0;
  RGCN_CUDA(cudaGetDevice(&dev));
  const void* fn = reinterpret_cast<const void*>(kernel);
  for (int i = 0; i < n_seen; ++i)
    if (seen[i].fn == fn && seen[i].dev == dev) return RGCN_OK;
  RGCN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (n_seen < 64) seen[n_seen++] = Seen{fn, dev};
  return RGCN_OK;
}

// weight planes (rgcn_prepare_weights): bf16 hi, then lo, each [K, wplane_ld(d_out)] row-major
static int64_t wplane_ld(int d_out) { return round_up(d_out, 8); }          // TMA row stride: a multiple of 16 bytes
static size_t wplane_bytes(int K, int d_out) { return align_up((size_t)K * wplane_ld(d_out) * 2, 1024); }

}  // namespace rgcn
