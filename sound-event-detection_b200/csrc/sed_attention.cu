// Scaled dot-product attention of the Transformer temporal block on the tensor cores (sm_100a).
//
// Reference: ScaledDotProductAttention.forward pytorch/models.py:808-820 (attn = q k^T / sqrt(d_k), softmax over
// keys, attn v) with the head split / merge of MultiHead.forward :863-875 (8 heads, d_k = d_v = 64).
//
// Work unit = (clip b, head h).  The QKV projection (sed_linear) leaves q | k | v as 16-bit rows of 1536 in the
// time-major order the conv stack produced its features in: row (t, b) = t * Bp + b.  Per unit one thread
//   * pulls Q_h, K_h, V_h [T x 64] with three 3-D TMA boxes (64 columns x 1 clip x 128 steps; steps >= T are
//     zero-filled by the TMA unit, so nothing is padded in memory) into SWIZZLE_128B tiles,
//   * issues S = Q K^T as 4 tcgen05.mma (M = 128 queries, N = 128 keys, K = 64) into TMEM;
// the four warps (thread = query row) read their row of S from TMEM, take the softmax in registers (exp2 on
// pre-scaled logits, keys >= T masked), and write the un-normalised probabilities as the 16-bit A operand P of
//   * O = P V: 8 tcgen05.mma (M = 128, N = 64, K = 128 keys) with V used AS LOADED -- [key][d], d contiguous -- through
//     an MN-major B descriptor (instruction-descriptor bit 16), so no transpose of V exists anywhere;
// O comes back from TMEM, is divided by the row sum in float32 and stored as the 16-bit row (t, b) of the context
// matrix, columns h*64.., i.e. the concatenation of heads that MultiHead.fc consumes (models.py:874-876).
// The logits are exponentiated, so operand rounding in Q K^T is amplified by their magnitude (digital silence drives
// the features far outside their usual range: 1e-2 on the output probabilities with plain 16-bit q, k).  With the
// residual tiles q_lo, k_lo of sed_linear_split16 the logits are accumulated as q_hi k_hi + q_lo k_hi + q_hi k_lo
// (12 instead of 4 MMAs of a kernel that is nowhere near the tensor pipe's limit).
// Two CTAs per SM (80 KB of shared memory, 256 TMEM columns each) overlap each other's load / MMA / softmax phases;
// P re-uses the Q tiles' storage (Q and K are dead once S is in TMEM).
#include "sed_common.cuh"
#include "sed_kernels.h"

namespace sed {

namespace attn {
constexpr int kThreads = 128;
constexpr int kTile = 128 * 128;             // one [128 rows x 64 cols] 16-bit SWIZZLE_128B tile
constexpr int kSmem = 1024 + 5 * kTile + 256;  // Q_hi, Q_lo (later P: two 64-key chunks), K_hi, K_lo, V + barriers
constexpr uint32_t kTmemCols = 256;          // S: columns 0..127, O: columns 128..191
}  // namespace attn

SED_DEVICE_INLINE void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

SED_DEVICE_INLINE float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename T, bool SPLIT>
__global__ void __launch_bounds__(attn::kThreads, 2)
mha_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmLo, int B, int Tn, long Bp,
              T* __restrict__ ctx) {
  using namespace attn;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sQ = smem;                // Q_hi, Q_lo
  uint8_t* sP = smem;                // [2 chunks of 64 keys][128 query rows][128 B], written after S is complete
  uint8_t* sK = smem + 2 * kTile;    // K_hi, K_lo
  uint8_t* sV = smem + 4 * kTile;
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + 5 * kTile);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    if (SPLIT) tma_prefetch_desc(&tmLo);
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_row = tmem + (static_cast<uint32_t>(warp * 32) << 16);  // this warp's 32 TMEM lanes
  const int row = warp * 32 + lane;                                        // query step of this thread

  constexpr uint32_t idesc_s = umma_idesc_f16(Elem16<T>::kFmt, 128, 128);
  constexpr uint32_t idesc_o = umma_idesc_f16(Elem16<T>::kFmt, 128, 64) | (1u << 16);  // B operand MN-major
  // softmax(s / sqrt(64)) = exp2((s - max) * log2(e) / 8)   (temperature: models.py:811, 843)
  constexpr float kScale = 0.125f * 1.4426950408889634f;

  const int units = B * 8;
  uint32_t ph_load = 0, ph_mma = 0;
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int b = u >> 3, h = u & 7;
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar_load, (SPLIT ? 5 : 3) * kTile);
      tma_load_3d(sQ, &tmQKV, bar_load, h * 64, b, 0);
      tma_load_3d(sK, &tmQKV, bar_load, 512 + h * 64, b, 0);
      tma_load_3d(sV, &tmQKV, bar_load, 1024 + h * 64, b, 0);
      if (SPLIT) {
        tma_load_3d(sQ + kTile, &tmLo, bar_load, h * 64, b, 0);
        tma_load_3d(sK + kTile, &tmLo, bar_load, 512 + h * 64, b, 0);
      }
      mbar_wait(bar_load, ph_load);
      tc_fence_after();
      const uint32_t aq = smem_u32(sQ), bk = smem_u32(sK);
#pragma unroll
      for (int k = 0; k < 4; ++k)  // S[q][key] += Q[q][16 d] * K[key][16 d]
        umma_f16(tmem, umma_desc_sw128(aq + k * 32, 1024), umma_desc_sw128(bk + k * 32, 1024), idesc_s, k ? 1u : 0u);
      if (SPLIT) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // + Q_lo K_hi + Q_hi K_lo
          umma_f16(tmem, umma_desc_sw128(aq + kTile + k * 32, 1024), umma_desc_sw128(bk + k * 32, 1024), idesc_s, 1u);
          umma_f16(tmem, umma_desc_sw128(aq + k * 32, 1024), umma_desc_sw128(bk + kTile + k * 32, 1024), idesc_s, 1u);
        }
      }
      umma_commit(bar_mma);
    }
    ph_load ^= 1;
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tc_fence_after();

    // ---- softmax over the keys of this thread's query row ----
    float s[128];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint32_t r[16];
      tmem_ld16(t_row + c * 16, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) s[c * 16 + j] = __uint_as_float(r[j]);
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 128; ++j)
      if (j < Tn) mx = fmaxf(mx, s[j]);
    const float off = mx * kScale;
    float l = 0.0f;
#pragma unroll
    for (int j = 0; j < 128; ++j) {
      const float p = (j < Tn) ? ex2f(fmaf(s[j], kScale, -off)) : 0.0f;
      s[j] = p;
      l += p;
    }
    // P as the K-major SWIZZLE_128B A operand: 16-byte unit q8 (8 keys) of row r in chunk q8 / 8 sits at
    // chunk*16 KB + r*128 + ((q8 % 8) ^ (r & 7)) * 16
#pragma unroll
    for (int q8 = 0; q8 < 16; ++q8) {
      uint4 pk;
      pk.x = Elem16<T>::pack2(s[8 * q8], s[8 * q8 + 1]);
      pk.y = Elem16<T>::pack2(s[8 * q8 + 2], s[8 * q8 + 3]);
      pk.z = Elem16<T>::pack2(s[8 * q8 + 4], s[8 * q8 + 5]);
      pk.w = Elem16<T>::pack2(s[8 * q8 + 6], s[8 * q8 + 7]);
      *reinterpret_cast<uint4*>(sP + (q8 >> 3) * kTile + row * 128 + (((q8 & 7) ^ (row & 7)) << 4)) = pk;
    }
    fence_proxy_async_smem();  // generic stores of P -> the tensor core's (async proxy) reads
    tc_fence_before();         // this thread's TMEM loads of S are done before O may be accumulated
    __syncthreads();

    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t ap = smem_u32(sP), bv = smem_u32(sV);
#pragma unroll
      for (int k = 0; k < 8; ++k)  // O[q][d] += P[q][16 keys] * V[16 keys][d]; V rows are keys: 16 keys = 2 KB
        umma_f16(tmem + 128, umma_desc_sw128(ap + (k >> 2) * kTile + (k & 3) * 32, 1024),
                 umma_desc_sw128(bv + k * 2048, 1024), idesc_o, k ? 1u : 0u);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tc_fence_after();

    const float inv = 1.0f / l;
    T* dst = ctx + (static_cast<size_t>(row) * Bp + b) * 512 + h * 64;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[16];
      tmem_ld16(t_row + 128 + c * 16, r);
      tmem_ld_wait();
      if (row < Tn) {
        uint4 q0, q1;
        q0.x = Elem16<T>::pack2(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv);
        q0.y = Elem16<T>::pack2(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv);
        q0.z = Elem16<T>::pack2(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv);
        q0.w = Elem16<T>::pack2(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv);
        q1.x = Elem16<T>::pack2(__uint_as_float(r[8]) * inv, __uint_as_float(r[9]) * inv);
        q1.y = Elem16<T>::pack2(__uint_as_float(r[10]) * inv, __uint_as_float(r[11]) * inv);
        q1.z = Elem16<T>::pack2(__uint_as_float(r[12]) * inv, __uint_as_float(r[13]) * inv);
        q1.w = Elem16<T>::pack2(__uint_as_float(r[14]) * inv, __uint_as_float(r[15]) * inv);
        reinterpret_cast<uint4*>(dst + c * 16)[0] = q0;
        reinterpret_cast<uint4*>(dst + c * 16)[1] = q1;
      }
    }
    tc_fence_before();
    __syncthreads();  // every warp has drained O and read Q / K / V's successor may land
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, kTmemCols);
  }
}

typedef CUresult (*EncodeTiledFnA)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_qkv_map(CUtensorMap* tm, const void* base, int cols, long Bp, int Tn, int dtype) {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return SED_ERR_DRIVER;
  }
  EncodeTiledFnA enc = reinterpret_cast<EncodeTiledFnA>(fp);
  const cuuint64_t gdim[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(Bp), static_cast<cuuint64_t>(Tn)};
  const cuuint64_t gstr[2] = {static_cast<cuuint64_t>(cols) * 2, static_cast<cuuint64_t>(Bp) * cols * 2};
  const cuuint32_t box[3] = {64, 1, 128};  // steps >= T are zero-filled
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                   const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("mha_tc: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
    return SED_ERR_DRIVER;
  }
  return SED_OK;
}

template <typename T, bool SPLIT>
static cudaError_t launch_mha_tc(const CUtensorMap& tm, const CUtensorMap& tml, int grid, int B, int Tn, long Bp,
                                 void* ctx16, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(mha_tc_kernel<T, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::kSmem);
  if (e == cudaSuccess)
    mha_tc_kernel<T, SPLIT><<<grid, attn::kThreads, attn::kSmem, stream>>>(tm, tml, B, Tn, Bp, static_cast<T*>(ctx16));
  return e;
}

// qkv16: [T * Bp, 1536] 16-bit, row (t, b) = t * Bp + b, columns [q | k | v], head h = 64 columns at h*64 of each part.
// qk_lo16: optional [T * Bp, 1024] residuals of q | k.  ctx16: [T * Bp, 512] 16-bit, same row order; rows of clips >= B
// are not written.
int mha_tc_launch(const void* qkv16, const void* qk_lo16, int B, int Tn, long Bp, void* ctx16, int dtype,
                  cudaStream_t stream) {
  if (B <= 0 || Tn <= 0 || Tn > 128 || Bp < B) {
    set_error("mha_tc: unsupported shape B=%d T=%d Bp=%ld (T <= 128 pooled steps)", B, Tn, Bp);
    return SED_ERR_BAD_SHAPE;
  }
  if (dtype != 0 && dtype != 1) {
    set_error("mha_tc: dtype must be 0 (fp16) or 1 (bf16)");
    return SED_ERR_UNSUPPORTED;
  }
  CUtensorMap tm, tml;
  int rc = encode_qkv_map(&tm, qkv16, 1536, Bp, Tn, dtype);
  if (rc != SED_OK) return rc;
  tml = tm;
  if (qk_lo16 != nullptr && (rc = encode_qkv_map(&tml, qk_lo16, 1024, Bp, Tn, dtype)) != SED_OK) return rc;
  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm_count <= 0)
      sm_count = 148;
  }
  const int units = B * 8;
  const int grid = units < 2 * sm_count ? units : 2 * sm_count;
  const bool split = qk_lo16 != nullptr;
  cudaError_t e;
  if (dtype == 0)
    e = split ? launch_mha_tc<__half, true>(tm, tml, grid, B, Tn, Bp, ctx16, stream)
              : launch_mha_tc<__half, false>(tm, tml, grid, B, Tn, Bp, ctx16, stream);
  else
    e = split ? launch_mha_tc<__nv_bfloat16, true>(tm, tml, grid, B, Tn, Bp, ctx16, stream)
              : launch_mha_tc<__nv_bfloat16, false>(tm, tml, grid, B, Tn, Bp, ctx16, stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("mha_tc launch: %s", cudaGetErrorString(e));
    return SED_ERR_CUDA;
  }
  return SED_OK;
}

}  // namespace sed
