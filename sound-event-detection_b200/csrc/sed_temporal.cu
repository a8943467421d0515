// Temporal blocks and the frame-attention head for sm_100a.
//
//  * gru_kernel      : persistent bidirectional GRU recurrence (pytorch/models.py:614-615, 670; gate order
//                      r, z, n).  An 8-CTA cluster owns 128 clips of one direction for all T steps; each CTA
//                      keeps a 96-row [r|z|n] x 32-unit slice of W_hh resident in shared memory and the f32
//                      state of its units in registers; h_t is exchanged through L2 + TMA once per step.
//  * mha_core_kernel : softmax(q k^T / sqrt(64)) v per (clip, head)  (models.py:799-820, 863-875), float32,
//                      one query row per thread, K/V of the head resident in shared memory.
//  * attpool_kernel  : AttBlock + interpolate + pad_framewise_output (models.py:161-169, 84-95, 65-81).
#include <cstdio>
#include <cstdlib>

#include "sed_common.cuh"
#include "sed_kernels.h"

namespace sed {

// =================================================================================================
// GRU recurrence: one 8-CTA cluster per (128-clip block, direction)
// =================================================================================================
// CTA `q` of the cluster owns hidden units 32q..32q+31 of all three gates: its 96x256 slice of W_hh stays
// resident in shared memory for all T steps and its threads keep the float32 state of those units in
// registers.  Per step every CTA (1) TMA-loads the full 16-bit h_{t-1} [128 x 256] (the UMMA A operand),
// (2) issues 16 tcgen05.mma (M=128, N=96, K=256) into TMEM, (3) runs the gate math for its 32 units, writes
// h_t (f32) to the output and the 16-bit copy to a double-buffered exchange tensor in global memory (L2),
// and (4) signals the h_ready mbarrier of all 8 CTAs (remote arrive, release/acquire at cluster scope).
constexpr int kGruCluster = 8;
constexpr int kGruChunkRows = 96;                    // 32 hidden units x 3 gates
constexpr int kGruWBytes = 4 * kGruChunkRows * 128;  // 4 k-chunks x 96 rows x 128 B
constexpr int kGruABytes = 4 * 16384;                // 128 clips x 256 k x 2 B
constexpr int kGruSmem = 1024 + kGruWBytes + kGruABytes + 512;
constexpr int kGruThreads = 64 + 256;  // TMA warp, MMA warp, 8 gate-math warps (2 per TMEM lane quarter)

SED_DEVICE_INLINE float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
SED_DEVICE_INLINE float fast_tanh(float x) {
  // 1 - 2 / (exp(2x) + 1); |error| ~1e-7 absolute, far below the 16-bit operand rounding of h W_hh^T
  return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f);
}

template <typename T>
__global__ void __cluster_dims__(kGruCluster, 1, 1) __launch_bounds__(kGruThreads, 1)
gru_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH,
           const float* __restrict__ gi, const float* __restrict__ bhh, int B, int Bpad, int Tn,
           float* __restrict__ out, T* __restrict__ hx, long long* __restrict__ stamps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem;                 // [4][96 rows][128 B]   resident W_hh slice
  // profiling hook (stamps != nullptr): CTA 0 records clock64() at 12 points of steps 8..15
  const bool prof = stamps != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
#define GRU_STAMP(step, slot)                                                       \
  do {                                                                               \
    if (prof && (step) >= 8 && (step) < 16) stamps[((step) - 8) * 12 + (slot)] = clock64(); \
  } while (0)
  uint8_t* smem_a = smem + kGruWBytes;    // [4][128 rows][128 B]  h_{t-1}
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + kGruABytes);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;
  uint64_t* acc_full = bars + 2;
  uint64_t* h_ready = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  float* s_bias = reinterpret_cast<float*>(bars + 5);  // [3][32] b_hh of this CTA's units

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x % kGruCluster;  // == %cluster_ctarank for cluster dims (8,1,1)
  const int clip0 = (blockIdx.x / kGruCluster) * 128;
  const int dir = blockIdx.y;

  // h_{-1} = 0 (nn.GRU default h0)
  for (int i = threadIdx.x; i < kGruABytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_a)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (threadIdx.x < 96)
    s_bias[threadIdx.x] = bhh[dir * 768 + (threadIdx.x >> 5) * 256 + q * 32 + (threadIdx.x & 31)];
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmH);
    mbar_init(w_full, 1);
    mbar_init(a_full, 1);
    mbar_init(acc_full, 1);
    mbar_init(h_ready, kGruCluster);  // one arrival per source CTA
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // every CTA's barriers are initialised before anyone signals them remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(w_full, kGruWBytes);
#pragma unroll
      for (int kc = 0; kc < 4; ++kc)
        tma_load_2d(smem_w + kc * (kGruChunkRows * 128), &tmW, w_full, kc * 64, dir * 768 + q * kGruChunkRows);
      for (int s = 1; s < Tn; ++s) {
        mbar_wait_cluster(h_ready, (s - 1) & 1);  // all 8 slices of h_{s-1} are in the exchange buffer
        GRU_STAMP(s, 0);  // (the writers issued fence.proxy.async before their release-arrive)
        GRU_STAMP(s, 1);
        mbar_expect_tx(a_full, kGruABytes);
        const int row = (((s - 1) & 1) * 2 + dir) * Bpad + clip0;
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) tma_load_2d(smem_a + kc * 16384, &tmH, a_full, kc * 64, row);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(Elem16<T>::kFmt, 128, kGruChunkRows);
      mbar_wait(w_full, 0);
      const uint32_t a_base = smem_u32(smem_a), b_base = smem_u32(smem_w);
      for (int s = 0; s < Tn; ++s) {
        if (s > 0) mbar_wait(a_full, (s - 1) & 1);
        GRU_STAMP(s, 2);
        tc_fence_after();
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_f16(tmem_base, umma_desc_sw128(a_base + kc * 16384 + k * 32, 1024),
                     umma_desc_sw128(b_base + kc * (kGruChunkRows * 128) + k * 32, 1024), idesc, (kc | k) ? 1u : 0u);
          }
        }
        umma_commit(acc_full);
        GRU_STAMP(s, 3);
      }
    }
  } else {
    const int quarter = warp & 3;             // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;         // which 16 of the CTA's 32 hidden units this warp owns
    const int m = quarter * 32 + lane;
    const int clip = clip0 + m;
    const bool valid = clip < B;
    const int u0 = q * 32 + half * 16;        // first hidden unit of this thread
    float h[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) h[j] = 0.0f;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + half * 16;
    for (int s = 0; s < Tn; ++s) {
      const int t = dir ? (Tn - 1 - s) : s;
      const float* gi_row = gi + (static_cast<size_t>(clip) * Tn + t) * 1536 + dir * 768 + u0;
      float* out_row = out + (static_cast<size_t>(clip) * Tn + t) * 512 + dir * 256 + u0;
      // input projections for this step: issued before the accumulator wait so they overlap the MMA
      float4 gr[4], gz[4], gn[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        if (valid) {
          gr[v] = *reinterpret_cast<const float4*>(gi_row + v * 4);
          gz[v] = *reinterpret_cast<const float4*>(gi_row + 256 + v * 4);
          gn[v] = *reinterpret_cast<const float4*>(gi_row + 512 + v * 4);
        } else {
          gr[v] = gz[v] = gn[v] = make_float4(0, 0, 0, 0);
        }
      }
      if (warp == 2 && lane == 0) GRU_STAMP(s, 4);
      mbar_wait(acc_full, s & 1);
      if (warp == 2 && lane == 0) GRU_STAMP(s, 5);
      tc_fence_after();
      uint32_t ar[16], az[16], an[16];
      tmem_ld16(taddr, ar);
      tmem_ld16(taddr + 32, az);
      tmem_ld16(taddr + 64, an);
      tmem_ld_wait();
#pragma unroll
      for (int v4 = 0; v4 < 4; ++v4) {
        const float grv[4] = {gr[v4].x, gr[v4].y, gr[v4].z, gr[v4].w};
        const float gzv[4] = {gz[v4].x, gz[v4].y, gz[v4].z, gz[v4].w};
        const float gnv[4] = {gn[v4].x, gn[v4].y, gn[v4].z, gn[v4].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int jj = v4 * 4 + e;
          const int j = half * 16 + jj;
          const float r = fast_sigmoid(grv[e] + __uint_as_float(ar[jj]) + s_bias[j]);
          const float z = fast_sigmoid(gzv[e] + __uint_as_float(az[jj]) + s_bias[32 + j]);
          const float nn = fast_tanh(gnv[e] + r * (__uint_as_float(an[jj]) + s_bias[64 + j]));
          h[jj] = (1.0f - z) * nn + z * h[jj];
        }
      }
      if (warp == 2 && lane == 0) GRU_STAMP(s, 6);
      if (valid) {
#pragma unroll
        for (int v = 0; v < 4; ++v)
          *reinterpret_cast<float4*>(out_row + v * 4) = make_float4(h[4 * v], h[4 * v + 1], h[4 * v + 2], h[4 * v + 3]);
      }
      if (s + 1 < Tn) {
        // 16-bit copy of this thread's slice of h_t into exchange buffer (s & 1)
        T* hx_row = hx + (static_cast<size_t>((s & 1) * 2 + dir) * Bpad + clip) * 256 + u0;
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          uint4 pk;
          pk.x = Elem16<T>::pack2(h[8 * v], h[8 * v + 1]);
          pk.y = Elem16<T>::pack2(h[8 * v + 2], h[8 * v + 3]);
          pk.z = Elem16<T>::pack2(h[8 * v + 4], h[8 * v + 5]);
          pk.w = Elem16<T>::pack2(h[8 * v + 6], h[8 * v + 7]);
          reinterpret_cast<uint4*>(hx_row)[v] = pk;
        }
        if (warp == 2 && lane == 0) GRU_STAMP(s, 7);
        tc_fence_before();
        named_bar_sync(1, 256);  // all 8 gate warps have written their slice of h_t (and drained TMEM)
        if (warp == 2 && lane == 0) GRU_STAMP(s, 8);
        // warp (2 + r) publishes to CTA r: proxy fence (generic global writes -> the peers' TMA reads), then a
        // release-arrive at cluster scope; the eight destinations are signalled in parallel
        if (lane == 0) {
          fence_proxy_async_all();
          mbar_arrive_remote(h_ready, warp - 2);
        }
        if (warp == 2 && lane == 0) GRU_STAMP(s, 9);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_2d(EncodeTiledFn enc, CUtensorMap* m, int dtype, void* base, uint64_t cols, uint64_t rows,
                     uint32_t box_rows) {
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstr[1] = {cols * 2};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim,
                   gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SED_OK : SED_ERR_DRIVER;
}

size_t gru_workspace_bytes(int B) {
  const size_t bpad = (static_cast<size_t>(B) + 127) / 128 * 128;
  return 2 * 2 * bpad * 256 * 2;
}

int gru_launch(const float* gi, const void* whh_packed, const float* bhh, int B, int Tn, float* out, void* workspace,
               int dtype, cudaStream_t stream, long long* stamps) {
  if (B <= 0 || Tn <= 0) {
    set_error("gru: bad shape B=%d T=%d", B, Tn);
    return SED_ERR_BAD_SHAPE;
  }
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return SED_ERR_DRIVER;
  }
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  const int Bpad = (B + 127) / 128 * 128;
  CUtensorMap tmW, tmH;
  if (encode_2d(enc, &tmW, dtype, const_cast<void*>(whh_packed), 256, 2 * 768, kGruChunkRows) != SED_OK ||
      encode_2d(enc, &tmH, dtype, workspace, 256, 4ull * Bpad, 128) != SED_OK) {
    set_error("gru: cuTensorMapEncodeTiled failed");
    return SED_ERR_DRIVER;
  }
  dim3 grid(kGruCluster * (Bpad / 128), 2);
  cudaError_t e;
  if (getenv("SED_GRU_DBG")) {  // developer aid: how many 8-CTA clusters can be resident at once
    cudaFuncSetAttribute(gru_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGruSmem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(kGruThreads); cfg.dynamicSmemBytes = kGruSmem; cfg.stream = stream;
    int ncl = -1;
    cudaError_t qe = cudaOccupancyMaxActiveClusters(&ncl, gru_kernel<__half>, &cfg);
    fprintf(stderr, "[sed] gru: max active clusters = %d (%s), grid clusters = %d\n", ncl, cudaGetErrorString(qe),
            (Bpad / 128) * 2);
  }
  if (dtype == 0) {
    e = cudaFuncSetAttribute(gru_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGruSmem);
    if (e == cudaSuccess)
      gru_kernel<__half><<<grid, kGruThreads, kGruSmem, stream>>>(tmW, tmH, gi, bhh, B, Bpad, Tn, out,
                                                                  reinterpret_cast<__half*>(workspace), stamps);
  } else {
    e = cudaFuncSetAttribute(gru_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGruSmem);
    if (e == cudaSuccess)
      gru_kernel<__nv_bfloat16><<<grid, kGruThreads, kGruSmem, stream>>>(tmW, tmH, gi, bhh, B, Bpad, Tn, out,
                                                                         reinterpret_cast<__nv_bfloat16*>(workspace), stamps);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("gru launch: %s", cudaGetErrorString(e));
    return SED_ERR_CUDA;
  }
  return SED_OK;
}

// =================================================================================================
// Multi-head attention core (8 heads, d_k = d_v = 64)
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(128)
mha_core_kernel(const float* __restrict__ qkv, int Tn, T* __restrict__ out16) {
  extern __shared__ float smem_kv[];
  float* Ks = smem_kv;            // [Tn][64]
  float* Vs = smem_kv + Tn * 64;  // [Tn][64]
  const int head = blockIdx.x, b = blockIdx.y;
  const float* base = qkv + static_cast<size_t>(b) * Tn * 1536;
  for (int i = threadIdx.x; i < Tn * 16; i += blockDim.x) {
    const int row = i >> 4, c4 = i & 15;
    reinterpret_cast<float4*>(Ks)[i] =
        *reinterpret_cast<const float4*>(base + static_cast<size_t>(row) * 1536 + 512 + head * 64 + c4 * 4);
    reinterpret_cast<float4*>(Vs)[i] =
        *reinterpret_cast<const float4*>(base + static_cast<size_t>(row) * 1536 + 1024 + head * 64 + c4 * 4);
  }
  __syncthreads();
  for (int qi = threadIdx.x; qi < Tn; qi += blockDim.x) {
    float q[64], o[64];
    const float* qp = base + static_cast<size_t>(qi) * 1536 + head * 64;
#pragma unroll
    for (int d = 0; d < 64; d += 4) {
      const float4 v = *reinterpret_cast<const float4*>(qp + d);
      q[d] = v.x; q[d + 1] = v.y; q[d + 2] = v.z; q[d + 3] = v.w;
    }
#pragma unroll
    for (int d = 0; d < 64; ++d) o[d] = 0.0f;
    float mx = -INFINITY, l = 0.0f;
    for (int j = 0; j < Tn; ++j) {
      const float4* kr = reinterpret_cast<const float4*>(Ks + j * 64);
      float s = 0.0f;
#pragma unroll
      for (int d4 = 0; d4 < 16; ++d4) {
        const float4 kv = kr[d4];
        s = fmaf(q[4 * d4], kv.x, s);
        s = fmaf(q[4 * d4 + 1], kv.y, s);
        s = fmaf(q[4 * d4 + 2], kv.z, s);
        s = fmaf(q[4 * d4 + 3], kv.w, s);
      }
      s *= 0.125f;  // attn / temperature, temperature = sqrt(d_k) = 8 (models.py:811, 843)
      if (s > mx) {
        const float corr = expf(mx - s);
        l *= corr;
#pragma unroll
        for (int d = 0; d < 64; ++d) o[d] *= corr;
        mx = s;
      }
      const float pj = expf(s - mx);
      l += pj;
      const float4* vr = reinterpret_cast<const float4*>(Vs + j * 64);
#pragma unroll
      for (int d4 = 0; d4 < 16; ++d4) {
        const float4 vv = vr[d4];
        o[4 * d4] = fmaf(pj, vv.x, o[4 * d4]);
        o[4 * d4 + 1] = fmaf(pj, vv.y, o[4 * d4 + 1]);
        o[4 * d4 + 2] = fmaf(pj, vv.z, o[4 * d4 + 2]);
        o[4 * d4 + 3] = fmaf(pj, vv.w, o[4 * d4 + 3]);
      }
    }
    const float inv = 1.0f / l;
    T* op = out16 + (static_cast<size_t>(b) * Tn + qi) * 512 + head * 64;
#pragma unroll
    for (int d = 0; d < 64; d += 8) {
      uint4 pk;
      pk.x = Elem16<T>::pack2(o[d] * inv, o[d + 1] * inv);
      pk.y = Elem16<T>::pack2(o[d + 2] * inv, o[d + 3] * inv);
      pk.z = Elem16<T>::pack2(o[d + 4] * inv, o[d + 5] * inv);
      pk.w = Elem16<T>::pack2(o[d + 6] * inv, o[d + 7] * inv);
      *reinterpret_cast<uint4*>(op + d) = pk;
    }
  }
}

int mha_core_launch(const float* qkv, int B, int Tn, void* out16, int dtype, cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(Tn) * 64 * 2 * sizeof(float);
  if (B <= 0 || Tn <= 0 || smem > 200 * 1024) {
    set_error("mha_core: unsupported shape B=%d T=%d", B, Tn);
    return SED_ERR_BAD_SHAPE;
  }
  dim3 grid(8, B);
  cudaError_t e;
  if (dtype == 0) {
    e = cudaFuncSetAttribute(mha_core_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      mha_core_kernel<__half><<<grid, 128, smem, stream>>>(qkv, Tn, reinterpret_cast<__half*>(out16));
  } else {
    e = cudaFuncSetAttribute(mha_core_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      mha_core_kernel<__nv_bfloat16><<<grid, 128, smem, stream>>>(qkv, Tn, reinterpret_cast<__nv_bfloat16*>(out16));
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("mha_core launch: %s", cudaGetErrorString(e));
    return SED_ERR_CUDA;
  }
  return SED_OK;
}

// =================================================================================================
// Frame-attention pooling head
// =================================================================================================
constexpr int kCls = 25;  // AttBlock(512, 25): hard-coded in the reference (models.py:617, 1022)

__global__ void __launch_bounds__(128)
attpool_kernel(const float* __restrict__ x, int Tn, const float* __restrict__ w_att, const float* __restrict__ b_att,
               const float* __restrict__ w_cla, const float* __restrict__ b_cla, int ratio, int frames_out,
               float* __restrict__ clip, float* __restrict__ frame, float* __restrict__ cla_t,
               float* __restrict__ norm_att_t) {
  extern __shared__ float smem_h[];
  float* s_w = smem_h;                      // [2*25][512]  att rows then cla rows
  float* s_e = s_w + 2 * kCls * 512;        // [Tn][25]  exp(att)+1e-6
  float* s_c = s_e + Tn * kCls;             // [Tn][25]  sigmoid(cla)
  float* s_sum = s_c + Tn * kCls;           // [25]
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < kCls * 512; i += blockDim.x) {
    s_w[i] = w_att[i];
    s_w[kCls * 512 + i] = w_cla[i];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < Tn; t += blockDim.x) {
    float acc[2 * kCls];
#pragma unroll
    for (int c = 0; c < kCls; ++c) {
      acc[c] = b_att[c];
      acc[kCls + c] = b_cla[c];
    }
    const float4* xr = reinterpret_cast<const float4*>(x + (static_cast<size_t>(b) * Tn + t) * 512);
    for (int k4 = 0; k4 < 128; ++k4) {
      const float4 xv = xr[k4];
#pragma unroll
      for (int c = 0; c < 2 * kCls; ++c) {
        const float4 wv = *reinterpret_cast<const float4*>(s_w + c * 512 + k4 * 4);
        acc[c] = fmaf(xv.x, wv.x, acc[c]);
        acc[c] = fmaf(xv.y, wv.y, acc[c]);
        acc[c] = fmaf(xv.z, wv.z, acc[c]);
        acc[c] = fmaf(xv.w, wv.w, acc[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < kCls; ++c) {
      const float a = fminf(fmaxf(acc[c], -10.0f), 10.0f);        // models.py:164
      s_e[t * kCls + c] = expf(a) + 1e-6f;                         // models.py:165 (temperature 1)
      s_c[t * kCls + c] = 1.0f / (1.0f + expf(-acc[kCls + c]));    // models.py:167 sigmoid
    }
  }
  __syncthreads();
  if (threadIdx.x < kCls) {
    float s = 0.0f;
    for (int t = 0; t < Tn; ++t) s += s_e[t * kCls + threadIdx.x];
    s_sum[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x < kCls) {
    const int c = threadIdx.x;
    const float s = s_sum[c];
    float acc = 0.0f;
    for (int t = 0; t < Tn; ++t) acc += (s_e[t * kCls + c] / s) * s_c[t * kCls + c];  // models.py:166, 168
    clip[b * kCls + c] = acc;
  }
  // framewise: repeat each step `ratio` times, pad with the last frame (models.py:93-94, 74-78)
  float* fr = frame + static_cast<size_t>(b) * frames_out * kCls;
  for (int i = threadIdx.x; i < frames_out * kCls; i += blockDim.x) {
    const int f = i / kCls, c = i - f * kCls;
    int t = f / ratio;
    if (t > Tn - 1) t = Tn - 1;
    fr[i] = s_c[t * kCls + c];
  }
  if (cla_t != nullptr) {  // 'embedding' of the GRU model: cla [B, 25, T'] (models.py:686)
    float* o = cla_t + static_cast<size_t>(b) * kCls * Tn;
    for (int i = threadIdx.x; i < kCls * Tn; i += blockDim.x) {
      const int c = i / Tn, t = i - c * Tn;
      o[i] = s_c[t * kCls + c];
    }
  }
  if (norm_att_t != nullptr) {
    float* o = norm_att_t + static_cast<size_t>(b) * kCls * Tn;
    for (int i = threadIdx.x; i < kCls * Tn; i += blockDim.x) {
      const int c = i / Tn, t = i - c * Tn;
      o[i] = s_e[t * kCls + c] / s_sum[c];
    }
  }
}

int attpool_launch(const float* x, int B, int Tn, const float* w_att, const float* b_att, const float* w_cla,
                   const float* b_cla, int ratio, int frames_out, float* clip, float* frame, float* cla_t,
                   float* norm_att_t, cudaStream_t stream) {
  const size_t smem = sizeof(float) * (2 * kCls * 512 + 2 * static_cast<size_t>(Tn) * kCls + 32);
  if (B <= 0 || Tn <= 0 || ratio <= 0 || frames_out < Tn * ratio || smem > 220 * 1024) {
    set_error("attpool: unsupported shape B=%d T=%d ratio=%d frames_out=%d", B, Tn, ratio, frames_out);
    return SED_ERR_BAD_SHAPE;
  }
  cudaError_t e = cudaFuncSetAttribute(attpool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess)
    attpool_kernel<<<B, 128, smem, stream>>>(x, Tn, w_att, b_att, w_cla, b_cla, ratio, frames_out, clip, frame, cla_t,
                                             norm_att_t);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("attpool launch: %s", cudaGetErrorString(e));
    return SED_ERR_CUDA;
  }
  return SED_OK;
}

// =================================================================================================
// Linear + sigmoid + mean/max pooling head of the sibling models
// (Cnn_9layers_FrameAvg / FrameMax / Gru_FrameAvg / Transformer_FrameAvg, pytorch/models.py:276-288, 361-373,
//  547-556, 963-972): framewise = interpolate(sigmoid(fc(x)), 8); clipwise = mean | max over frames.
// =================================================================================================
constexpr int kFcMaxCls = 32;

__global__ void __launch_bounds__(128)
fcpool_kernel(const float* __restrict__ x, int Tn, const float* __restrict__ w, const float* __restrict__ b, int C,
              int ratio, int use_max, float* __restrict__ clip, float* __restrict__ frame) {
  extern __shared__ float smem_h[];
  float* s_w = smem_h;                        // [32][512], rows >= C are zero
  float* s_p = s_w + kFcMaxCls * 512;         // [Tn][C]
  float* s_b = s_p + Tn * C;                  // [32]
  const int bidx = blockIdx.x;
  for (int i = threadIdx.x; i < kFcMaxCls * 512; i += blockDim.x) s_w[i] = (i < C * 512) ? w[i] : 0.0f;
  if (threadIdx.x < kFcMaxCls) s_b[threadIdx.x] = (threadIdx.x < C) ? b[threadIdx.x] : 0.0f;
  __syncthreads();
  for (int t = threadIdx.x; t < Tn; t += blockDim.x) {
    float acc[kFcMaxCls];
#pragma unroll
    for (int c = 0; c < kFcMaxCls; ++c) acc[c] = s_b[c];
    const float4* xr = reinterpret_cast<const float4*>(x + (static_cast<size_t>(bidx) * Tn + t) * 512);
    for (int k4 = 0; k4 < 128; ++k4) {
      const float4 xv = xr[k4];
#pragma unroll
      for (int c = 0; c < kFcMaxCls; ++c) {
        const float4 wv = *reinterpret_cast<const float4*>(s_w + c * 512 + k4 * 4);
        acc[c] = fmaf(xv.x, wv.x, acc[c]);
        acc[c] = fmaf(xv.y, wv.y, acc[c]);
        acc[c] = fmaf(xv.z, wv.z, acc[c]);
        acc[c] = fmaf(xv.w, wv.w, acc[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < kFcMaxCls; ++c)
      if (c < C) s_p[t * C + c] = 1.0f / (1.0f + expf(-acc[c]));
  }
  __syncthreads();
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    float r = use_max ? -INFINITY : 0.0f;
    for (int t = 0; t < Tn; ++t) {
      const float v = s_p[t * C + c];
      r = use_max ? fmaxf(r, v) : r + v;
    }
    clip[bidx * C + c] = use_max ? r : r / static_cast<float>(Tn);
  }
  float* fr = frame + static_cast<size_t>(bidx) * Tn * ratio * C;
  for (int i = threadIdx.x; i < Tn * ratio * C; i += blockDim.x) {
    const int f = i / C, c = i - f * C;
    fr[i] = s_p[(f / ratio) * C + c];
  }
}

int fcpool_launch(const float* x, int B, int Tn, const float* w, const float* b, int C, int ratio, int use_max,
                  float* clip, float* frame, cudaStream_t stream) {
  const size_t smem = sizeof(float) * (kFcMaxCls * 512 + static_cast<size_t>(Tn) * C + kFcMaxCls);
  if (B <= 0 || Tn <= 0 || C <= 0 || C > kFcMaxCls || ratio <= 0 || smem > 220 * 1024) {
    set_error("fcpool: unsupported shape B=%d T=%d classes=%d (max %d) ratio=%d", B, Tn, C, kFcMaxCls, ratio);
    return SED_ERR_BAD_SHAPE;
  }
  cudaError_t e = cudaFuncSetAttribute(fcpool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) fcpool_kernel<<<B, 128, smem, stream>>>(x, Tn, w, b, C, ratio, use_max, clip, frame);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("fcpool launch: %s", cudaGetErrorString(e));
    return SED_ERR_CUDA;
  }
  return SED_OK;
}

}  // namespace sed
