// Temporal blocks and the frame-attention head for sm_100a.
//
//  * gru_kernel      : persistent bidirectional GRU recurrence (pytorch/models.py:614-615, 670; gate order
//                      r, z, n).  One CTA owns 128 clips of one direction for all T steps.  The hidden state
//                      is the UMMA A operand (16-bit, SWIZZLE_128B, double-buffered in shared memory); the
//                      recurrent weights stream through a TMA ring in 96-row blocks ordered [r|z|n] x 32
//                      hidden units, so the gate math for those units runs straight out of TMEM while the
//                      next block's MMA is in flight.  The float32 state lives in the output tensor.
//  * mha_core_kernel : softmax(q k^T / sqrt(64)) v per (clip, head)  (models.py:799-820, 863-875), float32,
//                      one query row per thread, K/V of the head resident in shared memory.
//  * attpool_kernel  : AttBlock + interpolate + pad_framewise_output (models.py:161-169, 84-95, 65-81).
#include "sed_common.cuh"
#include "sed_kernels.h"

namespace sed {

// =================================================================================================
// GRU recurrence
// =================================================================================================
constexpr int kGruChunkRows = 96;    // 32 hidden units x 3 gates per weight block
constexpr int kGruChunks = 8;        // 256 / 32
constexpr int kGruABuf = 4 * 16384;  // 128 clips x 256 k x 2 B
constexpr int kGruBStage = 4 * kGruChunkRows * 128;  // 4 k-chunks x 96 rows x 128 B
constexpr int kGruSB = 2;
constexpr int kGruSmem = 1024 + 2 * kGruABuf + kGruSB * kGruBStage + 256;

template <typename T>
__global__ void __launch_bounds__(192, 1)
gru_kernel(const __grid_constant__ CUtensorMap tmW, const float* __restrict__ gi, const float* __restrict__ bhh,
           int B, int Tn, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                       // [2][4][128 rows][128 B]
  uint8_t* smem_b = smem + 2 * kGruABuf;        // [SB][4][96 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + kGruSB * kGruBStage);
  uint64_t* b_full = bars;            // [2]
  uint64_t* b_empty = bars + 2;       // [2]
  uint64_t* acc_full = bars + 4;      // [2]
  uint64_t* acc_empty = bars + 6;     // [2]
  uint64_t* h_ready = bars + 8;       // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dir = blockIdx.y;
  const int clip0 = blockIdx.x * 128;

  // h_{-1} = 0 (nn.GRU default h0)
  for (int i = threadIdx.x; i < kGruABuf / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_a)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    mbar_init(h_ready, 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      uint32_t sb = 0, pb = 0;
      for (int s = 0; s < Tn; ++s) {
        for (int q = 0; q < kGruChunks; ++q) {
          mbar_wait(&b_empty[sb], pb ^ 1);
          mbar_expect_tx(&b_full[sb], kGruBStage);
#pragma unroll
          for (int kc = 0; kc < 4; ++kc)
            tma_load_2d(smem_b + sb * kGruBStage + kc * (kGruChunkRows * 128), &tmW, &b_full[sb], kc * 64,
                        dir * 768 + q * kGruChunkRows);
          if (++sb == kGruSB) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(Elem16<T>::kFmt, 128, kGruChunkRows);
      uint32_t sb = 0, pb = 0, acc = 0, pacc = 0;
      for (int s = 0; s < Tn; ++s) {
        if (s > 0) {
          mbar_wait(h_ready, (s - 1) & 1);
          tc_fence_after();
        }
        const uint32_t a_base = smem_u32(smem_a + (s & 1) * kGruABuf);
        for (int q = 0; q < kGruChunks; ++q) {
          mbar_wait(&acc_empty[acc], pacc ^ 1);
          mbar_wait(&b_full[sb], pb);
          tc_fence_after();
          const uint32_t b_base = smem_u32(smem_b + sb * kGruBStage);
#pragma unroll
          for (int kc = 0; kc < 4; ++kc) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_f16(tmem_base + acc * kGruChunkRows, umma_desc_sw128(a_base + kc * 16384 + k * 32, 1024),
                       umma_desc_sw128(b_base + kc * (kGruChunkRows * 128) + k * 32, 1024), idesc, (kc | k) ? 1u : 0u);
            }
          }
          umma_commit(&b_empty[sb]);
          umma_commit(&acc_full[acc]);
          if (++sb == kGruSB) { sb = 0; pb ^= 1; }
          if (++acc == 2) { acc = 0; pacc ^= 1; }
        }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const int clip = clip0 + m;
    const bool valid = clip < B;
    const float* bh = bhh + dir * 768;
    uint32_t acc = 0, pacc = 0;
    for (int s = 0; s < Tn; ++s) {
      const int t = dir ? (Tn - 1 - s) : s;
      const int t_prev = dir ? (t + 1) : (t - 1);
      const float* gi_row = gi + (static_cast<size_t>(clip) * Tn + t) * 1536 + dir * 768;
      float* out_row = out + (static_cast<size_t>(clip) * Tn + t) * 512 + dir * 256;
      const float* prev_row = out + (static_cast<size_t>(clip) * Tn + t_prev) * 512 + dir * 256;
      uint8_t* a_next = smem_a + ((s + 1) & 1) * kGruABuf;
      for (int q = 0; q < kGruChunks; ++q) {
        mbar_wait(&acc_full[acc], pacc);
        tc_fence_after();
        const uint32_t taddr = tmem_base + acc * kGruChunkRows + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t ar[16], az[16], an[16];
          tmem_ld16(taddr + half * 16, ar);
          tmem_ld16(taddr + 32 + half * 16, az);
          tmem_ld16(taddr + 64 + half * 16, an);
          tmem_ld_wait();
          const int j0 = q * 32 + half * 16;
          float hn[16];
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            float4 gr = make_float4(0, 0, 0, 0), gz = gr, gn = gr, hp = gr;
            if (valid) {
              gr = *reinterpret_cast<const float4*>(gi_row + j0 + v4 * 4);
              gz = *reinterpret_cast<const float4*>(gi_row + 256 + j0 + v4 * 4);
              gn = *reinterpret_cast<const float4*>(gi_row + 512 + j0 + v4 * 4);
              if (s > 0) hp = *reinterpret_cast<const float4*>(prev_row + j0 + v4 * 4);
            }
            const float grv[4] = {gr.x, gr.y, gr.z, gr.w}, gzv[4] = {gz.x, gz.y, gz.z, gz.w};
            const float gnv[4] = {gn.x, gn.y, gn.z, gn.w}, hpv[4] = {hp.x, hp.y, hp.z, hp.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int jj = v4 * 4 + e;
              const int j = j0 + jj;
              const float r = 1.0f / (1.0f + expf(-(grv[e] + __uint_as_float(ar[jj]) + __ldg(bh + j))));
              const float z = 1.0f / (1.0f + expf(-(gzv[e] + __uint_as_float(az[jj]) + __ldg(bh + 256 + j))));
              const float nn = tanhf(gnv[e] + r * (__uint_as_float(an[jj]) + __ldg(bh + 512 + j)));
              hn[jj] = (1.0f - z) * nn + z * hpv[e];
            }
            if (valid)
              *reinterpret_cast<float4*>(out_row + j0 + v4 * 4) =
                  make_float4(hn[v4 * 4], hn[v4 * 4 + 1], hn[v4 * 4 + 2], hn[v4 * 4 + 3]);
          }
          // 16-bit copy of h_t into the next step's A operand (SWIZZLE_128B K-major layout)
          const int kc = j0 >> 6;
          const int c16 = (j0 & 63) >> 3;
          uint8_t* rowp = a_next + kc * 16384 + m * 128;
          uint4 q0, q1;
          q0.x = Elem16<T>::pack2(hn[0], hn[1]);   q0.y = Elem16<T>::pack2(hn[2], hn[3]);
          q0.z = Elem16<T>::pack2(hn[4], hn[5]);   q0.w = Elem16<T>::pack2(hn[6], hn[7]);
          q1.x = Elem16<T>::pack2(hn[8], hn[9]);   q1.y = Elem16<T>::pack2(hn[10], hn[11]);
          q1.z = Elem16<T>::pack2(hn[12], hn[13]); q1.w = Elem16<T>::pack2(hn[14], hn[15]);
          if (!valid) { q0 = make_uint4(0, 0, 0, 0); q1 = q0; }
          *reinterpret_cast<uint4*>(rowp + ((c16 ^ (m & 7)) << 4)) = q0;
          *reinterpret_cast<uint4*>(rowp + (((c16 + 1) ^ (m & 7)) << 4)) = q1;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[acc]);
        if (++acc == 2) { acc = 0; pacc ^= 1; }
      }
      fence_proxy_async_smem();  // generic-proxy writes of h_t -> visible to the UMMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(h_ready);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int gru_launch(const float* gi, const void* whh_packed, const float* bhh, int B, int Tn, float* out, int dtype,
               cudaStream_t stream) {
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
  CUtensorMap tmW;
  const cuuint64_t gdim[2] = {256, 2 * 768};
  const cuuint64_t gstr[1] = {256 * 2};
  const cuuint32_t box[2] = {64, kGruChunkRows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(fp)(
      &tmW, dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
      const_cast<void*>(whh_packed), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gru: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return SED_ERR_DRIVER;
  }
  dim3 grid((B + 127) / 128, 2);
  cudaError_t e;
  if (dtype == 0) {
    e = cudaFuncSetAttribute(gru_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGruSmem);
    if (e == cudaSuccess) gru_kernel<__half><<<grid, 192, kGruSmem, stream>>>(tmW, gi, bhh, B, Tn, out);
  } else {
    e = cudaFuncSetAttribute(gru_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGruSmem);
    if (e == cudaSuccess) gru_kernel<__nv_bfloat16><<<grid, 192, kGruSmem, stream>>>(tmW, gi, bhh, B, Tn, out);
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

}  // namespace sed
