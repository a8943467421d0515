// Temporal blocks and the frame-attention head for sm_100a.
//
//  * gru_kernel      : persistent bidirectional GRU recurrence (pytorch/models.py:614-615, 670; gate order
//                      r, z, n).  A 4-CTA cluster owns 128 clips of one direction for all T steps; each CTA
//                      keeps a 192-row [r|z|n] x 64-unit slice of W_hh resident in shared memory and the f32
//                      state of its units in registers; h_t is exchanged through distributed shared memory
//                      (bulk shared->shared::cluster copies) once per step.
//  * mha_core_kernel : softmax(q k^T / sqrt(64)) v per (clip, head)  (models.py:799-820, 863-875), float32,
//                      one query row per thread, K/V of the head resident in shared memory (reference path of
//                      the tensor-core attention kernel in sed_attention.cu).
//  * attpool_kernel  : AttBlock + interpolate + pad_framewise_output (models.py:161-169, 84-95, 65-81).
#include <cstdio>
#include <cstdlib>

#include "sed_common.cuh"
#include "sed_kernels.h"

// Developer experiments (clock stamps, switched-off phases) exist only in a -DSED_PROFILE build (make profile);
// in the shipped library SED_DBG(x) is the constant 0 and the compiler drops every hook.
#ifdef SED_PROFILE
#define SED_DBG(x) (x)
#else
#define SED_DBG(x) 0
#endif

namespace sed {

// =================================================================================================
// GRU recurrence: one 4-CTA cluster per (128-clip block, direction)
// =================================================================================================
// CTA `q` of the cluster owns hidden units 64q..64q+63 of all three gates: its 192x256 slice of W_hh (two packed
// 96-row [r|z|n] x 32-unit blocks) stays resident in shared memory for all T steps and its 16 warps keep the
// float32 state of those units in registers.  Per step every CTA
//   (1) issues 16 tcgen05.mma (M = 128 clips, N = 192, K = 256) against the full 16-bit h_{t-1} [128 x 256] in its
//       shared memory (UMMA A operand, SWIZZLE_128B, double buffered).  K-chunk kc of that operand is exactly the
//       slice produced by CTA kc, so the four MMAs of a chunk go out as soon as that slice has landed;
//   (2) runs the gate math for its 64 units (one reciprocal shared by four sigmoids / two tanh: the phase is
//       MUFU-bound) and writes h_t (f32) to the output in coalesced 128-clip blocks;
//   (3) stores the 16-bit copy of its slice of h_t into its own A buffer; the last warp to finish ships the
//       contiguous 16 KB slice to the three peers with bulk shared->distributed-shared copies that complete bytes
//       on the peers' per-slice mbarriers.
// No global-memory round trip and no generic->async proxy fence over global memory (fence.proxy.async lowers to
// MEMBAR.GPU, several microseconds per step) sit on the step's critical path; the exchange itself runs at the
// ~12-17 B/clk per SM that distributed shared memory delivers.
// Sixteen clusters (B = 1024, both directions) are resident at once; an 8-CTA cluster design fits only 15.
constexpr int kGruCluster = 4;
constexpr int kGruRows = 192;                        // 64 hidden units x 3 gates = two 96-row packed blocks
constexpr int kGruWBytes = 4 * kGruRows * 128;       // 4 k-chunks x 192 rows x 128 B
constexpr int kGruABytes = 4 * 16384;                // 128 clips x 256 k x 2 B (one buffer)
constexpr int kGruSmem = 1024 + kGruWBytes + 2 * kGruABytes + 1024;
constexpr int kGruWarps = 16;
constexpr int kGruThreads = 32 * kGruWarps;          // 16 gate-math warps (a 17th warp would cap registers at 96)
constexpr int kGruIssueWarp = 8;                     // lane 0 of this warp also issues the MMAs

SED_DEVICE_INLINE float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
SED_DEVICE_INLINE float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sigmoid of four pre-activations with ONE reciprocal (the gate phase is MUFU-bound): 1/(1+e^-x) for each, the
// four denominators share rcp(d0 d1 d2 d3).  Inputs are clamped at -20 (sigmoid(-20) = 2e-9) so the product of
// the denominators stays below 2^116.
SED_DEVICE_INLINE void sigmoid4(float x0, float x1, float x2, float x3, float& s0, float& s1, float& s2, float& s3) {
  constexpr float kNegLog2e = -1.4426950408889634f;
  const float d0 = 1.0f + ex2_approx(fmaxf(x0, -20.0f) * kNegLog2e);
  const float d1 = 1.0f + ex2_approx(fmaxf(x1, -20.0f) * kNegLog2e);
  const float d2 = 1.0f + ex2_approx(fmaxf(x2, -20.0f) * kNegLog2e);
  const float d3 = 1.0f + ex2_approx(fmaxf(x3, -20.0f) * kNegLog2e);
  const float p01 = d0 * d1, p23 = d2 * d3;
  const float inv = rcp_approx(p01 * p23);
  const float i01 = inv * p23, i23 = inv * p01;
  s0 = i01 * d1; s1 = i01 * d0; s2 = i23 * d3; s3 = i23 * d2;
}
// tanh of two pre-activations with one reciprocal: 1 - 2 / (e^{2x} + 1); |x| clamped at 20 (tanh(20) = 1 - 8e-18)
SED_DEVICE_INLINE void tanh2(float x0, float x1, float& t0, float& t1) {
  constexpr float k2Log2e = 2.8853900817779268f;
  const float d0 = 1.0f + ex2_approx(fminf(x0, 20.0f) * k2Log2e);
  const float d1 = 1.0f + ex2_approx(fminf(x1, 20.0f) * k2Log2e);
  const float inv = rcp_approx(d0 * d1);
  t0 = fmaf(-2.0f * inv, d1, 1.0f);
  t1 = fmaf(-2.0f * inv, d0, 1.0f);
}

// 32 lanes x 8 consecutive 32-bit TMEM columns -> 8 registers per thread
SED_DEVICE_INLINE void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

// bulk copy own shared memory -> a peer CTA's shared memory, completing `bytes` on the peer's mbarrier
SED_DEVICE_INLINE void bulk_copy_to_peer(uint32_t dst_cluster_addr, const void* src_smem, uint32_t bytes,
                                         uint32_t mbar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dst_cluster_addr),
               "r"(smem_u32(src_smem)), "r"(bytes), "r"(mbar_cluster_addr)
               : "memory");
}
SED_DEVICE_INLINE void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <typename T>
__global__ void __cluster_dims__(kGruCluster, 1, 1) __launch_bounds__(kGruThreads, 1)
gru_kernel(const __grid_constant__ CUtensorMap tmW, const float* __restrict__ gi, const float* __restrict__ bhh, int B,
           int Tn, float* __restrict__ out, long long* __restrict__ stamps, int dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* smem_w = smem;                 // [4][192 rows][128 B]   resident W_hh slice
  uint8_t* smem_a = smem + kGruWBytes;    // [2][4][128 rows][128 B]  h (double buffered)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + 2 * kGruABytes);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;      // [2 buffers][4 source CTAs]: the 16 KB slice of h from CTA kc has landed
  uint64_t* acc_full = bars + 9;    // all 16 MMAs of the step have completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  unsigned int* s_done = reinterpret_cast<unsigned int*>(bars + 11);  // warps that finished their piece of h_t
  float* s_bias = reinterpret_cast<float*>(bars + 12);  // [3][64] b_hh of this CTA's units
  // profiling hook (stamps != nullptr): CTA 0 records clock64() at a few points of steps 8..15
  const bool prof = SED_DBG(stamps != nullptr && blockIdx.x == 0 && blockIdx.y == 0);
  dbg = SED_DBG(dbg);
#define GRU_STAMP(step, slot)                                                       \
  do {                                                                               \
    if (prof && (step) >= 8 && (step) < 16) stamps[((step) - 8) * 12 + (slot)] = clock64(); \
  } while (0)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x % kGruCluster;  // == %cluster_ctarank for cluster dims (4,1,1)
  const int blk = blockIdx.x / kGruCluster, nblk = gridDim.x / kGruCluster;
  const int dir = blockIdx.y;

  // h_{-1} = 0 (nn.GRU default h0): step 0 reads buffer 1
  for (int i = threadIdx.x; i < kGruABytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_a + kGruABytes)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (threadIdx.x < 192)
    s_bias[threadIdx.x] = bhh[dir * 768 + (threadIdx.x >> 6) * 256 + q * 64 + (threadIdx.x & 63)];
  if (warp == 0 && lane == 0) {
    *s_done = 0;
    tma_prefetch_desc(&tmW);
    mbar_init(w_full, 1);
    for (int i = 0; i < 8; ++i) mbar_init(a_full + i, 1);  // own slice: the shipping thread's arrive; peers: arming
    mbar_init(acc_full, 1);
    fence_barrier_init();
    // arm the peer slices of both h buffers: h_0 -> buffer 0 (read at step 1), h_1 -> buffer 1 (read at step 2)
    for (int kc = 0; kc < kGruCluster; ++kc) {
      if (kc == q) continue;
      if (Tn > 1) mbar_expect_tx(a_full + kc, 16384);
      if (Tn > 2) mbar_expect_tx(a_full + 4 + kc, 16384);
    }
    mbar_expect_tx(w_full, kGruWBytes);
#pragma unroll
    for (int kc = 0; kc < 4; ++kc)
      tma_load_2d(smem_w + kc * (kGruRows * 128), &tmW, w_full, kc * 64, dir * 768 + q * kGruRows);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // every CTA's barriers are initialised before any peer arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  {
    // ---------------- 16 gate-math warps; lane 0 of warp kGruIssueWarp also issues the MMAs ----------------
    const int quarter = warp & 3;             // TMEM lane quarter this warp may access
    const int g = warp >> 2;                  // which 16 of the CTA's 64 hidden units this warp owns
    const int m = quarter * 32 + lane;        // clip row of this thread
    const int ul = g * 16;                    // first local unit of this thread
    const int u0 = q * 64 + ul;               // first hidden unit of this thread
    const bool stamper = prof && warp == 0 && lane == 0;
    float h[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) h[j] = 0.0f;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + ul + (g >> 1) * 64;
    // exchange addresses: this thread's two 16-byte units of row m in chunk q of the A buffer (SWIZZLE_128B: unit u
    // of row m sits at u ^ (m & 7))
    const uint32_t xoff0 = q * 16384 + m * 128 + (((2 * g) ^ (m & 7)) << 4);
    const uint32_t xoff1 = q * 16384 + m * 128 + (((2 * g + 1) ^ (m & 7)) << 4);
    uint32_t a_remote[kGruCluster - 1], bar_remote[kGruCluster - 1];  // the three peers, starting at rank q + 1
#pragma unroll
    for (int d = 1; d < kGruCluster; ++d) {
      a_remote[d - 1] = map_to_cta(smem_a, (q + d) & (kGruCluster - 1));
      bar_remote[d - 1] = map_to_cta(a_full + q, (q + d) & (kGruCluster - 1));  // the peer's barrier for OUR slice
    }
    // gi / out are stored as 128-clip transposed blocks: float4 column c4 of clip row m of block (t, blk) sits at
    // ((tile * ncol4 + c4) * 128 + m), tile = t * nblk + blk, so every warp-level access is 512 contiguous bytes
    // (ncol4 = 384 for gi [dir][gate][256], 128 for the output [fwd | bwd]).
    const size_t gi_col = static_cast<size_t>((dir * 768 + u0) >> 2) * 128 + m;
    const size_t out_col = static_cast<size_t>((dir * 256 + u0) >> 2) * 128 + m;
    const float4* gi4 = reinterpret_cast<const float4*>(gi);
    float4* out4 = reinterpret_cast<float4*>(out);
    const long tstep = dir ? -1 : 1;
    long tile = (dir ? static_cast<long>(Tn - 1) : 0L) * nblk + blk;  // tile of the current step
    const long dtile = tstep * nblk;

    // input projections are software-pipelined through the same registers: the values of step s + 1 are loaded as
    // soon as the gate math of step s has consumed a chunk, so the loads spread over the whole gate phase instead
    // of bursting into the memory pipe right when the slice has to be shipped
    float4 gr[4], gz[4], gn[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const float4* p0 = gi4 + static_cast<size_t>(tile) * 384 * 128 + gi_col;
      gr[v] = p0[v * 128];
      gz[v] = p0[(64 + v) * 128];
      gn[v] = p0[(128 + v) * 128];
    }

    for (int s = 0; s < Tn; ++s) {
      const bool xchg = s + 1 < Tn;
      const uint32_t boff = (s & 1) * kGruABytes;
      const float4* gi_next = gi4 + static_cast<size_t>(tile + dtile) * 384 * 128 + gi_col;  // step s + 1 (if any)
      float4* out_t = out4 + static_cast<size_t>(tile) * 128 * 128 + out_col;
      if (s + 3 < Tn && (lane & 7) == 0 && !(dbg & 8)) {
        // pull the blocks of step s + 3 into L2 (one 128-byte line per 8 lanes)
        const float4* pf = gi4 + static_cast<size_t>(tile + 3 * dtile) * 384 * 128 + gi_col;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          prefetch_l2(pf + v * 128);
          prefetch_l2(pf + (64 + v) * 128);
          prefetch_l2(pf + (128 + v) * 128);
        }
      }
      if (warp == kGruIssueWarp) {
        if (lane == 0) {
          // K-chunk kc of the A operand is exactly the slice of h_{s-1} produced by CTA kc: the MMAs of a chunk are
          // issued as soon as that slice has landed (own slice first), so only the last slice's quarter of the
          // tensor work is exposed after the exchange.  (The gi loads of this warp were issued during the previous
          // gate phase, so nothing of its own delays this thread on its way here.)
          constexpr uint32_t idesc = umma_idesc_f16(Elem16<T>::kFmt, 128, kGruRows);
          const int buf = (s + 1) & 1;  // h_{s-1}
          const uint32_t a_base = smem_u32(smem_a + buf * kGruABytes), b_base = smem_u32(smem_w);
          if (s == 0) mbar_wait(w_full, 0);
#pragma unroll
          for (int i = 0; i < kGruCluster; ++i) {
            const int kc = (q + i) & (kGruCluster - 1);
            if (s > 0) {
              if (i == 0) GRU_STAMP(s, 3);
              mbar_wait_cluster(a_full + buf * 4 + kc, ((s - 1) >> 1) & 1);
              if (kc != q && s + 2 < Tn) mbar_expect_tx(a_full + buf * 4 + kc, 16384);  // re-arm for h_{s+1}
              // no proxy fence needed: peer slices are written by the async proxy (bulk copies) and the local
              // warps fenced their generic stores before arriving
            }
            if (i == 0) GRU_STAMP(s, 0);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_base, umma_desc_sw128(a_base + kc * 16384 + k * 32, 1024),
                       umma_desc_sw128(b_base + kc * (kGruRows * 128) + k * 32, 1024), idesc, (i | k) ? 1u : 0u);
            if (i == 0) GRU_STAMP(s, 1);
          }
          umma_commit(acc_full);
          GRU_STAMP(s, 2);
        }
        __syncwarp();
      }
      if (prof && s == 12 && lane == 0) stamps[(warp & 7) * 12 + 6 + (warp >> 3)] = clock64();  // per-warp top
      mbar_wait(acc_full, s & 1);
      if (stamper) GRU_STAMP(s, 4);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {  // two chunks of 8 units
        uint32_t ar[8], az[8], an[8];
        tmem_ld8(taddr + c * 8, ar);
        tmem_ld8(taddr + 32 + c * 8, az);
        tmem_ld8(taddr + 64 + c * 8, an);
        tmem_ld_wait();
        if (c == 1) tc_fence_before();  // this thread's TMEM loads of step s are done before its arrive below
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const float4 grv = gr[2 * c + v], gzv = gz[2 * c + v], gnv = gn[2 * c + v];
          const float gre[4] = {grv.x, grv.y, grv.z, grv.w};
          const float gze[4] = {gzv.x, gzv.y, gzv.z, gzv.w};
          const float gne[4] = {gnv.x, gnv.y, gnv.z, gnv.w};
#pragma unroll
          for (int e = 0; e < 4; e += 2) {
            const int jc = v * 4 + e;          // index within the chunk
            const int jj = c * 8 + jc;         // index within the thread's 16 units
            const int j = ul + jj;             // local unit (bias index)
            float r0, z0, r1, z1, n0, n1;
            sigmoid4(gre[e] + __uint_as_float(ar[jc]) + s_bias[j], gze[e] + __uint_as_float(az[jc]) + s_bias[64 + j],
                     gre[e + 1] + __uint_as_float(ar[jc + 1]) + s_bias[j + 1],
                     gze[e + 1] + __uint_as_float(az[jc + 1]) + s_bias[64 + j + 1], r0, z0, r1, z1);
            tanh2(fmaf(r0, __uint_as_float(an[jc]) + s_bias[128 + j], gne[e]),
                  fmaf(r1, __uint_as_float(an[jc + 1]) + s_bias[128 + j + 1], gne[e + 1]), n0, n1);
            h[jj] = fmaf(z0, h[jj] - n0, n0);              // (1 - z) n + z h
            h[jj + 1] = fmaf(z1, h[jj + 1] - n1, n1);
          }
        }
        if (xchg) {
          // 16-bit copy of these 8 units of h_t into this CTA's own buffer (s & 1), chunk q (swizzled layout); rows
          // of padded clips >= B carry well-defined values too
          uint4 pk;
          pk.x = Elem16<T>::pack2(h[8 * c], h[8 * c + 1]);     pk.y = Elem16<T>::pack2(h[8 * c + 2], h[8 * c + 3]);
          pk.z = Elem16<T>::pack2(h[8 * c + 4], h[8 * c + 5]); pk.w = Elem16<T>::pack2(h[8 * c + 6], h[8 * c + 7]);
          *reinterpret_cast<uint4*>(smem_a + boff + (c ? xoff1 : xoff0)) = pk;
          if (!(dbg & 1)) {  // this chunk's gi registers are free: fetch the values of step s + 1
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              gr[2 * c + v] = gi_next[(2 * c + v) * 128];
              gz[2 * c + v] = gi_next[(64 + 2 * c + v) * 128];
              gn[2 * c + v] = gi_next[(128 + 2 * c + v) * 128];
            }
          }
        }
        if (!(dbg & 4)) {  // h_t (f32) of these 8 units
          out_t[(2 * c) * 128] = make_float4(h[8 * c], h[8 * c + 1], h[8 * c + 2], h[8 * c + 3]);
          out_t[(2 * c + 1) * 128] = make_float4(h[8 * c + 4], h[8 * c + 5], h[8 * c + 6], h[8 * c + 7]);
        }
      }
      if (stamper) GRU_STAMP(s, 5);
      if (xchg) {
        // the last warp to finish ships the CTA's complete 16 KB slice (chunk q is contiguous) to the three peers
        // with bulk shared->distributed-shared copies that complete bytes on the peers' per-slice barriers
        fence_proxy_async_smem();  // this thread's generic stores -> the bulk copy's (async proxy) reads
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          const unsigned int prev = atomicAdd(s_done, 1u);
          if (prev == static_cast<unsigned int>(kGruWarps * (s + 1) - 1)) {
            __threadfence_block();
            fence_proxy_async_smem();
            const uint8_t* src = smem_a + boff + q * 16384;
#pragma unroll
            for (int d = 0; d < kGruCluster - 1; ++d)
              bulk_copy_to_peer(a_remote[d] + boff + q * 16384, src, 16384, bar_remote[d] + (s & 1) * 32);
            mbar_arrive(a_full + (s & 1) * 4 + q);  // own slice in place (every local warp drained its TMEM loads)
            GRU_STAMP(s, 9);
          }
        }
        if (stamper) GRU_STAMP(s, 8);
        if (prof && s == 12 && lane == 0) stamps[(warp & 7) * 12 + 10 + (warp >> 3)] = clock64();  // per-warp finish
      }
      tile += dtile;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
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
  (void)workspace;  // h is exchanged through distributed shared memory; the workspace argument is kept for ABI stability
  CUtensorMap tmW;
  if (encode_2d(enc, &tmW, dtype, const_cast<void*>(whh_packed), 256, 2 * 768, kGruRows) != SED_OK) {
    set_error("gru: cuTensorMapEncodeTiled failed");
    return SED_ERR_DRIVER;
  }
  dim3 grid(kGruCluster * (Bpad / 128), 2);
  cudaError_t e;
  int dbg = 0;
#ifdef SED_PROFILE
  const char* e_dbg2 = getenv("SED_GRU_DBG2");  // developer experiments: 1 = no gi loads, 4 = no out, 8 = no L2 prefetch
  dbg = e_dbg2 ? atoi(e_dbg2) : 0;
  if (getenv("SED_GRU_DBG")) {  // developer aid: how many 4-CTA clusters can be resident at once
    cudaFuncSetAttribute(gru_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGruSmem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(kGruThreads); cfg.dynamicSmemBytes = kGruSmem; cfg.stream = stream;
    int ncl = -1;
    cudaError_t qe = cudaOccupancyMaxActiveClusters(&ncl, gru_kernel<__half>, &cfg);
    fprintf(stderr, "[sed] gru: max active clusters = %d (%s), grid clusters = %d\n", ncl, cudaGetErrorString(qe),
            (Bpad / 128) * 2);
  }
#else
  (void)stamps;
#endif
  if (dtype == 0) {
    e = cudaFuncSetAttribute(gru_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGruSmem);
    if (e == cudaSuccess)
      gru_kernel<__half><<<grid, kGruThreads, kGruSmem, stream>>>(tmW, gi, bhh, B, Tn, out, stamps, dbg);
  } else {
    e = cudaFuncSetAttribute(gru_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGruSmem);
    if (e == cudaSuccess)
      gru_kernel<__nv_bfloat16><<<grid, kGruThreads, kGruSmem, stream>>>(tmW, gi, bhh, B, Tn, out, stamps, dbg);
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
mha_core_kernel(const float* __restrict__ qkv, int Tn, long rs_t, long rs_b, T* __restrict__ out16) {
  extern __shared__ float smem_kv[];
  float* Ks = smem_kv;            // [Tn][64]
  float* Vs = smem_kv + Tn * 64;  // [Tn][64]
  const int head = blockIdx.x, b = blockIdx.y;
  // row of (clip b, step t) = b*rs_b + t*rs_t: (Tn, 1) clip-major, (1, Bp) time-major over the padded batch
  const float* base = qkv + static_cast<size_t>(b) * rs_b * 1536;
  const size_t rstep = static_cast<size_t>(rs_t) * 1536;
  for (int i = threadIdx.x; i < Tn * 16; i += blockDim.x) {
    const int row = i >> 4, c4 = i & 15;
    reinterpret_cast<float4*>(Ks)[i] =
        *reinterpret_cast<const float4*>(base + row * rstep + 512 + head * 64 + c4 * 4);
    reinterpret_cast<float4*>(Vs)[i] =
        *reinterpret_cast<const float4*>(base + row * rstep + 1024 + head * 64 + c4 * 4);
  }
  __syncthreads();
  for (int qi = threadIdx.x; qi < Tn; qi += blockDim.x) {
    float q[64], o[64];
    const float* qp = base + qi * rstep + head * 64;
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
    T* op = out16 + (static_cast<size_t>(b) * rs_b + static_cast<size_t>(qi) * rs_t) * 512 + head * 64;
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

int mha_core_launch(const float* qkv, int B, int Tn, long rs_t, long rs_b, void* out16, int dtype,
                    cudaStream_t stream) {
  if (rs_t <= 0 || rs_b <= 0) {
    rs_t = 1;
    rs_b = Tn;
  }
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
      mha_core_kernel<__half><<<grid, 128, smem, stream>>>(qkv, Tn, rs_t, rs_b, reinterpret_cast<__half*>(out16));
  } else {
    e = cudaFuncSetAttribute(mha_core_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      mha_core_kernel<__nv_bfloat16><<<grid, 128, smem, stream>>>(qkv, Tn, rs_t, rs_b,
                                                                   reinterpret_cast<__nv_bfloat16*>(out16));
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

// Shared tail of the pooling head for one clip: s_e[t][25] = exp(att)+1e-6 and s_c[t][25] = sigmoid(cla) are in shared
// memory (all threads arrive here); computes the attention-weighted clip output and writes every output tensor.
__device__ __forceinline__ void attpool_tail(const float* s_e, const float* s_c, float* s_sum, int b, int Tn, int ratio,
                                             int frames_out, float* __restrict__ clip, float* __restrict__ frame,
                                             float* __restrict__ cla_t, float* __restrict__ norm_att_t) {
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
  // framewise: repeat each step `ratio` times, pad with the last frame (models.py:93-94, 74-78).  The clip's rows
  // are one contiguous run: 16-byte stores (a warp instruction covers 512 contiguous bytes) -- the destination may
  // be another GPU's memory (dist.PeerGather), where the number of store requests in flight is what limits the rate.
  float* fr = frame + static_cast<size_t>(b) * frames_out * kCls;
  const int total = frames_out * kCls;
  if ((total & 3) == 0 && (reinterpret_cast<uintptr_t>(fr) & 15) == 0) {
    float4* fr4 = reinterpret_cast<float4*>(fr);
    for (int i4 = threadIdx.x; i4 < (total >> 2); i4 += blockDim.x) {
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = 4 * i4 + e;
        const int f = i / kCls, c = i - f * kCls;
        int t = f / ratio;
        if (t > Tn - 1) t = Tn - 1;
        v[e] = s_c[t * kCls + c];
      }
      fr4[i4] = make_float4(v[0], v[1], v[2], v[3]);
    }
  } else {
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int f = i / kCls, c = i - f * kCls;
      int t = f / ratio;
      if (t > Tn - 1) t = Tn - 1;
      fr[i] = s_c[t * kCls + c];
    }
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
  attpool_tail(s_e, s_c, s_sum, b, Tn, ratio, frames_out, clip, frame, cla_t, norm_att_t);
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

// ---- blocks variant: x arrives as 128-clip transposed blocks (what sed_bigru / sed_linear out_layout 1 write) ----
// Kernel A: one warp per (32 consecutive clips, step t), lane = clip: every x load is 512 contiguous bytes and the
// 50 weight rows are broadcast reads from shared memory.  e = exp(clamp(att)) + 1e-6 and c = sigmoid(cla) go to a
// scratch tensor S[t][50][Bp] (clip fastest: coalesced).  Kernel B: one block per clip gathers its [T][50] slice and
// runs the same tail as attpool_kernel.  The sums over t stay sequential per clip: results do not depend on how
// the batch is split.
constexpr int kApWarps = 8;
constexpr int kApTM = 4;  // clips per lane: every broadcast weight load (the shared-memory pipe is the limiter)
                          // feeds 4 x kApTM FMAs; a warp covers 32 * kApTM clips of one 128-clip block

__global__ void __launch_bounds__(kApWarps * 32, 1)
attproj_blocks_kernel(const float4* __restrict__ xb, int Bp, int Tn, const float* __restrict__ w_att,
                      const float* __restrict__ b_att, const float* __restrict__ w_cla,
                      const float* __restrict__ b_cla, float* __restrict__ S) {
  extern __shared__ float smem_h[];
  float* s_w = smem_h;  // [2*25][512]  att rows then cla rows
  for (int i = threadIdx.x; i < kCls * 512; i += blockDim.x) {
    s_w[i] = w_att[i];
    s_w[kCls * 512 + i] = w_cla[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int GPB = 4 / kApTM;  // warp tasks per 128-clip block
  const int groups = (Bp / 128) * GPB, nblk = Bp / 128;
  const long tasks = static_cast<long>(groups) * Tn;
  for (long task = static_cast<long>(blockIdx.x) * kApWarps + warp; task < tasks;
       task += static_cast<long>(gridDim.x) * kApWarps) {
    const int t = static_cast<int>(task / groups), cg = static_cast<int>(task - static_cast<long>(t) * groups);
    const int blk = cg / GPB, m0 = (cg % GPB) * (32 * kApTM) + lane;  // this lane's clips: m0 + 32 i
    const float4* xr = xb + (static_cast<size_t>(t) * nblk + blk) * 128 * 128 + m0;
    float acc[kApTM][2 * kCls];
#pragma unroll
    for (int c = 0; c < kCls; ++c) {
      const float ba = b_att[c], bc = b_cla[c];
#pragma unroll
      for (int i = 0; i < kApTM; ++i) {
        acc[i][c] = ba;
        acc[i][kCls + c] = bc;
      }
    }
    float4 xv[kApTM];
#pragma unroll
    for (int i = 0; i < kApTM; ++i) xv[i] = xr[32 * i];
    for (int k4 = 0; k4 < 128; ++k4) {
      float4 xn[kApTM];
#pragma unroll
      for (int i = 0; i < kApTM; ++i)
        xn[i] = (k4 + 1 < 128) ? xr[static_cast<size_t>(k4 + 1) * 128 + 32 * i] : make_float4(0, 0, 0, 0);
#pragma unroll
      for (int c = 0; c < 2 * kCls; ++c) {
        const float4 wv = *reinterpret_cast<const float4*>(s_w + c * 512 + k4 * 4);
#pragma unroll
        for (int i = 0; i < kApTM; ++i) {
          acc[i][c] = fmaf(xv[i].x, wv.x, acc[i][c]);
          acc[i][c] = fmaf(xv[i].y, wv.y, acc[i][c]);
          acc[i][c] = fmaf(xv[i].z, wv.z, acc[i][c]);
          acc[i][c] = fmaf(xv[i].w, wv.w, acc[i][c]);
        }
      }
#pragma unroll
      for (int i = 0; i < kApTM; ++i) xv[i] = xn[i];
    }
    float* So = S + static_cast<size_t>(t) * 2 * kCls * Bp + blk * 128 + m0;
#pragma unroll
    for (int c = 0; c < kCls; ++c) {
#pragma unroll
      for (int i = 0; i < kApTM; ++i) {
        const float a = fminf(fmaxf(acc[i][c], -10.0f), 10.0f);                                    // models.py:164
        So[static_cast<size_t>(c) * Bp + 32 * i] = expf(a) + 1e-6f;                                  // models.py:165
        So[static_cast<size_t>(kCls + c) * Bp + 32 * i] = 1.0f / (1.0f + expf(-acc[i][kCls + c]));   // models.py:167
      }
    }
  }
}

__global__ void __launch_bounds__(128)
attpool_tail_kernel(const float* __restrict__ S, int Bp, int Tn, int ratio, int frames_out, int clip0,
                    float* __restrict__ clip, float* __restrict__ frame, float* __restrict__ cla_t,
                    float* __restrict__ norm_att_t) {
  extern __shared__ float smem_h[];
  float* s_e = smem_h;               // [Tn][25]
  float* s_c = s_e + Tn * kCls;      // [Tn][25]
  float* s_sum = s_c + Tn * kCls;    // [25]
  const int b = clip0 + blockIdx.x;  // outputs are indexed by the absolute clip
  for (int i = threadIdx.x; i < Tn * 2 * kCls; i += blockDim.x) {
    const int t = i / (2 * kCls), c = i - t * 2 * kCls;
    const float v = S[(static_cast<size_t>(t) * 2 * kCls + c) * Bp + b];
    if (c < kCls) s_e[t * kCls + c] = v; else s_c[t * kCls + c - kCls] = v;
  }
  attpool_tail(s_e, s_c, s_sum, b, Tn, ratio, frames_out, clip, frame, cla_t, norm_att_t);
}

size_t attpool_blocks_scratch_bytes(int B, int Tn) {
  const size_t Bp = (static_cast<size_t>(B) + 127) / 128 * 128;
  return sizeof(float) * static_cast<size_t>(Tn) * 2 * kCls * Bp;
}

// stage 1 (all clips) and stage 2 (clips [clip0, clip0 + n)) can be launched separately so that a caller can overlap
// the device->host copy of one range of clips with the tail kernel of the next
int attpool_blocks_launch(const float* x_blocks, int B, int Tn, const float* w_att, const float* b_att,
                          const float* w_cla, const float* b_cla, int ratio, int frames_out, void* scratch, float* clip,
                          float* frame, float* cla_t, float* norm_att_t, int stage, int clip0, int n,
                          cudaStream_t stream) {
  const size_t smem_b = sizeof(float) * (2 * static_cast<size_t>(Tn) * kCls + 32);
  if (B <= 0 || Tn <= 0 || ratio <= 0 || frames_out < Tn * ratio || smem_b > 220 * 1024 || stage < 0 || stage > 2 ||
      clip0 < 0 || n < 0 || clip0 + n > B) {
    set_error("attpool_blocks: unsupported shape B=%d T=%d ratio=%d frames_out=%d stage=%d clips [%d, %d)", B, Tn, ratio,
              frames_out, stage, clip0, clip0 + n);
    return SED_ERR_BAD_SHAPE;
  }
  const int Bp = (B + 127) / 128 * 128;
  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm_count <= 0)
      sm_count = 148;
  }
  cudaError_t e = cudaSuccess;
  if (stage == 0 || stage == 1) {
    const size_t smem_a = sizeof(float) * 2 * kCls * 512;
    e = cudaFuncSetAttribute(attproj_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a);
    if (e == cudaSuccess) {
      const long tasks = static_cast<long>(Bp / (32 * kApTM)) * Tn;
      long blocks = (tasks + kApWarps - 1) / kApWarps;
      if (blocks > static_cast<long>(sm_count)) blocks = sm_count;
      attproj_blocks_kernel<<<static_cast<unsigned>(blocks), kApWarps * 32, smem_a, stream>>>(
          reinterpret_cast<const float4*>(x_blocks), Bp, Tn, w_att, b_att, w_cla, b_cla,
          reinterpret_cast<float*>(scratch));
      e = cudaGetLastError();
    }
  }
  if (e == cudaSuccess && (stage == 0 || stage == 2) && n > 0) {
    e = cudaFuncSetAttribute(attpool_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
    if (e == cudaSuccess) {
      attpool_tail_kernel<<<n, 128, smem_b, stream>>>(reinterpret_cast<const float*>(scratch), Bp, Tn, ratio,
                                                      frames_out, clip0, clip, frame, cla_t, norm_att_t);
      e = cudaGetLastError();
    }
  }
  if (e != cudaSuccess) {
    set_error("attpool_blocks launch: %s", cudaGetErrorString(e));
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
  const int total = Tn * ratio * C;
  if ((total & 3) == 0 && (reinterpret_cast<uintptr_t>(fr) & 15) == 0) {  // 16-byte stores (see attpool_tail)
    float4* fr4 = reinterpret_cast<float4*>(fr);
    for (int i4 = threadIdx.x; i4 < (total >> 2); i4 += blockDim.x) {
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = 4 * i4 + e;
        const int f = i / C, c = i - f * C;
        v[e] = s_p[(f / ratio) * C + c];
      }
      fr4[i4] = make_float4(v[0], v[1], v[2], v[3]);
    }
  } else {
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int f = i / C, c = i - f * C;
      fr[i] = s_p[(f / ratio) * C + c];
    }
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
