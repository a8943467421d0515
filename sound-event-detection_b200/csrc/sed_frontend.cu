// Fused audio front-end for sm_100a: reflect-pad -> frame -> window -> real DFT -> power -> mel
// projection -> clamped 10*log10 -> bn0 affine, one kernel, float32 end to end.
//
// Reference: STFT.forward pytorch/stft.py:223-247 (two Conv1d with the windowed DFT matrix),
// Spectrogram.forward :651-670, LogmelFilterBank.forward/power_to_db :698-734, bn0 models.py:642-644.
//
// The reference evaluates the DFT as a dense [n_fft x (n_fft/2+1)] contraction (0.53 GFLOP per 10 s
// clip at 16 kHz).  Here two consecutive real frames are packed as one complex sequence and pushed
// through an in-shared-memory Stockham FFT (three passes of radix 4 / 8 / 16), 25x fewer flops, in
// float32 with float64-derived twiddles -- the accuracy class of the reference's float32 conv1d, which
// the 1e-4 log-mel tolerance needs (single-pass 16-bit tensor-core DFTs do not reach it).  The window
// is read from the loaded conv_real kernel (row 0), so a checkpoint's window is honoured; the host
// wrapper verifies that the loaded kernels are a windowed DFT before selecting this path.
// The mel projection uses the loaded melW in banded form (first/last non-zero per mel bin).
#include <atomic>
#include <cstdlib>

#include "sed_common.cuh"
#include "sed_kernels.h"

namespace sed {

template <int NFFT>
struct FrontCfg {
  static constexpr int WARPS = 8;
  // frames per work item (even: frames are transformed in pairs).  A persistent block loops over work items =
  // (clip, chunk of FPB frames); the waveform segment of the next item is staged with cp.async while the current
  // one is transformed.
  static constexpr int FPB = (NFFT == 1024) ? 16 : 32;
  static constexpr int MELV = 1024;  // banded mel weights cached in shared memory (falls back to global beyond)
  // window taps of the first pass: shared memory (one conflict-free LDS per tap).  A register copy does not survive
  // the 128-register cap of two resident blocks: the compiler re-loaded it from global memory for every frame pair
  // (ncu source view: 4.7 % of all instructions, long-scoreboard stalls)
  static constexpr bool WIN_REGS = false;
  static constexpr int MIN_BLOCKS = (NFFT == 1024) ? 1 : 2;
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// Shared-memory FFT buffers are XOR-swizzled (low four bits of the complex index ^= bits 3..6): the unit-stride reads of
// a Stockham pass stay conflict-free and its strided autosort writes spread over the banks.  Simulated 64-bit
// wavefronts per 512-point transform: 212 against 324 with one-in-eight padding (ideal 196); 1024: 612 vs 964.
__device__ __forceinline__ int pidx(int i) { return i ^ ((i >> 3) & 15); }

template <int R>
__device__ __forceinline__ void butterfly(float2 (&v)[R]);

template <>
__device__ __forceinline__ void butterfly<2>(float2 (&v)[2]) {
  const float2 a0 = v[0], a1 = v[1];
  v[0] = make_float2(a0.x + a1.x, a0.y + a1.y);
  v[1] = make_float2(a0.x - a1.x, a0.y - a1.y);
}

template <>
__device__ __forceinline__ void butterfly<4>(float2 (&v)[4]) {
  const float2 a0 = make_float2(v[0].x + v[2].x, v[0].y + v[2].y);
  const float2 a1 = make_float2(v[0].x - v[2].x, v[0].y - v[2].y);
  const float2 a2 = make_float2(v[1].x + v[3].x, v[1].y + v[3].y);
  const float2 a3 = make_float2(v[1].x - v[3].x, v[1].y - v[3].y);
  v[0] = make_float2(a0.x + a2.x, a0.y + a2.y);
  v[1] = make_float2(a1.x + a3.y, a1.y - a3.x);  // a1 - i*a3
  v[2] = make_float2(a0.x - a2.x, a0.y - a2.y);
  v[3] = make_float2(a1.x - a3.y, a1.y + a3.x);  // a1 + i*a3
}

template <>
__device__ __forceinline__ void butterfly<8>(float2 (&v)[8]) {
  constexpr float kS = 0.70710678118654752440f;
  float2 a[4], b[4];
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    a[n] = make_float2(v[n].x + v[n + 4].x, v[n].y + v[n + 4].y);
    b[n] = make_float2(v[n].x - v[n + 4].x, v[n].y - v[n + 4].y);
  }
  // b[n] *= W8^n  (W8 = exp(-2 pi i / 8))
  b[1] = make_float2((b[1].x + b[1].y) * kS, (b[1].y - b[1].x) * kS);
  b[2] = make_float2(b[2].y, -b[2].x);
  b[3] = make_float2((b[3].y - b[3].x) * kS, -(b[3].x + b[3].y) * kS);
  butterfly<4>(a);
  butterfly<4>(b);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[2 * k] = a[k];
    v[2 * k + 1] = b[k];
  }
}

template <>
__device__ __forceinline__ void butterfly<16>(float2 (&v)[16]) {
  // n = 4a + i, k = m + 4p:  X[m + 4p] = sum_i W4^{ip} W16^{im} sum_a x[4a + i] W4^{am}
  constexpr float kC1 = 0.92387953251128675613f, kS1 = 0.38268343236508977173f;  // cos, sin(pi / 8)
  constexpr float kS = 0.70710678118654752440f;
  float2 u[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t[4] = {v[i], v[i + 4], v[i + 8], v[i + 12]};
    butterfly<4>(t);
#pragma unroll
    for (int m = 0; m < 4; ++m) u[i][m] = t[m];
  }
  // u[i][m] *= W16^{i m},  W16 = exp(-2 pi i / 16)
  u[1][1] = cmul(u[1][1], make_float2(kC1, -kS1));
  u[1][2] = make_float2((u[1][2].x + u[1][2].y) * kS, (u[1][2].y - u[1][2].x) * kS);      // W16^2 = W8
  u[1][3] = cmul(u[1][3], make_float2(kS1, -kC1));
  u[2][1] = make_float2((u[2][1].x + u[2][1].y) * kS, (u[2][1].y - u[2][1].x) * kS);      // W16^2
  u[2][2] = make_float2(u[2][2].y, -u[2][2].x);                                            // W16^4 = -i
  u[2][3] = make_float2((u[2][3].y - u[2][3].x) * kS, -(u[2][3].x + u[2][3].y) * kS);     // W16^6
  u[3][1] = cmul(u[3][1], make_float2(kS1, -kC1));                                         // W16^3
  u[3][2] = make_float2((u[3][2].y - u[3][2].x) * kS, -(u[3][2].x + u[3][2].y) * kS);     // W16^6
  u[3][3] = cmul(u[3][3], make_float2(-kC1, kS1));                                         // W16^9 = -W16^1
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    float2 t[4] = {u[0][m], u[1][m], u[2][m], u[3][m]};
    butterfly<4>(t);
#pragma unroll
    for (int pq = 0; pq < 4; ++pq) v[m + 4 * pq] = t[pq];
  }
}

// Per-lane twiddles of one Stockham pass (radix R, stride Ns) of an N-point transform.  They depend only on the
// lane, not on the frame, so a persistent warp keeps them in registers; the widest case (N = 1024, Ns = 128) reads
// the shared-memory table instead.
template <int N, int R, int Ns>
struct PassTw {
  static constexpr int NB = N / R, BPL = NB / 32;
  static constexpr int NK = (Ns > 32) ? BPL : 1;  // distinct k = (lane + 32 q) % Ns over the lane's butterflies
  static constexpr bool IN_REGS = (Ns > 1) && (NK * (R - 1) <= 14);
  static constexpr int TSTRIDE = N / (Ns * R);
  float2 t[IN_REGS ? NK : 1][R - 1];
  __device__ __forceinline__ void init(const float2* __restrict__ tw, int lane) {
    if (IN_REGS) {
#pragma unroll
      for (int q = 0; q < NK; ++q) {
        const int k = (lane + 32 * q) % Ns;
#pragma unroll
        for (int r = 1; r < R; ++r) t[q][r - 1] = tw[r * k * TSTRIDE];
      }
    }
  }
};

// int16 PCM input follows the reference's HDF5 path: x = q / 32767 (utils/utilities.py:78-79).  The three-operation
// sequence below (multiply by the rounded reciprocal, exact residual, correction) returns the correctly rounded
// float32 quotient for every int16 q (checked exhaustively on the host) -- the value numpy's
// (q / 32767.).astype(float32) produces.
__device__ __forceinline__ float load_sample(const float* p) { return *p; }
__device__ __forceinline__ float load_sample(const short* p) {
  const float q = static_cast<float>(*p);
  constexpr float kInv = 1.0f / 32767.0f;
  const float r = q * kInv;
  const float e = fmaf(-r, 32767.0f, q);
  return fmaf(e, kInv, r);
}
__device__ __forceinline__ float load_sample_global(const float* p) { return __ldg(p); }
__device__ __forceinline__ float load_sample_global(const short* p) { return static_cast<float>(__ldg(p)); }
__device__ __forceinline__ void store_raw(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_raw(short* p, float v) { *p = static_cast<short>(v); }

// 10*log10(max(x, amin)) - db_offset (stft.py:726-727) as one MUFU.LG2 and one FMA.  lg2.approx is accurate to 2^-22
// (absolute inside (0.5, 2), relative elsewhere): at most 2.4e-5 dB over the whole range, against a tolerance of
// 1e-4 * max(|dB|, 1).  Values at or below amin return `db_floor`, the float32 value of 10*log10(amin) - db_offset the
// host computed with the reference's own operations (digital silence is exactly -100 dB).
__device__ __forceinline__ float power_to_db(float x, float amin, float db_offset, float db_floor) {
  float l;
  asm("lg2.approx.f32 %0, %1;" : "=f"(l) : "f"(x));
  return x > amin ? fmaf(l, 3.010299956639812f, -db_offset) : db_floor;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// One Stockham pass of radix R over N complex points, IN PLACE in (padded) shared memory, executed by one warp:
// every lane first pulls the inputs of all its butterflies into registers, the warp synchronises, then the
// autosorted outputs are written back.  FIRST: inputs come from the windowed frame pair instead (re = frame a,
// im = frame b), staged as raw input samples.
template <int N, int R, int Ns, bool FIRST, bool WINREG, typename TIn>
__device__ __forceinline__ void fft_pass(float2* __restrict__ buf, const PassTw<N, R, Ns>& tws,
                                         const float2* __restrict__ s_tw, const TIn* __restrict__ seg_a,
                                         const TIn* __restrict__ seg_b, const float* __restrict__ s_win,
                                         const float (&wreg)[N / 32], int lane) {
  constexpr int NB = N / R;
  constexpr int BPL = NB / 32;  // butterflies per lane
  float2 v[BPL][R];
#pragma unroll
  for (int q = 0; q < BPL; ++q) {
    const int j = lane + 32 * q;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int idx = j + r * NB;
      if (FIRST) {
        const float w = WINREG ? wreg[q * R + r] : s_win[idx];
        v[q][r] = make_float2(w * load_sample(seg_a + idx), w * load_sample(seg_b + idx));
      } else {
        v[q][r] = buf[pidx(idx)];
      }
    }
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < BPL; ++q) {
    const int j = lane + 32 * q;
    const int k = j % Ns;
    if (Ns > 1) {
#pragma unroll
      for (int r = 1; r < R; ++r) {
        const float2 w = PassTw<N, R, Ns>::IN_REGS ? tws.t[PassTw<N, R, Ns>::NK == 1 ? 0 : q][r - 1]
                                                   : s_tw[r * k * PassTw<N, R, Ns>::TSTRIDE];
        v[q][r] = cmul(v[q][r], w);
      }
    }
    butterfly<R>(v[q]);
    const int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) buf[pidx(j0 + r * Ns)] = v[q][r];
  }
  __syncwarp();
}

// Radix schedule of the N-point transform, three passes each: 256 = 4*8*8, 512 = 8*8*8, 1024 = 8*8*16 (the earlier
// 2*8*8*8 cost a fourth trip through shared memory for a radix-2 pass).
template <int N> struct Sched;
template <> struct Sched<256> { static constexpr int R0 = 4, R1 = 8, R2 = 8; };
template <> struct Sched<512> { static constexpr int R0 = 8, R1 = 8, R2 = 8; };
template <> struct Sched<1024> { static constexpr int R0 = 8, R1 = 8, R2 = 16; };

// Stage the raw waveform segment [q0, q0 + seg_len) of clip b (reflect padding at the clip ends, stft.py:236-237;
// zero beyond total_len, pad_truncate_sequence utils/utilities.py:66-70) into shared memory.  Interior, 16-byte
// aligned segments go through cp.async; edge segments through guarded loads.
template <int NFFT, typename TIn>
__device__ __forceinline__ void stage_segment(TIn* __restrict__ dst, const TIn* __restrict__ wave, long clip_stride,
                                              const long* __restrict__ clip_offset, long total_len, int L, int hop,
                                              int seg_len, int item, int chunks, bool aligned) {
  constexpr int FPB = FrontCfg<NFFT>::FPB;
  const int b = item / chunks, c = item - b * chunks;
  const long clip_base = clip_offset ? clip_offset[b] : static_cast<long>(b) * clip_stride;
  aligned = aligned && ((clip_base * static_cast<long>(sizeof(TIn))) & 15) == 0;
  const long q0 = static_cast<long>(c) * FPB * hop - NFFT / 2;
  const TIn* w = wave + clip_base;
  if (aligned && q0 >= 0 && q0 + seg_len <= L && clip_base + q0 + seg_len <= total_len) {
    const char* src = reinterpret_cast<const char*>(w + q0);
    char* d = reinterpret_cast<char*>(dst);
    const int nvec = seg_len * static_cast<int>(sizeof(TIn)) / 16;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) cp_async16(d + 16 * i, src + 16 * i);
  } else {
    // edge segments (first / last chunk of a clip: 2 of ~32 items): four independent loads in flight per thread
    for (int s0 = threadIdx.x; s0 < seg_len; s0 += 4 * blockDim.x) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int s = s0 + u * blockDim.x;
        long i = q0 + s;
        if (i < 0) i = -i;
        if (i >= L) i = 2L * (L - 1) - i;
        v[u] = (s < seg_len && i >= 0 && i < L && clip_base + i < total_len) ? load_sample_global(w + i) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int s = s0 + u * blockDim.x;
        if (s < seg_len) store_raw(dst + s, v[u]);
      }
    }
  }
  cp_async_commit();
}

// mode 0: out = log-mel (+ optional bn0 affine) [B, T, n_mels];  mode 1: out = power spectrogram [B, T, F]
// Clip b starts at wave + b * clip_stride (or wave + clip_offset[b] when a table is given) and is L samples long; samples at or beyond total_len (counted from
// `wave`) read as zero.  clip_stride < L gives overlapping windows of one long recording (predict.py:297-307)
// without materialising them.
template <int NFFT, typename TIn>
__global__ void __launch_bounds__(FrontCfg<NFFT>::WARPS * 32, FrontCfg<NFFT>::MIN_BLOCKS)
frontend_kernel(const TIn* __restrict__ wave, long clip_stride, const long* __restrict__ clip_offset, long total_len,
                int B, int L, int T, int hop,
                const float* __restrict__ window, const float2* __restrict__ twiddle,
                const int* __restrict__ mel_lo, const int* __restrict__ mel_len, const int* __restrict__ mel_off,
                const float* __restrict__ mel_val, int n_mels, float amin, float db_offset, int is_log,
                const float* __restrict__ bn_scale, const float* __restrict__ bn_shift, float* __restrict__ out,
                int mode, int aligned, int dbg, int* __restrict__ work_counter) {
#ifndef SED_PROFILE
  dbg = 0;  // the experiment switches fold away in the shipped library
#endif
  constexpr int WARPS = FrontCfg<NFFT>::WARPS;
  constexpr int FPB = FrontCfg<NFFT>::FPB;
  constexpr int F = NFFT / 2 + 1;
  constexpr int BUF = NFFT;  // swizzled in place (see pidx)
  const float db_floor = 10.0f * log10f(amin) - db_offset;  // once per thread: the value every clamped bin takes
  constexpr bool WINREG = FrontCfg<NFFT>::WIN_REGS;
  constexpr int R0 = Sched<NFFT>::R0, R1 = Sched<NFFT>::R1, R2 = Sched<NFFT>::R2;
  static_assert(R0 * R1 * R2 == NFFT, "radix schedule");
  constexpr int NS1 = R0, NS2 = R0 * R1;  // strides of the second and third pass
  extern __shared__ float4 smem_f4[];
  const int seg_len = (FPB - 1) * hop + NFFT;
  const int seg_bytes = ((seg_len * static_cast<int>(sizeof(TIn)) + 15) & ~15);
  uint8_t* sp = reinterpret_cast<uint8_t*>(smem_f4);
  TIn* s_stage0 = reinterpret_cast<TIn*>(sp);
  TIn* s_stage1 = reinterpret_cast<TIn*>(sp + seg_bytes);
  float2* s_buf = reinterpret_cast<float2*>(sp + 2 * seg_bytes);     // [WARPS][NFFT]
  float2* s_tw = s_buf + WARPS * BUF;                                 // [NFFT]
  float* s_win = reinterpret_cast<float*>(s_tw + NFFT);               // [NFFT]
  float* s_melv = s_win + NFFT;                                       // [MELV] banded mel weights
  int* s_meli = reinterpret_cast<int*>(s_melv + FrontCfg<NFFT>::MELV);  // [3][n_mels] lo, len, off

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = (T + FPB - 1) / FPB;
  const int items = B * chunks;
  int item = blockIdx.x;
  if (item < items)
    stage_segment<NFFT, TIn>(s_stage0, wave, clip_stride, clip_offset, total_len, L, hop, seg_len, item, chunks,
                             aligned != 0);

  for (int i = threadIdx.x; i < NFFT; i += blockDim.x) {
    s_win[i] = window[i];
    s_tw[i] = twiddle[i];
  }
  int mel_total = 0;
  if (mode == 0) {
    mel_total = mel_off[n_mels - 1] + mel_len[n_mels - 1];
    for (int i = threadIdx.x; i < n_mels; i += blockDim.x) {
      s_meli[i] = mel_lo[i];
      s_meli[n_mels + i] = mel_len[i];
      s_meli[2 * n_mels + i] = mel_off[i];
    }
    if (mel_total <= FrontCfg<NFFT>::MELV)
      for (int i = threadIdx.x; i < mel_total; i += blockDim.x) s_melv[i] = mel_val[i];
  }
  const bool mel_in_smem = mel_total <= FrontCfg<NFFT>::MELV;

  // per-lane constants of the transform, loaded once per persistent warp
  PassTw<NFFT, R0, 1> tw0;  // first pass: no twiddles
  PassTw<NFFT, R1, NS1> tw1;
  PassTw<NFFT, R2, NS2> tw2;
  tw1.init(twiddle, lane);
  tw2.init(twiddle, lane);
  float wreg[NFFT / 32];
  if (WINREG) {
    constexpr int NB0 = NFFT / R0, BPL0 = NB0 / 32;
#pragma unroll
    for (int q = 0; q < BPL0; ++q)
#pragma unroll
      for (int r = 0; r < R0; ++r) wreg[q * R0 + r] = __ldg(window + lane + 32 * q + r * NB0);
  }

  float2* buf = s_buf + warp * BUF;
  // Work items are claimed dynamically (block b starts with item b, every further item comes from an atomic counter):
  // a block that starts late -- its SM was still busy with another stream's kernel -- simply claims fewer items, so
  // the launch ends when the work is done, not when the slowest static share is.
  __shared__ int s_next[2];
  int sel = 0;
  for (int nxt = 0; item < items; item = nxt, sel ^= 1) {
    cp_async_wait_all();
    if (threadIdx.x == 0) s_next[sel] = static_cast<int>(gridDim.x) + atomicAdd(work_counter, 1);
    __syncthreads();  // this item's segment and the claimed item are visible; every warp has finished the previous item
    nxt = s_next[sel];
    if (nxt < items)
      stage_segment<NFFT, TIn>(sel ? s_stage0 : s_stage1, wave, clip_stride, clip_offset, total_len, L, hop, seg_len,
                               nxt, chunks, aligned != 0);
    const TIn* s_seg = sel ? s_stage1 : s_stage0;
    const int b = item / chunks;
    const int f_base = (item - b * chunks) * FPB;

    for (int pair = warp; pair < FPB / 2; pair += WARPS) {
      const int fa = f_base + 2 * pair;
      if (fa >= T) break;
      const TIn* seg_a = s_seg + (2 * pair) * hop;
      const TIn* seg_b = seg_a + hop;

      // ---- 2 real frames -> 1 complex FFT of size NFFT (Stockham autosort, natural-order output) ----
      fft_pass<NFFT, R0, 1, true, WINREG, TIn>(buf, tw0, s_tw, seg_a, seg_b, s_win, wreg, lane);
      if (!(dbg & 1)) {
        fft_pass<NFFT, R1, NS1, false, WINREG, TIn>(buf, tw1, s_tw, seg_a, seg_b, s_win, wreg, lane);
        fft_pass<NFFT, R2, NS2, false, WINREG, TIn>(buf, tw2, s_tw, seg_a, seg_b, s_win, wreg, lane);
      }
      // ---- split the two real spectra and take the power (stft.py:663), in place: P[0..F) = |A|^2,
      //      P[F..2F) = |B|^2 overwrite the spectrum after every lane has read its bins ----
      constexpr int KPL = (F + 31) / 32;
      float pa[KPL], pb[KPL];
#pragma unroll
      for (int i = 0; i < KPL; ++i) {
        const int k = lane + 32 * i;
        if (k < F) {
          const float2 zk = buf[pidx(k & (NFFT - 1))];
          const float2 zn = buf[pidx((NFFT - k) & (NFFT - 1))];
          const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);
          const float br = 0.5f * (zk.y + zn.y), bi = 0.5f * (zn.x - zk.x);
          pa[i] = ar * ar + ai * ai;
          pb[i] = br * br + bi * bi;
        }
      }
      __syncwarp();
      float2* P2 = buf;  // P2[k] = (|A_k|^2, |B_k|^2): both frames of the pair side by side
#pragma unroll
      for (int i = 0; i < KPL; ++i) {
        const int k = lane + 32 * i;
        if (k < F) P2[k] = make_float2(pa[i], pb[i]);
      }
      __syncwarp();

      const bool has_b = fa + 1 < T;
      if (dbg & 2) {
        if (lane == 0) out[(static_cast<size_t>(b) * T + fa) * n_mels] = P2[3].x + P2[3].y;
      } else if (mode == 1) {
        float* o = out + (static_cast<size_t>(b) * T + fa) * F;
        for (int k = lane; k < F; k += 32) {
          const float2 p = P2[k];
          o[k] = p.x;
          if (has_b) o[F + k] = p.y;
        }
      } else {
        float* o = out + (static_cast<size_t>(b) * T + fa) * n_mels;
        // bins are visited as (lane, n_mels-1-lane, lane+32, ...): narrow low bands pair with wide high bands;
        // both frames of the pair share every weight load
        for (int mi = lane; mi < n_mels; mi += 32) {
          const int pr = mi >> 5;
          const int m = (pr & 1) ? (n_mels - 1 - (mi - 32 * pr) - 32 * (pr >> 1)) : (lane + 32 * (pr >> 1));
          if (m < 0 || m >= n_mels) continue;
          const int lo = s_meli[m], len = s_meli[n_mels + m], off = s_meli[2 * n_mels + m];
          const float2* Pm = P2 + lo;
          float a0 = 0.0f, a1 = 0.0f, b0 = 0.0f, b1 = 0.0f;  // stft.py:709 restricted to the band of non-zero weights
          if (mel_in_smem) {
            const float* mv = s_melv + off;
            int i = 0;
            for (; i + 3 < len; i += 4) {  // four bins per trip (same summation order as two trips of the loop below)
              const float2 p = Pm[i], q = Pm[i + 1], p2 = Pm[i + 2], q2 = Pm[i + 3];
              const float w0 = mv[i], w1 = mv[i + 1], w2 = mv[i + 2], w3 = mv[i + 3];
              a0 = fmaf(p.x, w0, a0);
              b0 = fmaf(p.y, w0, b0);
              a1 = fmaf(q.x, w1, a1);
              b1 = fmaf(q.y, w1, b1);
              a0 = fmaf(p2.x, w2, a0);
              b0 = fmaf(p2.y, w2, b0);
              a1 = fmaf(q2.x, w3, a1);
              b1 = fmaf(q2.y, w3, b1);
            }
            for (; i + 1 < len; i += 2) {
              const float2 p = Pm[i], q = Pm[i + 1];
              const float w0 = mv[i], w1 = mv[i + 1];
              a0 = fmaf(p.x, w0, a0);
              b0 = fmaf(p.y, w0, b0);
              a1 = fmaf(q.x, w1, a1);
              b1 = fmaf(q.y, w1, b1);
            }
            if (i < len) {
              const float2 p = Pm[i];
              const float w0 = mv[i];
              a0 = fmaf(p.x, w0, a0);
              b0 = fmaf(p.y, w0, b0);
            }
          } else {
            const float* mv = mel_val + off;
            for (int i = 0; i < len; ++i) {
              const float2 p = Pm[i];
              const float w0 = __ldg(mv + i);
              a0 = fmaf(p.x, w0, a0);
              b0 = fmaf(p.y, w0, b0);
            }
          }
          float ya = a0 + a1, yb = b0 + b1;
          if (is_log) {  // stft.py:726-727
            ya = power_to_db(ya, amin, db_offset, db_floor);
            yb = power_to_db(yb, amin, db_offset, db_floor);
          }
          if (bn_scale != nullptr) {  // models.py:642-644
            const float sc = bn_scale[m], sh = bn_shift[m];
            ya = fmaf(ya, sc, sh);
            yb = fmaf(yb, sc, sh);
          }
          o[m] = ya;
          if (has_b) o[n_mels + m] = yb;
        }
      }
      __syncwarp();
    }
  }
}

// Standalone LogmelFilterBank.forward (stft.py:698-718): rows [R, F] -> [R, n_mels]
__global__ void logmel_rows_kernel(const float* __restrict__ spec, long rows, int F, const int* __restrict__ mel_lo,
                                   const int* __restrict__ mel_len, const int* __restrict__ mel_off,
                                   const float* __restrict__ mel_val, int n_mels, float amin, float db_offset,
                                   int is_log, float* __restrict__ out) {
  const long gw = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= rows) return;
  const float* Pf = spec + gw * F;
  float* o = out + gw * n_mels;
  for (int m = lane; m < n_mels; m += 32) {
    const int lo = mel_lo[m], len = mel_len[m];
    const float* mv = mel_val + mel_off[m];
    float acc = 0.0f;
    for (int i = 0; i < len; ++i) acc = fmaf(__ldg(Pf + lo + i), __ldg(mv + i), acc);
    o[m] = is_log ? 10.0f * log10f(fmaxf(acc, amin)) - db_offset : acc;
  }
}

// claim counters of the launches in flight (one per launch, round robin; zeroed on the launch's stream)
__device__ int g_frontend_work_counters[64];
static std::atomic<unsigned> g_frontend_launch_seq{0};

template <int NFFT, typename TIn>
static int launch_frontend_t(const FrontendArgs& a, cudaStream_t stream) {
  constexpr int WARPS = FrontCfg<NFFT>::WARPS;
  constexpr int FPB = FrontCfg<NFFT>::FPB;
  const int seg_len = (FPB - 1) * a.hop + NFFT;
  const int seg_bytes = (seg_len * static_cast<int>(sizeof(TIn)) + 15) & ~15;
  const size_t smem = 2 * static_cast<size_t>(seg_bytes) + sizeof(float2) * (WARPS * NFFT + NFFT) +
                      sizeof(float) * (NFFT + FrontCfg<NFFT>::MELV) + sizeof(int) * 3 * (a.n_mels > 0 ? a.n_mels : 1);
  if (smem > 227 * 1024) return SED_ERR_UNSUPPORTED;
  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm_count <= 0)
      sm_count = 148;
  }
  cudaError_t e =
      cudaFuncSetAttribute(frontend_kernel<NFFT, TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return SED_ERR_CUDA;
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frontend_kernel<NFFT, TIn>, WARPS * 32, smem) !=
          cudaSuccess || per_sm < 1)
    per_sm = 1;
  const long items = static_cast<long>(a.B) * ((a.T + FPB - 1) / FPB);
  long blocks = static_cast<long>(sm_count) * per_sm;
  if (blocks > items) blocks = items;
  // cp.async staging needs 16-byte aligned interior segments
  const size_t es = sizeof(TIn);
  const int aligned = (reinterpret_cast<uintptr_t>(a.wave) % 16 == 0) &&
                      (a.clip_offset != nullptr || (a.clip_stride * es) % 16 == 0) &&
                      ((static_cast<size_t>(FPB) * a.hop * es) % 16 == 0) && ((NFFT / 2 * es) % 16 == 0) &&
                      ((seg_len * es) % 16 == 0);
  int dbg = 0;
#ifdef SED_PROFILE
  const char* e_dbg = getenv("SED_FE_DBG");  // developer experiments: 1 = first FFT pass only, 2 = no mel projection
  dbg = e_dbg ? atoi(e_dbg) : 0;
#endif
  int* counters = nullptr;
  if (cudaGetSymbolAddress(reinterpret_cast<void**>(&counters), g_frontend_work_counters) != cudaSuccess)
    return SED_ERR_CUDA;
  int* counter = counters + (g_frontend_launch_seq.fetch_add(1) & 63u);
  if (cudaMemsetAsync(counter, 0, sizeof(int), stream) != cudaSuccess) return SED_ERR_CUDA;
  frontend_kernel<NFFT, TIn><<<static_cast<unsigned>(blocks), WARPS * 32, smem, stream>>>(
      reinterpret_cast<const TIn*>(a.wave), a.clip_stride, a.clip_offset, a.total_len, a.B, a.L, a.T, a.hop, a.window,
      reinterpret_cast<const float2*>(a.twiddle), a.mel_lo, a.mel_len, a.mel_off, a.mel_val, a.n_mels, a.amin,
      a.db_offset, a.is_log, a.bn_scale, a.bn_shift, a.out, a.mode, aligned, dbg, counter);
  return cudaGetLastError() == cudaSuccess ? SED_OK : SED_ERR_CUDA;
}

template <int NFFT>
static int launch_frontend(const FrontendArgs& a, cudaStream_t stream) {
  if (a.wave_dtype == 1) return launch_frontend_t<NFFT, short>(a, stream);
  return launch_frontend_t<NFFT, float>(a, stream);
}

int frontend_launch(const FrontendArgs& a, cudaStream_t stream) {
  if (a.B <= 0 || a.L <= a.n_fft / 2 || a.hop <= 0 || a.clip_stride <= 0 || a.total_len <= 0) return SED_ERR_BAD_SHAPE;
  if (a.wave_dtype != 0 && a.wave_dtype != 1) return SED_ERR_UNSUPPORTED;
  switch (a.n_fft) {
    case 256: return launch_frontend<256>(a, stream);
    case 512: return launch_frontend<512>(a, stream);
    case 1024: return launch_frontend<1024>(a, stream);
    default: return SED_ERR_UNSUPPORTED;
  }
}

int logmel_rows_launch(const float* spec, long rows, int F, const int* mel_lo, const int* mel_len, const int* mel_off,
                       const float* mel_val, int n_mels, float amin, float db_offset, int is_log, float* out,
                       cudaStream_t stream) {
  if (rows <= 0) return SED_ERR_BAD_SHAPE;
  const int threads = 256;
  const long blocks = (rows * 32 + threads - 1) / threads;
  logmel_rows_kernel<<<(unsigned)blocks, threads, 0, stream>>>(spec, rows, F, mel_lo, mel_len, mel_off, mel_val, n_mels,
                                                               amin, db_offset, is_log, out);
  return cudaGetLastError() == cudaSuccess ? SED_OK : SED_ERR_CUDA;
}

}  // namespace sed
