// Fused audio front-end for sm_100a: reflect-pad -> frame -> window -> real DFT -> power -> mel
// projection -> clamped 10*log10 -> bn0 affine, one kernel, float32 end to end.
//
// Reference: STFT.forward pytorch/stft.py:223-247 (two Conv1d with the windowed DFT matrix),
// Spectrogram.forward :651-670, LogmelFilterBank.forward/power_to_db :698-734, bn0 models.py:642-644.
//
// The reference evaluates the DFT as a dense [n_fft x (n_fft/2+1)] contraction (0.53 GFLOP per 10 s
// clip at 16 kHz).  Here two consecutive real frames are packed as one complex sequence and pushed
// through an in-shared-memory Stockham FFT (radix-4 passes + one radix-2 pass), 25x fewer flops, in
// float32 with float64-derived twiddles -- the accuracy class of the reference's float32 conv1d, which
// the 1e-4 log-mel tolerance needs (single-pass 16-bit tensor-core DFTs do not reach it).  The window
// is read from the loaded conv_real kernel (row 0), so a checkpoint's window is honoured; the host
// wrapper verifies that the loaded kernels are a windowed DFT before selecting this path.
// The mel projection uses the loaded melW in banded form (first/last non-zero per mel bin).
#include <cstdlib>

#include "sed_common.cuh"
#include "sed_kernels.h"

namespace sed {

template <int NFFT>
struct FrontCfg {
  static constexpr int WARPS = 4;
  static constexpr int FPB = 16;  // frames per block (even: frames are transformed in pairs); small blocks so
                                  // that several are resident per SM and their load phases overlap
  static constexpr int MELV = 1024;  // banded mel weights cached in shared memory (falls back to global beyond)
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// Shared-memory FFT buffers are padded by one complex element per 8 so the strided writes of the
// Stockham passes spread over the banks.
__device__ __forceinline__ int pidx(int i) { return i + (i >> 3); }

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

// One Stockham pass of radix R over N complex points, IN PLACE in (padded) shared memory, executed by one warp:
// every lane first pulls the inputs of all its butterflies into registers, the warp synchronises, then the
// autosorted outputs are written back.  first == true: inputs come from the windowed frame pair instead
// (re = frame a, im = frame b).
template <int N, int R>
__device__ __forceinline__ void fft_pass(float2* __restrict__ buf, bool first, int Ns, const float2* __restrict__ tw,
                                         const float* __restrict__ seg_a, const float* __restrict__ seg_b,
                                         const float* __restrict__ win, int lane) {
  constexpr int NB = N / R;
  constexpr int BPL = NB / 32;  // butterflies per lane
  float2 v[BPL][R];
#pragma unroll
  for (int q = 0; q < BPL; ++q) {
    const int j = lane + 32 * q;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int idx = j + r * NB;
      if (first) {
        const float w = win[idx];
        v[q][r] = make_float2(w * seg_a[idx], w * seg_b[idx]);
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
      const int tstride = N / (Ns * R);
#pragma unroll
      for (int r = 1; r < R; ++r) v[q][r] = cmul(v[q][r], tw[r * k * tstride]);
    }
    butterfly<R>(v[q]);
    const int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) buf[pidx(j0 + r * Ns)] = v[q][r];
  }
  __syncwarp();
}

// mode 0: out = log-mel (+ optional bn0 affine) [B, T, n_mels];  mode 1: out = power spectrogram [B, T, F]
// int16 PCM input follows the reference's HDF5 path: x = q / 32767 (utils/utilities.py:78-79); the float32
// division is correctly rounded and equals numpy's float64-then-float32 result for every int16 q.
__device__ __forceinline__ float load_sample(const float* p) { return __ldg(p); }
__device__ __forceinline__ float load_sample(const short* p) { return static_cast<float>(__ldg(p)) / 32767.0f; }

// Clip b starts at wave + b * clip_stride and is L samples long; samples at or beyond total_len (counted from
// `wave`) read as zero (pad_truncate_sequence, utils/utilities.py:66-70).  clip_stride < L gives overlapping
// windows of one long recording (predict.py:297-307) without materialising them.
template <int NFFT, typename TIn>
__global__ void __launch_bounds__(FrontCfg<NFFT>::WARPS * 32)
frontend_kernel(const TIn* __restrict__ wave, long clip_stride, long total_len, int L, int T, int hop,
                const float* __restrict__ window,
                const float2* __restrict__ twiddle, const int* __restrict__ mel_lo, const int* __restrict__ mel_len,
                const int* __restrict__ mel_off, const float* __restrict__ mel_val, int n_mels, float amin,
                float db_offset, int is_log, const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                float* __restrict__ out, int mode, int dbg) {
  constexpr int WARPS = FrontCfg<NFFT>::WARPS;
  constexpr int FPB = FrontCfg<NFFT>::FPB;
  constexpr int F = NFFT / 2 + 1;
  extern __shared__ float smem_f[];
  const int seg_len = (FPB - 1) * hop + NFFT;
  float* s_seg = smem_f;                                            // [seg_len]
  float* s_win = s_seg + ((seg_len + 3) & ~3);                      // [NFFT]
  float2* s_tw = reinterpret_cast<float2*>(s_win + NFFT);           // [NFFT]
  float2* s_buf = s_tw + NFFT;                                      // [WARPS][NFFT + NFFT/8]
  constexpr int BUF = NFFT + NFFT / 8;                              // padded (see pidx)
  float* s_melv = reinterpret_cast<float*>(s_buf + WARPS * BUF);    // [MELV] banded mel weights
  int* s_meli = reinterpret_cast<int*>(s_melv + FrontCfg<NFFT>::MELV);  // [3][n_mels] lo, len, off

  const int b = blockIdx.y;
  const int f_base = blockIdx.x * FPB;
  const long clip_base = static_cast<long>(b) * clip_stride;
  const TIn* w = wave + clip_base;

  // stage the waveform segment with reflect padding (stft.py:236-237)
  const long q0 = static_cast<long>(f_base) * hop - NFFT / 2;
  for (int s = threadIdx.x; s < seg_len; s += blockDim.x) {
    long i = q0 + s;
    if (i < 0) i = -i;
    if (i >= L) i = 2L * (L - 1) - i;
    s_seg[s] = (i >= 0 && i < L && clip_base + i < total_len) ? load_sample(w + i) : 0.0f;
  }
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
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* buf = s_buf + warp * BUF;

  for (int pair = warp; pair < FPB / 2; pair += WARPS) {
    const int fa = f_base + 2 * pair;
    if (fa >= T) break;
    const float* seg_a = s_seg + (2 * pair) * hop;
    const float* seg_b = seg_a + hop;

    // ---- 2 real frames -> 1 complex FFT of size NFFT (Stockham autosort, natural-order output) ----
    // radix schedule: 256 = 4*8*8, 512 = 8*8*8, 1024 = 2*8*8*8
    int Ns = 1;
    bool first = true;
    if (NFFT == 1024) {
      fft_pass<NFFT, 2>(buf, first, Ns, s_tw, seg_a, seg_b, s_win, lane);
      Ns = 2;
      first = false;
    } else if (NFFT == 256) {
      fft_pass<NFFT, 4>(buf, first, Ns, s_tw, seg_a, seg_b, s_win, lane);
      Ns = 4;
      first = false;
    }
    while (Ns < NFFT) {
      if ((dbg & 1) && Ns > 1) break;
      fft_pass<NFFT, 8>(buf, first, Ns, s_tw, seg_a, seg_b, s_win, lane);
      Ns *= 8;
      first = false;
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
    float* P = reinterpret_cast<float*>(buf);
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      const int k = lane + 32 * i;
      if (k < F) {
        P[k] = pa[i];
        P[F + k] = pb[i];
      }
    }
    __syncwarp();

#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const int f = fa + which;
      if (f >= T) break;
      const float* Pf = P + which * F;
      if (dbg & 2) {
        if (lane == 0) out[(static_cast<size_t>(b) * T + f) * n_mels] = Pf[3];
      } else if (mode == 1) {
        float* o = out + (static_cast<size_t>(b) * T + f) * F;
        for (int k = lane; k < F; k += 32) o[k] = Pf[k];
      } else {
        float* o = out + (static_cast<size_t>(b) * T + f) * n_mels;
        // bins are visited as (lane, n_mels-1-lane, lane+32, ...): narrow low bands pair with wide high bands
        for (int mi = lane; mi < n_mels; mi += 32) {
          const int pr = mi >> 5;
          const int m = (pr & 1) ? (n_mels - 1 - (mi - 32 * pr) - 32 * (pr >> 1)) : (lane + 32 * (pr >> 1));
          if (m < 0 || m >= n_mels) continue;
          const int lo = s_meli[m], len = s_meli[n_mels + m], off = s_meli[2 * n_mels + m];
          float acc = 0.0f, acc2 = 0.0f;  // stft.py:709 restricted to the band of non-zero weights
          if (mel_in_smem) {
            const float* mv = s_melv + off;
            int i = 0;
            for (; i + 1 < len; i += 2) {
              acc = fmaf(Pf[lo + i], mv[i], acc);
              acc2 = fmaf(Pf[lo + i + 1], mv[i + 1], acc2);
            }
            if (i < len) acc = fmaf(Pf[lo + i], mv[i], acc);
          } else {
            const float* mv = mel_val + off;
            for (int i = 0; i < len; ++i) acc = fmaf(Pf[lo + i], __ldg(mv + i), acc);
          }
          acc += acc2;
          float y = acc;
          if (is_log) y = 10.0f * log10f(fmaxf(acc, amin)) - db_offset;  // stft.py:726-727
          if (bn_scale != nullptr) y = fmaf(y, bn_scale[m], bn_shift[m]);  // models.py:642-644
          o[m] = y;
        }
      }
    }
    __syncwarp();
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

template <int NFFT, typename TIn>
static int launch_frontend_t(const FrontendArgs& a, cudaStream_t stream) {
  constexpr int WARPS = FrontCfg<NFFT>::WARPS;
  constexpr int FPB = FrontCfg<NFFT>::FPB;
  const int seg_len = (FPB - 1) * a.hop + NFFT;
  const size_t smem = sizeof(float) * (((seg_len + 3) & ~3) + NFFT + FrontCfg<NFFT>::MELV) +
                      sizeof(float2) * (NFFT + WARPS * (NFFT + NFFT / 8)) + sizeof(int) * 3 * (a.n_mels > 0 ? a.n_mels : 1);
  if (smem > 227 * 1024) return SED_ERR_UNSUPPORTED;
  cudaError_t e =
      cudaFuncSetAttribute(frontend_kernel<NFFT, TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return SED_ERR_CUDA;
  dim3 grid((a.T + FPB - 1) / FPB, a.B);
  const char* e_dbg = getenv("SED_FE_DBG");
  const int dbg = e_dbg ? atoi(e_dbg) : 0;
  frontend_kernel<NFFT, TIn><<<grid, WARPS * 32, smem, stream>>>(
      reinterpret_cast<const TIn*>(a.wave), a.clip_stride, a.total_len, a.L, a.T, a.hop, a.window, reinterpret_cast<const float2*>(a.twiddle), a.mel_lo, a.mel_len, a.mel_off,
      a.mel_val, a.n_mels, a.amin, a.db_offset, a.is_log, a.bn_scale, a.bn_shift, a.out, a.mode, dbg);
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
