// Weight preparation for the kernels: eval-mode BatchNorm folding and the repacking of reference-layout parameters
// (OIHW conv weights, nn.GRU weight_hh, nn.Linear weights) into the layouts include/sed_b200.h documents.  One-time
// work per checkpoint, device pointers in and out, so that a caller of the C ABI needs no tensor library.
#include <cmath>
#include <cstdint>

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/sed_b200.h"
#include "sed_kernels.h"

namespace sed {
namespace {

template <typename T>
__device__ __forceinline__ T to16(float v);
template <>
__device__ __forceinline__ __half to16<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 to16<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// y = (x - mean) / sqrt(var + eps) * weight + bias  ->  y = x * scale + shift, folded in float64 with separately
// rounded operations (no fused multiply-add), the order the host reference of this fold uses.
__global__ void fold_bn_kernel(const float* w, const float* b, const float* mean, const float* var, int n, double eps,
                               float* scale, float* shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double s = __ddiv_rn(static_cast<double>(w[i]), __dsqrt_rn(__dadd_rn(static_cast<double>(var[i]), eps)));
  const float sf = static_cast<float>(s);
  scale[i] = sf;
  // the shift is folded with the float64 scale (before its rounding to float32)
  shift[i] = static_cast<float>(__dsub_rn(static_cast<double>(b[i]), __dmul_rn(static_cast<double>(mean[i]), s)));
}

template <typename T>
__global__ void pack_conv3x3_kernel(const float* w, int cout, int cin, T* out) {
  // out[o][tap][c] = w[o][c][tap]
  const long total = static_cast<long>(cout) * 9 * cin;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cin);
    const int tap = static_cast<int>((i / cin) % 9);
    const long o = i / (9L * cin);
    out[i] = to16<T>(w[(o * cin + c) * 9 + tap]);
  }
}

__global__ void pack_conv_first_kernel(const float* w, const float* scale, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // 64 x 9
  if (i >= 64 * 9) return;
  out[i] = static_cast<float>(__dmul_rn(static_cast<double>(w[i]), static_cast<double>(scale[i / 9])));
}

template <typename T>
__global__ void pack_gru_whh_kernel(const float* fwd, const float* bwd, T* out) {
  // out row (dir*768 + 96*q + 32*g + jj) = W_hh[dir][g*256 + 32*q + jj]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // 1536 x 256
  if (i >= 1536 * 256) return;
  const int col = i & 255, row = i >> 8;
  const int dir = row / 768, r = row % 768;
  const int q = r / 96, g = (r % 96) / 32, jj = r % 32;
  const float* src = dir ? bwd : fwd;
  out[i] = to16<T>(src[(g * 256 + 32 * q + jj) * 256 + col]);
}

template <typename T>
__global__ void cast16_kernel(const float* src, long n, T* dst) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * blockDim.x)
    dst[i] = to16<T>(src[i]);
}

int finish(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s launch: %s", what, cudaGetErrorString(e));
    return SED_ERR_CUDA;
  }
  return SED_OK;
}

int grid_for(long n) {
  long g = (n + 255) / 256;
  return static_cast<int>(g < 1 ? 1 : (g > 4096 ? 4096 : g));
}

}  // namespace

int fold_bn_launch(const float* w, const float* b, const float* mean, const float* var, int n, double eps, float* scale,
                   float* shift, cudaStream_t stream) {
  if (n <= 0 || !(eps >= 0.0)) {
    set_error("fold_bn: channels=%d eps=%g", n, eps);
    return SED_ERR_BAD_SHAPE;
  }
  fold_bn_kernel<<<(n + 127) / 128, 128, 0, stream>>>(w, b, mean, var, n, eps, scale, shift);
  return finish("fold_bn");
}

int pack_conv3x3_launch(const float* w, int cout, int cin, void* out, int dtype, cudaStream_t stream) {
  if (cout <= 0 || cin <= 0) {
    set_error("pack_conv3x3: cout=%d cin=%d", cout, cin);
    return SED_ERR_BAD_SHAPE;
  }
  const long n = static_cast<long>(cout) * cin * 9;
  if (dtype == SED_DTYPE_F16) pack_conv3x3_kernel<<<grid_for(n), 256, 0, stream>>>(w, cout, cin, static_cast<__half*>(out));
  else if (dtype == SED_DTYPE_BF16)
    pack_conv3x3_kernel<<<grid_for(n), 256, 0, stream>>>(w, cout, cin, static_cast<__nv_bfloat16*>(out));
  else {
    set_error("pack_conv3x3: dtype must be 0 (fp16) or 1 (bf16)");
    return SED_ERR_UNSUPPORTED;
  }
  return finish("pack_conv3x3");
}

int pack_conv_first_launch(const float* w, const float* scale, float* out, cudaStream_t stream) {
  pack_conv_first_kernel<<<(64 * 9 + 127) / 128, 128, 0, stream>>>(w, scale, out);
  return finish("pack_conv_first");
}

int pack_gru_whh_launch(const float* fwd, const float* bwd, void* out, int dtype, cudaStream_t stream) {
  const int n = 1536 * 256;
  if (dtype == SED_DTYPE_F16) pack_gru_whh_kernel<<<n / 256, 256, 0, stream>>>(fwd, bwd, static_cast<__half*>(out));
  else if (dtype == SED_DTYPE_BF16)
    pack_gru_whh_kernel<<<n / 256, 256, 0, stream>>>(fwd, bwd, static_cast<__nv_bfloat16*>(out));
  else {
    set_error("pack_gru_whh: dtype must be 0 (fp16) or 1 (bf16)");
    return SED_ERR_UNSUPPORTED;
  }
  return finish("pack_gru_whh");
}

int cast16_launch(const float* src, long n, void* dst, int dtype, cudaStream_t stream) {
  if (n <= 0) {
    set_error("cast_16: n=%ld", n);
    return SED_ERR_BAD_SHAPE;
  }
  if (dtype == SED_DTYPE_F16) cast16_kernel<<<grid_for(n), 256, 0, stream>>>(src, n, static_cast<__half*>(dst));
  else if (dtype == SED_DTYPE_BF16) cast16_kernel<<<grid_for(n), 256, 0, stream>>>(src, n, static_cast<__nv_bfloat16*>(dst));
  else {
    set_error("cast_16: dtype must be 0 (fp16) or 1 (bf16)");
    return SED_ERR_UNSUPPORTED;
  }
  return finish("cast_16");
}

// ---------------------------------------------------------------- range guard of the 16-bit path
// Every float32 -> 16-bit conversion of the hot path saturates (cvt.rn.satfinite): a value beyond the format's range
// is stored as exactly +-MAX (fp16: 65504).  Counting stored values with |x| >= MAX (or non-finite) therefore counts
// the conversions that clipped -- after the fact, with no work added to the conv epilogues.
__global__ void count_saturated16_kernel(const uint16_t* __restrict__ x, long n, uint16_t max_bits,
                                         unsigned long long* __restrict__ count) {
  unsigned int local = 0;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * blockDim.x)
    local += ((x[i] & 0x7FFFu) >= max_bits) ? 1u : 0u;
  local = __reduce_add_sync(0xffffffffu, local);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, static_cast<unsigned long long>(local));
}

int count_saturated16_launch(const void* x, long n, int dtype, unsigned long long* count, cudaStream_t stream) {
  if (n <= 0) {
    set_error("count_saturated16: n=%ld", n);
    return SED_ERR_BAD_SHAPE;
  }
  if (dtype != SED_DTYPE_F16 && dtype != SED_DTYPE_BF16) {
    set_error("count_saturated16: dtype must be 0 (fp16) or 1 (bf16)");
    return SED_ERR_UNSUPPORTED;
  }
  const uint16_t max_bits = dtype == SED_DTYPE_F16 ? 0x7BFFu : 0x7F7Fu;
  long blocks = (n + 256L * 8 - 1) / (256L * 8);
  if (blocks > 148L * 16) blocks = 148L * 16;
  count_saturated16_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<const uint16_t*>(x), n,
                                                                                max_bits, count);
  return finish("count_saturated16");
}

// ---------------------------------------------------------------- host-side front-end constants
int frontend_twiddle_host(int n_fft, float* out) {
  if (n_fft != 256 && n_fft != 512 && n_fft != 1024) {
    set_error("frontend_twiddle: n_fft=%d (256, 512 or 1024)", n_fft);
    return SED_ERR_UNSUPPORTED;
  }
  const double two_pi = 6.283185307179586476925286766559;
  for (int k = 0; k < n_fft; ++k) {
    const double ang = -two_pi * static_cast<double>(k) / static_cast<double>(n_fft);
    out[2 * k] = static_cast<float>(std::cos(ang));
    out[2 * k + 1] = static_cast<float>(std::sin(ang));
  }
  return SED_OK;
}

int band_mel_host(const float* melW, int F, int M, int* lo, int* len, int* off, float* val, int cap, int* n_val) {
  if (F <= 0 || M <= 0 || cap < 1) {
    set_error("band_mel: F=%d n_mels=%d capacity=%d", F, M, cap);
    return SED_ERR_BAD_SHAPE;
  }
  int pos = 0;
  for (int m = 0; m < M; ++m) {
    int first = -1, last = -1;
    for (int f = 0; f < F; ++f)
      if (melW[static_cast<long>(f) * M + m] != 0.0f) {
        if (first < 0) first = f;
        last = f;
      }
    lo[m] = first < 0 ? 0 : first;
    len[m] = first < 0 ? 0 : last - first + 1;
    off[m] = pos;
    if (pos + len[m] > cap) {
      set_error("band_mel: %d band values do not fit the capacity %d", pos + len[m], cap);
      return SED_ERR_BAD_SHAPE;
    }
    for (int k = 0; k < len[m]; ++k) val[pos + k] = melW[static_cast<long>(lo[m] + k) * M + m];
    pos += len[m];
  }
  if (pos == 0) val[0] = 0.0f;
  *n_val = pos;
  return SED_OK;
}

}  // namespace sed
