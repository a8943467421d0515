// Streaming post-processing on the device: overlap-add of per-window framewise probabilities and the
// reference's block-wise averaging (utils/utilities.py:405-446 merge / avg_merge, driven by the window loop of
// pytorch/predict.py:297-349).  Pure data movement + one division per element; bit-exact with the numpy code:
// windows are accumulated in ascending order in float32 and divided by float32(num_overlaps).
#include "sed_common.cuh"
#include "sed_kernels.h"

namespace sed {

__global__ void window_merge_kernel(const float* __restrict__ frames, int n_windows, int fpw, int classes, int oi,
                                    int sample_duration, int total_frames, float* __restrict__ merged) {
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long>(total_frames) * classes) return;
  const int f = static_cast<int>(idx / classes);
  const int c = static_cast<int>(idx - static_cast<long>(f) * classes);
  // merge(): window k (0-based) occupies frames [k*oi, k*oi + fpw); later windows are added to the running sum
  int k_lo = (f - fpw + oi) / oi;  // smallest k with f - k*oi < fpw  (ceil((f - fpw + 1) / oi))
  if (f - fpw + 1 <= 0) k_lo = 0;
  int k_hi = f / oi;
  if (k_hi > n_windows - 1) k_hi = n_windows - 1;
  float acc = 0.0f;
  bool first = true;
  for (int k = k_lo; k <= k_hi; ++k) {
    const float v = frames[(static_cast<long>(k) * fpw + (f - k * oi)) * classes + c];
    acc = first ? v : acc + v;
    first = false;
  }
  // avg_merge(): blocks start at i = oi, 2*oi, ... < total - oi; the divisor depends on the block start i
  const int interval = sample_duration * 100 - oi;
  if (f >= oi) {
    const int i = (f / oi) * oi;
    if (i < total_frames - oi) {
      int div;
      if (i < interval) div = i / oi + 1;
      else if (i >= total_frames - interval) div = (total_frames - i) / oi + 1;
      else div = sample_duration;
      acc = acc / static_cast<float>(div);
    }
  }
  merged[idx] = acc;
}

int window_merge_launch(const float* frames, int n_windows, int frames_per_window, int classes, int overlap_interval,
                        int sample_duration, float* merged, cudaStream_t stream) {
  if (n_windows <= 0 || frames_per_window <= 0 || classes <= 0 || overlap_interval <= 0 || sample_duration <= 0 ||
      overlap_interval > frames_per_window) {
    set_error("window_merge: bad shape n_windows=%d frames=%d classes=%d overlap_interval=%d duration=%d", n_windows,
              frames_per_window, classes, overlap_interval, sample_duration);
    return SED_ERR_BAD_SHAPE;
  }
  const int total = (n_windows - 1) * overlap_interval + frames_per_window;
  const long n = static_cast<long>(total) * classes;
  window_merge_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      frames, n_windows, frames_per_window, classes, overlap_interval, sample_duration, total, merged);
  return cudaGetLastError() == cudaSuccess ? SED_OK : SED_ERR_CUDA;
}

}  // namespace sed
