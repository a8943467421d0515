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
  frames += static_cast<size_t>(blockIdx.y) * n_windows * fpw * classes;   // recording blockIdx.y
  merged += static_cast<size_t>(blockIdx.y) * total_frames * classes;
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
                        int sample_duration, int n_recordings, float* merged, cudaStream_t stream) {
  if (n_recordings <= 0 || n_recordings > 65535 || n_windows <= 0 || frames_per_window <= 0 || classes <= 0 || overlap_interval <= 0 || sample_duration <= 0 ||
      overlap_interval > frames_per_window) {
    set_error("window_merge: bad shape n_windows=%d frames=%d classes=%d overlap_interval=%d duration=%d", n_windows,
              frames_per_window, classes, overlap_interval, sample_duration);
    return SED_ERR_BAD_SHAPE;
  }
  const int total = (n_windows - 1) * overlap_interval + frames_per_window;
  const long n = static_cast<long>(total) * classes;
  window_merge_kernel<<<dim3(static_cast<unsigned>((n + 255) / 256), n_recordings), 256, 0, stream>>>(
      frames, n_windows, frames_per_window, classes, overlap_interval, sample_duration, total, merged);
  return cudaGetLastError() == cudaSuccess ? SED_OK : SED_ERR_CUDA;
}


// =================================================================================================
// Event extraction: double-threshold hysteresis + smoothing + salt removal per (clip, class).
// utils/vad.py:11-45 activity_detection and its helpers :108-199, as called by
// utils/utilities.py:82-153 / pytorch/predict.py:57-121 (frame_prediction_to_event_prediction*).
// One thread per (clip, class) runs the reference's four list passes as a chain of O(1)-state streaming stages:
//   runs of x > high  -> [bgn, fin] with the reference's asymmetric +1s (find_bgn_fin_pairs)
//   -> extension while x >= low (activity_detection_with_second_thres) -> smooth(1) -> smooth(n_smooth)
//   -> drop fin - bgn <= n_salt -> events[clip][class][0..count) = (bgn, fin) in frames.
// Comparisons are done in float64 like numpy does for float32 data against the float64 thresholds of the
// shipped opt_thresholds pickles.
// =================================================================================================
struct SmoothState {
  int has, mem_bgn, last_fin, n;
};

template <typename Emit>
__device__ __forceinline__ void smooth_push(SmoothState& st, int bgn, int fin, Emit emit) {
  if (st.has && (bgn - st.last_fin > st.n)) {
    emit(st.mem_bgn, st.last_fin);
    st.mem_bgn = bgn;
  } else if (!st.has) {
    st.has = 1;
    st.mem_bgn = bgn;
  }
  st.last_fin = fin;
}
template <typename Emit>
__device__ __forceinline__ void smooth_flush(SmoothState& st, Emit emit) {
  if (st.has) emit(st.mem_bgn, st.last_fin);
}

__global__ void events_kernel(const float* __restrict__ frames, int n_clips, int n_frames, int classes,
                              const double* __restrict__ high, const double* __restrict__ low, int use_low,
                              const int* __restrict__ n_smooth, const int* __restrict__ n_salt, int max_events,
                              int* __restrict__ events, int* __restrict__ counts) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_clips * classes) return;
  const int clip = idx / classes, c = idx - clip * classes;
  const float* x = frames + static_cast<size_t>(clip) * n_frames * classes + c;
  auto X = [&](int i) { return static_cast<double>(x[static_cast<size_t>(i) * classes]); };
  const double th = high[c], tl = use_low ? low[c] : 0.0;
  const int salt = n_salt[c];
  int* ev = events + static_cast<size_t>(idx) * max_events * 2;
  int count = 0;

  auto emit_final = [&](int bgn, int fin) {  // remove_salt_noise (vad.py:187-199)
    if (fin - bgn <= salt) return;
    if (count < max_events) {
      ev[2 * count] = bgn;
      ev[2 * count + 1] = fin;
    }
    ++count;
  };
  SmoothState s2{0, 0, 0, n_smooth[c]};  // smooth(n_smooth) in activity_detection (vad.py:38)
  auto emit_s2 = [&](int bgn, int fin) { smooth_push(s2, bgn, fin, emit_final); };
  SmoothState s1{0, 0, 0, 1};            // smooth(n_smooth=1) inside the second-threshold pass (vad.py:154)
  auto emit_s1 = [&](int bgn, int fin) { smooth_push(s1, bgn, fin, emit_s2); };
  auto emit_raw = [&](int bgn, int fin) {  // activity_detection_with_second_thres (vad.py:133-152)
    if (use_low) {
      while (bgn != -1) {
        if (bgn >= n_frames || X(bgn) < tl) break;  // (the reference would raise IndexError at bgn == len(x))
        --bgn;
      }
      while (fin != n_frames) {
        if (X(fin) < tl) break;
        ++fin;
      }
      emit_s1(bgn + 1, fin);
    } else {
      emit_s2(bgn, fin);
    }
  };

  // find_bgn_fin_pairs (vad.py:108-130) over locts = where(x > high)
  int run_start = -1, run_end = -1, runs = 0;
  for (int i = 0; i < n_frames; ++i) {
    if (X(i) > th) {
      if (run_start >= 0 && i - run_end > 1) {  // a gap: the pending run is not the last one
        emit_raw(run_start + (runs > 0 ? 1 : 0), run_end + 1);
        ++runs;
        run_start = i;
      } else if (run_start < 0) {
        run_start = i;
      }
      run_end = i;
    }
  }
  if (run_start >= 0) emit_raw(run_start + (runs > 0 ? 1 : 0), run_end);  // last pair: fin without the +1
  if (use_low) smooth_flush(s1, emit_s2);
  smooth_flush(s2, emit_final);
  counts[idx] = count;
}

int events_launch(const float* frames, int n_clips, int n_frames, int classes, const double* high, const double* low,
                  const int* n_smooth, const int* n_salt, int max_events, int* events, int* counts,
                  cudaStream_t stream) {
  if (n_clips <= 0 || n_frames <= 0 || classes <= 0 || max_events <= 0) {
    set_error("events: bad shape clips=%d frames=%d classes=%d max_events=%d", n_clips, n_frames, classes, max_events);
    return SED_ERR_BAD_SHAPE;
  }
  const int n = n_clips * classes;
  events_kernel<<<(n + 127) / 128, 128, 0, stream>>>(frames, n_clips, n_frames, classes, high, low, low != nullptr,
                                                     n_smooth, n_salt, max_events, events, counts);
  return cudaGetLastError() == cudaSuccess ? SED_OK : SED_ERR_CUDA;
}

}  // namespace sed
