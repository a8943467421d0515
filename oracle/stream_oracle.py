"""CPU oracle for the streaming prediction path (window loop + merge + avg_merge).

TEST INFRASTRUCTURE ONLY.  Restates pytorch/predict.py:279-349 (window schedule) and utils/utilities.py:405-446
(merge, avg_merge) in numpy, quirks included; pinned against the imported reference functions in
tests/test_stream_oracle.py (live) and through tests/golden/stream_merge.npz (fixtures from the reference).
"""
import numpy as np
import torch

import sed_oracle


def window_starts(audio_duration, sample_duration, overlap=True):
    """Window start times in seconds.  predict.py:279-281 (start = end = 0), :297 (`while end <= audio_duration`),
    :334-339 (start += 1 with --overlap -- a literal 1 s, not overlap_value -- else += sample_duration)."""
    starts = []
    start, end = 0, 0
    while end <= audio_duration:
        starts.append(start)
        start += 1 if overlap else sample_duration
        end = start + sample_duration
    return starts


def merge(prev, curr, sample_duration, num_segment, overlap_value=1):
    """utils/utilities.py:405-414: overlap-add `curr` at frame (num_segment-1)*int(100*overlap_value)."""
    overlap_interval = int(100 * overlap_value)
    front_cutoff = (num_segment - 1) * overlap_interval
    back_cutoff = prev.shape[1] - front_cutoff
    merged = prev[:, front_cutoff:] + curr[:, :back_cutoff]
    return np.concatenate((prev[:, :front_cutoff], merged, curr[:, back_cutoff:]), axis=1)


def avg_merge(merged, sample_duration, overlap_value=1):
    """utils/utilities.py:425-436.  Bug-compatible: the first and last block are never divided, tail blocks are
    divided by (true overlap count + 1), and a single window is divided by 2, 3, 4 in its middle blocks."""
    merged = merged.copy()
    overlap_interval = int(100 * overlap_value)
    interval = (sample_duration * 100) - overlap_interval
    for i in range(overlap_interval, merged.shape[1] - overlap_interval, overlap_interval):
        if i < interval:
            num_overlaps = i // overlap_interval + 1
        elif i >= merged.shape[1] - interval:
            num_overlaps = ((merged.shape[1] - i) // overlap_interval) + 1
        else:
            num_overlaps = sample_duration
        merged[:, i:i + overlap_interval] /= num_overlaps
    return merged


def merge_windows(frames, sample_duration, overlap_value=1):
    """frames [n_windows, frames_per_window, classes] -> merged [1, total, classes] exactly as the loop of
    predict.py:323-329 followed by :349 builds it."""
    merged = prev = None
    for k in range(frames.shape[0]):
        curr = frames[k:k + 1]
        num_segment = k + 1
        if num_segment == 2:
            merged = merge(prev, curr, sample_duration, num_segment, overlap_value)
        elif num_segment > 2:
            merged = merge(merged, curr, sample_duration, num_segment, overlap_value)
        else:
            merged = curr
        prev = curr
    return avg_merge(merged, sample_duration, overlap_value)


def pad_truncate_sequence(x, max_len):
    """utils/utilities.py:66-70"""
    if len(x) < max_len:
        return np.concatenate((x, np.zeros(max_len - len(x))))
    return x[0:max_len]


def streaming_predict(sd, audio_full, model_type, sample_rate, n_fft, hop, sample_duration=5, overlap_value=1):
    """predict.py:297-349 with --overlap: one forward per window (B = 1), merge, avg_merge.
    Returns (merged [1, total, 25] float32, per-window framewise [n_windows, frames, 25])."""
    audio_full = np.asarray(audio_full, dtype=np.float32)
    audio_duration = len(audio_full) / float(sample_rate)
    audio_samples = sample_rate * sample_duration
    outs = []
    for start in window_starts(audio_duration, sample_duration, overlap=True):
        start_index = int(start * sample_rate)
        end_index = int((sample_duration * sample_rate) + start_index)
        audio = pad_truncate_sequence(audio_full[start_index:end_index], audio_samples)
        wave = torch.Tensor(audio).reshape(1, -1)
        out = sed_oracle.model_forward(sd, wave, model_type, n_fft, hop)
        outs.append(out["framewise_output"].numpy())
    frames = np.concatenate(outs, axis=0)
    return merge_windows(frames, sample_duration, overlap_value), frames


# ----------------------------------------------------------------------------------------------
# Event extraction: utils/vad.py:11-45 activity_detection and helpers :108-199
# ----------------------------------------------------------------------------------------------
def overlap_eval_predict(sd, audio_full, audio_duration, model_type, sample_rate, n_fft, hop, sample_duration,
                         overlap_value):
    """The per-file loop of pytorch/main_strong.py:768-835: `audio_full` is already padded / truncated to 10 s (:785);
    windows advance by `overlap_value` seconds (:826), are sliced without padding (:790-792) and run with batch size 1;
    merge from the second window on (:813-818), avg_merge at the end (:835).  Returns (merged, list of window frames)."""
    import torch
    import sed_oracle
    num_segment, start, end = 1, 0, 0
    merged = prev = None
    frames = []
    while end <= audio_duration:
        start_index = int(start * sample_rate)
        end_index = int((sample_duration * sample_rate) + start_index)
        audio = torch.from_numpy(np.asarray(audio_full[start_index:end_index], dtype=np.float32))[None]
        curr = sed_oracle.model_forward(sd, audio, model_type, n_fft, hop)["framewise_output"].numpy()
        frames.append(curr)
        if num_segment == 2:
            merged = merge(prev, curr, sample_duration, num_segment, overlap_value)
        elif num_segment > 2:
            merged = merge(merged, curr, sample_duration, num_segment, overlap_value)
        else:
            merged = curr
        prev = curr
        start += overlap_value
        end = start + sample_duration
        num_segment += 1
    merged = avg_merge(merged.copy(), sample_duration, overlap_value)
    return merged, frames


def find_bgn_fin_pairs(locts):
    """vad.py:108-130.  Gap handling adds +1 to the closing fin AND to the next bgn; the last fin has no +1."""
    if len(locts) == 0:
        return []
    bgns, fins = [locts[0]], []
    for i in range(1, len(locts)):
        if locts[i] - locts[i - 1] > 1:
            fins.append(locts[i - 1] + 1)
            bgns.append(locts[i] + 1)
    fins.append(locts[-1])
    return [[b, f] for b, f in zip(bgns, fins)]


def smooth(pairs, n_smooth):
    """vad.py:159-184"""
    if len(pairs) == 0:
        return []
    out = []
    mem_bgn, fin = pairs[0]
    for n in range(1, len(pairs)):
        pre_fin = pairs[n - 1][1]
        bgn, fin = pairs[n]
        if bgn - pre_fin > n_smooth:
            out.append([mem_bgn, pre_fin])
            mem_bgn = bgn
    out.append([mem_bgn, fin])
    return out


def second_threshold(x, pairs, thres):
    """vad.py:133-156 (the reference raises IndexError when a pair starts at len(x); treated as a stop here)."""
    out = []
    for bgn, fin in pairs:
        while bgn != -1:
            if bgn >= len(x) or x[bgn] < thres:
                break
            bgn -= 1
        while fin != len(x):
            if x[fin] < thres:
                break
            fin += 1
        out.append([bgn + 1, fin])
    return smooth(out, 1)


def activity_detection(x, thres, low_thres=None, n_smooth=1, n_salt=0):
    """vad.py:11-45"""
    x = np.asarray(x)
    locts = np.where(x > thres)[0]
    pairs = find_bgn_fin_pairs([int(v) for v in locts])
    if low_thres is not None:
        pairs = second_threshold(x, pairs, low_thres)
    pairs = smooth(pairs, n_smooth)
    return [[b, f] for b, f in pairs if f - b > n_salt]  # remove_salt_noise vad.py:187-199
