"""Streaming prediction: the window loop of the reference's `predict.py` as ONE batched call.

Reference behaviour (pytorch/predict.py:297-349 with --overlap): a recording is cut into `sample_duration`-second
windows advancing by a literal 1 s, each window is run through the model with batch size 1 (a host sync per window),
the framewise outputs are overlap-added (`merge`) and block-averaged (`avg_merge`, utils/utilities.py:405-446).
Here the windows become the batch dimension -- the front-end kernel reads them in place from the recording with a
1-second clip stride -- and merge/avg_merge run as one device kernel, bug-compatible with the numpy code.
"""
import torch

from . import engine


def window_starts(audio_duration, sample_duration, overlap=True):
    """Start times (s) of the windows the reference runs: predict.py:279-281, 297, 334-339."""
    starts = []
    start, end = 0, 0
    while end <= audio_duration:
        starts.append(start)
        start += 1 if overlap else sample_duration
        end = start + sample_duration
    return starts


def predict_framewise(model, recording, sample_rate, sample_duration=5, overlap_value=1, return_windows=False):
    """model: a sed_b200 drop-in model in eval mode (or an engine.PackedModel); recording: 1-D float32 / int16
    waveform on the model's CUDA device.  Returns merged framewise probabilities [1, total_frames, 25]
    (what `merged` holds after predict.py:349)."""
    if isinstance(model, engine.PackedModel):
        packed = model
        micro_batch, variant = engine.DEFAULT_MICRO_BATCH, 4
    else:
        if model.training:
            raise RuntimeError("inference only -- call .eval()")
        if not recording.is_cuda:
            raise RuntimeError("recording is on %s; the B200 path has no CPU fallback" % (recording.device,))
        packed = model._packed_for(recording.device)
        micro_batch, variant = model.micro_batch, model.conv_variant
    audio_duration = recording.numel() / float(sample_rate)
    starts = window_starts(audio_duration, sample_duration, overlap=True)
    n_windows = len(starts)
    window_samples = int(sample_rate * sample_duration)
    with torch.no_grad():
        out = packed.forward_windows(recording, window_samples, int(sample_rate), n_windows, micro_batch=micro_batch,
                                     variant=variant)
        merged = engine.window_merge_avg(out["framewise_output"], int(100 * overlap_value), int(sample_duration))
    if return_windows:
        return merged, out
    return merged


def predict_framewise_many(model, recordings, sample_rate, sample_duration=5, overlap_value=1):
    """predict_framewise for several recordings in ONE batch (the per-file loop of pytorch/predict.py:264 folded into the
    batch dimension): `recordings` is a list of 1-D float32 / int16 CUDA tensors of one dtype.  Every recording is laid
    out in one buffer, zero padded up to the end of its last window (the reference zero-pads the last window,
    predict.py:302-305), the windows of all recordings are read in place through an offset table, and recordings with
    the same number of windows are merged by one launch.  Returns a list of [1, total_frames, classes] tensors equal,
    bit for bit, to calling predict_framewise per recording."""
    if isinstance(model, engine.PackedModel):
        packed, micro_batch, variant = model, engine.DEFAULT_MICRO_BATCH, 4
    else:
        if model.training:
            raise RuntimeError("inference only -- call .eval()")
        packed = model._packed_for(recordings[0].device)
        micro_batch, variant = model.micro_batch, model.conv_variant
    dev, dtype = recordings[0].device, recordings[0].dtype
    if any(r.dim() != 1 or r.device != dev or r.dtype != dtype for r in recordings):
        raise ValueError("recordings must be 1-D tensors of one dtype on one CUDA device")
    window_samples = int(sample_rate * sample_duration)
    counts, seg_len = [], []
    for r in recordings:
        starts = window_starts(r.numel() / float(sample_rate), sample_duration, overlap=True)
        counts.append(len(starts))
        need = max(r.numel(), starts[-1] * int(sample_rate) + window_samples)
        seg_len.append((need + 7) // 8 * 8)  # keep every segment 16-byte aligned for the cp.async staging
    flat = torch.zeros(sum(seg_len), dtype=dtype, device=dev)
    offsets, base = [], 0
    for r, nw, sl in zip(recordings, counts, seg_len):
        flat[base:base + r.numel()].copy_(r)
        offsets.extend(base + k * int(sample_rate) for k in range(nw))
        base += sl
    offsets = torch.tensor(offsets, dtype=torch.int64, device=dev)
    with torch.no_grad():
        out = packed.forward_windows(flat, window_samples, 0, int(offsets.numel()), micro_batch=micro_batch,
                                     variant=variant, offsets=offsets)
        frames = out["framewise_output"]
        merged = [None] * len(recordings)
        firsts, first = [], 0
        for nw in counts:
            firsts.append(first)
            first += nw
        oi = int(100 * overlap_value)
        for nw in sorted(set(counts)):
            idx = [f for f in range(len(recordings)) if counts[f] == nw]
            rows = torch.tensor([firsts[f] + k for f in idx for k in range(nw)], dtype=torch.int64, device=dev)
            group = frames.index_select(0, rows).view(len(idx), nw, frames.shape[1], frames.shape[2])
            m = engine.window_merge_avg(group, oi, int(sample_duration))
            for j, f in enumerate(idx):
                merged[f] = m[j:j + 1]
    return merged


class HostStreamer:
    """predict_framewise_many for recordings that live in HOST memory, end to end: every recording of a call is copied
    by DMA from the caller's (pinned) buffer straight to its place in one flat device buffer (zero padded up to the end
    of its last window by a device memset; nothing is repacked on the host), the window offset table is built on the
    host, and the merged frames of all recordings come back with one device->host copy per window-count group.  Two staging slots alternate, so the copy in of call k+1 can be queued while call k
    computes (`submit` returns at once, `result` blocks for the oldest call).  Replaces the per-file, per-window loop of
    pytorch/predict.py:264-349, which reads each window back to the host before it runs the next one."""

    def __init__(self, model, sample_rate, sample_duration=5, overlap_value=1):
        self.model, self.sr, self.dur, self.ov = model, int(sample_rate), int(sample_duration), overlap_value
        self.packed = model if isinstance(model, engine.PackedModel) else None
        self.slots = [dict(), dict()]
        self.calls = 0
        self.pending = []
        self.copy_stream = None

    def _packed(self, device):
        if self.packed is not None:
            return self.packed, engine.DEFAULT_MICRO_BATCH, 4
        if self.model.training:
            raise RuntimeError("inference only -- call .eval()")
        return self.model._packed_for(device), self.model.micro_batch, self.model.conv_variant

    def submit(self, recordings, device):
        """recordings: list of 1-D CPU tensors (float32 or int16 PCM, one dtype).  Queues the call; returns a ticket."""
        device = torch.device(device)
        packed, micro_batch, variant = self._packed(device)
        dtype = recordings[0].dtype
        if any(r.dim() != 1 or r.is_cuda or r.dtype != dtype for r in recordings) or dtype not in (torch.float32, torch.int16):
            raise ValueError("recordings must be 1-D float32 or int16 CPU tensors of one dtype")
        window_samples = self.sr * self.dur
        counts, seg_len = [], []
        for r in recordings:
            starts = window_starts(r.numel() / float(self.sr), self.dur, overlap=True)
            counts.append(len(starts))
            need = max(r.numel(), starts[-1] * self.sr + window_samples)
            seg_len.append((need + 7) // 8 * 8)
        total = sum(seg_len)
        slot = self.slots[self.calls % 2]
        if "done" in slot:
            slot["done"].synchronize()
        with torch.cuda.device(device):
            if self.copy_stream is None:
                self.copy_stream = torch.cuda.Stream(device)
            key = (total, dtype, len(recordings))
            if slot.get("key") != key:
                slot["dev"] = torch.empty(total, dtype=dtype, device=device)
                slot["off_host"] = torch.zeros(sum(counts), dtype=torch.int64).pin_memory()
                slot["off_dev"] = torch.empty(sum(counts), dtype=torch.int64, device=device)
                slot["key"] = key
                slot["out"] = {}
            base, offs = 0, []
            for nw, sl in zip(counts, seg_len):
                offs.extend(base + k * self.sr for k in range(nw))
                base += sl
            slot["off_host"].copy_(torch.tensor(offs, dtype=torch.int64))
            slot["src"] = list(recordings)  # keeps pinned sources alive until their asynchronous copies have run
            compute = torch.cuda.current_stream(device)
            cs = self.copy_stream  # the slot's previous call has finished (host-synchronised above): copy at once,
            with torch.cuda.stream(cs):  # under the kernels of the call queued before this one
                # the DMA engine does the layout: every recording goes straight from the caller's (pinned) buffer to its
                # place in the flat device buffer -- no host-side repacking; the zero padding is a device memset
                base = 0
                for r, sl in zip(recordings, seg_len):
                    slot["dev"][base:base + r.numel()].copy_(r, non_blocking=True)
                    if r.numel() < sl:
                        slot["dev"][base + r.numel():base + sl].zero_()
                    base += sl
                slot["off_dev"].copy_(slot["off_host"], non_blocking=True)
            compute.wait_stream(cs)
            with torch.no_grad():
                out = packed.forward_windows(slot["dev"], window_samples, 0, len(offs), micro_batch=micro_batch,
                                             variant=variant, offsets=slot["off_dev"])
                frames = out["framewise_output"]
                oi = int(100 * self.ov)
                groups, first = {}, 0
                for f, nw in enumerate(counts):
                    groups.setdefault(nw, []).append((f, first))
                    first += nw
                results = []
                for nw, members in sorted(groups.items()):
                    if all(members[i + 1][1] == members[i][1] + nw for i in range(len(members) - 1)):
                        group = frames[members[0][1]:members[0][1] + nw * len(members)]  # contiguous rows: a view
                    else:
                        rows = torch.tensor([fst + k for _, fst in members for k in range(nw)], dtype=torch.int64).to(device)
                        group = frames.index_select(0, rows)
                    merged = engine.window_merge_avg(group.view(len(members), nw, frames.shape[1], frames.shape[2]), oi, self.dur)
                    host_out = slot["out"].get(tuple(merged.shape))
                    if host_out is None:
                        host_out = slot["out"][tuple(merged.shape)] = torch.empty(merged.shape, dtype=torch.float32).pin_memory()
                    host_out.copy_(merged, non_blocking=True)
                    results.append((members, host_out))
            slot["done"] = torch.cuda.Event()
            slot["done"].record(compute)
        self.calls += 1
        self.pending.append((slot, results, len(recordings)))
        return self.calls - 1

    def result(self):
        """Merged frames of the oldest queued call: list of [1, total_frames, classes] CPU tensors, one per recording
        (views of the slot's pinned buffers: valid until two further calls have been submitted)."""
        slot, results, n = self.pending.pop(0)
        slot["done"].synchronize()
        merged = [None] * n
        for members, host_out in results:
            for j, (f, _) in enumerate(members):
                merged[f] = host_out[j:j + 1]
        return merged


def overlap_window_counts(audio_durations, sample_duration, overlap_value):
    """Windows the loop of main_strong.py:786-834 runs per file: start k*overlap_value for k = 0 and every k with
    k*overlap_value + sample_duration <= audio_duration (the `while end <= audio_duration` rule, end updated after
    each window)."""
    counts = []
    for d in audio_durations:
        n, start, end = 0, 0.0, 0.0
        while end <= d:
            n += 1
            start += overlap_value
            end = start + sample_duration
        counts.append(n)
    return counts


def predict_framewise_overlap(model, clips, sample_rate, sample_duration, overlap_value, audio_durations=None):
    """The overlap evaluation loop of pytorch/main_strong.py:768-834 for a batch of files at once.

    clips: (n_files, 10 * sample_rate) float32 / int16 CUDA tensor, every file already padded / truncated to 10 s
    (`pad_truncate_sequence`, main_strong.py:785); audio_durations: the files' real durations in seconds (default
    10.0 each) -- they decide how many windows a file gets.  All windows of all files run as ONE batch (the front-end
    reads them in place through an offset table), each file's windows are overlap-added and block-averaged on the
    device.  Returns a list of per-file tensors [1, total_frames, classes] (what `merged` holds after :835)."""
    if isinstance(model, engine.PackedModel):
        packed, micro_batch, variant = model, engine.DEFAULT_MICRO_BATCH, 4
    else:
        if model.training:
            raise RuntimeError("inference only -- call .eval()")
        if not clips.is_cuda:
            raise RuntimeError("clips are on %s; the B200 path has no CPU fallback" % (clips.device,))
        packed = model._packed_for(clips.device)
        micro_batch, variant = model.micro_batch, model.conv_variant
    if clips.dim() != 2:
        raise ValueError("clips must be (n_files, samples)")
    n_files, L = clips.shape
    if audio_durations is None:
        audio_durations = [L / float(sample_rate)] * n_files
    counts = overlap_window_counts(audio_durations, sample_duration, overlap_value)
    window_samples = int(sample_duration * sample_rate)
    starts = []
    for f, nw in enumerate(counts):
        for k in range(nw):
            s0 = int(k * overlap_value * sample_rate)  # int(start * sample_rate), main_strong.py:790
            if s0 + window_samples > L:
                raise ValueError("file %d: window %d runs past the padded clip (the reference would feed a short "
                                 "window to the model)" % (f, k))
            starts.append(f * L + s0)
    offsets = torch.tensor(starts, dtype=torch.int64, device=clips.device)
    flat = clips.contiguous().view(-1)
    with torch.no_grad():
        out = packed.forward_windows(flat, window_samples, 0, len(starts), micro_batch=micro_batch, variant=variant,
                                     offsets=offsets)
        frames = out["framewise_output"]
        merged = [None] * n_files
        oi = int(100 * overlap_value)
        first = 0
        firsts = []
        for nw in counts:
            firsts.append(first)
            first += nw
        for nw in sorted(set(counts)):  # files with the same number of windows are merged by one launch
            idx = [f for f in range(n_files) if counts[f] == nw]
            rows = torch.tensor([firsts[f] + k for f in idx for k in range(nw)], dtype=torch.int64, device=clips.device)
            group = frames.index_select(0, rows).view(len(idx), nw, frames.shape[1], frames.shape[2])
            m = engine.window_merge_avg(group, oi, int(sample_duration))
            for j, f in enumerate(idx):
                merged[f] = m[j:j + 1]
    return merged


# Class names of the 25-class strong-label task (utils/config.py:26)
LABELS = ['Applause', 'Breathing', 'Chatter', 'Cheering', 'Child_speech_kid_speaking', 'Clapping', 'Conversation',
          'Cough', 'Crowd', 'Crying_sobbing', 'Female_speech_woman_speaking', 'Laughter',
          'Male_speech_man_speaking', 'Run', 'Screaming', 'Shout', 'Sneeze', 'Walk_footsteps', 'Whispering',
          'Air_horn_truck_horn', 'Car_alarm', 'Emergency_vehicle', 'Explosion', 'Gunshot_gunfire', 'Siren']


def frame_prediction_to_event_prediction(framewise_output, sed_params_dict, frames_per_second=100, audio_names=None,
                                         labels=LABELS):
    """Device version of utils/utilities.py:82-153 (and predict.py:57-121, which uses filename 'test').

    framewise_output: (audios_num, frames_num, classes_num) float32 CUDA tensor; sed_params_dict as loaded from the
    shipped opt_thresholds pickles or the scalar defaults of predict.py:252-257.  Returns the reference's list of
    {'filename', 'onset', 'offset', 'event_label'} dicts in the reference's order (clip-major, class, event)."""
    events, counts = engine.extract_events(framewise_output, sed_params_dict['sed_high_threshold'],
                                           sed_params_dict.get('sed_low_threshold'), sed_params_dict['n_smooth'],
                                           sed_params_dict['n_salt'])
    events = events.cpu().numpy()
    counts = counts.cpu().numpy()
    event_list = []
    for n in range(events.shape[0]):
        name = 'test' if audio_names is None else audio_names[n]
        for k in range(events.shape[1]):
            for e in range(int(counts[n, k])):
                event_list.append({'filename': name, 'onset': events[n, k, e, 0] / float(frames_per_second),
                                   'offset': events[n, k, e, 1] / float(frames_per_second), 'event_label': labels[k]})
    return event_list
