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
        micro_batch, variant = 148, 2
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
