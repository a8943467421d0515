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
