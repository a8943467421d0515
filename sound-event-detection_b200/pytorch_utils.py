"""Batch inference driver: drop-in for `forward` / `move_data_to_device` of the reference's pytorch/pytorch_utils.py
(:6-15, :25-78), the loop `Evaluator.evaluate` runs over a DataLoader (pytorch/evaluate.py:64).

Same arguments, same returned dict of concatenated numpy arrays.  Differences that do not change results
(SURVEY.md 8f-2): a batch's `waveform` may stay int16 PCM (the HDF5 storage format, utils/utilities.py:78-79) -- the
division by 32767 then happens inside the front-end kernel instead of on the host -- and, on a CUDA model, the
outputs of all batches stay on the device and come back with ONE device->host copy at the end instead of a
synchronising `.data.cpu().numpy()` per batch (pytorch_utils.py:57-62).
"""
import numpy as np
import torch


def move_data_to_device(x, device):
    """pytorch_utils.py:6-15 (float arrays -> FloatTensor, integer arrays -> LongTensor, anything else unchanged);
    int16 arrays are kept as int16 PCM for the front-end."""
    if isinstance(x, torch.Tensor):
        return x.to(device, non_blocking=True)
    if 'float' in str(x.dtype):
        x = torch.Tensor(x)
    elif str(x.dtype) == 'int16':
        x = torch.from_numpy(np.ascontiguousarray(x))
    elif 'int' in str(x.dtype):
        x = torch.LongTensor(x)
    else:
        return x
    return x.to(device, non_blocking=True)


def append_to_dict(dict, key, value):
    if key in dict.keys():
        dict[key].append(value)
    else:
        dict[key] = [value]


def forward(model, data_loader, return_input=False, return_target=False):
    """Forward data to model (pytorch_utils.py:25-78).

    Returns:
      output_dict: {'audio_name': (N,), 'clipwise_output': (N, classes_num),
                    'framewise_output': (N, frames_num, classes_num),
                    (optional) 'waveform', 'target': (N, classes_num), 'strong_target': (N, frames_num, classes_num)}
    """
    device = next(model.parameters()).device
    output_dict = {}
    on_device = {}  # key -> list of device tensors, fetched once at the end
    model.eval()
    for n, batch_data_dict in enumerate(data_loader):
        batch_waveform = move_data_to_device(batch_data_dict['waveform'], device)
        with torch.no_grad():
            batch_output = model(batch_waveform)
        append_to_dict(output_dict, 'audio_name', batch_data_dict['audio_name'])
        append_to_dict(on_device, 'clipwise_output', batch_output['clipwise_output'])
        if 'framewise_output' in batch_output.keys():
            append_to_dict(on_device, 'framewise_output', batch_output['framewise_output'])
        if return_input:
            append_to_dict(output_dict, 'waveform', batch_data_dict['waveform'])
        if return_target:
            if 'target' in batch_data_dict.keys():
                append_to_dict(output_dict, 'target', batch_data_dict['target'])
            if 'strong_target' in batch_data_dict.keys():
                append_to_dict(output_dict, 'strong_target', batch_data_dict['strong_target'])
    for key in output_dict.keys():
        output_dict[key] = np.concatenate(output_dict[key], axis=0)
    for key, parts in on_device.items():
        output_dict[key] = torch.cat([p.data for p in parts], dim=0).cpu().numpy()
    # key order of the reference: audio_name, clipwise_output, framewise_output, then the optional entries
    order = ['audio_name', 'clipwise_output', 'framewise_output', 'waveform', 'target', 'strong_target']
    return {k: output_dict[k] for k in order if k in output_dict}
