"""Batch inference driver: drop-in for `forward` / `move_data_to_device` of the reference's pytorch/pytorch_utils.py
(:6-15, :25-78), the loop `Evaluator.evaluate` runs over a DataLoader (pytorch/evaluate.py:64).

Same arguments, same returned dict of concatenated numpy arrays.  What differs does not change results
(SURVEY.md 8f-2):
  * a batch's `waveform` may stay int16 PCM (the HDF5 storage format, utils/utilities.py:78-79): the division by
    32767 then happens inside the front-end kernel instead of on the host;
  * on one of this package's models the loop runs through `pipeline.HostPipeline`: the host->device copy of batch
    k+1 and the device->host copy of batch k-1 overlap the kernels of batch k, instead of the reference's three
    serialised phases with a synchronising `.data.cpu().numpy()` per batch (pytorch_utils.py:57-62);
  * for any other module (e.g. a `torch.nn.DataParallel` wrapper) outputs stay on the device for at most
    `flush_every` batches before they are fetched, so device memory stays bounded however long the loader is.
"""
import collections

import numpy as np
import torch

_RESULT_ORDER = ('audio_name', 'clipwise_output', 'framewise_output', 'waveform', 'target', 'strong_target')


def move_data_to_device(x, device):
    """pytorch_utils.py:6-15 (float arrays -> FloatTensor, integer arrays -> LongTensor, anything else unchanged);
    int16 arrays are kept as int16 PCM for the front-end."""
    if isinstance(x, torch.Tensor):
        return x.to(device, non_blocking=True)
    kind = str(x.dtype)
    if 'float' in kind:
        x = torch.Tensor(x)
    elif kind == 'int16':
        x = torch.from_numpy(np.ascontiguousarray(x))
    elif 'int' in kind:
        x = torch.LongTensor(x)
    else:
        return x
    return x.to(device, non_blocking=True)


def _host_waveform(x):
    """A batch's waveform as the CPU tensor the host pipeline takes: float32, or int16 PCM left as it is."""
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
    if t.dtype != torch.int16:
        t = t.float()
    return t.contiguous()


def _packed_engine(model, device):
    """The PackedModel behind one of this package's models on a CUDA device, else None."""
    get = getattr(model, "_packed_for", None)
    if get is None or device.type != "cuda":
        return None
    return get(device)


def forward(model, data_loader, return_input=False, return_target=False, flush_every=8, pipeline_depth=2):
    """Forward data to model (pytorch_utils.py:25-78).

    Returns:
      output_dict: {'audio_name': (N,), 'clipwise_output': (N, classes_num),
                    'framewise_output': (N, frames_num, classes_num),
                    (optional) 'waveform', 'target': (N, classes_num), 'strong_target': (N, frames_num, classes_num)}
    """
    device = next(model.parameters()).device
    model.eval()
    gathered = collections.defaultdict(list)   # key -> per-batch numpy arrays, in loader order

    def keep_side_data(batch):
        gathered['audio_name'].append(batch['audio_name'])
        if return_input:
            gathered['waveform'].append(batch['waveform'])
        if return_target:
            for key in ('target', 'strong_target'):
                if key in batch:
                    gathered[key].append(batch[key])

    packed = _packed_engine(model, device)
    if packed is not None:
        pipe = packed.host_pipeline(depth=pipeline_depth, micro_batch=model.micro_batch, variant=model.conv_variant)
        pipe.drain()

        def collect():
            res = pipe.result()
            # fresh arrays per batch (the reference's `.data.cpu().numpy()`): the pipeline's buffers rotate
            gathered['clipwise_output'].append(np.array(res['clipwise_output'].numpy()))
            gathered['framewise_output'].append(np.array(res['framewise_output'].numpy()))

        for batch in data_loader:
            if pipe.in_flight >= pipe.depth:
                collect()
            pipe.submit(_host_waveform(batch['waveform']))
            keep_side_data(batch)
        while pipe.in_flight:
            collect()
    else:
        waiting = []  # outputs still on the device

        def flush():
            for out in waiting:
                for key in ('clipwise_output', 'framewise_output'):
                    if key in out:
                        gathered[key].append(out[key].data.cpu().numpy())
            waiting.clear()

        for batch in data_loader:
            wave = move_data_to_device(batch['waveform'], device)
            with torch.no_grad():
                waiting.append(model(wave))
            keep_side_data(batch)
            if len(waiting) >= flush_every:
                flush()
        flush()
    return {key: np.concatenate(gathered[key], axis=0) for key in _RESULT_ORDER if key in gathered}
