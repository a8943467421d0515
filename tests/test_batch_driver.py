"""Batch driver: sed_b200.pytorch_utils.forward against the reference's pytorch_utils.forward
(pytorch/pytorch_utils.py:25-78) on the same model and loader (CPU, reference tree present), and on the GPU with
float32 and int16 batches."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

import ref_import
import sed_oracle as so
from conftest import synthetic_sd
from sed_b200 import pytorch_utils as pu
from sed_b200 import synth

MT = "Cnn_9layers_Gru_FrameAtt"


def make_loader(n_batches, batch, L, int16=False, seed=3):
    out = []
    for i in range(n_batches):
        w = synth.synthetic_waveform(batch, L, seed=seed + i, kind="events").numpy()
        if int16:
            w = np.round(w * 32767.0).astype(np.int16)
        out.append({"audio_name": np.array(["clip_%d_%d.wav" % (i, j) for j in range(batch)]), "waveform": w,
                    "target": np.full((batch, 25), i, dtype=np.float32),
                    "strong_target": np.zeros((batch, 4, 25), dtype=np.float32)})
    return out


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_driver_equals_reference_driver_on_the_reference_model():
    _, rm = ref_import.load()
    saved = list(sys.path)
    sys.path.insert(0, os.path.join(ref_import.REF_ROOT, "pytorch"))
    try:
        sys.modules.pop("pytorch_utils", None)
        ref_pu = importlib.import_module("pytorch_utils")
    finally:
        sys.path[:] = saved
        sys.modules.pop("pytorch_utils", None)
    model = getattr(rm, MT)(16000, 512, 160, 64, 25, 7000, 25, "logmel").eval()
    model.load_state_dict(synthetic_sd(MT), strict=True)
    loader = make_loader(2, 2, 16000)
    for kw in ({}, {"return_input": True, "return_target": True}):
        a = ref_pu.forward(model, loader, **kw)
        b = pu.forward(model, loader, **kw)
        assert list(a.keys()) == list(b.keys())
        for k in a:
            assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, k
            assert np.array_equal(a[k], b[k]), k


def test_move_data_to_device_dtypes():
    dev = torch.device("cpu")
    assert pu.move_data_to_device(np.zeros((2, 3), np.float64), dev).dtype == torch.float32   # torch.Tensor(x)
    assert pu.move_data_to_device(np.zeros((2, 3), np.int64), dev).dtype == torch.int64       # LongTensor
    assert pu.move_data_to_device(np.zeros((2, 3), np.int16), dev).dtype == torch.int16       # PCM stays PCM
    s = np.array(["a", "b"])
    assert pu.move_data_to_device(s, dev) is s


@pytest.mark.gpu
@pytest.mark.parametrize("int16", [False, True])
def test_driver_on_the_b200_model_matches_oracle(int16):
    from sed_b200 import models
    model = getattr(models, MT)(16000, 512, 160, 64, 25, 7000, 25, "logmel")
    model.load_state_dict(synthetic_sd(MT))
    model = model.to("cuda:0")
    loader = make_loader(3, 2, 32000, int16=int16)
    out = pu.forward(model, loader, return_target=True)
    assert list(out.keys()) == ["audio_name", "clipwise_output", "framewise_output", "target", "strong_target"]
    assert out["framewise_output"].shape == (6, 200, 25) and out["audio_name"].shape == (6,)
    wave = np.concatenate([b["waveform"] for b in loader], 0)
    wave = torch.from_numpy(wave.astype(np.float32) / 32767.0 if int16 else wave)
    ref = so.model_forward(synthetic_sd(MT), wave, MT, 512, 160)
    assert np.abs(out["framewise_output"] - ref["framewise_output"].numpy()).max() <= 2e-3
    assert np.abs(out["clipwise_output"] - ref["clipwise_output"].numpy()).max() <= 2e-3


@pytest.mark.gpu
def test_driver_generic_path_flushes_in_bounded_groups():
    """A wrapped model (torch.nn.DataParallel has no packed engine behind it) takes the generic loop: outputs stay on the
    device for at most `flush_every` batches; results equal the pipelined path bit for bit."""
    from sed_b200 import models
    model = getattr(models, MT)(16000, 512, 160, 64, 25, 7000, 25, "logmel")
    model.load_state_dict(synthetic_sd(MT))
    model = model.to("cuda:0")
    loader = make_loader(5, 2, 32000)
    piped = pu.forward(model, loader)
    wrapped = torch.nn.DataParallel(model, device_ids=[0])
    for flush_every in (1, 2, 8):
        plain = pu.forward(wrapped, loader, flush_every=flush_every)
        assert list(plain.keys()) == list(piped.keys())
        for k in ("clipwise_output", "framewise_output"):
            assert np.array_equal(plain[k], piped[k]), (flush_every, k)
    assert np.array_equal(piped["audio_name"], np.concatenate([b["audio_name"] for b in loader]))
