"""Sibling heads of the Cnn_9layers trunk (reference pytorch/models.py:213-561, 880-978):
oracle against reference goldens / the live reference (CPU), CUDA path against the goldens (GPU)."""
import inspect

import numpy as np
import pytest
import torch

import ref_import
import sed_oracle as so
from conftest import load_golden
from sed_b200 import models, synth

SIBS = list(synth.SIBLING_TYPES)


def classes_of(mt):
    return 10 if mt == "Cnn_9layers_FrameMax" else 25  # as in oracle/gen_golden.py::sibling_goldens


def ctor_args(mt):
    args = (16000, 512, 160, 64, 25, 7000, classes_of(mt))
    return args + ("logmel",) if mt in ("Cnn_9layers_Gru_FrameAvg", "Cnn_9layers_Transformer_FrameAvg") else args


@pytest.mark.parametrize("mt", SIBS)
def test_sibling_oracle_matches_reference_golden(mt):
    g = load_golden("model_siblings_16k.npz")
    wave = torch.from_numpy(g["wave_i16"]).float() / 32767.0
    sd = synth.synthetic_state_dict(mt, 16000, classes_num=classes_of(mt))
    out = so.model_forward(sd, wave, mt, 512, 160)
    for k in ("framewise_output", "clipwise_output", "embedding"):
        ref = g["%s.%s" % (mt, k)]
        assert tuple(out[k].shape) == ref.shape, k
        assert np.abs(out[k].numpy() - ref).max() < 2e-4, k
    assert g[mt + ".framewise_output"].shape == (4, 144, classes_of(mt))  # 151 frames -> T' = 18 -> 144, never padded


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("mt", SIBS)
def test_sibling_boundary_matches_live_reference(mt):
    """Constructor arity and state_dict layout equal the reference class."""
    _, rm = ref_import.load()
    ref_cls = getattr(rm, mt)
    assert list(inspect.signature(getattr(models, mt).__init__).parameters) == \
        list(inspect.signature(ref_cls.__init__).parameters)
    ref = ref_cls(*ctor_args(mt))
    mine = getattr(models, mt)(*ctor_args(mt))
    want = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    got = {k: tuple(v.shape) for k, v in mine.state_dict().items()}
    assert list(got) == list(want) and got == want
    res = mine.load_state_dict(synth.synthetic_state_dict(mt, 16000, classes_num=classes_of(mt)))
    assert not res.missing_keys and not res.unexpected_keys


@pytest.mark.gpu
@pytest.mark.parametrize("mt", SIBS)
def test_sibling_cuda_matches_reference_golden(mt):
    g = load_golden("model_siblings_16k.npz")
    wave = (torch.from_numpy(g["wave_i16"]).float() / 32767.0).to("cuda:0")
    model = getattr(models, mt)(*ctor_args(mt))
    model.load_state_dict(synth.synthetic_state_dict(mt, 16000, classes_num=classes_of(mt)))
    out = model.to("cuda:0").eval()(wave)
    for k in ("framewise_output", "clipwise_output"):
        got, ref = out[k].cpu().numpy(), g["%s.%s" % (mt, k)]
        assert got.shape == ref.shape and got.dtype == np.float32
        # clips 0-2: north_star tolerance; clip 3 = digital silence (outside the bn0 calibration range), looser
        assert np.abs(got[:3] - ref[:3]).max() <= 2e-3, (k, np.abs(got[:3] - ref[:3]).max())
        assert np.abs(got[3] - ref[3]).max() <= (2e-2 if "Transformer" in mt else 5e-3), (k, np.abs(got[3] - ref[3]).max())
    emb, ref = out["embedding"].cpu().numpy(), g[mt + ".embedding"]
    assert emb.shape == ref.shape
    # cla (probabilities) for FrameAtt; un-squashed features (|x| up to ~5, 16-bit operands upstream) otherwise
    assert np.abs(emb[:3] - ref[:3]).max() <= (2e-3 if mt == "Cnn_9layers_FrameAtt" else 3e-2)


@pytest.mark.gpu
def test_sibling_host_entry_and_shards():
    from sed_b200 import engine
    mt = "Cnn_9layers_FrameAvg"
    pm = engine.PackedModel(synth.synthetic_state_dict(mt, 16000), mt, 512, 160, torch.device("cuda:0"))
    wave = synth.synthetic_waveform(5, 48000, seed=21, kind="events").pin_memory()
    host = pm.forward_host(wave, micro_batch=2)
    dev = pm.forward(wave.to("cuda:0"), micro_batch=3)
    assert torch.equal(host["framewise_output"], dev["framewise_output"].cpu())
    assert torch.equal(host["clipwise_output"], dev["clipwise_output"].cpu())
    one = pm.forward(wave[2:3].to("cuda:0"))
    assert torch.equal(one["framewise_output"], dev["framewise_output"][2:3])
