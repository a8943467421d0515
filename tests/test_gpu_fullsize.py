"""BASELINE.json configs[1] at its full size (batch 1024 x 10 s, 16 kHz) on the GPU: size-independent properties
(shard invariance, determinism, host-buffer entry == device entry) plus an oracle spot check on a few clips."""
import numpy as np
import pytest
import torch

import sed_oracle as so
from conftest import synthetic_sd
from sed_b200 import engine, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MT = "Cnn_9layers_Gru_FrameAtt"


@pytest.fixture(scope="module")
def full_run():
    pm = engine.PackedModel(synthetic_sd(MT), MT, 512, 160, torch.device(DEV))
    wave = synth.synthetic_waveform(1024, 160000, seed=1234)
    wave[5] = synth.synthetic_waveform(1, 160000, seed=9, kind="events")[0]
    wave[1023] = synth.synthetic_waveform(1, 160000, seed=10, kind="events")[0]
    out = pm.forward(wave.to(DEV))
    torch.cuda.synchronize()
    return pm, wave, out


def test_full_batch_shapes_and_determinism(full_run):
    pm, wave, out = full_run
    assert out["framewise_output"].shape == (1024, 1000, 25)
    assert out["clipwise_output"].shape == (1024, 25)
    assert out["embedding"].shape == (1024, 25, 125)
    assert torch.isfinite(out["framewise_output"]).all()
    again = pm.forward(wave.to(DEV))
    for k in ("framewise_output", "clipwise_output", "embedding"):
        assert torch.equal(out[k], again[k]), k  # no atomics / order-dependent reductions anywhere on the path


def test_full_batch_equals_its_shards(full_run):
    """run(1024) == concat(run(shards)): micro-batches (148), GRU clusters (128 clips) and pooling tasks never mix clips."""
    pm, wave, out = full_run
    for (a, b) in ((0, 148), (148, 200), (1000, 1024), (511, 513)):
        part = pm.forward(wave[a:b].to(DEV))
        for k in ("framewise_output", "clipwise_output"):
            assert torch.equal(part[k], out[k][a:b]), (k, a, b)


def test_full_batch_host_entry_matches(full_run):
    pm, wave, out = full_run
    host = pm.forward_host(wave.pin_memory())
    assert torch.equal(host["framewise_output"], out["framewise_output"].cpu())
    assert torch.equal(host["clipwise_output"], out["clipwise_output"].cpu())
    q = torch.round(wave * 32767.0).to(torch.int16)
    host16 = pm.forward_host(q.pin_memory())  # int16 PCM in: x = q / 32767 reproduces the float input exactly
    assert torch.equal(host16["framewise_output"], host["framewise_output"])
    ref_frames = host["framewise_output"].clone()
    two = pm.forward_host(wave.pin_memory(), result_parts=2)  # temporal block + head per part of the batch
    assert torch.equal(two["framewise_output"], ref_frames)
    assert torch.equal(two["clipwise_output"], out["clipwise_output"].cpu())


def test_full_batch_spot_check_against_oracle(full_run, thresholds):
    pm, wave, out = full_run
    idx = [0, 5, 147, 148, 1023]
    ref = so.model_forward(synthetic_sd(MT), wave[idx], MT, 512, 160)
    fw = out["framewise_output"][idx].cpu().numpy()
    assert np.abs(fw - ref["framewise_output"].numpy()).max() <= 2e-3
    assert np.abs(out["clipwise_output"][idx].cpu().numpy() - ref["clipwise_output"].numpy()).max() <= 2e-3
    thr = np.asarray(thresholds["%s/best_logmel_16k.sed.valid.pkl" % MT]["sed_high_threshold"])
    agree = ((fw > thr[None, None, :]) == (ref["framewise_output"].numpy() > thr[None, None, :])).mean()
    assert agree >= 0.999, agree


def test_long_clips_60s():
    """60 s clips (T = 6001, T' = 750 GRU steps, 6000 framewise rows): nothing on the path is specialised to 10 s."""
    pm = engine.PackedModel(synthetic_sd(MT), MT, 512, 160, torch.device(DEV))
    wave = synth.synthetic_waveform(2, 960000, seed=71, kind="events")
    out = pm.forward(wave.to(DEV))
    assert out["framewise_output"].shape == (2, 6000, 25)
    ref = so.model_forward(synthetic_sd(MT), wave, MT, 512, 160)
    assert np.abs(out["framewise_output"].cpu().numpy() - ref["framewise_output"].numpy()).max() <= 2e-3
    assert np.abs(out["clipwise_output"].cpu().numpy() - ref["clipwise_output"].numpy()).max() <= 2e-3


def test_too_short_clip_is_rejected():
    pm = engine.PackedModel(synthetic_sd(MT), MT, 512, 160, torch.device(DEV))
    with pytest.raises(ValueError):
        pm.forward(torch.zeros(1, 800, device=DEV))  # T = 6 frames < one pooled step


def test_long_clip_batches_are_split_by_the_launch_group_cap():
    """60 s clips: the launch group shrinks to 148 clips (workspace and tile arithmetic stay where 1036 x 10 s puts
    them); a batch of 150 therefore runs as 148 + 2 and equals the separately computed pieces."""
    assert engine.clamp_micro_batch(engine.DEFAULT_MICRO_BATCH, 6001) == 148
    pm = engine.PackedModel(synthetic_sd(MT), MT, 512, 160, torch.device(DEV))
    wave = synth.synthetic_waveform(150, 960000, seed=72).to(DEV)
    out = pm.forward(wave)
    tail = pm.forward(wave[148:150])
    head = pm.forward(wave[0:3])
    assert torch.equal(out["framewise_output"][148:150], tail["framewise_output"])
    assert torch.equal(out["framewise_output"][0:3], head["framewise_output"])
    assert torch.isfinite(out["clipwise_output"]).all()
