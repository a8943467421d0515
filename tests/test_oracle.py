"""The CPU oracle against (a) fixtures produced by the real reference and (b) the live reference when its
tree is present (build container only)."""
import numpy as np
import pytest
import torch

import ref_import
import sed_oracle as so
from conftest import load_golden, logmel_close, synthetic_sd
from sed_b200 import synth

PRESETS = synth.PRESETS


@pytest.mark.parametrize("sr", [8000, 16000, 32000])
def test_oracle_frontend_matches_reference_golden(sr):
    g = load_golden("frontend_%dk.npz" % (sr // 1000))
    n_fft, hop, fmin, fmax = PRESETS[sr]
    wave = torch.from_numpy(g["wave_i16"]).float() / 32767.0
    wr, wi = so.stft_conv_weights(n_fft)
    spec = so.spectrogram(wave, wr, wi, n_fft, hop)
    lm = so.logmel(spec, torch.from_numpy(g["melW"]))[:, 0].numpy()
    assert lm.shape == g["logmel"].shape
    assert logmel_close(lm, g["logmel"], rtol=5e-5).all()
    lm80 = so.logmel(spec, torch.from_numpy(g["melW"]), top_db=80.0)[:, 0].numpy()
    assert logmel_close(lm80, g["logmel_top80"], rtol=5e-5).all()
    np.testing.assert_allclose(spec[:, 0, ::25].numpy(), g["spec_rows"], rtol=2e-4, atol=1e-9)
    # silence maps to exactly -100 dB (amin = 1e-10)
    assert np.all(lm[3] == -100.0)


@pytest.mark.parametrize("sr", [8000, 16000, 32000])
def test_oracle_frontend_vs_float64_fft(sr):
    """Independent cross-check (method of stft.py:925-1177 debug()): float64 rfft log-mel."""
    g = load_golden("frontend_%dk.npz" % (sr // 1000))
    n_fft, hop, _, _ = PRESETS[sr]
    wave = g["wave_i16"].astype(np.float64) / 32767.0
    f64 = so.logmel_float64_fft(wave, n_fft, hop, g["melW"])
    # noise-like signals: the float32 reference sits within ~1e-4 dB of the exact transform
    for i in (0, 1):
        assert np.abs(g["logmel"][i] - f64[i]).max() < 2e-4


def test_mel_filterbank_matches_torchaudio():
    torchaudio = pytest.importorskip("torchaudio")
    import melbank
    from sed_b200 import melbank as pkg_melbank
    for sr, (n_fft, hop, fmin, fmax) in PRESETS.items():
        a = melbank.slaney_mel_filterbank(sr, n_fft, 64, fmin, fmax).T
        b = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, fmin, fmax, 64, sr, norm="slaney",
                                                  mel_scale="slaney").numpy()
        c = pkg_melbank.mel_filterbank(sr, n_fft, 64, fmin, fmax).numpy()
        assert np.abs(a - b).max() < 2e-7
        assert np.abs(a - c).max() < 1e-9


@pytest.mark.parametrize("mt", synth.MODEL_TYPES)
@pytest.mark.parametrize("sr", [8000, 16000, 32000])
def test_oracle_model_matches_reference_golden(mt, sr, golden_meta):
    tag = "gru" if "Gru" in mt else "transformer"
    g = load_golden("model_%s_%dk.npz" % (tag, sr // 1000))
    sd = synthetic_sd(mt, sr)
    fp = golden_meta["ckpt_fingerprint"]["%s_%d" % (mt, sr)]
    var_sum = float(sum(v.double().sum() for k, v in sd.items() if k.endswith("running_var")))
    assert abs(var_sum - fp["bn_var_sum"]) <= 1e-4 * abs(fp["bn_var_sum"]), "synthetic checkpoint drifted"
    n_fft, hop, _, _ = PRESETS[sr]
    wave = torch.from_numpy(g["wave_i16"]).float() / 32767.0
    out = so.model_forward(sd, wave, mt, n_fft, hop)
    for k in ("framewise_output", "clipwise_output", "embedding"):
        assert tuple(out[k].shape) == g[k].shape
        assert np.abs(out[k].numpy() - g[k]).max() < 2e-4, k


@pytest.mark.parametrize("mt", synth.MODEL_TYPES)
def test_oracle_output_shapes_5s_and_10s(mt):
    """GRU model pads 496 -> 500 frames (models.py:680-681); Transformer returns 496 (models.py:1069-1070)."""
    g = load_golden("model_%s_full.npz" % ("gru" if "Gru" in mt else "transformer"))
    assert g["framewise_10s"].shape == (3, 1000, 25)
    assert g["framewise_5s"].shape == (3, 500 if "Gru" in mt else 496, 25)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("mt", synth.MODEL_TYPES)
def test_oracle_matches_live_reference(mt):
    rs, rm = ref_import.load()
    sd = synthetic_sd(mt, 16000)
    model = getattr(rm, mt)(16000, 512, 160, 64, 25, 7000, 25, "logmel").eval()
    model.load_state_dict(sd, strict=True)
    wave = torch.cat([synth.synthetic_waveform(2, 48000, seed=5, kind="events"), synth.synthetic_waveform(1, 48000, seed=6)])
    with torch.no_grad():
        ref = model(wave)
    out, stages = so.model_forward(sd, wave, mt, 512, 160, return_stages=True)
    for k in ref:
        assert (out[k] - ref[k]).abs().max().item() < 2e-5, k
    # per-stage: reference front-end modules
    with torch.no_grad():
        lm = model.logmel_extractor(model.spectrogram_extractor(wave))
    assert torch.equal(lm, stages["logmel"])


def test_threshold_fixture_matches_shipped_pickles(thresholds):
    assert "Cnn_9layers_Gru_FrameAtt/best_logmel_16k.sed.valid.pkl" in thresholds
    for key, d in thresholds.items():
        assert len(d["sed_high_threshold"]) == 25 and len(d["sed_low_threshold"]) == 25
        assert d["n_smooth"] == 10 and d["n_salt"] == 10
