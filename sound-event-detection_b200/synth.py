"""Seeded synthetic inputs and checkpoints in the reference's formats.

The shipped `checkpoints/main_strong/**/best_*.pth` files are not available, so
parity tests and the benchmark use a seeded stand-in with the exact `state_dict` layout of
`pytorch/models.py:564-688` / `:981-1077`, written in the reference's checkpoint format
`{'iteration', 'model', 'optimizer'}` (`pytorch/main_strong.py:326-333`).

BatchNorm running statistics are calibrated on a short synthetic batch so every layer sees
"trained-like" unit-scale activations (otherwise eight un-normalised layers collapse the signal and
parity would be vacuous).  Calibration is a one-off CPU helper, not part of the inference path.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .melbank import mel_filterbank, windowed_dft_kernels

PRESETS = {
    # sample_rate: (window_size, hop_size, fmin, fmax)   pytorch/predict.py:186-203
    8000: (256, 80, 12, 3500),
    16000: (512, 160, 25, 7000),
    32000: (1024, 320, 50, 14000),
}

MODEL_TYPES = ("Cnn_9layers_Gru_FrameAtt", "Cnn_9layers_Transformer_FrameAtt")
# sibling heads on the same trunk (pytorch/models.py:213-561, 880-978): (temporal block, head)
SIBLING_TYPES = {
    "Cnn_9layers_FrameMax": (None, "fc"), "Cnn_9layers_FrameAvg": (None, "fc"), "Cnn_9layers_FrameAtt": (None, "att"),
    "Cnn_9layers_Gru_FrameAvg": ("gru", "fc"), "Cnn_9layers_Transformer_FrameAvg": ("mha", "fc"),
}
_PLANS = dict({"Cnn_9layers_Gru_FrameAtt": ("gru", "att"), "Cnn_9layers_Transformer_FrameAtt": ("mha", "att")},
              **SIBLING_TYPES)


def synthetic_waveform(batch, samples, seed=1234, rank=0, kind="noise", sample_rate=16000):
    """float32 [batch, samples] on CPU, quantised to int16 steps like the HDF5 path
    (`utils/features.py:370`, `utils/utilities.py:73-79`).  kind='noise' is SURVEY.md 8(d)'s bench
    input; kind='events' adds tone bursts / chirps / silence so framewise outputs vary in time."""
    g = torch.Generator().manual_seed(seed + rank)
    x = torch.clamp(0.1 * torch.randn(batch, samples, generator=g), -1.0, 1.0)
    if kind == "events":
        t = torch.arange(samples, dtype=torch.float64) / sample_rate
        for b in range(batch):
            env = torch.zeros(samples, dtype=torch.float64)
            sig = torch.zeros(samples, dtype=torch.float64)
            n_ev = 4 + (b % 3)
            for e in range(n_ev):
                r = torch.rand(4, generator=g).double()
                t0 = r[0] * t[-1] * 0.9
                dur = 0.2 + r[1] * 1.5
                f0 = 80.0 * (2.0 ** (r[2] * math.log2(0.4 * sample_rate / 80.0)))
                sweep = (r[3] - 0.5) * 0.8 * f0
                mask = ((t >= t0) & (t < t0 + dur)).double()
                ph = 2 * math.pi * (f0 * (t - t0) + 0.5 * sweep * (t - t0) ** 2 / dur)
                sig += mask * 0.3 * torch.sin(ph)
                env += mask
            quiet = (torch.rand(1, generator=g).item() * 0.5, 0.5 + torch.rand(1, generator=g).item() * 0.4)
            gate = torch.ones(samples, dtype=torch.float64)
            gate[int(quiet[0] * samples):int(quiet[0] * samples) + int(0.08 * samples)] = 0.02
            x[b] = torch.clamp(x[b].double() * 0.3 * gate + sig, -1.0, 1.0).float()
    elif kind != "noise":
        raise ValueError(kind)
    return (torch.round(x * 32767.0) / 32767.0).contiguous()


def _xavier_uniform(shape, g, gain=1.0):
    fan_out = shape[0] * int(np.prod(shape[2:])) if len(shape) > 2 else shape[0]
    fan_in = shape[1] * int(np.prod(shape[2:])) if len(shape) > 2 else shape[1]
    bound = gain * math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=g) * 2 - 1) * bound


def _bn_fold(sd, prefix):
    scale = sd[prefix + ".weight"] / torch.sqrt(sd[prefix + ".running_var"] + 1e-5)
    shift = sd[prefix + ".bias"] - sd[prefix + ".running_mean"] * scale
    return scale, shift


def synthetic_state_dict(model_type, sample_rate=16000, seed=0, calib_seconds=3.0, calib_clips=2, classes_num=25,
                         calib_silence=False):
    """Reference-layout `state_dict` (float32 CPU tensors) with calibrated BN statistics.
    calib_silence=True adds a half-silent clip to the calibration batch, so that the statistics of every BatchNorm
    cover digital silence the way a checkpoint trained on real recordings does (the default calibration batch has no
    silent stretch: silence then sits ~8 sigma outside bn0's range, the regime tests/test_gpu_decisions.py examines)."""
    if model_type not in _PLANS:
        raise ValueError("unsupported model_type %r" % (model_type,))
    temporal, head = _PLANS[model_type]
    n_fft, hop, fmin, fmax = PRESETS[sample_rate]
    g = torch.Generator().manual_seed(1000003 * seed + 17)
    sd = {}
    wr, wi = windowed_dft_kernels(n_fft, n_fft, "hann")
    sd["spectrogram_extractor.stft.conv_real.weight"] = wr
    sd["spectrogram_extractor.stft.conv_imag.weight"] = wi
    sd["logmel_extractor.melW"] = mel_filterbank(sample_rate, n_fft, 64, fmin, fmax)

    def bn(prefix, c):
        sd[prefix + ".weight"] = 0.8 + 0.4 * torch.rand(c, generator=g)
        sd[prefix + ".bias"] = 0.1 * torch.randn(c, generator=g)
        sd[prefix + ".running_mean"] = torch.zeros(c)
        sd[prefix + ".running_var"] = torch.ones(c)
        sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    bn("bn0", 64)
    cin = 1
    for i, c in enumerate((64, 128, 256, 512), start=1):
        p = "conv_block%d" % i
        sd[p + ".conv1.weight"] = _xavier_uniform((c, cin, 3, 3), g)
        sd[p + ".conv2.weight"] = _xavier_uniform((c, c, 3, 3), g)
        bn(p + ".bn1", c)
        bn(p + ".bn2", c)
        cin = c

    # ---- calibrate BN running statistics on a short synthetic batch (CPU, one-off) ----
    L = int(calib_seconds * sample_rate)
    wave = torch.cat([synthetic_waveform(calib_clips, L, seed=977 + seed, kind="events", sample_rate=sample_rate),
                      synthetic_waveform(1, L, seed=978 + seed, kind="noise")], 0)
    if calib_silence:
        half = synthetic_waveform(1, L, seed=979 + seed, kind="events", sample_rate=sample_rate)
        half[:, :L // 2] = 0.0
        wave = torch.cat([wave, half], 0)

    def calibrate(prefix, x, dims):
        mean = x.mean(dim=dims)
        var = x.var(dim=dims, unbiased=False)
        c = mean.numel()
        sd[prefix + ".running_mean"] = (mean + 0.05 * var.sqrt() * torch.randn(c, generator=g)).contiguous()
        sd[prefix + ".running_var"] = (var * (0.8 + 0.45 * torch.rand(c, generator=g)) + 1e-4).contiguous()

    with torch.no_grad():
        x = F.pad(wave[:, None, :], (n_fft // 2, n_fft // 2), mode="reflect")
        re = F.conv1d(x, wr, stride=hop)
        im = F.conv1d(x, wi, stride=hop)
        spec = (re ** 2 + im ** 2).transpose(1, 2)  # [B,T,F]
        lm = 10.0 * torch.log10(torch.clamp(spec @ sd["logmel_extractor.melW"], min=1e-10))
        calibrate("bn0", lm, (0, 1))
        s, b = _bn_fold(sd, "bn0")
        x = (lm * s + b)[:, None]  # [B,1,T,64]
        for i in range(1, 5):
            p = "conv_block%d" % i
            for j in (1, 2):
                x = F.conv2d(x, sd[p + ".conv%d.weight" % j], padding=1)
                calibrate(p + ".bn%d" % j, x, (0, 2, 3))
                s, b = _bn_fold(sd, p + ".bn%d" % j)
                x = torch.relu(x * s[None, :, None, None] + b[None, :, None, None])
            if i < 4:
                x = F.avg_pool2d(x, 2)

    if temporal == "gru":
        for suffix in ("", "_reverse"):
            sd["gru.weight_ih_l0" + suffix] = (torch.rand(768, 512, generator=g) * 2 - 1) * math.sqrt(3.0 / 512)
            sd["gru.weight_hh_l0" + suffix] = (torch.rand(768, 256, generator=g) * 2 - 1) * math.sqrt(3.0 / 256)
            sd["gru.bias_ih_l0" + suffix] = 0.1 * torch.randn(768, generator=g)
            sd["gru.bias_hh_l0" + suffix] = 0.1 * torch.randn(768, generator=g)
    elif temporal == "mha":
        for name in ("w_qs", "w_ks", "w_vs"):
            sd["multihead.%s.weight" % name] = torch.randn(512, 512, generator=g) * math.sqrt(2.0 / (512 + 64))
            sd["multihead.%s.bias" % name] = 0.05 * torch.randn(512, generator=g)
        sd["multihead.layer_norm.weight"] = torch.ones(512)
        sd["multihead.layer_norm.bias"] = torch.zeros(512)
        sd["multihead.fc.weight"] = torch.randn(512, 512, generator=g) * math.sqrt(2.0 / 1024)
        sd["multihead.fc.bias"] = 0.05 * torch.randn(512, generator=g)

    # head gain chosen so framewise probabilities span roughly [0.02, 0.98] on the synthetic inputs
    gain = 1.0 if temporal == "gru" else 0.5
    if head == "fc":
        sd["fc.weight"] = _xavier_uniform((classes_num, 512), g, gain=gain)
        sd["fc.bias"] = 0.5 * torch.randn(classes_num, generator=g)
        return {k: (v.float().contiguous() if v.is_floating_point() else v) for k, v in sd.items()}
    sd["att_block.att.weight"] = _xavier_uniform((25, 512, 1), g, gain=gain)
    sd["att_block.att.bias"] = 0.3 * torch.randn(25, generator=g)
    sd["att_block.cla.weight"] = _xavier_uniform((25, 512, 1), g, gain=gain)
    sd["att_block.cla.bias"] = 0.5 * torch.randn(25, generator=g)
    sd["att_block.bn_att.weight"] = torch.ones(25)
    sd["att_block.bn_att.bias"] = torch.zeros(25)
    sd["att_block.bn_att.running_mean"] = torch.zeros(25)
    sd["att_block.bn_att.running_var"] = torch.ones(25)
    sd["att_block.bn_att.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    return {k: (v.float().contiguous() if v.is_floating_point() else v) for k, v in sd.items()}


def synthetic_checkpoint(model_type, sample_rate=16000, seed=0):
    """Checkpoint dict in the reference's on-disk format (`pytorch/main_strong.py:326-333`)."""
    return {"iteration": 0, "model": synthetic_state_dict(model_type, sample_rate, seed), "optimizer": {}}
