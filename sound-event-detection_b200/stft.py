"""Drop-in `Spectrogram` / `LogmelFilterBank` (reference: pytorch/stft.py:636-734) on the B200 kernels.

Same constructor signatures, same parameters (`stft.conv_real.weight`, `stft.conv_imag.weight`, `melW`)
and therefore the same `state_dict` keys and shapes as the reference modules; `forward` runs the fused
FFT front-end kernels through the C ABI.  Inputs must live on a CUDA device -- there is no CPU path.
"""
import torch
import torch.nn as nn

from . import engine
from .melbank import mel_filterbank, windowed_dft_kernels


class STFT(nn.Module):
    """Parameter container mirroring reference STFT (pytorch/stft.py:157-221)."""

    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window='hann', center=True,
                 pad_mode='reflect', freeze_parameters=True):
        super().__init__()
        assert pad_mode in ['constant', 'reflect']  # stft.py:175
        self.n_fft = n_fft
        self.win_length = n_fft if win_length is None else win_length  # stft.py:185-186
        self.hop_length = int(self.win_length // 4) if hop_length is None else hop_length  # stft.py:189-190
        self.window = window
        self.center = center
        self.pad_mode = pad_mode
        out_channels = n_fft // 2 + 1
        self.conv_real = nn.Conv1d(1, out_channels, n_fft, stride=self.hop_length, padding=0, bias=False)
        self.conv_imag = nn.Conv1d(1, out_channels, n_fft, stride=self.hop_length, padding=0, bias=False)
        wr, wi = windowed_dft_kernels(n_fft, self.win_length, window)
        self.conv_real.weight.data = wr
        self.conv_imag.weight.data = wi
        if freeze_parameters:
            for p in self.parameters():
                p.requires_grad = False


class _PlanCache:
    """Per-device FrontendPlan cache invalidated when the backing parameters change."""

    def __init__(self):
        self._plans = {}

    def get(self, key, tensors, build):
        sig = tuple((t.data_ptr(), t._version) for t in tensors)
        hit = self._plans.get(key)
        if hit is None or hit[0] != sig:
            hit = (sig, build())
            self._plans[key] = hit
        return hit[1]


def _require_cuda(x, who):
    if not x.is_cuda:
        raise RuntimeError("%s: input is on %s; the B200 path has no CPU fallback" % (who, x.device))


class Spectrogram(nn.Module):
    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window='hann', center=True,
                 pad_mode='reflect', power=2.0, freeze_parameters=True):
        super().__init__()
        self.power = power
        self.stft = STFT(n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window, center=center,
                         pad_mode=pad_mode, freeze_parameters=True)
        self._cache = _PlanCache()

    def forward(self, input):
        """input (batch_size, data_length) -> (batch_size, 1, time_steps, n_fft // 2 + 1)."""
        _require_cuda(input, "Spectrogram")
        st = self.stft
        if not st.center or st.pad_mode != 'reflect':
            raise NotImplementedError("Spectrogram: only center=True, pad_mode='reflect' is built")
        x = input.float().contiguous()
        plan = self._cache.get(x.device, (st.conv_real.weight, st.conv_imag.weight), lambda: engine.FrontendPlan(
            st.conv_real.weight, st.conv_imag.weight, st.n_fft, st.hop_length, None, x.device))
        spec = engine.spectrogram_forward(plan, x)
        if self.power != 2.0:
            spec = spec ** (self.power / 2.0)  # stft.py:665-668
        return spec


class LogmelFilterBank(nn.Module):
    def __init__(self, sr=22050, n_fft=2048, n_mels=64, fmin=0.0, fmax=None, is_log=True, ref=1.0, amin=1e-10,
                 top_db=80.0, freeze_parameters=True):
        super().__init__()
        self.is_log = is_log
        self.ref = ref
        self.amin = amin
        self.top_db = top_db
        self.melW = nn.Parameter(mel_filterbank(sr, n_fft, n_mels, fmin, fmax))  # (n_fft // 2 + 1, mel_bins)
        if freeze_parameters:
            for p in self.parameters():
                p.requires_grad = False
        self._cache = _PlanCache()

    def forward(self, input):
        """input (*, n_fft // 2 + 1) -> (*, mel_bins)."""
        _require_cuda(input, "LogmelFilterBank")
        x = input.float()
        plan = self._cache.get((x.device, self.is_log, self.ref, self.amin), (self.melW,),
                               lambda: _mel_only_plan(self.melW, x.device, self.amin, self.ref, self.is_log))
        out = engine.logmel_rows_forward(plan, x)
        if self.is_log and self.top_db is not None:
            if self.top_db < 0:
                raise ValueError('top_db must be non-negative')  # stft.py:730-731
            out = torch.clamp(out, min=out.max().item() - self.top_db)  # stft.py:732 (batch-global, host sync)
        return out


def _mel_only_plan(melW, device, amin, ref, is_log):
    plan = engine.FrontendPlan.__new__(engine.FrontendPlan)
    import numpy as np
    lo, ln, off, val = engine.band_mel_c(melW)
    plan.mel_lo, plan.mel_len, plan.mel_off, plan.mel_val = (t.to(device) for t in (lo, ln, off, val))
    plan.n_mels = int(melW.shape[1])
    plan.F = int(melW.shape[0])
    plan.amin = float(amin)
    plan.db_offset = float(10.0 * np.log10(np.maximum(amin, ref)))
    plan.is_log = 1 if is_log else 0
    return plan
