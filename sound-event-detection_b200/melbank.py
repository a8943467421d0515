"""Constructor-time constants of the audio front-end (window, DFT kernels, mel filterbank).

The reference builds these with librosa/numpy at `pytorch/stft.py:192-212` (periodic Hann window,
windowed DFT matrix as two Conv1d weights) and `pytorch/stft.py:688` (`librosa.filters.mel`, Slaney
scale, area-normalised).  They are ordinary parameters in the reference `state_dict`, so a loaded
checkpoint overrides them and the CUDA path always consumes the loaded tensors.
"""

import numpy as np
import torch


def hann_periodic(n):
    # scipy.signal.get_window('hann', n, fftbins=True) == 0.5 - 0.5 cos(2 pi k / n)
    k = np.arange(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)


def get_window(window, win_length):
    if window == "hann":
        return hann_periodic(win_length)
    import scipy.signal
    return scipy.signal.get_window(window, win_length, fftbins=True)


def windowed_dft_kernels(n_fft, win_length, window):
    """(conv_real.weight, conv_imag.weight), each float32 [n_fft//2+1, 1, n_fft] (stft.py:207-212)."""
    win = get_window(window, win_length)
    lpad = (n_fft - win_length) // 2
    win = np.pad(win, (lpad, n_fft - win_length - lpad))
    F = n_fft // 2 + 1
    # W[n, k] = omega ** (n * k), omega = exp(-2 pi i / N): same expression as stft.py:21-25 so the
    # float32 kernels are bit-identical to a freshly constructed reference module
    n_idx, k_idx = np.meshgrid(np.arange(n_fft), np.arange(F), indexing="ij")
    W = np.power(np.exp(-2 * np.pi * 1j / n_fft), n_idx * k_idx) * win[:, None]
    wr = np.ascontiguousarray(np.real(W).T).astype(np.float32)
    wi = np.ascontiguousarray(np.imag(W).T).astype(np.float32)
    return torch.from_numpy(wr)[:, None, :].contiguous(), torch.from_numpy(wi)[:, None, :].contiguous()


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f * 3.0 / 200.0
    log = 15.0 + np.log(np.maximum(f, 1e-30) / 1000.0) * (27.0 / np.log(6.4))
    return np.where(f >= 1000.0, log, lin)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    lin = m * 200.0 / 3.0
    log = 1000.0 * np.exp((m - 15.0) * (np.log(6.4) / 27.0))
    return np.where(m >= 15.0, log, lin)


def mel_filterbank(sr, n_fft, n_mels, fmin, fmax):
    """float32 [n_fft//2+1, n_mels] == librosa.filters.mel(...).T (Slaney, area-normalised)."""
    if fmax is None:
        fmax = sr // 2  # stft.py:685-686
    F = n_fft // 2 + 1
    freqs = np.linspace(0.0, float(sr) / 2, F)
    edges = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    lo, ce, hi = edges[:-2, None], edges[1:-1, None], edges[2:, None]
    up = (freqs[None, :] - lo) / (ce - lo)
    down = (hi - freqs[None, :]) / (hi - ce)
    w = np.maximum(0.0, np.minimum(up, down)).astype(np.float32)
    w *= (2.0 / (hi - lo)).astype(np.float64)
    return torch.from_numpy(np.ascontiguousarray(w.T.astype(np.float32)))
