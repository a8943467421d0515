"""Minimal stand-in for the `librosa` calls the reference makes at constructor time.

TEST INFRASTRUCTURE ONLY (used by oracle/ref_import.py to import /root/reference unchanged in the
build container, where librosa is not installed).  Call sites in the reference:
  pytorch/stft.py:192  librosa.filters.get_window(window, win_length, fftbins=True)
  pytorch/stft.py:195  librosa.util.pad_center(fft_window, n_fft)
  pytorch/stft.py:688  librosa.filters.mel(sr=, n_fft=, n_mels=, fmin=, fmax=)
  pytorch/stft.py:730  librosa.util.exceptions.ParameterError
The arithmetic is the published librosa algorithm (Slaney mel scale, area normalisation); it lives in
oracle/melbank.py so the product package and this shim share one restatement.
"""
from . import filters, util  # noqa: F401
