import os, sys
import scipy.signal
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from melbank import slaney_mel_filterbank  # noqa: E402


def get_window(window, Nx, fftbins=True):
    return scipy.signal.get_window(window, Nx, fftbins=fftbins)


def mel(sr=22050, n_fft=2048, n_mels=128, fmin=0.0, fmax=None, **kw):
    return slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
