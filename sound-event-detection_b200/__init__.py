"""B200-native sound-event-detection inference hot path (drop-in for the reference's
`pytorch/stft.py` Spectrogram/LogmelFilterBank and `pytorch/models.py`
Cnn_9layers_Gru_FrameAtt / Cnn_9layers_Transformer_FrameAtt).

The directory name carries a hyphen (repo layout contract); import it as `sed_b200`.
"""
__all__ = ["stft", "models", "synth", "engine", "capi", "dist", "streaming", "pytorch_utils"]
