"""CPU oracle for the sound-event-detection inference hot path.

TEST INFRASTRUCTURE ONLY.  This is a stateless float32 restatement (torch CPU ops + explicit loops)
of the reference algorithm, each function citing the reference file:line it follows (paths relative
to /root/reference).  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it; the product package never does.

Parity pin: the reference has no automated tests or golden vectors of its own, so this
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, imported unchanged in the build container
(`oracle/ref_import.py`), two ways:
  * live, in `tests/test_oracle.py::test_oracle_matches_live_reference*` (skipped where
    /root/reference is absent, i.e. on the GPU box);
  * through committed fixtures `tests/golden/*.npz` written by `oracle/gen_golden.py` from the
    imported reference (those travel to the GPU box).

All functions take the reference's `state_dict` layout so the same checkpoint feeds
the reference, this oracle and the CUDA path.
"""

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # torch.nn.BatchNorm{1,2}d default, used by every BN in models.py


# --------------------------------------------------------------------------------------
# a1  STFT.__init__  (pytorch/stft.py:158-221, DFT matrix :21-25)
# --------------------------------------------------------------------------------------
def stft_conv_weights(n_fft, win_length=None, window="hann"):
    """conv_real/conv_imag.weight [F,1,n_fft] = Re/Im(W[:, :F] * window[:, None]).T as float32."""
    import scipy.signal

    if win_length is None:
        win_length = n_fft
    win = scipy.signal.get_window(window, win_length, fftbins=True)  # stft.py:192 (periodic hann)
    lpad = (n_fft - win_length) // 2  # stft.py:195 pad_center
    win = np.pad(win, (lpad, n_fft - win_length - lpad))
    (x, y) = np.meshgrid(np.arange(n_fft), np.arange(n_fft))  # stft.py:21-25
    omega = np.exp(-2 * np.pi * 1j / n_fft)
    W = np.power(omega, x * y)
    out_channels = n_fft // 2 + 1
    wr = torch.Tensor(np.real(W[:, 0:out_channels] * win[:, None]).T)[:, None, :]  # stft.py:207-208
    wi = torch.Tensor(np.imag(W[:, 0:out_channels] * win[:, None]).T)[:, None, :]  # stft.py:211-212
    return wr, wi


# --------------------------------------------------------------------------------------
# a2/a3  STFT.forward + Spectrogram.forward  (pytorch/stft.py:223-247, 651-670)
# --------------------------------------------------------------------------------------
def spectrogram(wave, conv_real_w, conv_imag_w, n_fft, hop, center=True, pad_mode="reflect", power=2.0):
    """wave [B,L] -> power spectrogram [B,1,T,F]."""
    x = wave[:, None, :]
    if center:
        x = F.pad(x, pad=(n_fft // 2, n_fft // 2), mode=pad_mode)  # stft.py:236-237
    real = F.conv1d(x, conv_real_w, stride=hop)  # stft.py:239
    imag = F.conv1d(x, conv_imag_w, stride=hop)  # stft.py:240
    real = real[:, None, :, :].transpose(2, 3)  # stft.py:243
    imag = imag[:, None, :, :].transpose(2, 3)
    spec = real ** 2 + imag ** 2  # stft.py:663
    if power != 2.0:
        spec = spec ** (power / 2.0)  # stft.py:665-668
    return spec


# --------------------------------------------------------------------------------------
# a5  LogmelFilterBank.forward + power_to_db  (pytorch/stft.py:698-734)
# --------------------------------------------------------------------------------------
def logmel(spec, melW, is_log=True, ref=1.0, amin=1e-10, top_db=None):
    mel = torch.matmul(spec, melW)  # stft.py:709
    if not is_log:
        return mel
    log_spec = 10.0 * torch.log10(torch.clamp(mel, min=amin, max=np.inf))  # stft.py:726
    log_spec = log_spec - 10.0 * np.log10(np.maximum(amin, ref))  # stft.py:727
    if top_db is not None:
        if top_db < 0:
            raise ValueError("top_db must be non-negative")  # stft.py:730-731
        log_spec = torch.clamp(log_spec, min=log_spec.max().item() - top_db, max=np.inf)  # stft.py:732
    return log_spec


def bn_eval(x, sd, prefix, channel_dim=1):
    """Eval-mode BatchNorm: y = (x - mean) / sqrt(var + eps) * weight + bias."""
    shape = [1] * x.dim()
    shape[channel_dim] = -1
    mean = sd[prefix + ".running_mean"].view(shape)
    var = sd[prefix + ".running_var"].view(shape)
    w = sd[prefix + ".weight"].view(shape)
    b = sd[prefix + ".bias"].view(shape)
    return (x - mean) / torch.sqrt(var + BN_EPS) * w + b


# --------------------------------------------------------------------------------------
# a6  bn0 over the mel axis  (pytorch/models.py:642-644)
# --------------------------------------------------------------------------------------
def bn0(x, sd):
    """x [B,1,T,64]: BatchNorm2d(64) applied with mel as the channel axis."""
    return bn_eval(x, sd, "bn0", channel_dim=3)


# --------------------------------------------------------------------------------------
# a7  ConvBlock.forward  (pytorch/models.py:98-141)
# --------------------------------------------------------------------------------------
def conv_block(x, sd, prefix, pool_size):
    x = F.conv2d(x, sd[prefix + ".conv1.weight"], padding=1)  # models.py:103-106 (bias=False)
    x = torch.relu(bn_eval(x, sd, prefix + ".bn1"))  # models.py:128
    x = F.conv2d(x, sd[prefix + ".conv2.weight"], padding=1)
    x = torch.relu(bn_eval(x, sd, prefix + ".bn2"))  # models.py:129
    return F.avg_pool2d(x, kernel_size=pool_size)  # models.py:133 (floors odd dims)


def conv_stack(x, sd, return_stages=False):
    """x [B,1,T,64] (after bn0) -> [B,512,T//8,8]  (models.py:663-666)."""
    stages = []
    for i, pool in ((1, (2, 2)), (2, (2, 2)), (3, (2, 2)), (4, (1, 1))):
        x = conv_block(x, sd, "conv_block%d" % i, pool)
        stages.append(x)
    return (x, stages) if return_stages else x


# --------------------------------------------------------------------------------------
# a9  nn.GRU(512, 256, bidirectional, batch_first)  (pytorch/models.py:614-615, 670)
#     PyTorch gate order r, z, n; h0 = 0.
# --------------------------------------------------------------------------------------
def _gru_direction(x, w_ih, w_hh, b_ih, b_hh, reverse):
    B, T, _ = x.shape
    Hd = w_hh.shape[1]
    gi = x @ w_ih.t() + b_ih  # [B,T,3H]
    h = x.new_zeros(B, Hd)
    out = x.new_zeros(B, T, Hd)
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        gh = h @ w_hh.t() + b_hh
        i_r, i_z, i_n = gi[:, t].chunk(3, dim=1)
        h_r, h_z, h_n = gh.chunk(3, dim=1)
        r = torch.sigmoid(i_r + h_r)
        z = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        h = (1.0 - z) * n + z * h
        out[:, t] = h
    return out


def bigru(x, sd, prefix="gru"):
    """x [B,T,512] -> [B,T,512] = [forward | backward]."""
    fwd = _gru_direction(x, sd[prefix + ".weight_ih_l0"], sd[prefix + ".weight_hh_l0"],
                         sd[prefix + ".bias_ih_l0"], sd[prefix + ".bias_hh_l0"], False)
    bwd = _gru_direction(x, sd[prefix + ".weight_ih_l0_reverse"], sd[prefix + ".weight_hh_l0_reverse"],
                         sd[prefix + ".bias_ih_l0_reverse"], sd[prefix + ".bias_hh_l0_reverse"], True)
    return torch.cat([fwd, bwd], dim=2)


# --------------------------------------------------------------------------------------
# a9' MultiHead + ScaledDotProductAttention  (pytorch/models.py:799-877)
#     8 heads, d_k = d_v = 64, temperature sqrt(64); relu(fc(concat)); no residual, no LayerNorm.
# --------------------------------------------------------------------------------------
def multihead(x, sd, prefix="multihead", n_head=8, d_k=64, d_v=64):
    B, T, _ = x.shape
    q = F.linear(x, sd[prefix + ".w_qs.weight"], sd[prefix + ".w_qs.bias"]).view(B, T, n_head, d_k)
    k = F.linear(x, sd[prefix + ".w_ks.weight"], sd[prefix + ".w_ks.bias"]).view(B, T, n_head, d_k)
    v = F.linear(x, sd[prefix + ".w_vs.weight"], sd[prefix + ".w_vs.bias"]).view(B, T, n_head, d_v)
    q = q.permute(2, 0, 1, 3).reshape(-1, T, d_k)  # models.py:867-869 (head-major)
    k = k.permute(2, 0, 1, 3).reshape(-1, T, d_k)
    v = v.permute(2, 0, 1, 3).reshape(-1, T, d_v)
    attn = torch.bmm(q, k.transpose(1, 2)) / float(np.power(d_k, 0.5))  # models.py:810-811, 843
    attn = torch.softmax(attn, dim=2)  # models.py:816
    out = torch.bmm(attn, v)  # models.py:818
    out = out.view(n_head, B, T, d_v).permute(1, 2, 0, 3).reshape(B, T, n_head * d_v)  # :874-875
    return torch.relu(F.linear(out, sd[prefix + ".fc.weight"], sd[prefix + ".fc.bias"]))  # :876


# --------------------------------------------------------------------------------------
# a10 AttBlock.forward  (pytorch/models.py:161-169), activation='sigmoid', temperature=1
# --------------------------------------------------------------------------------------
def att_block(x, sd, prefix="att_block", temperature=1.0):
    """x [B,512,T'] -> clip [B,25], norm_att [B,25,T'], cla [B,25,T']."""
    tmp = F.conv1d(x, sd[prefix + ".att.weight"], sd[prefix + ".att.bias"])  # models.py:163
    tmp = torch.clamp(tmp, -10, 10)  # models.py:164
    att = torch.exp(tmp / temperature) + 1e-6  # models.py:165
    norm_att = att / torch.sum(att, dim=2)[:, :, None]  # models.py:166
    cla = torch.sigmoid(F.conv1d(x, sd[prefix + ".cla.weight"], sd[prefix + ".cla.bias"]))  # :167
    clip = torch.sum(norm_att * cla, dim=2)  # models.py:168
    return clip, norm_att, cla


# --------------------------------------------------------------------------------------
# a11 interpolate / pad_framewise_output / roundup  (pytorch/models.py:62-95)
# --------------------------------------------------------------------------------------
def roundup(x):
    return x if x % 100 == 0 else x + 100 - x % 100


def interpolate(x, ratio):
    (B, T, C) = x.shape
    return x[:, :, None, :].repeat(1, 1, ratio, 1).reshape(B, T * ratio, C)


def pad_framewise_output(fw, frames_num):
    pad = fw[:, -1:, :].repeat(1, frames_num - fw.shape[1], 1)
    return torch.cat((fw, pad), dim=1)


# --------------------------------------------------------------------------------------
# whole-model forward (models.py:625-688 GRU, :1029-1077 Transformer), eval mode only
# --------------------------------------------------------------------------------------
# sibling heads on the same trunk: model_type -> (temporal, head)
SIBLING_PLANS = {
    "Cnn_9layers_FrameMax": (None, "max"),               # models.py:213-295
    "Cnn_9layers_FrameAvg": (None, "avg"),               # models.py:298-380
    "Cnn_9layers_FrameAtt": (None, "att"),               # models.py:383-463
    "Cnn_9layers_Gru_FrameAvg": ("gru", "avg"),          # models.py:466-561
    "Cnn_9layers_Transformer_FrameAvg": ("mha", "avg"),  # models.py:880-978
}


def sibling_forward(sd, wave, model_type, n_fft, hop):
    """The five alternative heads of the Cnn_9layers trunk; same output dict."""
    temporal, head = SIBLING_PLANS[model_type]
    with torch.no_grad():
        spec = spectrogram(wave, sd["spectrogram_extractor.stft.conv_real.weight"],
                           sd["spectrogram_extractor.stft.conv_imag.weight"], n_fft, hop)
        x = conv_stack(bn0(logmel(spec, sd["logmel_extractor.melW"], top_db=None), sd), sd)
        x = torch.mean(x, dim=3)  # [B,512,T']   models.py:275, 360, 445, 543, 959
        if temporal == "gru":
            x = bigru(x.transpose(1, 2), sd).transpose(1, 2)  # models.py:544-546
        elif temporal == "mha":
            x = multihead(x.transpose(1, 2), sd).transpose(1, 2)  # models.py:960-962
        if head == "att":
            clip, _, cla = att_block(x, sd)  # models.py:448
            fw = interpolate(cla.transpose(1, 2), 8)  # models.py:452-453 (no padding)
            return {"framewise_output": fw, "clipwise_output": clip, "embedding": cla}
        fw = torch.sigmoid(F.linear(x.transpose(1, 2), sd["fc.weight"], sd["fc.bias"]))  # models.py:280, 365, 551
        fw = interpolate(fw, 8)
        clip = torch.max(fw, dim=1)[0] if head == "max" else torch.mean(fw, dim=1)  # models.py:284 / 369
        return {"framewise_output": fw, "clipwise_output": clip, "embedding": x}


def model_forward(sd, wave, model_type, n_fft, hop, return_stages=False):
    """sd: reference-layout state_dict (float32 CPU tensors); wave [B,L] float32.

    model_type in {'Cnn_9layers_Gru_FrameAtt', 'Cnn_9layers_Transformer_FrameAtt'} (with stages) or one of
    SIBLING_PLANS (outputs only).
    """
    if model_type in SIBLING_PLANS and not return_stages:
        return sibling_forward(sd, wave, model_type, n_fft, hop)
    stages = {}
    with torch.no_grad():
        spec = spectrogram(wave, sd["spectrogram_extractor.stft.conv_real.weight"],
                           sd["spectrogram_extractor.stft.conv_imag.weight"], n_fft, hop)
        lm = logmel(spec, sd["logmel_extractor.melW"], top_db=None)  # models.py:575 top_db=None
        stages["logmel"] = lm
        x = bn0(lm, sd)
        stages["bn0"] = x
        x, conv_stages = conv_stack(x, sd, return_stages=True)
        for i, s in enumerate(conv_stages):
            stages["conv_block%d" % (i + 1)] = s
        x = torch.mean(x, dim=3).transpose(1, 2)  # models.py:668-669 -> [B,T',512]
        stages["feat"] = x
        if model_type == "Cnn_9layers_Gru_FrameAtt":
            x = bigru(x, sd)  # models.py:670
        elif model_type == "Cnn_9layers_Transformer_FrameAtt":
            x = multihead(x, sd)  # models.py:1061
        else:
            raise ValueError("oracle does not cover model_type=%r" % (model_type,))
        stages["temporal"] = x
        x = x.transpose(1, 2)
        clip, norm_att, cla = att_block(x, sd)
        stages["norm_att"] = norm_att
        fw = interpolate(cla.transpose(1, 2), 8)  # models.py:677-678 / :1069-1070
        if model_type == "Cnn_9layers_Gru_FrameAtt":
            if fw.size()[1] != 1000:  # models.py:680-681
                fw = pad_framewise_output(fw, roundup(fw.size()[1]))
            emb = cla  # models.py:686
        else:
            emb = x  # models.py:1061-1063, 1075
        out = {"framewise_output": fw, "clipwise_output": clip, "embedding": emb}
    return (out, stages) if return_stages else out


# --------------------------------------------------------------------------------------
# Independent float64 cross-check of the front-end (method of stft.py:925-1177 debug(): compare
# against numpy.fft), SURVEY.md 8c "independent cross-check".
# --------------------------------------------------------------------------------------
def logmel_float64_fft(wave, n_fft, hop, melW, window=None, amin=1e-10):
    import scipy.signal

    w = np.asarray(wave, dtype=np.float64)
    if window is None:
        window = scipy.signal.get_window("hann", n_fft, fftbins=True)
    pad = n_fft // 2
    wp = np.pad(w, ((0, 0), (pad, pad)), mode="reflect")
    T = w.shape[1] // hop + 1
    idx = np.arange(T)[:, None] * hop + np.arange(n_fft)[None, :]
    frames = wp[:, idx] * window[None, None, :]
    spec = np.abs(np.fft.rfft(frames, axis=-1)) ** 2
    mel = spec @ np.asarray(melW, dtype=np.float64)
    return 10.0 * np.log10(np.maximum(mel, amin))


# --------------------------------------------------------------------------------------
# thresholded decisions for the 99.9 % agreement metric (north_star): framewise > sed_high_threshold
# using the shipped opt_thresholds pickles (dict keys per utils/optimize_thresholds.py).
# --------------------------------------------------------------------------------------
def threshold_decisions(framewise, thresholds):
    thr = np.asarray(thresholds, dtype=np.float64).reshape(1, 1, -1)
    return np.asarray(framewise, dtype=np.float64) > thr


def flops_per_clip(T, n_fft, model_type="Cnn_9layers_Gru_FrameAtt"):
    """Algorithmic MACs*2 per clip. Returns dict of GFLOP."""
    Fb = n_fft // 2 + 1
    fe = 2 * (T * n_fft * Fb * 2 + T * Fb * 64)
    H, W = T, 64
    conv = 0
    cin = 1
    for i, cout in enumerate((64, 128, 256, 512)):
        conv += 2 * H * W * 9 * cin * cout
        conv += 2 * H * W * 9 * cout * cout
        cin = cout
        if i < 3:
            H, W = H // 2, W // 2
    Tp = H
    if model_type == "Cnn_9layers_Gru_FrameAtt":
        temporal = 2 * Tp * 2 * (768 * 512 + 768 * 256)
    else:
        temporal = 2 * (Tp * 3 * 512 * 512 + 8 * 2 * Tp * Tp * 64 + Tp * 512 * 512)
    head = 2 * Tp * 512 * 50
    return {"frontend": fe / 1e9, "conv": conv / 1e9, "temporal": temporal / 1e9, "head": head / 1e9,
            "total": (fe + conv + temporal + head) / 1e9}
