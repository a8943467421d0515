"""Generate tests/golden/* from the UNMODIFIED reference imported in the build container.

TEST INFRASTRUCTURE.  Run here (where /root/reference exists):  python oracle/gen_golden.py
The fixtures travel to the GPU box; the reference tree does not.  Everything is seeded; inputs are stored
as int16 (they are int16-quantised by construction, `utils/utilities.py:78-79` semantics: x = q / 32767).
"""
import glob
import json
import os
import pickle
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ref_import  # noqa: E402
from sed_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
PRESET_ARGS = {sr: (sr,) + synth.PRESETS[sr][:2] + (64,) + synth.PRESETS[sr][2:] + (25, "logmel") for sr in synth.PRESETS}


def frontend_signals(sr, seconds=1.0):
    """SURVEY.md 8(c) list: uniform(-1,1), 0.1*N(0,1) clipped, int16 tone+noise, zeros, impulse, full-scale square."""
    L = int(sr * seconds) + 1  # not a multiple of hop
    g = torch.Generator().manual_seed(4242 + sr)
    t = torch.arange(L, dtype=torch.float64) / sr
    sig = [
        torch.rand(L, generator=g, dtype=torch.float64) * 2 - 1,
        torch.clamp(0.1 * torch.randn(L, generator=g, dtype=torch.float64), -1, 1),
        0.5 * torch.sin(2 * np.pi * 0.11 * sr * t) + 0.003 * torch.randn(L, generator=g, dtype=torch.float64),
        torch.zeros(L, dtype=torch.float64),
        torch.zeros(L, dtype=torch.float64),
        torch.sign(torch.sin(2 * np.pi * 97.0 * t)),
    ]
    sig[4][L // 3] = 1.0
    q = torch.round(torch.stack(sig) * 32767.0).clamp(-32767, 32767).to(torch.int16)
    return q


def sibling_goldens(rm):
    """The five sibling heads of the Cnn_9layers trunk (models.py:213-561, 880-978) at 16 kHz, one file."""
    res = {}
    L = 24000
    wave = torch.cat([synth.synthetic_waveform(2, L, seed=41, kind="events"), synth.synthetic_waveform(1, L, seed=42),
                      torch.zeros(1, L)])
    q = torch.round(wave * 32767.0).to(torch.int16)
    wave = q.float() / 32767.0
    res["wave_i16"] = q.numpy()
    for mt in synth.SIBLING_TYPES:
        classes = 10 if mt == "Cnn_9layers_FrameMax" else 25  # fc heads honour classes_num (models.py:247)
        sd = synth.synthetic_state_dict(mt, 16000, classes_num=classes)
        args = PRESET_ARGS[16000][:6] + (classes,)
        if mt in ("Cnn_9layers_Gru_FrameAvg", "Cnn_9layers_Transformer_FrameAvg"):
            args = args + ("logmel",)
        model = getattr(rm, mt)(*args).eval()
        model.load_state_dict(sd, strict=True)
        with torch.no_grad():
            out = model(wave)
        for k in ("framewise_output", "clipwise_output", "embedding"):
            res["%s.%s" % (mt, k)] = out[k].numpy()
    np.savez_compressed(os.path.join(GOLD, "model_siblings_16k.npz"), **res)


def main():
    os.makedirs(GOLD, exist_ok=True)
    rs, rm = ref_import.load()
    torch.set_num_threads(8)
    sibling_goldens(rm)
    if "--only-siblings" in sys.argv:
        return
    meta = {"torch": torch.__version__, "state_dict_keys": {}, "ckpt_fingerprint": {}}

    # ---- front-end goldens (reference Spectrogram + LogmelFilterBank, top_db=None as in the models) ----
    for sr in (8000, 16000, 32000):
        n_fft, hop, fmin, fmax = synth.PRESETS[sr]
        q = frontend_signals(sr)
        wave = q.float() / 32767.0
        spec_mod = rs.Spectrogram(n_fft=n_fft, hop_length=hop, win_length=n_fft, window='hann', center=True,
                                  pad_mode='reflect', freeze_parameters=True)
        mel_mod = rs.LogmelFilterBank(sr=sr, n_fft=n_fft, n_mels=64, fmin=fmin, fmax=fmax, ref=1.0, amin=1e-10,
                                      top_db=None, freeze_parameters=True)
        with torch.no_grad():
            spec = spec_mod(wave)
            lm = mel_mod(spec)
            mel_db80 = rs.LogmelFilterBank(sr=sr, n_fft=n_fft, n_mels=64, fmin=fmin, fmax=fmax, top_db=80.0)
            lm80 = mel_db80(spec)
        np.savez_compressed(os.path.join(GOLD, "frontend_%dk.npz" % (sr // 1000)), wave_i16=q.numpy(),
                            logmel=lm[:, 0].numpy(), logmel_top80=lm80[:, 0].numpy(),
                            spec_rows=spec[:, 0, ::25].numpy(), melW=mel_mod.melW.detach().numpy())

    # ---- whole-model goldens ----
    for mt in synth.MODEL_TYPES:
        for sr in (16000, 8000, 32000):
            sd = synth.synthetic_state_dict(mt, sr)
            model = getattr(rm, mt)(*PRESET_ARGS[sr]).eval()
            model.load_state_dict(sd, strict=True)
            if sr == 16000:
                meta["state_dict_keys"][mt] = {k: list(v.shape) for k, v in model.state_dict().items()}
            meta["ckpt_fingerprint"]["%s_%d" % (mt, sr)] = {
                "bn_var_sum": float(sum(v.double().sum() for k, v in sd.items() if k.endswith("running_var"))),
                "conv_w_abs_sum": float(sum(v.double().abs().sum() for k, v in sd.items() if "conv_block" in k and k.endswith("weight") and v.dim() == 4)),
            }
            seconds = 2.0 if sr == 16000 else 1.5
            L = int(sr * seconds)
            wave = torch.cat([synth.synthetic_waveform(3, L, seed=31, kind="events", sample_rate=sr),
                              synth.synthetic_waveform(1, L, seed=32, kind="noise"),
                              0.01 * synth.synthetic_waveform(1, L, seed=33, kind="noise"), torch.zeros(1, L)])
            q = torch.round(wave * 32767.0).to(torch.int16)
            wave = q.float() / 32767.0
            with torch.no_grad():
                out = model(wave)
            np.savez_compressed(os.path.join(GOLD, "model_%s_%dk.npz" % ("gru" if "Gru" in mt else "transformer", sr // 1000)),
                                wave_i16=q.numpy(), framewise_output=out["framewise_output"].numpy(),
                                clipwise_output=out["clipwise_output"].numpy(), embedding=out["embedding"].numpy())
        # one full-size case (10 s and 5 s, 16 kHz); inputs regenerated from the seed on the test side
        sd = synth.synthetic_state_dict(mt, 16000)
        model = getattr(rm, mt)(*PRESET_ARGS[16000]).eval()
        model.load_state_dict(sd, strict=True)
        res = {}
        for name, L in (("10s", 160000), ("5s", 80000)):
            wave = torch.cat([synth.synthetic_waveform(2, L, seed=77, kind="events"), synth.synthetic_waveform(1, L, seed=78)])
            with torch.no_grad():
                out = model(wave)
            res["framewise_" + name] = out["framewise_output"].numpy()
            res["clipwise_" + name] = out["clipwise_output"].numpy()
            res["wave_checksum_" + name] = np.array([wave.double().sum().item(), wave.double().abs().sum().item()])
        np.savez_compressed(os.path.join(GOLD, "model_%s_full.npz" % ("gru" if "Gru" in mt else "transformer")), **res)

    # ---- streaming post-processing goldens: reference merge / avg_merge (utils/utilities.py:405-446) ----
    ref_util, ref_vad = ref_import.load_utilities()
    rng = np.random.RandomState(20260101)
    stream = {}
    for fpw in (500, 496):
        for (ov, dur) in ((1, 5), (0.5, 6)):
            for nw in (1, 2, 3, 5, 6, 8):
                frames = rng.rand(nw, fpw, 5).astype(np.float32)  # regenerated from the seed by the tests
                merged = prev = None
                for k in range(nw):
                    curr = frames[k:k + 1]
                    if k + 1 == 2:
                        merged = ref_util.merge(prev, curr, dur, k + 1, ov)
                    elif k + 1 > 2:
                        merged = ref_util.merge(merged, curr, dur, k + 1, ov)
                    else:
                        merged = curr.copy()
                    prev = curr
                merged = ref_util.avg_merge(merged.copy(), dur, ov)
                key = "f%d_ov%s_d%d_n%d" % (fpw, str(ov).replace(".", "p"), dur, nw)
                stream[key] = merged
    np.savez_compressed(os.path.join(GOLD, "stream_merge.npz"), **stream)

    # ---- event extraction goldens: reference vad.activity_detection on seeded smooth random sequences ----
    ev = {"cases": []}
    ev_x = {}
    rng = np.random.RandomState(77)
    for case in range(40):
        n_frames = int(rng.choice([100, 500, 1000, 1237]))
        base = rng.rand(n_frames + 40)
        width = int(rng.choice([1, 3, 9, 25]))
        x = np.convolve(base, np.ones(width) / width, mode="same")[:n_frames]
        x = ((x - x.min()) / (x.max() - x.min() + 1e-9)).astype(np.float32)
        if case % 7 == 0:
            x[-1] = 1.0  # event touching the end
        if case % 5 == 0:
            x[0] = 1.0
        hi = float(rng.uniform(0.3, 0.8)); lo = float(rng.uniform(-0.2, hi))
        ns = int(rng.choice([0, 1, 3, 10])); nsalt = int(rng.choice([0, 1, 5, 10]))
        use_low = bool(case % 4 != 3)
        try:
            pairs = ref_vad.activity_detection(x, np.float64(hi), np.float64(lo) if use_low else None, ns, nsalt)
        except IndexError:
            continue  # the reference itself crashes when a non-first event starts on the last frame
        ev_x["x%d" % len(ev["cases"])] = x
        ev["cases"].append({"seed_case": case, "hi": hi, "lo": lo if use_low else None,
                            "n_smooth": ns, "n_salt": nsalt, "pairs": [[int(a), int(b)] for a, b in pairs]})
    ev["example"] = ref_vad.find_bgn_fin_pairs([3, 4, 5, 9, 10, 20])
    with open(os.path.join(GOLD, "events.json"), "w") as fh:
        json.dump(ev, fh)
    np.savez_compressed(os.path.join(GOLD, "events_x.npz"), **ev_x)

    # ---- shipped thresholds (opt_thresholds/**/best_*.pkl) as a JSON fixture ----
    thr = {}
    base = os.path.join(ref_import.REF_ROOT, "opt_thresholds")
    for f in sorted(glob.glob(os.path.join(base, "**", "best_*.pkl"), recursive=True)):
        d = pickle.load(open(f, "rb"))
        mt = f.split("model_type=")[1].split("/")[0]
        key = "%s/%s" % (mt, os.path.basename(f))
        thr[key] = {k: (np.asarray(v).astype(float).tolist() if hasattr(v, "__len__") else v) for k, v in d.items()}
    with open(os.path.join(GOLD, "opt_thresholds.json"), "w") as fh:
        json.dump(thr, fh, indent=0)
    with open(os.path.join(GOLD, "meta.json"), "w") as fh:
        json.dump(meta, fh, indent=0)
    print("wrote", sorted(os.listdir(GOLD)))


if __name__ == "__main__":
    main()
