"""Event extraction (utils/vad.py activity_detection): oracle vs reference fixtures (CPU) and device kernel vs both (GPU)."""
import json
import os

import numpy as np
import pytest
import torch

import ref_import
import stream_oracle as so
from conftest import GOLD, load_golden


def cases():
    with open(os.path.join(GOLD, "events.json")) as f:
        meta = json.load(f)
    xs = load_golden("events_x.npz")
    return meta, [dict(c, x=xs["x%d" % i]) for i, c in enumerate(meta["cases"])]


def test_find_bgn_fin_pairs_quirk():
    meta, _ = cases()
    assert meta["example"] == [[3, 6], [10, 11], [21, 20]]
    assert so.find_bgn_fin_pairs([3, 4, 5, 9, 10, 20]) == meta["example"]


def test_oracle_activity_detection_matches_reference_golden():
    _, cs = cases()
    assert len(cs) >= 30
    for c in cs:
        got = so.activity_detection(c["x"], np.float64(c["hi"]), None if c["lo"] is None else np.float64(c["lo"]),
                                    c["n_smooth"], c["n_salt"])
        assert got == c["pairs"], c["seed_case"]


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_oracle_matches_live_reference_vad():
    _, ref_vad = ref_import.load_utilities()
    rng = np.random.RandomState(5)
    for _ in range(30):
        x = np.convolve(rng.rand(540), np.ones(7) / 7, mode="same")[:500].astype(np.float32)
        hi, lo = rng.uniform(0.4, 0.6), rng.uniform(0.2, 0.4)
        try:
            ref = ref_vad.activity_detection(x, np.float64(hi), np.float64(lo), 10, 10)
        except IndexError:
            continue
        assert so.activity_detection(x, np.float64(hi), np.float64(lo), 10, 10) == [list(p) for p in ref]


@pytest.mark.gpu
def test_device_events_match_reference_golden():
    from sed_b200 import engine
    _, cs = cases()
    for c in cs:
        x = torch.from_numpy(c["x"]).cuda()[None, :, None].contiguous()  # [1, frames, 1]
        ev, cnt = engine.extract_events(x, [c["hi"]], None if c["lo"] is None else [c["lo"]], [c["n_smooth"]],
                                        [c["n_salt"]], max_events=8)  # small buffer: exercises the rerun path
        n = int(cnt[0, 0])
        assert ev[0, 0, :n].cpu().tolist() == c["pairs"], c["seed_case"]


@pytest.mark.gpu
def test_device_events_with_shipped_thresholds(thresholds):
    """Whole tensor at once with the shipped per-class thresholds (negative low thresholds included)."""
    from sed_b200 import engine, streaming
    p = thresholds["Cnn_9layers_Gru_FrameAtt/best_logmel_16k.sed.valid.pkl"]
    rng = np.random.RandomState(9)
    fw = np.stack([np.convolve(rng.rand(1040), np.ones(15) / 15, mode="same")[:1000] for _ in range(3 * 25)])
    fw = ((fw - fw.min()) / (fw.max() - fw.min())).astype(np.float32).reshape(3, 25, 1000).transpose(0, 2, 1).copy()
    ev, cnt = engine.extract_events(torch.from_numpy(fw).cuda(), p["sed_high_threshold"], p["sed_low_threshold"],
                                    p["n_smooth"], p["n_salt"])
    for n in range(3):
        for k in range(25):
            ref = so.activity_detection(fw[n, :, k], np.float64(p["sed_high_threshold"][k]),
                                        np.float64(p["sed_low_threshold"][k]), p["n_smooth"], p["n_salt"])
            assert ev[n, k, :int(cnt[n, k])].cpu().tolist() == ref, (n, k)
    params = {"sed_high_threshold": p["sed_high_threshold"], "sed_low_threshold": p["sed_low_threshold"],
              "n_smooth": p["n_smooth"], "n_salt": p["n_salt"]}
    lst = streaming.frame_prediction_to_event_prediction(torch.from_numpy(fw).cuda(), params, 100, ["a", "b", "c"])
    assert len(lst) == int(cnt.sum()) and set(e["filename"] for e in lst) <= {"a", "b", "c"}
    assert all(e["event_label"] in streaming.LABELS for e in lst)
