"""Streaming post-processing oracle (merge / avg_merge / window schedule) against the reference's functions."""
import numpy as np
import pytest

import ref_import
import stream_oracle as so
from conftest import load_golden
from sed_b200 import streaming

CASES = [(fpw, ov, dur, nw) for fpw in (500, 496) for (ov, dur) in ((1, 5), (0.5, 6)) for nw in (1, 2, 3, 5, 6, 8)]


def golden_frames():
    """Same RandomState stream as oracle/gen_golden.py."""
    rng = np.random.RandomState(20260101)
    out = {}
    for (fpw, ov, dur, nw) in CASES:
        out[(fpw, ov, dur, nw)] = rng.rand(nw, fpw, 5).astype(np.float32)
    return out


def key(fpw, ov, dur, nw):
    return "f%d_ov%s_d%d_n%d" % (fpw, str(ov).replace(".", "p"), dur, nw)


def test_merge_avg_merge_match_reference_golden_bit_exact():
    g = load_golden("stream_merge.npz")
    frames = golden_frames()
    for case in CASES:
        fpw, ov, dur, nw = case
        got = so.merge_windows(frames[case], dur, ov)
        ref = g[key(*case)]
        assert got.shape == ref.shape == (1, (nw - 1) * int(100 * ov) + fpw, 5)
        assert np.array_equal(got, ref), case


def test_avg_merge_quirks_are_preserved():
    """a single 5 s window is divided by 2, 3, 4 in its middle blocks; first/last block untouched."""
    x = np.ones((1, 500, 1), np.float32)
    m = so.merge_windows(x, 5, 1)
    assert m[0, 0, 0] == 1 and m[0, 499, 0] == 1
    assert np.allclose(m[0, [150, 250, 350], 0], [1 / 2, 1 / 3, 1 / 4])
    frames = np.ones((8, 500, 1), np.float32)
    m = so.merge_windows(frames, 5, 1)  # true overlap counts 1,2,3,4,5,5,5,5,4,3,2,1
    assert np.allclose(m[0, ::100, 0], [1, 1, 1, 1, 1, 1, 1, 1, 4 / 5, 3 / 4, 2 / 3, 1])


def test_window_schedule_matches_reference_rule():
    # 60 s, 5 s windows, 1 s stride -> 56 windows; shorter than a window -> one window
    assert len(so.window_starts(60.0, 5)) == 56
    assert so.window_starts(3.2, 5) == [0]
    assert so.window_starts(5.0, 5) == [0]
    assert so.window_starts(6.0, 5) == [0, 1]
    assert so.window_starts(6.0, 5) == streaming.window_starts(6.0, 5)
    assert so.window_starts(59.99, 5) == streaming.window_starts(59.99, 5)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_oracle_matches_live_reference_utilities():
    ref_util, _ = ref_import.load_utilities()
    rng = np.random.RandomState(3)
    for nw, fpw, ov, dur in ((4, 500, 1, 5), (7, 496, 1, 5), (3, 600, 0.5, 6)):
        frames = rng.rand(nw, fpw, 25).astype(np.float32)
        merged = prev = None
        for k in range(nw):
            curr = frames[k:k + 1]
            if k == 1:
                merged = ref_util.merge(prev, curr, dur, 2, ov)
            elif k > 1:
                merged = ref_util.merge(merged, curr, dur, k + 1, ov)
            else:
                merged = curr.copy()
            prev = curr
        ref = ref_util.avg_merge(merged.copy(), dur, ov)
        assert np.array_equal(so.merge_windows(frames, dur, ov), ref)
    x = rng.rand(1000)
    assert np.array_equal(so.pad_truncate_sequence(x, 1500), ref_util.pad_truncate_sequence(x, 1500))
    q = (rng.rand(100) * 65535 - 32768).astype(np.int16)
    assert np.array_equal(ref_util.int16_to_float32(q), (q / 32767.).astype(np.float32))


def test_overlap_window_rule_matches_main_strong_loop():
    """`while end <= audio_duration` with end updated after each window (main_strong.py:789, 826-828)."""
    from sed_b200 import streaming
    assert streaming.overlap_window_counts([10.0], 5, 1) == [6]        # starts 0..5
    assert streaming.overlap_window_counts([10.0], 6, 0.5) == [9]      # starts 0, 0.5, ..., 4
    assert streaming.overlap_window_counts([10.0], 7, 1) == [4]
    assert streaming.overlap_window_counts([7.3, 4.0, 9.99], 5, 1) == [3, 1, 5]
