"""Host-side packing logic (no GPU): BN folding, banded mel matrix, DFT-structure check, GRU weight order."""
import numpy as np
import pytest
import torch

import sed_oracle as so
from conftest import synthetic_sd
from sed_b200 import engine, melbank, synth
from sed_b200 import dist as sdist


def test_fold_bn_equals_eval_batchnorm():
    sd = synthetic_sd("Cnn_9layers_Gru_FrameAtt")
    x = torch.randn(2, 64, 5, 7)
    s, b = engine.fold_bn(sd, "conv_block1.bn2")
    ref = so.bn_eval(x, sd, "conv_block1.bn2")
    assert torch.allclose(x * s[None, :, None, None] + b[None, :, None, None], ref, atol=2e-6, rtol=1e-5)


@pytest.mark.parametrize("sr", [8000, 16000, 32000])
def test_band_mel_reconstructs_matrix(sr):
    n_fft, hop, fmin, fmax = synth.PRESETS[sr]
    W = melbank.mel_filterbank(sr, n_fft, 64, fmin, fmax)
    lo, ln, off, val = engine.band_mel(W)
    R = np.zeros(W.shape, np.float32)
    for m in range(64):
        R[lo[m]:lo[m] + ln[m], m] = val[off[m]:off[m] + ln[m]].numpy()
    assert np.array_equal(R, W.numpy())
    assert int(ln.sum()) < 0.06 * W.numel()  # triangular filters: ~2.6 % dense


def test_band_mel_handles_dense_and_empty_columns():
    W = torch.rand(17, 5)
    W[:, 2] = 0
    lo, ln, off, val = engine.band_mel(W)
    assert ln[2] == 0 and ln[0] == 17


def test_windowed_dft_check_accepts_reference_and_rejects_others():
    wr, wi = melbank.windowed_dft_kernels(512, 512, "hann")
    win = engine.check_windowed_dft(wr, wi, 512)
    assert torch.allclose(win, torch.from_numpy(melbank.hann_periodic(512)).float(), atol=1e-7)
    with pytest.raises(NotImplementedError):
        engine.check_windowed_dft(wr + 0.01 * torch.randn_like(wr), wi, 512)
    # shorter window padded to n_fft is still window x DFT
    wr2, wi2 = melbank.windowed_dft_kernels(512, 400, "hann")
    engine.check_windowed_dft(wr2, wi2, 512)


def test_twiddle_table():
    tw = engine.twiddle_table(512).numpy()
    k = 37
    assert abs(tw[k, 0] - np.cos(2 * np.pi * k / 512)) < 1e-7 and abs(tw[k, 1] + np.sin(2 * np.pi * k / 512)) < 1e-7


def test_gru_weight_block_order():
    """Row (96*q + 32*g + jj) of the packed recurrent matrix is W_hh[g*256 + 32*q + jj] (sed_b200.h: sed_bigru)."""
    whh = torch.arange(768 * 256, dtype=torch.float32).view(768, 256)
    packed = whh.view(3, 8, 32, 256).permute(1, 0, 2, 3).reshape(768, 256)
    for q, g, jj in ((0, 0, 0), (3, 1, 7), (7, 2, 31)):
        assert torch.equal(packed[96 * q + 32 * g + jj], whh[g * 256 + 32 * q + jj])


def test_shard_bounds_cover_batch_contiguously():
    for B, n in ((4096, 8), (1024, 3), (5, 8), (7, 2)):
        spans = [sdist.shard_bounds(B, n, r) for r in range(n)]
        assert spans[0][0] == 0 and spans[-1][1] == B
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_synthetic_waveform_is_int16_quantised_and_seeded():
    a = synth.synthetic_waveform(2, 1000, seed=1234, rank=0)
    b = synth.synthetic_waveform(2, 1000, seed=1234, rank=0)
    c = synth.synthetic_waveform(2, 1000, seed=1234, rank=1)
    assert torch.equal(a, b) and not torch.equal(a, c)
    q = a * 32767.0
    assert torch.allclose(q, torch.round(q), atol=1e-3)


def test_flops_per_clip_match_survey():
    f = so.flops_per_clip(1001, 512)
    assert abs(f["conv"] - 26.031) < 0.01 and abs(f["total"] - 26.892) < 0.02


def test_c_abi_host_helpers_match_the_python_statements():
    """sed_frontend_twiddle / sed_band_mel (host entries of the C ABI, no GPU needed) against engine.twiddle_table /
    engine.band_mel."""
    from sed_b200 import capi
    lib = capi.load()
    for n_fft in (256, 512, 1024):
        tw = torch.empty((n_fft, 2), dtype=torch.float32)
        assert lib.sed_frontend_twiddle(n_fft, capi.ptr(tw)) == 0
        assert torch.equal(tw, engine.twiddle_table(n_fft))
    assert lib.sed_frontend_twiddle(500, capi.ptr(torch.empty((500, 2)))) != 0
    assert b"n_fft" in lib.sed_last_error_string()
    g = torch.Generator().manual_seed(3)
    W = torch.zeros((257, 64))
    for m in range(64):  # banded columns with a gap inside one band and an empty column
        a = int(torch.randint(0, 200, (1,), generator=g))
        W[a:a + 9, m] = torch.rand(9, generator=g) + 0.1
    W[5, 3] = 0.0
    W[:, 7] = 0.0
    ref = engine.band_mel(W)
    got = engine.band_mel_c(W)
    for a, b in zip(ref, got):
        assert torch.equal(a, b)


def test_host_micro_batch_plan_tiles_the_batch():
    """forward_host's schedule: contiguous cover of [0, B), no micro-batch above the cap, parts respected, and the
    copy of a micro-batch never more than ~1.8x (float32) / ~3.6x (int16) the clips of the one before it."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=300, deadline=None)
    @given(B=st.integers(1, 5000), i16=st.booleans(), mb=st.integers(1, 600), parts=st.sampled_from([1, 2]))
    def check(B, i16, mb, parts):
        ranges, plan = engine.plan_host_micro_batches(B, i16, mb, parts)
        assert len(ranges) == len(plan) and ranges[0][0] == 0 and ranges[-1][1] == B
        pos = 0
        for (p0, p1), spans in zip(ranges, plan):
            assert p0 == pos and spans[0][0] == p0 and spans[-1][1] == p1
            for (b0, b1) in spans:
                assert b0 == pos and b0 < b1 <= p1 and b1 - b0 <= mb
                pos = b1
        assert pos == B

    check()
    _, plan = engine.plan_host_micro_batches(1024, False)
    assert [b1 - b0 for b0, b1 in plan[0]] == [37, 37, 74, 111, 185, 296, 284]
    _, plan = engine.plan_host_micro_batches(1024, True)
    assert [b1 - b0 for b0, b1 in plan[0]] == [37, 111, 296, 580]
    ranges, plan = engine.plan_host_micro_batches(1024, False, result_parts=2)
    assert ranges == [(0, 839), (839, 1024)] and plan[1] == [(839, 1024)]


def test_steady_state_spans_tile_the_batch():
    from hypothesis import given, settings, strategies as st
    from sed_b200 import pipeline

    @settings(max_examples=300, deadline=None)
    @given(B=st.integers(1, 5000), mb=st.integers(1, 1200), span=st.integers(1, 1200))
    def check(B, mb, span):
        spans = pipeline.plan_steady_spans(B, mb, span)
        assert spans[0][0] == 0 and spans[-1][1] == B
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 == b0
        assert all(0 < b1 - b0 <= mb for b0, b1 in spans)

    check()
    assert pipeline.plan_steady_spans(1024, 1036, 370) == [(0, 370), (370, 740), (740, 1024)]
    assert pipeline.plan_steady_spans(2048, 1036, 370)[-1][1] == 2048


def test_micro_batch_cap_scales_with_clip_length():
    assert engine.clamp_micro_batch(1036, 1001) == 1036          # 10 s: the default launch group
    assert engine.clamp_micro_batch(1036, 501) == 1036           # shorter clips never exceed the request
    assert engine.clamp_micro_batch(1036, 6001) == 148           # 60 s: 172 -> whole waves of 37
    assert engine.clamp_micro_batch(2, 6001) == 2                # explicit small caps are honoured
    assert engine.clamp_micro_batch(1036, 10 ** 7) == 1
    for T in (1001, 3001, 6001, 60001):
        cap = engine.clamp_micro_batch(1036, T)
        tiles_per_clip = ((T + 15) // 16) * 8
        assert tiles_per_clip * (cap * tiles_per_clip + 4096) < 2 ** 32   # conv_block1's magic-division guard
    import pytest
    with pytest.raises(ValueError):
        engine.clamp_micro_batch(0, 1001)


def test_peer_view_addressing():
    """dist._PeerView: the pointer arithmetic behind a rank's slice of the destination's result buffer."""
    from sed_b200.dist import _PeerView
    v = _PeerView(1 << 20, (8 * 1024, 1000, 25))
    mine = v[3 * 1024:4 * 1024]
    assert mine.shape == (1024, 1000, 25)
    assert mine.data_ptr() == (1 << 20) + 3 * 1024 * 1000 * 25 * 4
    assert mine[10:12].data_ptr() == mine.data_ptr() + 10 * 100000 and mine[10:12].shape == (2, 1000, 25)
    cai = mine.__cuda_array_interface__
    assert cai["shape"] == (1024, 1000, 25) and cai["typestr"] == "<f4" and cai["data"] == (mine.data_ptr(), False)
    import pytest
    with pytest.raises(IndexError):
        v[::2]
    with pytest.raises(IndexError):
        v[3]


def test_saturation_and_stream_entries_reject_null_without_a_gpu():
    from sed_b200 import capi
    lib = capi.load()
    assert lib.sed_count_saturated16(None, 16, 0, None, None) == 4
    assert lib.sed_stream_write32(None, 1, None) == 4 and lib.sed_stream_wait_geq32(None, 1, None) == 4
    assert lib.sed_peer_copy(None, None, 4, None) == 4 and lib.sed_peer_alloc(0, None) == 4
    assert lib.sed_mha_attention(None, None, 1, 1, 128, None, 0, None) == 4
    assert lib.sed_linear_split16(None, 128, 512, None, None, 1536, None, None, 1024, 0, None) == 4
