"""CPU model of the index arithmetic inside `csrc/sed_frontend.cu` (no GPU needed).

The front-end kernel relies on three pieces of integer bookkeeping that are easy to get wrong and impossible to see in
a tolerance test: (1) swizzled shared-memory addresses formed as (per-lane base) XOR / + (compile-time constant),
(2) the lane / register-slot pairing of the power split after the last FFT pass, (3) the segment schedule of the mel
projection built by `build_mel_schedule`.  This file restates each of them in numpy exactly as the kernel computes
them and checks them against first principles (the plain swizzle function, numpy's FFT, a dense matmul)."""
import numpy as np
import pytest

from sed_b200 import melbank, synth

SCHED = {256: (4, 8, 8), 512: (8, 8, 8), 1024: (8, 8, 16)}          # Sched<N> in the kernel
SEG = {256: 3, 512: 7, 1024: 9}                                     # FrontCfg<N>::SEG
SEGS_MAX = {256: 96, 512: 160, 1024: 160}                           # FrontCfg<N>::SEGS_MAX


def pidx(i):
    return i ^ ((i >> 3) & 15)


def eaddr(base, c):
    p = pidx(c)
    return (base ^ ((p & 15) << 3)) + ((p & ~15) << 3)


@pytest.mark.parametrize("n_fft", [256, 512, 1024])
def test_pass_addresses_equal_the_plain_swizzle(n_fft):
    r0, r1, r2 = SCHED[n_fft]
    buf = 128 * 37  # any 128-byte aligned shared-memory address
    for pi, (R, Ns) in enumerate(((r0, 1), (r1, r0), (r2, r0 * r1))):
        NB = n_fft // R
        BPL = NB // 32
        for lane in range(32):
            ld_base = buf + 8 * pidx(lane)
            st_base = buf + 8 * pidx((lane // Ns) * Ns * R + lane % Ns) if Ns <= 32 else None
            for q in range(BPL):
                j = lane + 32 * q
                for r in range(R):
                    i = q + BPL * r
                    assert j + r * NB == lane + 32 * i                       # input index of butterfly j, leg r
                    if pi > 0:
                        assert eaddr(ld_base, 32 * i) == buf + 8 * pidx(j + r * NB)
                    out = (j // Ns) * Ns * R + j % Ns + r * Ns                # Stockham autosort output index
                    if pi < 2:
                        assert eaddr(st_base, r * Ns + 32 * R * q) == buf + 8 * pidx(out)
                    else:
                        assert out == lane + 32 * i                          # last pass: X[lane + 32 i] stays in the lane


@pytest.mark.parametrize("n_fft", [256, 512, 1024])
def test_power_split_pairing(n_fft):
    rng = np.random.default_rng(n_fft)
    a, b = rng.standard_normal(n_fft), rng.standard_normal(n_fft)
    Z = np.fft.fft(a + 1j * b)
    NS, H = n_fft // 32, n_fft // 64
    X = [[Z[lane + 32 * i] for i in range(NS)] for lane in range(32)]
    P = np.zeros((n_fft // 2 + 1, 2))
    for lane in range(32):
        src = (32 - lane) & 31
        for i in range(H):
            zk = X[lane][i]
            zn = X[src][NS - 1 - i]                  # the shuffle
            if lane == 0:
                zn = X[0][(NS - i) % NS]              # lane 0 pairs with itself
            ar, ai = zk.real + zn.real, zk.imag - zn.imag
            br, bi = zk.imag + zn.imag, zn.real - zk.real
            P[lane + 32 * i] = (0.25 * (ar * ar + ai * ai), 0.25 * (br * br + bi * bi))
        if lane == 0:
            z = X[0][H]
            P[n_fft // 2] = (z.real ** 2, z.imag ** 2)
    assert np.allclose(P[:, 0], np.abs(np.fft.rfft(a)) ** 2)
    assert np.allclose(P[:, 1], np.abs(np.fft.rfft(b)) ** 2)


@pytest.mark.parametrize("n_fft", [256, 512, 1024])
def test_two_pass_transform(n_fft):
    """frontend2_kernel: N = RA * 32; pass A = RA-point DFTs by lane n2 with the twiddle W_N^(n2 k1) applied by the
    producer, pass B = 32-point DFTs by lane (transform, k1) over a row-swizzled buffer, then the split pairing."""
    from collections import defaultdict
    RA, T = n_fft // 32, 32 // (n_fft // 32)
    rng = np.random.default_rng(n_fft)
    xs = [rng.standard_normal(n_fft) + 1j * rng.standard_normal(n_fft) for _ in range(T)]
    tw = np.exp(-2j * np.pi * np.arange(n_fft) / n_fft)
    buf = np.zeros(T * n_fft, complex)

    def wavefronts(addrs):  # 64-bit accesses: one wavefront per half-warp when the 16 slots mod 16 are distinct
        worst = 0
        for h in range(2):
            d = defaultdict(set)
            for a in addrs[16 * h:16 * h + 16]:
                d[a % 16].add(a)
            worst = max(worst, max(len(v) for v in d.values()))
        return worst

    for t in range(T):
        for k1 in range(RA):
            row, addrs = t * RA + k1, []
            for lane in range(32):
                y = np.fft.fft(np.array([xs[t][lane + 32 * n1] for n1 in range(RA)]))[k1] * tw[(lane * k1) % n_fft]
                byte = ((8 * lane) ^ (8 * (row & 15))) + 256 * row              # the kernel's store address
                assert byte == 8 * (32 * row + (lane ^ (row & 15)))
                buf[byte // 8] = y
                addrs.append(byte // 8)
            assert wavefronts(addrs) == 1
    regs = {}
    for n2 in range(32):
        addrs = []
        for lane in range(32):
            byte = ((256 * lane + 8 * (lane & 15)) ^ (8 * (n2 & 15))) + 8 * (n2 & 16)   # the kernel's load address
            assert byte == 8 * (32 * lane + (n2 ^ (lane & 15)))
            addrs.append(byte // 8)
        assert wavefronts(addrs) == 1
    for lane in range(32):
        regs[lane] = np.fft.fft(np.array([buf[32 * lane + (n2 ^ (lane & 15))] for n2 in range(32)]))
    X = [np.fft.fft(x) for x in xs]
    for lane in range(32):
        t, k1 = lane // RA, lane % RA
        src = (lane & ~(RA - 1)) | ((RA - k1) & (RA - 1))
        for k2 in range(32):
            assert np.isclose(regs[lane][k2], X[t][k1 + RA * k2])                # pass B leaves X_t[k1 + RA k2]
        for k2 in range(16):
            zn = regs[src][31 - k2] if k1 else regs[lane][(32 - k2) & 31]
            assert np.isclose(zn, X[t][(n_fft - (k1 + RA * k2)) % n_fft])        # the partner bin N - k
    # power spectra of the transforms do not collide in the warp's buffer, and their stores are conflict-free
    base = [t * n_fft + (8 * (t & 1) if RA == 8 else 0) for t in range(T)]
    for k2 in range(16):
        assert wavefronts([base[lane // RA] + lane % RA + RA * k2 for lane in range(32)]) == 1
    for t in range(T - 1):
        assert base[t] + n_fft // 2 + 1 + SEGS_MAX[n_fft] <= base[t + 1]
    assert base[T - 1] + n_fft // 2 + 1 + SEGS_MAX[n_fft] <= T * n_fft


def band(W):
    F, M = W.shape
    lo, ln, off, vals, pos = np.zeros(M, int), np.zeros(M, int), np.zeros(M, int), [], 0
    for m in range(M):
        nz = np.nonzero(W[:, m])[0]
        if nz.size:
            lo[m], ln[m] = nz[0], nz[-1] - nz[0] + 1
        off[m] = pos
        vals.append(W[lo[m]:lo[m] + ln[m], m])
        pos += ln[m]
    return lo, ln, off, (np.concatenate(vals) if pos else np.zeros(1))


def build_schedule(n_fft, lo, ln, off, val, n_mels):
    """build_mel_schedule of the kernel, statement for statement (the warp-synchronous window search included)."""
    seg, F = SEG[n_fft], n_fft // 2 + 1
    npm = [(max(ln[m], 0) + seg - 1) // seg for m in range(n_mels)]
    first = np.concatenate([[0], np.cumsum(npm)[:-1]]).astype(int)
    nseg = int(sum(npm))
    if nseg > SEGS_MAX[n_fft]:
        return None
    slots = ((nseg + 31) // 32) * 32
    segmj = [None] * slots
    for m in range(n_mels):
        for j in range(npm[m]):
            segmj[first[m] + j] = (m, j)
    seglo = np.zeros(slots, int)
    for r in range(slots // 32):
        start, lowest = [0] * 32, [0] * 32
        for lane in range(32):
            if segmj[32 * r + lane] is not None:
                m, j = segmj[32 * r + lane]
                start[lane] = min(lo[m] + j * seg, F - seg)
                lowest[lane] = max(lo[m] + min((j + 1) * seg, ln[m]) - seg, 0)
        for _ in range(seg):
            key = [(start[l] & 15) | (l & 16) | (0 if segmj[32 * r + l] is not None else 32 + l) for l in range(32)]
            new = list(start)
            for l in range(32):
                f = key.index(key[l])                                         # lowest lane with the same key
                if l != f and start[l] != start[f] and start[l] > lowest[l]:
                    new[l] = start[l] - 1
            start = new
        for lane in range(32):
            seglo[32 * r + lane] = start[lane] if segmj[32 * r + lane] is not None else start[lane & 16]
    segw = np.zeros(slots * seg)
    for s in range(slots):
        for i in range(seg):
            w = 0.0
            if segmj[s] is not None:
                m, j = segmj[s]
                rel = seglo[s] + i - lo[m]
                if j * seg <= rel < min((j + 1) * seg, ln[m]):
                    w = val[off[m] + rel]
            segw[((s >> 5) * seg + i) * 32 + (s & 31)] = w
    return nseg, first, npm, seglo, segw


def run_schedule(n_fft, sched, P, n_mels):
    seg = SEG[n_fft]
    nseg, first, npm, seglo, segw = sched
    rounds = (nseg + 31) >> 5
    part = np.zeros(32 * rounds)
    for r in range(rounds):
        for lane in range(32):
            acc = 0.0
            for i in range(seg):
                k = seglo[32 * r + lane] + i
                assert 0 <= k < n_fft // 2 + 1                                 # every read stays inside the spectrum
                acc += P[k] * segw[(r * seg + i) * 32 + lane]
            part[32 * r + lane] = acc
    return np.array([part[first[m]:first[m] + npm[m]].sum() for m in range(n_mels)])


@pytest.mark.parametrize("sr", [8000, 16000, 32000])
@pytest.mark.parametrize("n_mels", [40, 64, 128])
def test_mel_schedule_equals_dense_projection(sr, n_mels):
    n_fft, _, fmin, fmax = synth.PRESETS[sr]
    W = np.asarray(melbank.mel_filterbank(sr, n_fft, n_mels, fmin, fmax), dtype=np.float64)
    if W.shape[0] == n_mels:
        W = W.T
    lo, ln, off, val = band(W)
    sched = build_schedule(n_fft, lo, ln, off, val, n_mels)
    if sched is None:      # too many segments: the kernel projects one lane per band instead
        assert int(np.ceil(ln / SEG[n_fft]).sum()) > SEGS_MAX[n_fft]
        return
    P = np.random.default_rng(sr + n_mels).random(n_fft // 2 + 1)
    assert np.allclose(run_schedule(n_fft, sched, P, n_mels), P @ W, rtol=1e-12, atol=1e-15)
    if n_mels == 64:       # the shipped presets run the fully unrolled path (FrontCfg::ROUNDS_UNROLL)
        assert (sched[0] + 31) // 32 == (4 if n_fft == 1024 else 3)
        assert max(sched[2]) <= (6 if n_fft == 1024 else 3)                    # FrontCfg::NP_UNROLL


@pytest.mark.parametrize("n_fft", [256, 512, 1024])
def test_mel_schedule_random_banded_matrices(n_fft):
    rng = np.random.default_rng(n_fft + 1)
    F = n_fft // 2 + 1
    for _ in range(60):
        M = int(rng.integers(1, 70))
        W = np.zeros((F, M))
        for m in range(M):
            if rng.random() < 0.1:
                continue                                                       # empty band
            n = int(rng.integers(1, 12))
            a = F - n if rng.random() < 0.2 else int(rng.integers(0, F - n + 1))  # some bands touch the last bin
            W[a:a + n, m] = rng.random(n) + 0.1
            if n > 2 and rng.random() < 0.3:
                W[a + 1, m] = 0.0                                              # interior zero weight
        lo, ln, off, val = band(W)
        sched = build_schedule(n_fft, lo, ln, off, val, M)
        if sched is None:
            continue
        P = rng.random(F)
        assert np.allclose(run_schedule(n_fft, sched, P, M), P @ W, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("sr", [8000, 16000, 32000])
def test_kernel_arithmetic_in_float32_meets_the_reference_fixture(sr):
    """The kernel's arithmetic end to end, in numpy float32: two real frames packed into one complex transform, the
    power split with its folded 1/4, the segment mel schedule with quarter weights, 10*log10 through log2 -- against
    the log-mel fixture the unmodified reference produced (tests/golden/frontend_*.npz) at the north-star tolerance."""
    from conftest import load_golden, logmel_close
    n_fft, hop, _, _ = synth.PRESETS[sr]
    g = load_golden("frontend_%dk.npz" % (sr // 1000))
    x = (g["wave_i16"].astype(np.float64) / 32767.0).astype(np.float32)       # the correctly rounded q / 32767
    melW = g["melW"].astype(np.float64)
    lo, ln, off, val = band(melW)
    sched = build_schedule(n_fft, lo, ln, off, (0.25 * val.astype(np.float32)).astype(np.float32), 64)
    assert sched is not None
    nseg, first, npm, seglo, segw = sched
    seg = SEG[n_fft]
    rounds = (nseg + 31) >> 5
    win = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n_fft) / n_fft)).astype(np.float32)   # periodic Hann (stft.py:192-195)
    pad = np.pad(x, ((0, 0), (n_fft // 2, n_fft // 2)), mode="reflect")
    T = x.shape[1] // hop + 1
    out = np.zeros((x.shape[0], T, 64), np.float32)
    F = n_fft // 2 + 1
    k = np.arange(F)
    for b in range(x.shape[0]):
        for fa in range(0, T, 2):
            fb = min(fa + 1, T - 1)
            a = pad[b, fa * hop:fa * hop + n_fft] * win
            c = pad[b, fb * hop:fb * hop + n_fft] * win
            Z = np.fft.fft((a + 1j * c).astype(np.complex64)).astype(np.complex64)
            zk, zn = Z[k % n_fft], Z[(n_fft - k) % n_fft]
            ar, ai = zk.real + zn.real, zk.imag - zn.imag
            br, bi = zk.imag + zn.imag, zn.real - zk.real
            P4 = np.stack([ai * ai + ar * ar, bi * bi + br * br], 1).astype(np.float32)     # 4 |A|^2, 4 |B|^2
            part = np.zeros((32 * rounds, 2), np.float32)
            for r in range(rounds):
                for lane in range(32):
                    s = 32 * r + lane
                    acc = np.zeros(2, np.float32)
                    for i in range(seg):
                        acc = (acc + P4[seglo[s] + i] * np.float32(segw[(r * seg + i) * 32 + lane])).astype(np.float32)
                    part[s] = acc
            for m in range(64):
                y = np.zeros(2, np.float32)
                for j in range(npm[m]):
                    y = (y + part[first[m] + j]).astype(np.float32)
                db = np.where(y > 1e-10, np.float32(3.010299956639812) * np.log2(np.maximum(y, 1e-30)).astype(np.float32),
                              np.float32(-100.0)).astype(np.float32)
                out[b, fa, m] = db[0]
                if fa + 1 < T:
                    out[b, fa + 1, m] = db[1]
    ok = logmel_close(out, g["logmel"], rtol=1e-4)
    assert ok.all(), "violations %.3e, max |d| %.3e dB" % (1 - ok.mean(), np.abs(out - g["logmel"]).max())
    assert np.all(out[3] == -100.0)  # the silent clip of the fixture
