"""CPU-only: pins the oracle restatement (oracle/jrc_oracle.c) against known answers,
NumPy float64 / SciPy complex64 FFTs and hand-derived estimator internals."""
import numpy as np
import pytest
import scipy.fft

from mimo_ofdm_jrc import synth

C1 = dict(T=4, R=2, S=4, N=64, IR=8, IA=16)


def chain(orc, rx, tx, cfg, **kw):
    est = synth.default_estimator_params(cfg["N"], cfg["T"] * cfg["R"], cfg["IR"], cfg["IA"])
    return orc.chain_batch(rx, tx, cfg["N"], cfg["T"], cfg["R"], cfg["S"], cfg["IR"], cfg["IA"], est, **kw), est


# SURVEY.md section 4 table (shipped-sim parameters, noise free)
KATS = [((10, 0), (67, 64), dict(angle_null_idx=0, discard_range_idx=15, discard_angle_idx=5, n_noise=300)),
        ((10, -3), (67, 61), dict(angle_null_idx=126, discard_range_idx=15, discard_angle_idx=5, n_noise=300)),
        ((25, 30), (167, 96), dict(angle_null_idx=8, discard_range_idx=15, discard_angle_idx=16, n_noise=960)),
        ((40, -45), (267, 19), dict())]


@pytest.mark.parametrize("target,peak,internals", KATS)
def test_single_target_kat(orc, target, peak, internals):
    tx = synth.tx_symbols(C1["T"], C1["S"], C1["N"])
    rx = synth.rx_symbols(tx, C1["R"], [[target[0]]], [[target[1]]], [[1.0]])
    (m, cm, d), est = chain(orc, rx, tx, C1, want_cmap=True)
    assert (d[0]["range_idx"], d[0]["angle_idx"]) == peak
    assert peak == synth.expected_peak(target[0], target[1], 64, 8, 8, 16)
    det, dbg = orc.range_angle_estimate(cm[0], **est, want_dbg=True)
    for k, v in internals.items():
        got = det["n_noise"] if k == "n_noise" else dbg[k]
        assert got == v, (k, got, v)
    assert d[0]["flags"] == 1 and d[0]["snr_db"] > 40


def test_chain_matches_numpy_float64(orc):
    rng = np.random.default_rng(1)
    tx = synth.tx_symbols(4, 4, 64)
    r, a, amp = synth.random_scene(rng, 8, 3, 64, amp_db_span=20)
    rx = synth.rx_symbols(tx, 2, r, a, amp, snr_db=20, rng=rng)
    (m, cm, d), _ = chain(orc, rx, tx, C1, want_cmap=True)
    for c in range(8):
        H = np.einsum("rsk,tsk->rtk", rx[c].astype(np.complex128), np.conj(tx.astype(np.complex128))).reshape(8, 64)
        y = np.fft.ifft(H, n=512, axis=1) * 512
        M = np.fft.fftshift(np.fft.fft(y.T, n=128, axis=1), axes=1)
        ref = np.abs(M) ** 2
        assert np.abs(ref - m[c]).max() <= 2e-6 * ref.max()
        assert np.abs(M - cm[c]).max() <= 2e-6 * np.abs(M).max()
        assert np.unravel_index(np.argmax(ref), ref.shape) == (d[c]["range_idx"], d[c]["angle_idx"])


@pytest.mark.parametrize("n", [2, 8, 64, 128, 512, 1024, 4096])
@pytest.mark.parametrize("forward,shift", [(True, False), (True, True), (False, False), (False, True)])
def test_fft_vcc_semantics(orc, n, forward, shift):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal((3, n)) + 1j * rng.standard_normal((3, n))).astype(np.complex64)
    got = orc.fft_vcc(x, forward, shift)
    x64 = x.astype(np.complex128)
    if forward:
        ref = np.fft.fft(x64, axis=1)
        if shift:
            ref = np.fft.fftshift(ref, axes=1)
    else:
        xin = np.fft.ifftshift(x64, axes=1) if shift else x64     # gr-fft swaps the input halves
        ref = np.fft.ifft(xin, axis=1) * n
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 3e-6 * scale
    # second opinion in float32 arithmetic: pocketfft complex64
    sp = scipy.fft.fft(x, axis=1) if forward else scipy.fft.ifft(np.fft.ifftshift(x, axes=1) if shift else x, axis=1) * n
    if forward and shift:
        sp = np.fft.fftshift(sp, axes=1)
    assert np.abs(got - sp).max() <= 3e-6 * scale


def test_fft_non_power_of_two_uses_exact_dft(orc):
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(960) + 1j * rng.standard_normal(960)).astype(np.complex64)
    got = orc.fft_vcc(x[None], True, False)[0]
    ref = np.fft.fft(x.astype(np.complex128))
    assert np.abs(got - ref).max() <= 2e-7 * np.abs(ref).max()


def test_matrix_transpose_and_mag(orc):
    rng = np.random.default_rng(2)
    x = (rng.standard_normal((8, 512)) + 1j * rng.standard_normal((8, 512))).astype(np.complex64)
    t = orc.matrix_transpose(x, 8, 16)
    assert t.shape == (512, 128)
    assert np.array_equal(t[:, :8], x.T) and not t[:, 8:].any()
    m = orc.mag_squared(x)
    assert np.array_equal(m, (x.real * x.real + x.imag * x.imag).astype(np.float32))


def test_hypotf_is_double_sqrt(orc):
    """The CUDA side evaluates std::abs(complex<float>) as (float)sqrt((double)x*x+(double)y*y);
    check that this is what the oracle's libm hypotf does on this image."""
    rng = np.random.default_rng(3)
    z = ((rng.standard_normal(200000) + 1j * rng.standard_normal(200000)) *
         10.0 ** rng.uniform(-6, 6, 200000)).astype(np.complex64)
    # estimator with a 1 x n map: peak power = max pow(abs,2); compare per element through peak1d instead
    re, im = z.real.astype(np.float64), z.imag.astype(np.float64)
    mine = np.sqrt(re * re + im * im).astype(np.float32)
    for i in range(0, z.size, 4001):     # sample: fft_peak_detect returns abs(in[k]) of the arg-max
        k, f, ph, mag = orc.fft_peak_detect(z[i:i + 1], 1, 1.0, -400.0, 0)
        assert k == 0 and np.float32(mag) == mine[i]


def test_estimator_null_angle_rules(orc):
    est = synth.default_estimator_params(64, 8, 8, 16)
    ab = est["angle_bins"]
    m = np.zeros((512, 128), dtype=np.complex64)
    m[:] = 0.01
    # peak at broadside -> angle_null below the first bin -> lower_bound == begin -> idx 0
    m[100, 64] = 5.0
    det, dbg = orc.range_angle_estimate(m, **est, want_dbg=True)
    assert (det["range_idx"], det["angle_idx"]) == (100, 64) and dbg["angle_null_idx"] == 0
    # peak slightly negative -> angle_null above the last bin -> end() -> size-1 -> clamped to size-2
    m[100, 64] = 0.01
    m[100, 61] = 5.0
    det, dbg = orc.range_angle_estimate(m, **est, want_dbg=True)
    assert dbg["angle_null_idx"] == ab.size - 2
    # window wraps modulo the map (peak near the end of the range axis)
    m[:] = 0.02
    m[500, 30] = 3.0
    det, dbg = orc.range_angle_estimate(m, **est, want_dbg=True)
    assert dbg["start_range_idx"] == 500 + 256 - 15 and det["n_noise"] == 30 * 2 * dbg["discard_angle_idx"]
    assert abs(det["noise_power"] - 0.02 ** 2) < 1e-7
    # first maximum in row-major order wins
    m[:] = 0.0
    m[7, 9] = m[7, 100] = m[300, 2] = 2.0
    det = orc.range_angle_estimate(m, **est)
    assert (det["range_idx"], det["angle_idx"]) == (7, 9)
    # gate: zero noise -> snr = +inf passes any threshold (faithful to :226-234); finite noise does not
    est2 = dict(est, snr_threshold=np.float32(1e9))
    assert orc.range_angle_estimate(m, **est2)["flags"] == 1
    m[m == 0] = 0.1
    assert orc.range_angle_estimate(m, **est2)["flags"] == 0
    assert orc.range_angle_estimate(m, **dict(est, power_threshold=np.float32(5.0)))["flags"] == 0


def test_radar_block_background_and_interleave(orc):
    rng = np.random.default_rng(4)
    T, R, S, N, pre = 4, 2, 4, 64, 5
    def frame():
        return [(rng.standard_normal((pre + S) * N) + 1j * rng.standard_normal((pre + S) * N)).astype(np.complex64)
                for _ in range(T + R)]
    rad = orc.Radar(N, T, R, S, pre, True, True, 3, 8, False)
    raws, outs = [], []
    for i in range(6):
        f = frame()
        tx, rx = f[:T], f[T:]
        out = rad.work(tx, rx)
        H = np.zeros((R * T, N), dtype=np.complex128)
        for r in range(R):
            for t in range(T):
                a = rx[r].reshape(-1, N)[pre:pre + S].astype(np.complex128)
                b = tx[t].reshape(-1, N)[pre:pre + S].astype(np.complex128)
                H[r * T + t] = (a * np.conj(b)).sum(axis=0)
        mean = np.mean(raws[-3:], axis=0) if raws else 0
        assert np.abs(out[:, :N] - (H - mean)).max() < 2e-5 * np.abs(H).max()
        assert not out[:, N:].any()
        raws.append(H)
    # tx interleave permutes the rows: p = t*R + r
    a = orc.Radar(N, T, R, S, pre, False, False, 1, 1, False)
    b = orc.Radar(N, T, R, S, pre, False, False, 1, 1, True)
    f = frame()
    oa, ob = a.work(f[:T], f[T:]), b.work(f[:T], f[T:])
    for r in range(R):
        for t in range(T):
            assert np.array_equal(oa[r * T + t], ob[t * R + r])
    # stale TX frames are skipped by tx_skip_items (lib/mimo_ofdm_radar_impl.cc:189-197,260)
    f2 = frame()
    txcat = [np.concatenate([f[i], f2[i]]) for i in range(T)]
    o2 = a.work(txcat, f2[T:], tx_skip_items=pre + S)
    assert np.array_equal(o2, a.work(f2[:T], f2[T:]))


def test_fft_peak_detect_rules(orc):
    n = 1000
    x = np.zeros(n, dtype=np.complex64)
    x[3] = 10.0          # inside the protected zone
    x[700] = 2.0 + 1.0j
    x[200] = 2.0 + 1.0j  # equal magnitude, earlier -> wins
    k, f, ph, mag = orc.fft_peak_detect(x, 1000, 2.0, 0.0, 10)
    assert k == 200 and np.isclose(f, 200 / 1000 * 2000) and np.isclose(ph, np.arctan2(1, 2))
    x[200] = 0
    k, f, ph, mag = orc.fft_peak_detect(x, 1000, 2.0, 0.0, 10)
    assert k == 700 and np.isclose(f, -2000 + 700 * 2.0)     # negative-frequency branch
    k, *_ = orc.fft_peak_detect(x, 1000, 2.0, 30.0, 10)     # nothing above 30 dB
    assert k == -1


def test_zero_pad_and_cp_remove(orc):
    x = (np.arange(160) + 1j).astype(np.complex64)
    y = orc.zero_pad(x, 7, 240, seed=11)
    assert y.size == 160 + 247 and np.array_equal(y[7:167], x)
    pads = np.concatenate([y[:7], y[167:]])
    big = orc.zero_pad(x, 20000, 20000, seed=12)
    p = np.concatenate([big[:20000], big[-20000:]])
    assert abs(p.real.std() - 1e-2) < 3e-4 and abs(p.imag.std() - 1e-2) < 3e-4 and abs(p.mean()) < 3e-4
    assert pads.size == 247
    z = orc.cp_remove(x, 2, 64, 16)
    assert np.array_equal(z[0], x[16:80]) and np.array_equal(z[1], x[96:160])


def test_single_bin_dit(orc):
    """csrc/jrc_exact.cuh dit_bin_*: ONE bin of the oracle's radix-2 FFT of a zero-padded input as a pairwise reduction
    of the bit-reversed inputs with one twiddle and one sign per level.  NumPy model of exactly that recipe (float32
    products and sums rounded one by one) against the oracle's transform, every bin, identical bits."""
    f32 = np.float32

    def cmul(a, b):
        ar, ai, br, bi = f32(a.real), f32(a.imag), f32(b.real), f32(b.imag)
        return complex(f32(f32(ar * br) - f32(ai * bi)), f32(f32(ar * bi) + f32(ai * br)))

    def cadd(a, b, sgn):
        return complex(f32(f32(a.real) + sgn * f32(b.real)), f32(f32(a.imag) + sgn * f32(b.imag)))

    def rev(k, bits):
        return int(format(k, f"0{bits}b")[::-1], 2) if bits else 0

    def dit_bin(x, n, tw, o):
        m, ln = len(x).bit_length() - 1, n.bit_length() - 1
        A = [complex(x[rev(k, m)]) for k in range(len(x))]
        for s in range(m):
            lvl = ln - m + s
            w = complex(tw[(o & ((1 << lvl) - 1)) * (n >> (lvl + 1))])
            sgn = f32(-1.0) if (o >> lvl) & 1 else f32(1.0)
            A = [cadd(A[2 * k], cmul(A[2 * k + 1], w), sgn) for k in range(len(A) // 2)]
        return np.complex64(A[0])

    rng = np.random.default_rng(0)
    for n_in, n, fwd in ((64, 1024, False), (8, 64, True), (64, 512, False), (8, 128, True), (32, 256, True)):
        x = (rng.standard_normal(n_in) + 1j * rng.standard_normal(n_in)).astype(np.complex64)
        xp = np.zeros(n, np.complex64)
        xp[:n_in] = x
        ref = orc.fft_vcc(xp[None], fwd, fwd)[0]              # the angle FFT runs with shift=True, the range IFFT without
        a = (-1.0 if fwd else 1.0) * 2.0 * np.pi * np.arange(n // 2) / n
        tw = (np.cos(a).astype(f32) + 1j * np.sin(a).astype(f32)).astype(np.complex64)
        for i in range(0, n, 1 if n <= 128 else 7):
            o = (i + n // 2) % n if fwd else i
            assert dit_bin(x, n, tw, o).view(np.uint64) == ref[i].view(np.uint64), (n_in, n, i)


def test_speculative_sequential_sum():
    """The device's noise-window sum (jrc_common.cuh seq_sum_sq_warp) takes t = RN32(acc + RN32(v*v)) as the result of the
    reference's acc = (float)((double)acc + (double)v*v) (lib/range_angle_estimator_impl.cc:217) whenever the exact remainder
    stays clear of half an ulp of t, and redoes the step with the expression itself otherwise.  NumPy model of that test:
    no accepted step may differ from the expression -- random operands over 12 decades, sums placed ON and next to float
    midpoints -- and a noise-like accumulation must accept nearly every step."""
    f32 = np.float32

    def ulp_of(t):
        return ((t.view(np.uint32) & np.uint32(0x7f800000)) - np.uint32(23 << 23)).view(np.float32)

    def fast_step(s, a):
        ph = a * a
        pl = (a.astype(np.float64) * a.astype(np.float64) - ph.astype(np.float64)).astype(f32)     # = fma(a, a, -ph), exact
        t = s + ph
        bv = t - s
        e = (s - (t - bv)) + (ph - bv)
        d = e + pl
        tb = t.view(np.uint32)
        ex = (tb >> 23) & 0xff
        safe = (np.abs(d) <= f32(0.49999) * ulp_of(t)) & ((tb & np.uint32(0x007fffff)) != 0) & (ex > 30) & (ex < 250)
        return t, safe

    def ref_step(s, a):
        return (s.astype(np.float64) + a.astype(np.float64) * a.astype(np.float64)).astype(f32)

    rng = np.random.default_rng(1)
    n_safe = 0
    with np.errstate(all="ignore"):
        for scale_s in (1e-3, 1.0, 37.0, 1e4, 3e7):
            for scale_a in (1e-4, 1e-2, 0.3, 1.0, 20.0, 3e3):
                s = (np.abs(rng.standard_normal(100000)) * scale_s).astype(f32)
                a = (rng.standard_normal(100000) * scale_a).astype(f32)
                t, safe = fast_step(s, a)
                assert not (safe & (t != ref_step(s, a))).any()
                n_safe += int(safe.sum())
        s = (np.abs(rng.standard_normal(500000)) * 8).astype(f32)
        mid = s.astype(np.float64) + ulp_of(s).astype(np.float64) * (rng.integers(0, 64, s.size) + 0.5)
        a0 = np.sqrt(mid - s.astype(np.float64)).astype(f32)
        for a in (np.nextafter(a0, f32(0)), a0, np.nextafter(a0, f32(np.inf))):
            t, safe = fast_step(s, a)
            assert not (safe & (t != ref_step(s, a))).any()
            n_safe += int(safe.sum())
    assert n_safe > 2_000_000
    a = (rng.standard_normal(5000) * 3e-2).astype(f32)
    acc, unsafe = f32(0), 0
    for v in a:
        t, safe = fast_step(np.array([acc], f32), np.array([v], f32))
        r = ref_step(np.array([acc], f32), np.array([v], f32))[0]
        if safe[0]:
            assert t[0] == r
        else:
            unsafe += 1
        acc = r
    assert unsafe <= 25, unsafe
