"""Known-answer tests that pin the CPU oracle's micro-frontend restatement (oracle/frontend.c).

The reference holds no golden vectors for this path, so the oracle is pinned
against closed forms of the published algorithm, a float FFT, and
structural invariants (chunk-size invariance, warm-up length, silence).
"""

import math

import numpy as np
import pytest

import oracle
from conftest import edge_case_audio, synth_audio


@pytest.fixture(scope="module")
def fe():
    return oracle.Frontend()


@pytest.fixture(scope="module")
def tables(fe):
    return fe.tables()


def test_window_table_closed_form(tables):
    w = tables["window"].astype(np.int64)
    i = np.arange(480)
    ref = np.floor(0.5 + 4096 * (0.5 - 0.5 * np.cos(2 * np.pi * (i + 0.5) / 480))).astype(np.int64)
    assert np.abs(w - ref).max() <= 1          # float32 'arg' upstream may flip a rounding
    assert (w != ref).sum() <= 4
    assert w[0] == 0 and w[-1] == 0 and w[239] == 4096 and w[240] == 4096
    assert np.array_equal(w, w[::-1])            # Hann symmetry


def test_filterbank_layout(tables):
    s = tables["scalars"]
    assert s[0] == 5 and s[1] == 241             # start_index = int(1.5 + 125/31.25), end past mel(7500)
    cs = tables["chan_start"].astype(int)
    assert cs[0] == 5 and cs[41] == 241 and np.all(np.diff(cs) >= 1)   # 41 non-empty ranges
    bc, bw, bu = tables["bin_channel"], tables["bin_weight"].astype(int), tables["bin_unweight"].astype(int)
    inside = np.arange(257)
    assert np.all(bc[(inside < 5) | (inside >= 241)] == -1)
    assert np.all(np.abs(bw[5:241] + bu[5:241] - 4096) <= 1)
    assert np.all((bw[5:241] >= 0) & (bw[5:241] <= 4096))
    # closed-form mel centres
    mel = lambda f: 1127.0 * np.log1p(f / 700.0)
    lo, hi = mel(125.0), mel(7500.0)
    centers = lo + (hi - lo) / 41.0 * np.arange(1, 42)
    for ch in range(41):
        for b in range(cs[ch], cs[ch + 1]):
            assert bc[b] == ch
            prev = lo if ch == 0 else centers[ch - 1]
            w = (centers[ch] - mel(b * 31.25)) / (centers[ch] - prev)
            assert abs(bw[b] - math.floor(w * 4096 + 0.5)) <= 1
        # the first bin of the NEXT range lies above this centre
        assert mel(cs[ch + 1] * 31.25) > centers[ch] - 1e-3


def test_noise_and_pcan_constants(tables):
    s = tables["scalars"]
    assert list(s[2:8]) == [409, 983, 819, 6, 3, 6]   # int(x*16384); snr_shift; correction bits; scale_shift
    lut = tables["gain_lut"].astype(int)
    g = lambda x: min(32767, int(2 ** 21 * (x / 128.0 + 80.0) ** -0.95 + 0.5))
    assert abs(lut[0] - g(0)) <= 1 and abs(lut[1] - g(1)) <= 1
    for k in range(2, 33):
        x0 = 1 << (k - 1)
        assert abs(lut[4 * k - 6] - g(x0)) <= 1


def test_wide_dynamic_function(fe):
    g = lambda x: min(32767.0, 2 ** 21 * (x / 128.0 + 80.0) ** -0.95)
    xs = [0, 1, 2, 3, 4, 5, 7, 8, 100, 1000, 12345, 1 << 16, (1 << 20) + 12345, 1 << 28, (1 << 32) - 1]
    prev = 1 << 30
    for x in xs:
        v = fe.wdf(x)
        assert abs(v - g(x)) <= max(2.0, 0.01 * g(x)), (x, v, g(x))   # quadratic interpolation of the gain curve
        assert v <= prev
        prev = v


def test_log_lut_and_log(tables, fe):
    lut = tables["log_lut"].astype(int)
    assert list(lut[:8]) == [0, 224, 442, 654, 861, 1063, 1259, 1450]
    assert lut.max() == 5641 and lut[128] == 0
    assert round(math.log(2) * 65536) == 45426
    for x in [2, 3, 8, 100, 4097, 65535, 65536, 1 << 20, (1 << 31) + 12345, (1 << 32) - 1]:
        assert abs(fe.log_scaled(x) - 64.0 * math.log(x)) <= 1.0, x


def test_sqrt64_exact_rounding():
    L = oracle.lib()
    rng = np.random.default_rng(0)
    vals = [0, 1, 2, 3, 4, 6, 7, 255, 256, 65535 ** 2, 65535 ** 2 + 65535, 65535 ** 2 + 65536, (1 << 32) - 1,
            1 << 32, (1 << 32) + 1, (1 << 62) - 1, (1 << 63) + 12345, (1 << 64) - 1]
    vals += [int(x) for x in rng.integers(0, 1 << 62, 2000)] + [int(x) for x in rng.integers(0, 1 << 34, 2000)]
    for v in vals:
        r = math.isqrt(v)
        want = r + 1 if v - r * r > r else r
        if v < (1 << 32):
            want = min(want, 0xFFFF)             # 32-bit fast path cannot round up past 0xFFFF
        else:
            want = min(want, 0xFFFFFFFF)
        assert L.mwwo_sqrt64(v) == want, v


def test_fft_against_float(fe):
    rng = np.random.default_rng(1)
    for trial in range(20):
        amp = [30000, 8000, 1000, 100][trial % 4]
        x = np.zeros(512, np.int16)
        x[:480] = rng.integers(-amp, amp + 1, 480)
        out = fe.fftr(x).astype(np.float64)
        ref = np.fft.rfft(x.astype(np.float64)) / 512.0
        err = np.abs(out[:, 0] + 1j * out[:, 1] - ref)
        assert err.max() < 4.0, err.max()        # a handful of Q15 roundings
    # impulse at n=0: flat spectrum  x0/512
    x = np.zeros(512, np.int16)
    x[0] = 32000
    out = fe.fftr(x)
    assert np.all(np.abs(out[:, 0] - 62.5) <= 2) and np.all(np.abs(out[:, 1]) <= 2)
    # DC
    out = fe.fftr(np.full(512, 1000, np.int16))
    assert abs(int(out[0, 0]) - 1000) <= 2 and np.abs(out[1:]).max() <= 1


def test_fft_int16_wrap_is_reachable_and_deterministic(fe):
    """Full-scale complex-exponential patterns overflow int16 inside the radix-4 butterflies; the oracle
    must wrap exactly like int16 stores do (documented KissFFT behaviour), not saturate."""
    t = np.arange(512)
    x = (32767 * np.sign(np.sin(2 * np.pi * (t // 2) / 4.0 + np.pi / 4 + (t % 2) * np.pi / 2) + 1e-9)).astype(np.int16)
    a = fe.fftr(x)
    b = fe.fftr(x)
    assert np.array_equal(a, b)


def test_silence_and_warmup(fe):
    fe.reset()
    out, n = fe.process_samples(np.zeros(160, np.int16))
    assert out is None and n == 160
    out, n = fe.process_samples(np.zeros(160, np.int16))
    assert out is None and n == 160
    out, n = fe.process_samples(np.zeros(160, np.int16))
    assert out is not None and np.all(out == 0)
    # only 160 more samples fit after the hop
    out, n = fe.process_samples(np.zeros(400, np.int16))
    assert n == 160 and out is not None


@pytest.mark.parametrize("n_samples", [16000, 16001, 15999, 480, 481, 640, 641, 320, 0])
def test_clip_loop_row_count(n_samples):
    """m-3 rows for N = 160 m (strict '<' drops the last exact chunk, audio_utils.py:56)."""
    x = synth_audio(max(n_samples, 1), 5)[:n_samples]
    rows = oracle.generate_features_for_clip(x).shape[0]
    chunks_fed = 0
    idx = 0
    while idx + 320 < n_samples * 2:
        idx += 320
        chunks_fed += 1
    assert rows == max(chunks_fed - 2, 0)
    if n_samples % 160 == 0 and n_samples >= 640:
        assert rows == n_samples // 160 - 3


def test_chunking_invariance(fe):
    x = synth_audio(16000, 11)
    fe.reset()
    whole = fe.stream(x)
    assert whole.shape == (98, 40)
    fe.reset()
    parts = []
    pos = 0
    rng = np.random.default_rng(2)
    while pos < x.size:
        n = int(rng.integers(1, 700))
        parts.append(fe.stream(x[pos:pos + n]))
        pos += n
    assert np.array_equal(np.concatenate(parts), whole)
    # the clip loop sees the same rows minus the one the strict '<' drops
    assert np.array_equal(oracle.generate_features_for_clip(x), whole[:97])


def test_feature_range_and_float_model():
    """Features are ~64*ln(.)-scaled (25.6 per unit after the 0.0390625 float scale): a louder copy of the
    same noise must not be quieter in the first frame, and values stay inside the uint16 / [0, 26+] band."""
    rng = np.random.default_rng(3)
    base = rng.normal(0, 300, 1600)
    f1 = oracle.generate_features_for_clip(np.round(base).astype(np.int16))
    f2 = oracle.generate_features_for_clip(np.round(base * 20).astype(np.int16))
    assert f1.shape == f2.shape == (7, 40)
    assert f2[0].astype(int).sum() >= f1[0].astype(int).sum()
    assert f2.max() * 0.0390625 < 40.0


def test_edge_cases_run_and_are_bounded():
    a = edge_case_audio(4800)
    for row in a:
        f = oracle.generate_features_for_clip(row)
        assert f.shape == (27, 40)
        assert f.dtype == np.uint16
    assert np.all(oracle.generate_features_for_clip(a[0]) == 0)     # silence -> zeros


def test_stage_taps_consistent(fe):
    """energy -> work -> sqrt -> noise -> pcan taps obey the stage definitions on a real frame."""
    fe.reset()
    x = synth_audio(480, 21)
    for k in range(3):
        out, _ = fe.process_samples(x[160 * k:160 * (k + 1)])
    assert out is not None
    t, tb = fe.taps(), fe.tables()
    fo = t["fft_out"].astype(np.int64)
    e = fo[:, 0] ** 2 + fo[:, 1] ** 2
    assert np.array_equal(t["energy"][5:241].astype(np.int64), e[5:241])
    cs = tb["chan_start"].astype(int)
    w, u = tb["bin_weight"].astype(np.int64), tb["bin_unweight"].astype(np.int64)
    for ch in range(1, 41):
        want = sum(int(w[b]) * int(e[b]) for b in range(cs[ch], cs[ch + 1])) + \
               sum(int(u[b]) * int(e[b]) for b in range(cs[ch - 1], cs[ch]))
        assert int(t["work"][ch]) == want
        r = math.isqrt(want)
        r = r + 1 if want - r * r > r else r
        assert int(t["sqrt"][ch - 1]) == r >> int(t["shift"][0])
    # first frame from reset: estimate = (sig<<10)*smoothing >> 14
    sm = np.where(np.arange(40) % 2 == 0, 409, 983)
    est = ((t["sqrt"].astype(np.int64) << 10) * sm) >> 14
    assert np.array_equal(t["estimate"].astype(np.int64), est)
