"""VERDICT r01 next-7b / SURVEY.md section 7 hard part 1: the 16 kHz / 30 ms / 10 ms / 40-channel frontend pinned to its
MATHEMATICAL definition, independently of the integer tables and of the upstream unit-test configuration
(tests/test_upstream_kat.py pins the code on the library's own 1 kHz / 2-channel vectors).

A float64 model written from the definitions -- Hann window, 512-point DFT, triangular mel weights from the mel formula,
square root, one-pole noise estimate, spectral subtraction with a 5 % floor, the PCAN gain curve 2^21 (x/128 + 80)^-0.95 and
soft knee, 64 ln(.) -- is compared with the integer oracle stage by stage.  Each stage is compared on ITS OWN integer inputs
(the subtraction stage is ill-conditioned whenever signal ~ noise estimate, so chaining float stages end to end would only
measure that), with bounds that a wrong table entry, channel boundary, smoothing constant, shift or LUT segment breaks:
  stage A  audio -> filterbank square root  (window, FFT scaling 1/512, mel weights, sqrt, >> shift):  1 % of the value + 0.5 % of the frame's strongest channel
  stage B  noise estimate recursion, even/odd smoothing 0.025 / 0.06:  relative 2e-3
  stage C  subtraction + floor + PCAN + log on the integer (sqrt, estimate):  <= 4 feature LSB + the rounding of the integer gain, median of the per-frame worst <= 2
and end to end on the well-conditioned entries (signal well above its noise estimate): median <= 2 LSB."""

import numpy as np

import oracle
from conftest import edge_case_audio, synth_audio

N_FRAMES = 60


def _mel_weights():
    mel = lambda f: 1127.0 * np.log1p(f / 700.0)
    lo, hi = mel(125.0), mel(7500.0)
    centers = lo + (hi - lo) / 41.0 * np.arange(1, 42)
    bm = mel(np.arange(257) * (8000.0 / 256.0))
    w = np.zeros((42, 257))
    ch = 0
    for b in range(int(1.5 + 125.0 / 31.25), 257):
        while ch < 41 and bm[b] > centers[ch]:
            ch += 1
        if ch >= 41:
            break
        prev = lo if ch == 0 else centers[ch - 1]
        t = (centers[ch] - bm[b]) / (centers[ch] - prev)
        w[ch, b] += t                 # falling edge of half-band ch ...
        w[ch + 1, b] += 1.0 - t       # ... is the rising edge of the next one
    return w[1:41]                    # output channels = half-bands 1..40


def _frames(x):
    fe = oracle.Frontend()
    fe.reset()
    out = []
    for k in range(len(x) // 160):
        feat, _ = fe.process_samples(x[160 * k:160 * (k + 1)])
        if feat is not None:
            t = fe.taps()
            out.append(dict(feat=feat.astype(np.int64), v=t["sqrt"].astype(np.float64), est=t["estimate"].astype(np.float64),
                            start=160 * (k + 1) - 480))
    return out


def _stage_c(v, est):
    """subtraction + floor + PCAN + log, float64, on the integer sqrt values and the UPDATED integer estimate"""
    s = v * 1024.0
    e = np.minimum(est, s)
    v2 = np.maximum((s - e) / 1024.0, v * 0.05)
    gain = np.minimum(32767.0, 2.0 ** 21 * (est / 128.0 + 80.0) ** -0.95)
    snr = v2 * gain / 64.0
    o = np.where(snr < 8192.0, snr * snr / 2.0 ** 20, snr / 64.0 - 64.0) * 8.0
    return np.where(o > 1.0, 64.0 * np.log(np.maximum(o, 1e-300)), 0.0)


def _signals():
    rng = np.random.default_rng(11)
    t = np.arange(160 * (N_FRAMES + 3))
    sig = [synth_audio(t.size, 400 + i) for i in range(3)]
    sig.append(np.round(6000 * np.sin(2 * np.pi * 440.0 * t / 16000) + rng.normal(0, 200, t.size)).astype(np.int16))
    chirp = 9000 * np.sin(2 * np.pi * (200 + 3500 * t / t.size) * t / 16000)
    sig.append(np.round(chirp * (0.2 + 0.8 * (t % 4000 < 2000))).astype(np.int16))         # level steps: signal well above the estimate
    sig.append(np.round(rng.normal(0, 1, t.size) * np.exp(rng.normal(0, 1.0, t.size // 800 + 1).repeat(800)[:t.size]) * 1500).clip(-32768, 32767).astype(np.int16))
    return sig


def test_stage_a_filterbank_root_matches_float_dft_and_mel_definition():
    W = _mel_weights()
    win = 0.5 - 0.5 * np.cos(2 * np.pi * (np.arange(480) + 0.5) / 480)
    worst = 0.0
    for x in _signals():
        for fr in _frames(x)[:N_FRAMES]:
            seg = x[fr["start"]:fr["start"] + 480].astype(np.float64) * win
            e = np.abs(np.fft.rfft(seg, 512)) ** 2
            v = np.sqrt(W @ e) / 8.0                  # Q12 weights (x 64 after the root), FFT gain 1/512
            # the input is scaled to 15 bits by the frame's LARGEST sample, so the 16-bit FFT's rounding noise is a fixed
            # fraction of the frame's strongest channel (a dominant tone leaves the weak channels few bits): bound the error
            # by 1 % of the value + 0.5 % of the frame maximum; a wrong weight, channel edge, window or gain breaks that by far
            tol = 0.01 * v + 0.005 * v.max() + 3.0
            assert np.all(np.abs(fr["v"] - v) <= tol), (np.max(np.abs(fr["v"] - v) / tol), fr["start"])
            big = v >= 0.25 * v.max()
            if v.max() >= 200.0:
                worst = max(worst, float(np.median(np.abs(fr["v"][big] - v[big]) / v[big])))
    assert 0.0 < worst <= 0.006


def test_stage_b_noise_estimate_is_the_one_pole_filter():
    sm = np.where(np.arange(40) % 2 == 1, 0.06, 0.025)
    for x in _signals():
        est = np.zeros(40)
        for fr in _frames(x)[:N_FRAMES]:
            est = fr["v"] * 1024.0 * sm + est * (1.0 - sm)      # driven by the INTEGER roots: isolates the recursion
            ok = est > 2000.0
            assert np.all(np.abs(fr["est"][ok] - est[ok]) <= 2e-3 * est[ok])
            assert np.all(np.abs(fr["est"][~ok] - est[~ok]) <= 8.0)
            est = fr["est"].copy()                               # resynchronise: bound the per-frame error, not its accumulation


def test_stage_c_subtraction_pcan_log_match_float_definitions():
    worst = []
    for x in _signals() + [edge_case_audio(160 * (N_FRAMES + 3))[i] for i in range(4)]:
        for fr in _frames(x)[:N_FRAMES]:
            want = _stage_c(fr["v"], fr["est"])
            got = fr["feat"].astype(np.float64)
            # integer truncations (>> 10, >> 6, >> 20) matter when an intermediate value is a handful of LSBs -- the soft
            # knee's (snr^2 >> 20) is 0 or 1 right where the float curve crosses 1, i.e. feature 0 versus 133 -- so compare
            # where the knee's output is at least 32 (truncation <= 3 % = 2 LSB of 64 ln)
            sub = np.maximum(fr["v"] * 1024.0 - np.minimum(fr["est"], fr["v"] * 1024.0), 0.0) / 1024.0
            v2 = np.maximum(sub, fr["v"] * 0.05)
            ok = (v2 >= 64.0) & (want >= 64.0 * np.log(8.0 * 32.0))
            # the gain LUT holds integers: at large noise estimates the gain is a single digit and its rounding alone is worth
            # up to 64 * 2 * ln(1 + 0.5 / gain) (the knee squares the SNR)
            gain = np.minimum(32767.0, 2.0 ** 21 * (fr["est"] / 128.0 + 80.0) ** -0.95)
            tol = 4.0 + 128.0 * np.log1p(0.75 / np.maximum(gain, 1.0))
            d = np.abs(got - want)
            ok = ok & (gain >= 16.0)                      # below that the integer gain has less than 5 bits: nothing to compare
            if ok.any():
                assert np.all(d[ok] <= tol[ok]), (float(np.max(d[ok] / tol[ok])), fr["start"])
                worst.append(float(np.max(d[ok][gain[ok] >= 64.0])) if np.any(gain[ok] >= 64.0) else 0.0)
            assert np.all(got[fr["v"] == 0] == 0)
    assert len(worst) > 100 and np.median(worst) <= 2.0


def test_end_to_end_on_well_conditioned_entries():
    """Float model chained end to end (no integer intermediates); compared where the subtraction is well conditioned."""
    W = _mel_weights()
    win = 0.5 - 0.5 * np.cos(2 * np.pi * (np.arange(480) + 0.5) / 480)
    sm = np.where(np.arange(40) % 2 == 1, 0.06, 0.025)
    diffs = []
    for x in _signals():
        est = np.zeros(40)
        for fr in _frames(x)[:N_FRAMES]:
            seg = x[fr["start"]:fr["start"] + 480].astype(np.float64) * win
            v = np.sqrt(W @ (np.abs(np.fft.rfft(seg, 512)) ** 2)) / 8.0
            est = v * 1024.0 * sm + est * (1.0 - sm)
            want = _stage_c(v, est)
            ratio = np.maximum(v * 1024.0 - np.minimum(est, v * 1024.0), 0.0) / np.maximum(v * 1024.0, 1e-9)
            ok = (v >= 200.0) & ((ratio >= 0.4) | (ratio <= 0.02))     # clearly above the estimate, or clearly on the 5 % floor
            diffs.extend(np.abs(fr["feat"][ok] - want[ok]))
    diffs = np.asarray(diffs)
    assert diffs.size > 2000 and np.median(diffs) <= 2.0 and np.percentile(diffs, 95) <= 8.0
