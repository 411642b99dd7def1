"""Pins the CPU oracle's MixedNet restatements against each other, against structural properties of
the reference graph (streaming == non-streaming once rings are full, README.md:27-28) and against the
committed golden vectors."""

import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN
from microwakeword_b200 import model_file as MF
from oracle import mixednet_ref as R


@pytest.fixture(scope="module")
def f32():
    return MF.load(os.path.join(GOLDEN, "okay_nabu_synth_f32.mww"))


@pytest.fixture(scope="module")
def q8():
    return MF.load(os.path.join(GOLDEN, "okay_nabu_synth_int8.mww"))


def test_appendix_a_numbers():
    a = MF.OKAY_NABU
    assert a.state_elements == 4176 and a.macs_per_step == 24800 and a.first_conv_ring_rows == 2
    assert [a.block_ring_rows(i) for i in range(4)] == [4, 10, 14, 22]
    assert list(a.channel_ksizes(1)) == [7] * 32 + [11] * 32
    assert MF.Arch.decode(a.encode()) == a


def test_container_roundtrip_and_errors(f32):
    blob = MF.write_container(f32)
    back = MF.read_container(blob)
    assert set(back) == set(f32) and all(np.array_equal(back[k], f32[k]) for k in f32)
    with pytest.raises(ValueError):
        MF.read_container(b"garbage")
    with pytest.raises(ValueError):
        oracle.MixedNet(b"MWWB200\0" + b"\0" * 64)


def test_keras_form_equals_folded_form():
    spec = R.OKAY_NABU
    p = R.init_synthetic(spec, 0)
    t = R.fold_bn(spec, p)
    feats = np.load(os.path.join(GOLDEN, "config0_features.npy"))[:150]
    a = np.asarray(R.predict_spectrogram(R.KerasStreamingF32(spec, p), feats))
    b = np.asarray(R.predict_spectrogram(R.FoldedStreamingF32(t), feats))
    assert np.abs(a - b).max() < 1e-5
    # the committed container holds exactly fold_bn(init_synthetic(seed 0))
    g = MF.load(os.path.join(GOLDEN, "okay_nabu_synth_f32.mww"))
    assert all(np.array_equal(g[k], t[k]) for k in t)


def test_streaming_equals_nonstreaming_when_rings_are_full(f32):
    rng = np.random.default_rng(1)
    ff = rng.uniform(0, 26, (297, 40)).astype(np.float32)
    m = R.FoldedStreamingF32(f32)
    logits = np.asarray([m.step(ff[3 * j:3 * j + 3], want_logit=True) for j in range(99)])
    ns = R.nonstreaming_logits(f32, ff[1:])      # streaming windows start at rows 3j-2: one row of offset
    assert len(ns) == 32 and np.abs(logits[-32:] - ns).max() < 1e-4


def test_c_and_numpy_agree_and_match_golden(f32, q8):
    feats = np.load(os.path.join(GOLDEN, "config0_features.npy"))
    c32 = oracle.MixedNet(MF.write_container(f32)).predict_u16(feats)
    n32 = np.asarray(R.predict_spectrogram(R.FoldedStreamingF32(f32), feats), np.float32)
    assert np.abs(c32 - n32).max() < 2e-6
    assert np.array_equal(c32, np.load(os.path.join(GOLDEN, "config0_probs_f32.npy")))
    qm = R.StreamingInt8(q8)
    n8 = np.asarray(R.predict_spectrogram(qm, feats, quantized=True, in_scale=qm.input_scale, in_zp=qm.input_zero_point), np.float32)
    c8 = oracle.MixedNet(MF.write_container(q8)).predict_u16(feats)
    assert np.array_equal(c8, n8) and np.array_equal(c8, np.load(os.path.join(GOLDEN, "config0_probs_int8.npy")))
    assert np.array_equal(oracle.generate_features_for_clip(np.load(os.path.join(GOLDEN, "config0_audio.npy"))), feats)


def test_state_persists_across_calls(f32):
    feats = np.load(os.path.join(GOLDEN, "config0_features.npy"))
    m = oracle.MixedNet(MF.write_container(f32))
    whole = m.predict_u16(feats[:300])
    m.reset()
    a, b = m.predict_u16(feats[:150]), m.predict_u16(feats[150:300])
    assert np.array_equal(np.concatenate([a, b]), whole)


def test_tflite_integer_primitives():
    # QuantizeMultiplier: real = M0 * 2^(shift-31), M0 in [2^30, 2^31)
    for real in (0.75, 0.0003, 1.0, 0.4999999, 3.7):
        m, s = R.quantize_multiplier(real)
        assert (1 << 30) <= m < (1 << 31) and abs(m * 2.0 ** (s - 31) - real) < 1e-9 * max(real, 1)
    # MultiplyByQuantizedMultiplier is round-to-nearest of x*real (ties away via the double rounding)
    rng = np.random.default_rng(0)
    x = rng.integers(-(1 << 20), 1 << 20, 2000)
    for real in (0.0123, 0.6, 0.00071):
        m, s = R.quantize_multiplier(real)
        y = R.mbqm(x, m, s)
        assert np.abs(y - x * real).max() <= 0.51      # double rounding: 0.5 + 2^-(right_shift+1)
    assert int(R.srdhm(-(1 << 31), -(1 << 31))) == (1 << 31) - 1
    assert list(R.rounding_divide_by_pot(np.array([5, -5, 6, -6, 7, -7]), 2)) == [1, -1, 2, -2, 2, -2]
    # quantize_input: truncation toward zero, wrap instead of clamp (inference.py:146-147)
    q = R.quantize_input(np.array([0.0, 0.05, 26.0, 30.0], np.float32), 26 / 255, -128)
    assert list(q) == [-128, -127, 127, -90]     # -127.51 truncates toward zero; 166 wraps to -90


def test_upstream_quantize_multiplier_vectors():
    """Expected pairs of TFLite's own kernels/internal/quantization_util_test.cc
    (QuantizeMultiplierSmallerThanOneExp / QuantizeMultiplierGreaterThanOne); upstream sources are absent here,
    the pairs are restated from knowledge of that file and each is re-derivable as round(frexp(x) * 2^31)."""
    want = {0.25: (1073741824, -1), 0.50 - 5e-9: (2147483627, -1), 0.50 - 1e-10: (1073741824, 0),
            0.50: (1073741824, 0), 0.75: (1610612736, 0), 1 - 1e-9: (2147483646, 0),
            1 + 1e-9: (1073741825, 1), 1.5: (1610612736, 1), 2.0: (1073741824, 2), 3.0: (1610612736, 2),
            4.0: (1073741824, 3), 5.0: (1342177280, 3)}
    for real, pair in want.items():
        assert tuple(int(v) for v in R.quantize_multiplier(real)) == pair, real


def test_int8_reset_state_is_zero_point(q8):
    m = R.StreamingInt8(q8)
    assert np.all(m.st_first == m.zp["in"]) and np.all(m.st_head == m.zp["p4"])
    assert m.zp["in"] == -128 and abs(float(m.input_scale) - 26 / 255) < 1e-6
