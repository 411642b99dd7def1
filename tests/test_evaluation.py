"""ambient splitting (data.py:301-311) and the validation bookkeeping (train.py:41-163).
The golden was produced by executing the reference's own validate_nonstreaming (tests/golden/make_validation_golden.py)."""

import os

import numpy as np
import pytest

from conftest import GOLDEN
from microwakeword_b200 import evaluation as E


def test_validation_bookkeeping_equals_reference_output():
    g = np.load(os.path.join(GOLDEN, "validation_golden.npz"))
    for case in range(3):
        test = E.threshold_counts(g["c%d_p_test" % case], g["c%d_y_test" % case])
        amb = E.threshold_counts(g["c%d_p_amb" % case], np.zeros(g["c%d_p_amb" % case].size, bool))
        m = E.viable_recall_metrics(test["tp"], amb["fp"], test["fn"], float(g["c%d_hours" % case]))
        for k in ("recall_at_no_faph", "cutoff_for_no_faph", "ambient_false_positives", "ambient_false_positives_per_hour", "average_viable_recall"):
            assert abs(float(m[k]) - float(g["c%d_%s" % (case, k)])) <= 1e-12, (case, k)


@pytest.mark.parametrize("rows,length", [(0, 204), (204, 204), (205, 204), (234, 204), (235, 204), (1000, 204), (997, 150)])
def test_split_ambient_is_the_reference_range(rows, length):
    spec = (np.arange(rows * 40) % 65536).astype(np.uint16).reshape(rows, 40)
    got = E.split_ambient(spec, length, 0.01, 3)
    starts = list(range(0, rows - length, int(1000 * 0.01 * 3)))                        # data.py:301-305
    assert got.shape == (len(starts), length, 40) and got.dtype == np.float32
    for w, s in zip(got, starts):
        assert np.array_equal(w, spec[s:s + length].astype(np.float32) * np.float32(0.0390625))
    f = spec.astype(np.float32)
    assert np.array_equal(E.split_ambient(f, length, 0.02, 1), np.stack([f[s:s + length] for s in range(0, rows - length, 20)]) if rows - length > 0
                          else np.zeros((0, length, 40), np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["f32", "int8"])
def test_gpu_nonstreaming_windows_equal_streaming_oracle_last_step(torch_cuda, kind):
    """predict_nonstreaming on ambient windows (persistent 1 024-stream engine, several batches) == the oracle streaming model
    stepped over each window from reset state, last step; then the bookkeeping on top of it runs end to end."""
    import oracle
    from microwakeword.inference import Model
    path = os.path.join(GOLDEN, "okay_nabu_synth_%s.mww" % kind)
    feats = np.load(os.path.join(GOLDEN, "config0_features.npy"))                       # uint16 [997, 40]
    m = Model(path)
    assert m.nonstreaming_length() == 203          # 5 + 3 * 66; the reference's training windows are 204 rows (one spare)
    windows = E.split_ambient(feats, 204, 0.01, 3)                                       # 27 windows
    got = m.predict_nonstreaming(windows, batch_size=8)                                  # 4 batches through one engine
    blob = open(path, "rb").read()
    want = []
    for w in windows:
        ref = oracle.MixedNet(blob)
        rows = np.concatenate([np.zeros((1, 40), np.float32), w])                        # the alignment row (k0 - stride = 2, d = 1)
        u16 = np.round(rows / np.float32(0.0390625)).astype(np.uint16)
        want.append(ref.predict_u16(u16)[-1])
    want = np.asarray(want, np.float32)
    assert got.shape == want.shape and (np.array_equal(got, want) if kind == "int8" else np.abs(got - want).max() <= 1e-5)
    labels = np.arange(len(windows)) % 3 == 0
    metrics = E.validate_nonstreaming(m, windows, labels, ambient_spectrograms=[feats, feats[100:900]], ambient_duration_h=0.01)
    assert set(metrics) >= {"accuracy", "recall", "precision", "average_viable_recall", "cutoff_for_no_faph"}
