"""Detection post-processing.  The golden file was produced by executing the REFERENCE'S OWN
`compute_false_accepts_per_hour` (tests/golden/make_detection_golden.py), so this row's parity is pinned:
  * CPU: the oracle restatement (oracle/detection_ref.py) equals the reference output bit for bit;
  * GPU: the CUDA kernels behind include/mww.h equal it bit for bit too."""

import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import detection_ref as D


@pytest.fixture(scope="module")
def gold():
    g = np.load(os.path.join(GOLDEN, "detection_golden.npz"))
    lengths = g["lengths"]
    offs = np.concatenate([[0], np.cumsum(lengths)])
    tracks = [g["probs"][offs[i]:offs[i + 1]] for i in range(len(lengths))]
    return g, tracks


def test_oracle_equals_reference_output(gold):
    g, tracks = gold
    window, ignore = int(g["window"]), int(g["ignore"])
    moving = [D.moving_average(t, window) for t in tracks]
    assert np.array_equal(np.concatenate(moving), g["moving"])
    faph = D.compute_false_accepts_per_hour(moving, g["cutoffs"], ignore, stride=int(g["stride"]), step_s=float(g["step_s"]))
    assert np.array_equal(faph, g["faph"])
    pos = np.asarray([D.positive_score(t, window, ignore) for t in tracks], np.float32)
    assert np.array_equal(np.isnan(pos), np.isnan(g["pos_max"]))
    assert np.array_equal(pos[~np.isnan(pos)], g["pos_max"][~np.isnan(pos)])


def test_cooldown_rule_small_cases():
    # a detection needs `ignore` warm-up slices, then re-arms the cooldown (test.py:116-135)
    ones = np.ones(10, np.float32)
    assert list(D.false_accept_counts(ones, [0.5], 3)) == [3]        # fires at i = 2, 5, 8
    assert list(D.false_accept_counts(ones, [0.5], 11)) == [0]
    assert list(D.false_accept_counts(ones, [1.0], 1)) == [0]        # strictly greater
    assert D.moving_average([1, 2, 3], 5).size == 0


@pytest.mark.gpu
def test_gpu_detection_equals_reference_output(gold):
    from microwakeword_b200 import detection as G
    g, tracks = gold
    window, ignore = int(g["window"]), int(g["ignore"])
    moving = G.moving_average(tracks, window)
    assert np.array_equal(np.concatenate(moving), g["moving"])
    # the reference call site: averaged tracks in, window = 1
    faph = G.compute_false_accepts_per_hour(moving, g["cutoffs"], ignore, stride=int(g["stride"]), step_s=float(g["step_s"]))
    assert np.array_equal(faph, g["faph"])
    # fused variant: raw probabilities in, moving average inside the kernel
    faph2 = G.compute_false_accepts_per_hour(tracks, g["cutoffs"], ignore, stride=int(g["stride"]), step_s=float(g["step_s"]), window=window)
    assert np.array_equal(faph2, g["faph"])
    pos = G.positive_scores(tracks, window, ignore)
    assert np.array_equal(np.isnan(pos), np.isnan(g["pos_max"]))
    assert np.array_equal(pos[~np.isnan(pos)], g["pos_max"][~np.isnan(pos)])


@pytest.mark.gpu
def test_gpu_detection_on_model_output():
    """End of the chain: probabilities straight from the engine (device tensors) into the detection kernels."""
    import torch
    from microwakeword_b200 import detection as G
    from microwakeword_b200.engine import StreamEngine
    audio = np.load(os.path.join(GOLDEN, "batch_audio.npy"))
    eng = StreamEngine(os.path.join(GOLDEN, "okay_nabu_synth_int8.mww"), n_streams=audio.shape[0])
    probs = eng.predict_clip(torch.from_numpy(audio).cuda())
    tracks = [probs[i] for i in range(probs.shape[0])]
    got = G.false_accept_counts(tracks, np.arange(0, 1.01, 0.01), 4, window=5)
    want = np.stack([D.false_accept_counts(D.moving_average(t.cpu().numpy(), 5), np.arange(0, 1.01, 0.01), 4) for t in tracks])
    assert np.array_equal(got, want)
