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


def test_roc_curve_equals_reference_output(gold):
    """generate_roc_curve / false-rejection rates: the oracle restatement AND the product's host function against the output
    of the reference's own generate_roc_curve (three faph profiles: starting above max_faph, below it, reaching zero)."""
    from microwakeword_b200 import detection as G
    g, _ = gold
    for impl in (D, G):
        frr = impl.false_rejection_rates(g["pos_scores"], g["cutoffs"])
        assert np.array_equal(frr, g["frr"])
        for name in "abc":
            x, y, c = impl.generate_roc_curve(g["roc_%s_faph" % name], frr, g["cutoffs"])
            assert np.array_equal(x, g["roc_%s_x" % name]) and np.array_equal(y, g["roc_%s_y" % name]) and np.array_equal(c, g["roc_%s_c" % name]), name
            assert abs(impl.roc_auc(x, y) - float(g["roc_%s_auc" % name])) <= 1e-12
    with pytest.raises(ValueError):
        G.generate_roc_curve([1.0], [0.5, 0.5], [0.0])
    with pytest.raises(ValueError):
        G.false_rejection_rates([], g["cutoffs"])


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


@pytest.mark.gpu
def test_gpu_streaming_model_roc_equals_oracle_composition(torch_cuda):
    """test.py:293-403 end to end on spectrogram tracks: Model.predict_spectrogram (GPU) -> detection kernels -> ROC, against
    the same pipeline composed from the oracle's restatements on the oracle's probabilities (int8 model: bit-exact)."""
    import oracle
    from microwakeword.inference import Model
    from microwakeword_b200 import detection as G
    path = os.path.join(GOLDEN, "okay_nabu_synth_int8.mww")
    feats = np.load(os.path.join(GOLDEN, "config0_features.npy"))                       # uint16 [997, 40]
    ambient = [feats[:600], feats[300:997], feats[100:450]]
    positives = [feats[s:s + 204] for s in (0, 150, 400, 600, 790)]
    model = Model(path)
    got = G.streaming_model_roc(model, ambient, positives, stride=3, window_step_ms=10.0)
    ref = oracle.MixedNet(open(path, "rb").read())        # ONE interpreter for every track, in call order: the reference never
    probs = lambda track: np.asarray(ref.predict_u16(track), np.float32)      # resets it between clips (inference.py:52-64)
    cut = np.arange(0, 1.01, 0.01)
    amb = [D.moving_average(probs(t), 5) for t in ambient]
    faph = D.compute_false_accepts_per_hour(amb, cut, 25, stride=3, step_s=0.01)
    scores = np.asarray([D.positive_score(probs(t), 5, 25) for t in positives], np.float32)
    frr = D.false_rejection_rates(scores[~np.isnan(scores)], cut)
    x, y, c = D.generate_roc_curve(faph, frr, cut)
    assert np.array_equal(got["faph"], faph) and np.array_equal(got["frr"], frr)
    assert np.array_equal(got["roc"][0], x) and np.array_equal(got["roc"][1], y) and np.array_equal(got["roc"][2], c)
    assert abs(got["auc"] - D.roc_auc(x, y)) <= 1e-12
