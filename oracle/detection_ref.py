"""oracle/detection_ref.py -- CPU ORACLE (test infrastructure) for the detection post-processing that
follows the streaming model in the reference's evaluation harness:

  * moving average ......... microwakeword/test.py:337-341 (`sliding_window_view(p, L).mean(axis=-1)`, L = 5, :301)
  * false accepts per hour .. microwakeword/test.py:94-137 (`compute_false_accepts_per_hour`): per cutoff, a
                              detection fires when the cooldown is 0 and prob > cutoff; the cooldown starts at
                              `ignore_slices_after_accept` for every track and is re-armed by each detection
  * positive-sample score ... microwakeword/test.py:364-373 (max of the moving average after skipping the
                              first `ignore_slices_after_accept` probabilities)

PARITY PINNED: tests/golden/detection_golden.npz was produced by executing the reference's own function
(lifted from /root/reference/microwakeword/test.py with ast, see tests/golden/make_detection_golden.py), and
tests/test_detection.py checks this restatement against it bit for bit.
"""

import numpy as np


def moving_average(probs, window: int = 5) -> np.ndarray:
    """float32 sequential sum of `window` consecutive values divided by window (what NumPy's mean over the
    last axis of the sliding-window view computes for float32)."""
    p = np.asarray(probs, np.float32)
    n = p.shape[0] - window + 1
    if n <= 0:
        return np.zeros(0, np.float32)
    acc = p[0:n].copy()
    for j in range(1, window):
        acc = (acc + p[j:j + n]).astype(np.float32)
    return (acc / np.float32(window)).astype(np.float32)


def false_accept_counts(track, cutoffs, ignore_slices_after_accept: int) -> np.ndarray:
    """Detections per cutoff for one track of (moving-average) probabilities; test.py:120-135."""
    cutoffs = np.asarray(cutoffs, np.float64)
    counts = np.zeros(cutoffs.shape[0], np.int64)
    cooldown = np.full(cutoffs.shape[0], ignore_slices_after_accept, np.int64)
    for p in np.asarray(track, np.float32):
        cooldown = np.maximum(cooldown - 1, 0)
        fire = (cooldown == 0) & (np.float64(p) > cutoffs)
        counts += fire
        cooldown[fire] = ignore_slices_after_accept
    return counts


def compute_false_accepts_per_hour(streaming_probabilities_list, cutoffs, ignore_slices_after_accept=75, stride=1, step_s=0.02):
    cutoffs = np.asarray(cutoffs, np.float64)
    total = np.zeros(cutoffs.shape[0])
    duration_h = 0
    for track in streaming_probabilities_list:
        duration_h += len(track) * stride * step_s / 3600.0
        total += false_accept_counts(track, cutoffs, ignore_slices_after_accept)
    return total / duration_h


def positive_score(probs, window: int = 5, ignore_slices_after_accept: int = 25) -> np.float32:
    m = moving_average(np.asarray(probs, np.float32)[ignore_slices_after_accept:], window)
    return np.float32(np.nan) if m.size == 0 else np.float32(m.max())


def false_rejection_rates(positive_scores, cutoffs) -> np.ndarray:
    """test.py:376-381: share of positive samples whose score does NOT exceed each cutoff (strict >)."""
    s = np.asarray(positive_scores, np.float64)
    return np.asarray([1 - int((s > c).sum()) / len(s) for c in np.asarray(cutoffs, np.float64)], np.float64)


def generate_roc_curve(false_accepts_per_hour, false_rejections, cutoffs, max_faph: float = 2.0):
    """test.py:140-204 restated: ROC coordinates (faph ascending) from per-cutoff faph / false-rejection rates.

    Quirks of the reference that a drop-in keeps: the interpolated first point takes BOTH ordinates from the last
    cutoff above max_faph (test.py:168-171 reads index - 1 twice, so it is that cutoff's rejection rate, no interpolation
    in y), the interpolation constant is the literal 2.0 rather than max_faph (:173), its cutoff is the midpoint of the
    two neighbouring cutoffs (:174-176), a repeated faph keeps only its first (= lowest-cutoff) point (:190-196), and a
    curve that never reaches 0 faph gets the closing point (0, 1) with cutoff 0 (:198-202)."""
    faph = np.asarray(false_accepts_per_hour, np.float64)
    frr = np.asarray(false_rejections, np.float64)
    cut = np.asarray(cutoffs, np.float64)
    if faph[0] > max_faph:
        k = 1
        while faph[k] > max_faph:
            k += 1
        x0, x1, y = faph[k - 1], faph[k], frr[k - 1]
        first = ((y * (x1 - 2.0) + y * (2.0 - x0)) / (x1 - x0), (cut[k] + cut[k - 1]) / 2.0)
    else:
        k = 0
        first = (frr[0], cut[0])
    xs, ys, cs = [max_faph], [first[0]], [first[1]]
    for i in range(k, len(frr)):
        if faph[i] != xs[-1]:
            xs.append(faph[i]); ys.append(frr[i]); cs.append(cut[i])
    if xs[-1] > 0:
        xs.append(0.0); ys.append(1.0); cs.append(0.0)
    return np.asarray(xs[::-1]), np.asarray(ys[::-1]), np.asarray(cs[::-1])


def roc_auc(x, y) -> float:
    """test.py:391 np.trapz(y, x)"""
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    return float(np.sum((x[1:] - x[:-1]) * (y[1:] + y[:-1]) / 2.0))
