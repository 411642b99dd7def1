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
