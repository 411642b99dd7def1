"""Non-streaming batched evaluation around the clip kernel: the data side of
``microwakeword.data.FeatureHandler.get_data(..., truncation_strategy="split")`` (data.py:301-311) and the metric
bookkeeping of ``microwakeword.train.validate_nonstreaming`` (train.py:41-163), with ``model.evaluate(batch_size=1024)``
replaced by ``Model.predict_nonstreaming`` (the GPU clip kernel evaluated on whole windows).

Training itself (train.py:166-), the Keras metric objects and the feature store are out of scope; these functions take
plain arrays so the reference's harness can call them with what it already has."""

from __future__ import annotations

import numpy as np

from .model_file import FEATURE_SCALE


def split_ambient(spectrogram: np.ndarray, features_length: int, step_s: float = 0.01, stride: int = 3):
    """data.py:297-311, truncation_strategy == "split": windows of `features_length` rows starting every
    int(1000 * step * stride) rows (30 rows = 300 ms for the 10 ms / stride-3 models), for start in
    range(0, T - features_length, hop) -- so a window that would end exactly at the last row is NOT produced.
    uint16 spectrograms are scaled by 0.0390625 first (data.py:297-298).  Returns float32 [n, features_length, 40]."""
    spec = np.asarray(spectrogram)
    if np.issubdtype(spec.dtype, np.uint16):
        spec = spec.astype(np.float32) * np.float32(FEATURE_SCALE)
    hop = int(1000 * step_s * stride)
    if hop < 1:
        raise ValueError("range() arg 3 must not be zero")          # what the reference's range() raises
    starts = range(0, spec.shape[0] - features_length, hop)
    if len(starts) == 0:
        return np.zeros((0, features_length, spec.shape[1]), np.float32)
    return np.stack([spec[s:s + features_length] for s in starts]).astype(np.float32, copy=False)


def threshold_counts(predictions, labels, cutoffs=None):
    """What the Keras TruePositives / FalsePositives / FalseNegatives metrics hold for thresholds np.linspace(0, 1, 101)
    (train.py:44-56): a prediction is positive when it is strictly greater than the threshold."""
    cut = np.linspace(0.0, 1.0, 101) if cutoffs is None else np.asarray(cutoffs, np.float64)
    p = np.asarray(predictions, np.float32).reshape(-1).astype(np.float64)
    y = np.asarray(labels).reshape(-1).astype(bool)
    pos = p[None, :] > cut[:, None]
    return dict(tp=(pos & y[None, :]).sum(1).astype(np.float64), fp=(pos & ~y[None, :]).sum(1).astype(np.float64),
                fn=(~pos & y[None, :]).sum(1).astype(np.float64), cutoffs=cut)


def viable_recall_metrics(tp, fp_ambient, fn, ambient_duration_h: float) -> dict:
    """train.py:99-161: recall / false accepts per hour per cutoff, the cutoff with no false accepts, and the average recall
    over 0..2 false accepts per hour (trapezoid of the recall-vs-faph curve, interpolated at 2 faph, divided by 2)."""
    tp, fp, fn = (np.asarray(v, np.float64) for v in (tp, fp_ambient, fn))
    recall = tp / (tp + fn)
    faph = fp / ambient_duration_h
    cutoffs = np.linspace(0.0, 1.0, 101)
    cutoff_no_faph, recall_no_faph = 1.0, None
    for i, c in enumerate(cutoffs):
        if faph[i] == 0:
            cutoff_no_faph, recall_no_faph = c, recall[i]
            break
    if recall_no_faph is None:
        raise UnboundLocalError("no cutoff reaches 0 false accepts per hour (the reference fails here as well, train.py:158)")
    if faph[0] > 2:
        k = 1
        while faph[k] > 2:
            k += 1
        x0, y0, x1, y1 = faph[k - 1], recall[k - 1], faph[k], recall[k]
        first = (y0 * (x1 - 2.0) + y1 * (2.0 - x0)) / (x1 - x0)
    else:
        k, first = 0, recall[0]
    xs, ys = [2.0], [first]
    for i in range(k, len(recall)):
        if faph[i] != xs[-1]:
            xs.append(faph[i])
            ys.append(recall[i])
    x, y = np.asarray(xs[::-1]), np.asarray(ys[::-1])
    average = float(np.sum((x[1:] - x[:-1]) * (y[1:] + y[:-1]) / 2.0)) / 2.0
    return dict(recall_at_no_faph=recall_no_faph, cutoff_for_no_faph=cutoff_no_faph, ambient_false_positives=fp[50],
                ambient_false_positives_per_hour=faph[50], average_viable_recall=average)


def validate_nonstreaming(model, test_fingerprints, test_ground_truth, ambient_spectrograms=None, ambient_duration_h: float = 0.0,
                          features_length: int | None = None, step_s: float = 0.01, stride: int = 3, batch_size: int = 1024) -> dict:
    """train.py:41-163 on arrays: `test_fingerprints` [n, features_length, 40] with boolean labels, and (optionally) the
    ambient set as a list of long spectrograms that is split like data.py:301-311.  Accuracy / recall / precision come from
    the 0.5 threshold (Keras defaults); auc and loss are Keras-internal and not reproduced."""
    features_length = features_length or model.nonstreaming_length()
    probs = model.predict_nonstreaming(np.asarray(test_fingerprints), batch_size=batch_size)
    counts = threshold_counts(probs, test_ground_truth)
    y = np.asarray(test_ground_truth).reshape(-1).astype(bool)
    tp, fp, fn = counts["tp"][50], counts["fp"][50], counts["fn"][50]
    tn = float((~y).sum()) - fp
    metrics = dict(accuracy=(tp + tn) / max(len(y), 1), recall=tp / max(tp + fn, 1.0), precision=tp / max(tp + fp, 1.0),
                   recall_at_no_faph=0, cutoff_for_no_faph=0, ambient_false_positives=0, ambient_false_positives_per_hour=0,
                   average_viable_recall=0)
    if ambient_spectrograms:
        windows = [split_ambient(sp, features_length, step_s, stride) for sp in ambient_spectrograms]
        windows = np.concatenate([w for w in windows if len(w)], 0)
        amb = threshold_counts(model.predict_nonstreaming(windows, batch_size=batch_size), np.zeros(len(windows), bool))
        # the reference accumulates both sets in one metric object and subtracts the test set's false positives (:95-97)
        metrics.update(viable_recall_metrics(counts["tp"], amb["fp"], counts["fn"], ambient_duration_h))
    return metrics
