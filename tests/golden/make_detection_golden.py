"""Golden vectors for the detection post-processing row, produced by EXECUTING THE
REFERENCE'S OWN CODE: `compute_false_accepts_per_hour` is lifted verbatim (ast) out of
/root/reference/microwakeword/test.py:94-137 -- the module itself cannot be imported because it imports
TensorFlow at the top -- and run with NumPy; the moving average is the exact expression of
test.py:337-341 / :364-373.  Unlike the frontend / TFLite rows, parity for this row is therefore PINNED.

    python tests/golden/make_detection_golden.py      (needs /root/reference; run in the authoring container)
"""

import ast
import os

import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/microwakeword/test.py"


def reference_function(name):
    src = open(REF).read()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"np": np, "List": list}
    exec(compile(ast.Module(body=[node], type_ignores=[]), REF, "exec"), ns)
    return ns[name]


def main():
    faph_ref = reference_function("compute_false_accepts_per_hour")
    rng = np.random.default_rng(42)
    lengths = [400, 37, 29, 1000, 5, 260]
    tracks = []
    for i, n in enumerate(lengths):
        base = rng.beta(0.4, 2.5, n).astype(np.float32)                 # mostly low, occasional peaks
        bursts = rng.integers(0, n, max(n // 60, 1))
        for b in bursts:
            base[b:b + 6] = np.float32(rng.uniform(0.6, 1.0))
        tracks.append(list(base))                                         # Model.predict_spectrogram returns a Python list
    window, ignore, stride, step_s = 5, 25, 3, 0.01                      # test.py:301-302, notebook stride 3 / 10 ms
    moving = [sliding_window_view(t, window).mean(axis=-1) for t in tracks]           # test.py:337-341
    cutoffs = np.arange(0, 1.01, 0.01)                                                  # test.py:343
    faph = faph_ref(moving, cutoffs, ignore, stride=stride, step_s=step_s)             # test.py:346-352
    pos_max = np.asarray([np.max(sliding_window_view(t[ignore:], window).mean(axis=-1)) if len(t) - ignore >= window else np.nan
                          for t in tracks], np.float32)                                # test.py:364-373
    flat = np.concatenate([np.asarray(t, np.float32) for t in tracks])
    # false-rejection side + ROC curve (test.py:375-388), again by executing the reference's own generate_roc_curve
    roc_ref = reference_function("generate_roc_curve")
    pos = [float(v) for v in pos_max if not np.isnan(v)] + [float(v) for v in rng.beta(2.0, 1.2, 40)]      # positive-sample scores
    frr = []
    for cutoff in cutoffs:                                                                 # test.py:376-381, verbatim arithmetic
        true_accepts = sum(i > cutoff for i in pos)
        frr.append(1 - true_accepts / len(pos))
    roc = {}
    for name, f in (("a", faph), ("b", faph * 0.01), ("c", np.maximum(faph - faph[60], 0.0))):   # above / below max_faph, reaching 0
        x, y, c = roc_ref(false_accepts_per_hour=f, false_rejections=frr, cutoffs=cutoffs)
        roc["roc_%s_faph" % name], roc["roc_%s_x" % name], roc["roc_%s_y" % name], roc["roc_%s_c" % name] = f, x, y, c
        roc["roc_%s_auc" % name] = np.trapz(y, x) if hasattr(np, "trapz") else np.trapezoid(y, x)     # test.py:391
    np.savez(os.path.join(HERE, "detection_golden.npz"), probs=flat, lengths=np.asarray(lengths, np.int32),
             moving=np.concatenate(moving).astype(np.float32), cutoffs=cutoffs, faph=faph, pos_max=pos_max,
             window=window, ignore=ignore, stride=stride, step_s=step_s, pos_scores=np.asarray(pos, np.float64), frr=np.asarray(frr, np.float64), **roc)
    print("tracks", lengths, "faph[0,50,90,100] =", faph[[0, 50, 90, 100]], "pos_max", pos_max)
    print("roc points", {k: v.shape for k, v in roc.items() if k.endswith("_x")}, "auc", [float(roc["roc_%s_auc" % n]) for n in "abc"])


if __name__ == "__main__":
    main()
