"""Golden for SURVEY.md section 8 row f-4's metric bookkeeping, produced by EXECUTING THE REFERENCE'S OWN
`validate_nonstreaming` (lifted with ast out of /root/reference/microwakeword/train.py:41-163; the module imports
TensorFlow at the top) against stand-ins for the Keras model and the FeatureHandler that return prescribed predictions:
what is pinned is everything train.py computes FROM the tp / fp / fn arrays.

    python tests/golden/make_validation_golden.py      (needs /root/reference)
"""

import ast
import contextlib
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/microwakeword/train.py"


class _T:                      # a tensor with .numpy()
    def __init__(self, a):
        self.a = np.asarray(a, np.float64)

    def numpy(self):
        return self.a


def counts(p, y):
    cut = np.linspace(0.0, 1.0, 101)
    pos = np.asarray(p, np.float32).astype(np.float64)[None, :] > cut[:, None]      # Keras: prediction > threshold
    y = np.asarray(y, bool)
    return (pos & y).sum(1).astype(np.float64), (pos & ~y).sum(1).astype(np.float64), (~pos & y).sum(1).astype(np.float64)


class _Model:
    """model.evaluate accumulates tp / fp / fn across calls unless reset_metrics() runs (train.py:73-84)"""

    def __init__(self, predictions):
        self.pred, self.calls = predictions, 0
        self.reset_metrics()

    def reset_metrics(self):
        self.tp = self.fp = self.fn = np.zeros(101)

    def evaluate(self, x, y, batch_size, return_dict, verbose):
        assert batch_size == 1024 and return_dict
        self.reset_metrics()                                  # Keras' evaluate() resets first; swap_attribute disables it for the 2nd call
        p = self.pred[self.calls]
        self.calls += 1
        tp, fp, fn = counts(p, y.reshape(-1))
        self.tp, self.fp, self.fn = self.tp + tp, self.fp + fp, self.fn + fn
        return dict(accuracy=0.0, recall=0.0, precision=0.0, auc=0.0, loss=0.0, tp=_T(self.tp), fp=_T(self.fp), fn=_T(self.fn))


class _Data:
    def __init__(self, test, ambient, hours):
        self.sets, self.hours = {"testing": test, "testing_ambient": ambient}, hours

    def get_data(self, name, batch_size, features_length, truncation_strategy):
        x, y = self.sets[name]
        return x, np.asarray(y), None

    def get_mode_size(self, mode):
        return 1

    def get_mode_duration(self, mode):
        return self.hours * 3600.0


def main():
    tree = ast.parse(open(REF).read())
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "validate_nonstreaming")

    @contextlib.contextmanager
    def swap_attribute(obj, name, value):
        old = getattr(obj, name)
        setattr(obj, name, value)
        try:
            yield
        finally:
            setattr(obj, name, old)

    ns = {"np": np, "swap_attribute": swap_attribute}
    if not hasattr(np, "trapz"):
        np.trapz = np.trapezoid
    exec(compile(ast.Module(body=[node], type_ignores=[]), REF, "exec"), ns)
    rng = np.random.default_rng(7)
    out = {}
    for case, (n_pos, n_neg, n_amb, hours, sharp) in enumerate(((300, 500, 4000, 1.5, 3.0), (120, 80, 9000, 10.0, 2.5), (50, 50, 2600, 0.2, 4.0))):
        y_test = np.concatenate([np.ones(n_pos, bool), np.zeros(n_neg, bool)])
        p_test = np.where(y_test, rng.beta(sharp, 1.5, y_test.size), rng.beta(1.2, sharp, y_test.size)).astype(np.float32)
        p_amb = rng.beta(0.6, sharp * 1.5, n_amb).astype(np.float32)
        y_amb = np.zeros(n_amb, bool)
        model = _Model([p_test, p_amb])
        # the second evaluate() must ACCUMULATE on top of the first (swap_attribute makes reset_metrics a no-op): emulate by
        # letting evaluate() call self.reset_metrics(), which the reference replaces with a lambda for that call
        data = _Data((np.zeros((y_test.size, 1, 1)), y_test), (np.zeros((n_amb, 1, 1)), y_amb), hours)
        metrics = ns["validate_nonstreaming"]({"batch_size": 128, "spectrogram_length": 204}, data, model, "testing")
        for k in ("recall_at_no_faph", "cutoff_for_no_faph", "ambient_false_positives", "ambient_false_positives_per_hour", "average_viable_recall"):
            out["c%d_%s" % (case, k)] = np.float64(metrics[k])
        out["c%d_p_test" % case], out["c%d_y_test" % case], out["c%d_p_amb" % case], out["c%d_hours" % case] = p_test, y_test, p_amb, hours
        print(case, {k: float(metrics[k]) for k in ("recall_at_no_faph", "cutoff_for_no_faph", "ambient_false_positives_per_hour", "average_viable_recall")})
    np.savez(os.path.join(HERE, "validation_golden.npz"), **out)


if __name__ == "__main__":
    main()
