"""CPU: the host logic of zs3_b200.utils.metrics.Evaluator_seen_unseen (row selections of ONE confusion matrix instead of
the reference's 3 + num_class masked histograms per image, zs3/utils/metrics.py:88-200) against golden vectors produced
by the reference class (tests/golden/make_golden_metrics_seen_unseen.py).  The device part (confusion-matrix
accumulation, zs3_confusion_from_pred) is replaced by its numpy definition here; tests/test_metrics_gpu.py holds the
kernel against the same definition."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "metrics_seen_unseen.npz"))


def _flat(metrics):
    out = []

    def walk(x):
        if isinstance(x, (tuple, list)):
            for y in x:
                walk(y)
        else:
            out.append(float(x))
    walk(metrics)
    return np.array(out, dtype=np.float64)


class _HostEvaluator:
    """numpy stand-in for Evaluator: conf[gt][pred] += 1 for 0 <= gt < C (metrics.py:73-77)"""

    def __init__(self, num_class, *a, **k):
        self.num_class = num_class
        self.confusion_matrix = np.zeros((num_class, num_class), dtype=np.int64)

    def add_batch(self, gt, pred):
        gt, pred = np.asarray(gt).reshape(-1), np.asarray(pred).reshape(-1)
        m = (gt >= 0) & (gt < self.num_class)
        self.confusion_matrix += np.bincount(self.num_class * gt[m].astype(int) + pred[m].astype(int),
                                             minlength=self.num_class ** 2).reshape(self.num_class, self.num_class)


def test_seen_unseen_scores_match_reference_golden(monkeypatch):
    from zs3_b200.utils import metrics as M
    monkeypatch.setattr(M, "Evaluator", _HostEvaluator)
    for i, (C, n, hw) in enumerate(GOLD["cases"]):
        unseen = [int(u) for u in GOLD[f"c{i}_unseen"]]
        ev = M.Evaluator_seen_unseen(int(C), unseen)
        gts, preds = list(GOLD[f"c{i}_gt"]), list(GOLD[f"c{i}_pred"])
        for by_class, key in ((False, "metrics"), (True, "metrics_by_class")):
            got = _flat(ev.label_accuracy_score(gts, preds, by_class=by_class))
            ref = GOLD[f"c{i}_{key}"]
            assert got.shape == ref.shape, (i, key, got.shape, ref.shape)
            assert np.array_equal(np.isnan(got), np.isnan(ref)), (i, key)
            assert np.allclose(got[~np.isnan(ref)], ref[~np.isnan(ref)], rtol=1e-12, atol=0), (i, key)
