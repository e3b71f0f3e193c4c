"""Generates tests/golden/metrics_seen_unseen.npz by running the REAL Evaluator_seen_unseen of the reference
(/root/reference/zs3/utils/metrics.py:88-200, pure numpy) on seeded label / prediction maps.

Run in the build container only:  python tests/golden/make_golden_metrics_seen_unseen.py
"""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_FILE = "/root/reference/zs3/utils/metrics.py"


def load():
    spec = importlib.util.spec_from_file_location("zs3_ref_metrics", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.Evaluator_seen_unseen


def flat(metrics):
    """nested tuples / lists of floats -> 1-D float64 array (NaN kept)"""
    out = []

    def walk(x):
        if isinstance(x, (tuple, list)):
            for y in x:
                walk(y)
        else:
            out.append(float(x))
    walk(metrics)
    return np.array(out, dtype=np.float64)


def main():
    E = load()
    rs = np.random.RandomState(3)
    out = {}
    cases = [(21, [15, 16, 17, 18, 19], 3, 40), (21, [10, 14], 2, 65), (60, [14, 36], 2, 33), (7, [], 2, 20), (5, [4], 1, 16)]
    out["cases"] = np.array([(c, n, hw) for c, _, n, hw in cases])
    for i, (C, unseen, n, hw) in enumerate(cases):
        gts, preds = [], []
        for _ in range(n):
            gt = rs.randint(0, C, size=(hw, hw)).astype(np.float64)
            gt[rs.rand(hw, hw) < 0.07] = 255
            if C > 3:
                gt[gt == 2] = 1                       # a class that never occurs: NaN rows in the per-class metrics
            pred = np.where(rs.rand(hw, hw) < 0.6, np.where(gt == 255, 0, gt), rs.randint(0, C, size=(hw, hw))).astype(np.int64)
            gts.append(gt)
            preds.append(pred)
        out[f"c{i}_gt"], out[f"c{i}_pred"], out[f"c{i}_unseen"] = np.stack(gts), np.stack(preds), np.array(unseen, dtype=np.int64)
        ev = E(C, unseen)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out[f"c{i}_metrics"] = flat(ev.label_accuracy_score(gts, preds))
            out[f"c{i}_metrics_by_class"] = flat(ev.label_accuracy_score(gts, preds, by_class=True))
    np.savez_compressed(os.path.join(HERE, "metrics_seen_unseen.npz"), **out)
    print("wrote metrics_seen_unseen.npz", os.path.getsize(os.path.join(HERE, "metrics_seen_unseen.npz")), "bytes")


if __name__ == "__main__":
    main()
