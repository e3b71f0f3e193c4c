"""Generates tests/golden/metrics.npz with the REAL Evaluator of the reference (/root/reference/zs3/utils/metrics.py).
Run in the build container only:  python tests/golden/make_golden_metrics.py"""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_metrics", "/root/reference/zs3/utils/metrics.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

rng = np.random.RandomState(5)
C = 21
seen = [c for c in range(C) if c not in (10, 14)]
unseen = [10, 14]
gt = rng.randint(0, C, size=(3, 48, 40)).astype(np.float32)
gt[gt == 7] = 3                       # class 7 never occurs: NaN rows (nan_to_num -> 0 in the class means)
gt[rng.rand(*gt.shape) < 0.08] = 255  # ignore pixels
pred = np.where(rng.rand(*gt.shape) < 0.7, np.clip(gt, 0, C - 1), rng.randint(0, C, size=gt.shape)).astype(np.int64)
ev = ref.Evaluator(C, seen, unseen)
ev.add_batch(gt[:2], pred[:2])
ev.add_batch(gt[2:], pred[2:])
acc = ev.Pixel_Accuracy()
acc_c = ev.Pixel_Accuracy_Class()
miou = ev.Mean_Intersection_over_Union()
fw = ev.Frequency_Weighted_Intersection_over_Union()
np.savez_compressed(os.path.join(HERE, "metrics.npz"), gt=gt, pred=pred, confusion=ev.confusion_matrix,
                    pixel_acc=np.array(acc), class_acc=np.array([acc_c[0], acc_c[2], acc_c[3]]), class_acc_by_class=acc_c[1],
                    miou=np.array([miou[0], miou[2], miou[3]]), miou_by_class=miou[1], fwiou=np.array(fw))
print("confusion sum", ev.confusion_matrix.sum(), "mIoU", miou[0], miou[2], miou[3])
