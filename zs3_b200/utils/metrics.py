"""Evaluator with the reference's surface (zs3/utils/metrics.py:4-82) on a device-resident confusion matrix.

`add_batch(gt_image, pre_image)` keeps the reference signature (label map + predicted label map);
`add_batch_logits(target, output)` is the fast path for the validation loops (zs3/train_pascal_GMMN.py:358-375):
the argmax over the class axis and the confusion-matrix update happen in one kernel on the logits where they
are, instead of copying them to the host.  The scores are derived from the (tiny) confusion matrix on the host.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib as L


def _ratio(num, den):
    with np.errstate(divide="ignore", invalid="ignore"):
        return num / den


class Evaluator:
    def __init__(self, num_class, seen_classes_idx=None, unseen_classes_idx=None):
        if not 1 <= num_class <= 64:
            raise ValueError("1 <= num_class <= 64")
        self.num_class = num_class
        self.seen_classes_idx = seen_classes_idx
        self.unseen_classes_idx = unseen_classes_idx
        self._conf = None          # device int64 [C, C], created on the first batch's device

    # ---------------------------------------------------------------------------------- accumulation
    def _device_conf(self, dev):
        if self._conf is None:
            self._conf = torch.zeros((self.num_class, self.num_class), dtype=torch.int64, device=dev)
        return self._conf

    @staticmethod
    def _cuda(t):
        if isinstance(t, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(t))
        if not torch.cuda.is_available():
            raise RuntimeError("zs3_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        return t.cuda()

    def add_batch(self, gt_image, pre_image):
        """metrics.py:79-81; numpy arrays (as the unchanged trainer passes them) or tensors"""
        assert tuple(gt_image.shape) == tuple(pre_image.shape)
        gt = self._cuda(gt_image).float().contiguous().view(-1)
        pre = self._cuda(pre_image).to(torch.int32).contiguous().view(-1)
        conf = self._device_conf(gt.device)
        L.check(L.lib().zs3_confusion_from_pred(L.ptr(pre), L.ptr(gt), gt.numel(), self.num_class, L.ptr(conf),
                                                L.stream_ptr()), "zs3_confusion_from_pred")

    def add_batch_logits(self, target, output, want_pred=False):
        """target [B, H, W] float labels, output [B, C, H, W] fp32 logits (CUDA).  Returns the uint8 prediction
        map when want_pred."""
        if not output.is_cuda:
            raise RuntimeError("zs3_b200 runs on CUDA (sm_100a) tensors only; there is no CPU path")
        B, Cn = output.shape[0], output.shape[1]
        if Cn != self.num_class:
            raise ValueError(f"{Cn} logit channels for an evaluator over {self.num_class} classes")
        output = output.float().contiguous()
        target = target.float().contiguous()
        hw = output[0, 0].numel()
        assert target.numel() == B * hw
        pred = torch.empty((B,) + tuple(output.shape[2:]), dtype=torch.uint8, device=output.device) if want_pred else None
        conf = self._device_conf(output.device)
        L.check(L.lib().zs3_argmax_confusion(L.ptr(output), L.ptr(target), B, Cn, hw, L.ptr(pred), L.ptr(conf),
                                             L.stream_ptr()), "zs3_argmax_confusion")
        return pred

    def reset(self):
        if self._conf is not None:
            self._conf.zero_()

    @property
    def confusion_matrix(self):
        """float64 [C, C] like the reference's attribute (rows = ground truth, columns = prediction)"""
        if self._conf is None:
            return np.zeros((self.num_class,) * 2)
        return self._conf.cpu().numpy().astype(np.float64)

    @confusion_matrix.setter
    def confusion_matrix(self, value):
        """the reference's attribute is a plain ndarray callers may assign (metrics.py:9,85)"""
        value = np.asarray(value)
        if value.shape != (self.num_class, self.num_class):
            raise ValueError(f"confusion matrix must be [{self.num_class}, {self.num_class}]")
        if not torch.cuda.is_available():
            raise RuntimeError("zs3_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        dev = self._conf.device if self._conf is not None else torch.device("cuda")
        self._conf = torch.from_numpy(np.ascontiguousarray(value).astype(np.int64)).to(dev)

    # ---------------------------------------------------------------------------------------- scores
    def _split(self):
        return bool(self.seen_classes_idx) and bool(self.unseen_classes_idx)

    def Pixel_Accuracy(self):
        cm = self.confusion_matrix
        d = np.diag(cm)
        acc = _ratio(d.sum(), cm.sum())
        if not self._split():
            return acc
        s, u = self.seen_classes_idx, self.unseen_classes_idx
        return acc, _ratio(d[s].sum(), cm[s, :].sum()), _ratio(d[u].sum(), cm[u, :].sum())

    def Pixel_Accuracy_Class(self):
        cm = self.confusion_matrix
        by_class = _ratio(np.diag(cm), cm.sum(axis=1))
        mean = lambda v: np.nanmean(np.nan_to_num(v))  # noqa: E731  (metrics.py:28: NaN rows count as 0)
        if not self._split():
            return mean(by_class), by_class
        return mean(by_class), by_class, mean(by_class[self.seen_classes_idx]), mean(by_class[self.unseen_classes_idx])

    def _iou(self, cm):
        d = np.diag(cm)
        return _ratio(d, cm.sum(axis=1) + cm.sum(axis=0) - d)

    def Mean_Intersection_over_Union(self):
        iou = self._iou(self.confusion_matrix)
        mean = lambda v: np.nanmean(np.nan_to_num(v))  # noqa: E731
        if not self._split():
            return mean(iou), iou
        return mean(iou), iou, mean(iou[self.seen_classes_idx]), mean(iou[self.unseen_classes_idx])

    def Frequency_Weighted_Intersection_over_Union(self):
        cm = self.confusion_matrix
        freq = _ratio(cm.sum(axis=1), cm.sum())
        iou = self._iou(cm)

        def fw(idx):
            f, i = (freq, iou) if idx is None else (freq[idx], iou[idx])
            return (f[f > 0] * i[f > 0]).sum()

        if not self._split():
            return fw(None)
        return fw(None), fw(self.seen_classes_idx), fw(self.unseen_classes_idx)


class Evaluator_seen_unseen:
    """zs3/utils/metrics.py:88-200 (used by eval_pascal.py:83 / eval_context.py).  Same constructor and
    `label_accuracy_score(label_trues, label_preds, by_class=False)` return structure.  The reference builds up to
    3 + num_class masked histograms per image on the host; every one of those masks selects ROWS of the one
    confusion matrix (they only test the ground-truth label), so here the matrix is accumulated once on the device
    (`zs3_confusion_from_pred`) and the seen / unseen / per-class histograms are row selections of it."""

    def __init__(self, num_class, unseen_classes_idx):
        self.num_class = num_class
        self.unseen_classes_idx = unseen_classes_idx

    @staticmethod
    def _hist_to_metrics(hist):
        """metrics.py:127-139: (overall acc, mean class acc, mean IoU, frequency-weighted IoU); NaN classes skipped"""
        d = np.diag(hist)
        total = hist.sum()
        acc = 0.0 if total == 0 else d.sum() / total
        acc_cls = np.nanmean(_ratio(d, hist.sum(axis=1)))
        iu = _ratio(d, hist.sum(axis=1) + hist.sum(axis=0) - d)
        freq = _ratio(hist.sum(axis=1), total)
        pos = freq > 0
        return acc, acc_cls, np.nanmean(iu), (freq[pos] * iu[pos]).sum()

    def _rows(self, hist, rows):
        out = np.zeros_like(hist)
        rows = [r for r in rows if 0 <= r < self.num_class]
        out[rows] = hist[rows]
        return out

    def label_accuracy_score(self, label_trues, label_preds, by_class=False):
        ev = Evaluator(self.num_class)
        for lt, lp in zip(label_trues, label_preds):
            ev.add_batch(np.asarray(lt).reshape(-1), np.asarray(lp).reshape(-1))
        hist = ev.confusion_matrix
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)  # nanmean of an all-NaN slice, as in the reference
            metrics = self._hist_to_metrics(hist)
            if self.unseen_classes_idx:
                unseen = list(self.unseen_classes_idx)
                seen = [c for c in range(self.num_class) if c not in unseen]
                metrics = metrics, self._hist_to_metrics(self._rows(hist, seen)), self._hist_to_metrics(self._rows(hist, unseen))
            if by_class:
                return metrics, [self._hist_to_metrics(self._rows(hist, [c])) for c in range(self.num_class)]
        return metrics
