"""SegmentationLosses / GMMNLoss with the reference's interface (zs3/utils/loss.py), on fused CUDA kernels."""
import torch

from .. import functional as ZF


class SegmentationLosses:
    """zs3/utils/loss.py:5-81"""

    def __init__(self, weight=None, size_average=True, batch_average=True, ignore_index=255, cuda=False):
        self.ignore_index = ignore_index
        self.weight = weight
        self.size_average = size_average
        self.batch_average = batch_average
        self.cuda = cuda

    def build_loss(self, mode="ce"):
        modes = {"ce": self.CrossEntropyLoss, "focal": self.FocalLoss, "ce_finetune": self.CrossEntropyLossFinetune}
        if mode not in modes:
            raise NotImplementedError(mode)
        return modes[mode]

    def _ce(self, logit, target, weight, div=1.0):
        if not self.size_average:
            raise NotImplementedError("size_average=False is not used by any reference trainer")
        if not logit.is_cuda:
            raise RuntimeError("zs3_b200 losses run on CUDA tensors only")
        w = weight
        if w is not None and w.device != logit.device:
            w = w.to(logit.device)
        return ZF.CrossEntropy.apply(logit, target, w, self.ignore_index, float(div))

    def CrossEntropyLoss(self, logit, target):
        # the division by the batch size (loss.py:43-44) is folded into the kernel
        return self._ce(logit, target, self.weight, logit.shape[0] if self.batch_average else 1.0)

    def UpsampledCrossEntropyLoss(self, scores, num_classes, target):
        """== CrossEntropyLoss(F.interpolate(scores, size=target.shape[-2:], bilinear, align_corners=True), target)
        for NHWC bf16 low-resolution `scores` (DeepLab.forward_scores): deeplab.py:44 fused into loss.py:31-46."""
        if not self.size_average:
            raise NotImplementedError("size_average=False is not used by any reference trainer")
        w = self.weight
        if w is not None and w.device != scores.device:
            w = w.to(scores.device)
        div = scores.shape[0] if self.batch_average else 1.0
        return ZF.UpsampleCrossEntropy.apply(scores, num_classes, target, w, self.ignore_index, float(div))

    def CrossEntropyLossFinetune(self, logit, target):
        return self._ce(logit, target, None, logit.shape[0] if self.batch_average else 1.0)

    def FocalLoss(self, logit, target, gamma=2, alpha=0.5):
        n = logit.shape[0]
        logpt = -self._ce(logit, target, self.weight)
        pt = torch.exp(logpt)
        if alpha is not None:
            logpt = logpt * alpha
        loss = -((1 - pt) ** gamma) * logpt
        return loss / n if self.batch_average else loss


class GMMNLoss:
    """zs3/utils/loss.py:84-115.  build_loss() returns a callable (gen_samples, x) -> differentiable scalar."""

    def __init__(self, sigma=[2, 5, 10, 20, 40, 80], cuda=False):
        self.sigma = sigma
        self.cuda = cuda

    def build_loss(self):
        return self.moment_loss

    def get_scale_matrix(self, M, N):
        """kept for API compatibility (loss.py:92-97); the kernel applies the same [+1/N]*N ++ [-1/M]*M weights"""
        s1 = torch.ones((N, 1)) * 1.0 / N
        s2 = torch.ones((M, 1)) * -1.0 / M
        if self.cuda:
            s1, s2 = s1.cuda(), s2.cuda()
        return torch.cat((s1, s2), 0)

    def moment_loss(self, gen_samples, x):
        if not gen_samples.is_cuda:
            raise RuntimeError("zs3_b200 losses run on CUDA tensors only")
        from ..gmmn_ops import MomentLoss
        return MomentLoss.apply(gen_samples, x, tuple(self.sigma))
