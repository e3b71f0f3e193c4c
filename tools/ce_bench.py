"""Times the fused upsample + cross-entropy loss kernels at the training shape (16 x 21 x 129x129 -> 513x513).
Usage: python tools/ce_bench.py [C] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from zs3_b200.utils.loss import SegmentationLosses  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 21
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
g = torch.Generator().manual_seed(1)
x = (torch.randn(16, 129, 129, 64, generator=g) * 2).to(torch.bfloat16).cuda()
x[..., C:] = 0
x.requires_grad_(True)
t = torch.randint(0, C, (16, 513, 513), generator=g).float()
t[torch.rand(16, 513, 513, generator=g) < 0.02] = 255
t = t.cuda()
losses = SegmentationLosses(weight=None, cuda=True)
for mode in ("1", "0"):
    os.environ["ZS3_CE_BWD_X4"] = mode
    for _ in range(2):
        losses.UpsampledCrossEntropyLoss(x, C, t).backward()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for _ in range(reps):
        e[0].record()
        l = losses.UpsampledCrossEntropyLoss(x, C, t)
        e[1].record()
        l.backward()
        e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1])
        tb += e[1].elapsed_time(e[2])
    print(f"C={C} ZS3_CE_BWD_X4={mode}: forward {tf / reps * 1e3:.1f} us, backward {tb / reps * 1e3:.1f} us (incl. launch gaps)")
