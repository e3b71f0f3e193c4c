"""Times the BatchNorm streaming kernels on the DeepLab activation shapes (bs=16, 513x513): CUDA events over `reps`
back-to-back launches (warm L2 for the small tensors, like in the real step right after the producing conv)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from zs3_b200 import kernels as K  # noqa: E402

N = 16
SHAPES = [("stem 64@257", 257, 64), ("l1 64@129", 129, 64), ("l1 256@129", 129, 256), ("l2 128@65", 65, 128),
          ("l2 512@65", 65, 512), ("l3 256@33", 33, 256), ("l3 1024@33", 33, 1024), ("l4 2048@33", 33, 2048)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(2_000_000)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


print("| tensor | MB | stats us | apply(fused finalize) us | apply(precomputed) us | apply+res us | bwd_reduce us | bwd_apply us |")
print("|---|---:|---:|---:|---:|---:|---:|---:|")
for name, hw, c in SHAPES:
    dev = "cuda"
    y = torch.randn(N, hw, hw, c, device=dev).to(torch.bfloat16)
    res = torch.randn_like(y)
    out = torch.empty_like(y)
    dout = torch.randn_like(y)
    dy = torch.empty_like(y)
    mb = y.numel() * 2 / 1e6
    stats = torch.zeros(2, c, dtype=torch.float64, device=dev)
    other = torch.zeros(2, c, dtype=torch.float64, device=dev)
    K.bn_stats(y, stats)
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    coef = torch.empty(4, c, device=dev)
    fin = dict(stats=(stats[0], stats[1]), count=N * hw * hw, gamma=gamma, beta=beta, eps=1e-5, momentum=0.1,
               running_mean=rm, running_var=rv, c_real=c, coef=coef, reset=(other[0], other[1], c))
    t_stats = timed(lambda: K.bn_stats(y, other))
    t_apply_f = timed(lambda: K.bn_apply(y, None, None, True, out=out, finalize=fin))
    t_apply_p = timed(lambda: K.bn_apply(y, coef[0], coef[1], True, out=out))
    t_apply_r = timed(lambda: K.bn_apply(y, coef[0], coef[1], True, residual=res, out=out))
    sums = torch.zeros(2, c, dtype=torch.float64, device=dev)
    lib_reduce = lambda: K.bn_backward(dout, out, y, coef[2], coef[3], coef[0], True, dy=dy, shift=coef[1], sums=sums)  # noqa: E731
    t_bwd = timed(lib_reduce)  # reduce + apply together
    print(f"| {name} | {mb:.1f} | {t_stats:.1f} | {t_apply_f:.1f} | {t_apply_p:.1f} | {t_apply_r:.1f} | {t_bwd:.1f} (both) | |")
