"""Micro-benchmark of the tcgen05 conv kernels on the layer shapes that dominate the DeepLab step.
GPU-bound timing: a spin kernel holds the stream while the launches are queued, then CUDA events bracket `reps`
back-to-back launches.  Usage: python tools/conv_bench.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from zs3_b200 import kernels as K  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
N = 16
SHAPES = [
    # name, H, Cin, Cout, R, dil
    ("l3.conv2 3x3 256->256 @33", 33, 256, 256, 3, 1),
    ("l3.conv3 1x1 256->1024 @33", 33, 256, 1024, 1, 1),
    ("l3.conv1 1x1 1024->256 @33", 33, 1024, 256, 1, 1),
    ("l4.conv2 3x3 512->512 d4 @33", 33, 512, 512, 3, 4),
    ("aspp 3x3 2048->256 d12 @33", 33, 2048, 256, 3, 12),
    ("dec 3x3 256->256 @129", 129, 256, 256, 3, 1),
    ("l1.conv3 1x1 64->256 @129", 129, 64, 256, 1, 1),
    ("l1.conv1 1x1 256->64 @129", 129, 256, 64, 1, 1),
    ("l1.conv2 3x3 64->64 @129", 129, 64, 64, 3, 1),
    ("l2.conv3 1x1 128->512 @65", 65, 128, 512, 1, 1),
]


def timed(fn):
    fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(int(2e7))  # ~10 ms head start for the host
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


print(f"cluster={os.environ.get('ZS3_CLUSTER', 'default')} reps={reps}")
print("| layer | fprop us | TF/s | fprop+stats us | dgrad us | TF/s | wgrad us | TF/s |")
print("|---|---:|---:|---:|---:|---:|---:|---:|")
for name, H, cin, cout, R, dil in SHAPES:
    pad = dil * (R - 1) // 2
    x = torch.randn(N, H, H, cin, device="cuda").to(torch.bfloat16)
    w = torch.randn(cout, cin, R, R, device="cuda") * 0.05
    wp = K.pack_weight(w, cout, cin)
    wt = K.pack_weight(w, cout, cin, mode=1)
    dy = torch.randn(N, H, H, cout, device="cuda").to(torch.bfloat16)
    y = torch.empty(N, H, H, cout, device="cuda", dtype=torch.bfloat16)
    dx = torch.empty(N, H, H, cin, device="cuda", dtype=torch.bfloat16)
    dw = torch.zeros(cout, R * R, cin, device="cuda")
    stats = torch.zeros(2, cout, dtype=torch.float64, device="cuda")
    fl = 2.0 * N * H * H * cin * cout * R * R
    t_f = timed(lambda: K.conv_fprop([(x, wp)], R, R, 1, pad, dil, cout, out=y))
    t_fs = timed(lambda: K.conv_fprop([(x, wp)], R, R, 1, pad, dil, cout, out=y, stats=(stats[0], stats[1])))
    t_d = timed(lambda: K.conv_fprop([(dy, wt)], R, R, 1, pad, dil, cin, out=dx))
    t_w = timed(lambda: K.conv_wgrad(x, dy, R, R, 1, pad, dil, cin, cout, dw=dw))
    print(f"| {name} | {t_f:.1f} | {fl / t_f / 1e6:.0f} | {t_fs:.1f} | {t_d:.1f} | {fl / t_d / 1e6:.0f} | {t_w:.1f} | "
          f"{fl / t_w / 1e6:.0f} |")
