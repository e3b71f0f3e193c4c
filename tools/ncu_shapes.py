"""Launches one instance of each hot kernel on its dominant shape (for `ncu --set full`):
decoder 3x3 256->256 @129 (fprop, dgrad, wgrad), layer3 3x3 and 1x1 @33, and the BatchNorm passes on [16,129,129,256]."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from zs3_b200 import kernels as K  # noqa: E402

N = 16


def conv_set(H, cin, cout, R, dil):
    pad = dil * (R - 1) // 2
    x = torch.randn(N, H, H, cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(cout, cin, R, R, device="cuda") * 0.05).contiguous(memory_format=torch.channels_last)
    wp = K.pack_weight(w, cout, cin)
    dy = torch.randn(N, H, H, cout, device="cuda").to(torch.bfloat16)
    g = torch.zeros_like(w)
    stats = torch.zeros(2, cout, dtype=torch.float64, device="cuda")
    for _ in range(2):  # first round warms tensor maps / caches, second is the one to look at
        K.conv_fprop([(x, wp)], R, R, 1, pad, dil, cout, stats=(stats[0], stats[1]) if R * R * cin >= 1152 else None)
        K.conv_fprop([(dy, wp)], R, R, 1, pad, dil, cin, w_forward_layout=True)
        K.conv_wgrad(x, dy, R, R, 1, pad, dil, cin, cout, dw=g, dw_view=(cin, 0, cout, cin))
    torch.cuda.synchronize()


def bn_set(H, C):
    y = torch.randn(N, H, H, C, device="cuda").to(torch.bfloat16)
    dout = torch.randn(N, H, H, C, device="cuda").to(torch.bfloat16)
    gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    for _ in range(2):
        st = torch.zeros(2, 2048, dtype=torch.float64, device="cuda")
        other = torch.zeros(2, 2048, dtype=torch.float64, device="cuda")
        K.bn_stats(y, (st[0, :C], st[1, :C]))
        coef = torch.empty(4, C, device="cuda")
        out = K.bn_apply(y, None, None, True, finalize=dict(stats=(st[0, :C], st[1, :C]), count=N * H * H, gamma=gamma,
                                                          beta=beta, eps=1e-5, momentum=0.1, running_mean=rm,
                                                          running_var=rv, coef=coef, c_real=C,
                                                          reset=(other[0], other[1], C)))
        K.bn_backward(dout, out, y, coef[2], coef[3], coef[0], True, shift=coef[1])
    torch.cuda.synchronize()


conv_set(129, 256, 256, 3, 1)
conv_set(33, 256, 256, 3, 1)
conv_set(33, 256, 1024, 1, 1)
bn_set(129, 256)
bn_set(33, 1024)
print("done")
