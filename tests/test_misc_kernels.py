"""HBM-bound glue kernels vs their torch equivalents (the ops the reference calls)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.to(torch.bfloat16).float()


def test_stem_im2col_conv():
    from zs3_b200 import kernels as K
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 33, 37, generator=g).cuda()
    w = _bf(torch.randn(64, 3, 7, 7, generator=g) * 0.1).cuda()
    ref = F.conv2d(_bf(x), w, stride=2, padding=3)
    ho, wo = ref.shape[2], ref.shape[3]
    cols = K.stem_im2col(x, 7, 2, 3, ho, wo, 192)
    wp = K.pack_weight(w.reshape(64, 147, 1, 1), 64, 192)
    y = K.conv_fprop([(cols, wp)], 1, 1, 1, 0, 1, 64, out_f32=True)
    assert rel_l2(y.permute(0, 3, 1, 2), ref) < 2e-5


def test_maxpool():
    from zs3_b200 import kernels as K
    g = torch.Generator().manual_seed(1)
    x = _bf(torch.relu(torch.randn(2, 64, 17, 19, generator=g))).cuda().requires_grad_(True)
    ref = F.max_pool2d(x, 3, 2, 1)
    dy = _bf(torch.randn(ref.shape, generator=g)).cuda()
    (dx_ref,) = torch.autograd.grad(ref, x, dy)
    xh = K.nchw_to_nhwc(x.detach(), 64)
    y, arg = K.maxpool_fwd(xh, 3, 2, 1)
    assert torch.equal(K.nhwc_to_nchw(y, 64), ref.detach())
    dx = K.maxpool_bwd(K.nchw_to_nhwc(dy, 64), arg, xh.shape, 3, 2, 1)
    assert rel_l2(K.nhwc_to_nchw(dx, 64), dx_ref) < 4e-3  # bf16 rounding of sums of up to 4 gradients


@pytest.mark.parametrize("hi,ho", [(5, 17), (9, 33), (33, 129), (7, 7), (1, 5)])
def test_bilinear(hi, ho):
    from zs3_b200 import kernels as K
    g = torch.Generator().manual_seed(2)
    x = _bf(torch.randn(2, 64, hi, hi, generator=g)).cuda().requires_grad_(True)
    ref = F.interpolate(x, size=(ho, ho), mode="bilinear", align_corners=True)
    dy = _bf(torch.randn(ref.shape, generator=g)).cuda()
    (dx_ref,) = torch.autograd.grad(ref, x, dy)
    y = K.bilinear_fwd(K.nchw_to_nhwc(x.detach(), 64), ho, ho)
    assert rel_l2(K.nhwc_to_nchw(y, 64), ref) < 4e-3
    dx = K.bilinear_bwd(K.nchw_to_nhwc(dy, 64), hi, hi)
    assert rel_l2(K.nhwc_to_nchw(dx, 64), dx_ref) < 4e-3


@pytest.mark.parametrize("C,hi,ho", [(21, 17, 65), (60, 9, 33), (21, 129, 513)])
def test_upsample_logits(C, hi, ho):
    from zs3_b200 import kernels as K
    g = torch.Generator().manual_seed(3)
    x = _bf(torch.randn(2, C, hi, hi, generator=g)).cuda().requires_grad_(True)
    ref = F.interpolate(x, size=(ho, ho), mode="bilinear", align_corners=True)
    dy = torch.randn(ref.shape, generator=g).cuda()
    (dx_ref,) = torch.autograd.grad(ref, x, dy)
    xh = K.nchw_to_nhwc(x.detach(), 64)
    y = K.upsample_logits_fwd(xh, C, ho, ho)
    assert rel_l2(y, ref) < 1e-5
    dx = K.upsample_logits_bwd(dy, xh.shape, C)
    assert rel_l2(K.nhwc_to_nchw(dx, C), dx_ref) < 4e-3
    assert dx[..., C:].abs().max() == 0


def test_spatial_sum_broadcast():
    from zs3_b200 import kernels as K
    g = torch.Generator().manual_seed(4)
    x = _bf(torch.randn(3, 128, 9, 9, generator=g)).cuda()
    xh = K.nchw_to_nhwc(x, 128)
    m = K.spatial_sum(xh, 1.0 / 81)
    assert rel_l2(m.view(3, 128).float(), x.mean(dim=(2, 3))) < 4e-3
    b = K.spatial_broadcast(m, 9, 9, 1.0)
    assert torch.equal(b, m.expand(3, 9, 9, 128).contiguous())


@pytest.mark.parametrize("C,weighted", [(21, False), (21, True), (60, True)])
def test_cross_entropy(C, weighted):
    from zs3_b200.utils.loss import SegmentationLosses
    g = torch.Generator().manual_seed(5)
    N, H = 3, 37
    logit = (torch.randn(N, C, H, H, generator=g) * 3).cuda().requires_grad_(True)
    target = torch.randint(0, C, (N, H, H), generator=g).float()
    target[torch.rand(N, H, H, generator=g) < 0.05] = 255
    target = target.cuda()
    w = None
    if weighted:
        w = torch.ones(C)
        w[[3, 7]] = 100.0
        w = w.cuda()
    ref = F.cross_entropy(logit, target.long(), weight=w, ignore_index=255) / N
    (g_ref,) = torch.autograd.grad(ref, logit)
    crit = SegmentationLosses(weight=w, cuda=True).build_loss("ce")
    lg = logit.detach().clone().requires_grad_(True)
    loss = crit(lg, target)
    loss.backward()
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())
    assert rel_l2(lg.grad, g_ref) < 1e-5
    # all-ignored image contributes nothing; fully ignored batch would be 0/0 like the reference
    target2 = target.clone()
    target2[0] = 255
    l2 = crit(lg.detach(), target2)
    r2 = F.cross_entropy(logit.detach(), target2.long(), weight=w, ignore_index=255) / N
    assert abs(l2.item() - r2.item()) < 1e-5 * abs(r2.item())


@pytest.mark.parametrize("x4", [True, False], ids=["x4_fast_path", "generic"])
@pytest.mark.parametrize("C,hi,ho,weighted", [(21, 33, 129, False), (21, 17, 65, True), (5, 9, 33, True), (21, 129, 513, False),
                                              (60, 17, 65, True), (60, 129, 513, False), (33, 33, 129, True),
                                              (21, 20, 65, True), (7, 2, 5, False)])
def test_fused_upsample_cross_entropy(C, hi, ho, weighted, x4, monkeypatch):
    """loss = CE(interpolate(scores)) straight from the low-resolution NHWC bf16 scores (training-loss fusion) against
    torch's interpolate + cross_entropy in fp32 on the same bf16-representable scores; tolerances: loss 1e-5 rel,
    d loss / d scores 4e-3 rel-L2 (the gradient is stored in bf16, 2^-9 per element)."""
    from zs3_b200 import kernels as K
    from zs3_b200.utils.loss import SegmentationLosses
    # the backward has two kernels: the exact-x4 geometry of DeepLab (one softmax per output pixel, deterministic
    # combine) and the generic one; (20 -> 65) is not x4 and always takes the generic kernel
    monkeypatch.setenv("ZS3_CE_BWD_X4", "1" if x4 else "0")
    g = torch.Generator().manual_seed(11)
    N = 2
    x = _bf(torch.randn(N, C, hi, hi, generator=g) * 2).cuda().requires_grad_(True)
    target = torch.randint(0, C, (N, ho, ho), generator=g).float()
    target[torch.rand(N, ho, ho, generator=g) < 0.05] = 255
    target = target.cuda()
    w = None
    if weighted:
        w = torch.ones(C)
        w[[1, 3]] = 7.0
        w = w.cuda()
    up = F.interpolate(x, size=(ho, ho), mode="bilinear", align_corners=True)
    ref = F.cross_entropy(up, target.long(), weight=w, ignore_index=255) / N
    (g_ref,) = torch.autograd.grad(ref, x)
    losses = SegmentationLosses(weight=w, cuda=True)
    xh = K.nchw_to_nhwc(x.detach(), 64).requires_grad_(True)   # C = 60 is Pascal-Context (the C <= 64 instantiation)
    loss = losses.UpsampledCrossEntropyLoss(xh, C, target)
    loss.backward()
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())
    assert rel_l2(K.nhwc_to_nchw(xh.grad, C), g_ref) < 4e-3
    assert xh.grad[..., C:].abs().max() == 0
    # and against the unfused path of this package (upsample kernel + CE kernel)
    xh2 = xh.detach().clone().requires_grad_(True)
    from zs3_b200 import functional as ZF
    l2 = losses.build_loss("ce")(ZF.UpsampleLogits.apply(xh2, C, ho, ho), target)
    l2.backward()
    assert abs(loss.item() - l2.item()) < 1e-6 * abs(l2.item())
    print(f"fused vs unfused d scores rel_l2 = {rel_l2(xh.grad.float(), xh2.grad.float()):.3e}")
    assert rel_l2(xh.grad.float(), xh2.grad.float()) < 4e-3


def test_optimizers():
    from zs3_b200 import kernels as K
    torch.manual_seed(0)
    p = torch.randn(1000, device="cuda", requires_grad=True)
    q = p.detach().clone()
    buf = torch.zeros_like(q)
    opt = torch.optim.SGD([p], lr=0.07, momentum=0.9, weight_decay=5e-4)
    for it in range(3):
        g = torch.randn(1000, device="cuda")
        p.grad = g.clone()
        opt.step()
        K.sgd_step(q, g, buf, 0.07, 0.9, 5e-4, False, it == 0)
    assert torch.allclose(p.detach(), q, atol=1e-6)
    p = torch.randn(1000, device="cuda", requires_grad=True)
    q = p.detach().clone()
    m, v = torch.zeros_like(q), torch.zeros_like(q)
    opt = torch.optim.Adam([p], lr=2e-4)
    for it in range(1, 4):
        g = torch.randn(1000, device="cuda")
        p.grad = g.clone()
        opt.step()
        K.adam_step(q, g, m, v, 2e-4, 0.9, 0.999, 1e-8, it)
    assert torch.allclose(p.detach(), q, atol=1e-6)
