"""BatchNorm(+ReLU/+residual/+Dropout) fused passes vs torch (F.batch_norm semantics of the reference)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _nhwc(t):  # NCHW fp32 -> NHWC fp32 view
    return t.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("C,hw,relu,res", [(64, 17, True, False), (256, 9, False, True), (48, 12, True, True),
                                           (2048, 5, True, False)])
def test_bn_train_forward_backward(C, hw, relu, res):
    from zs3_b200 import kernels as K
    N = 4
    Cp = K.cpad(C)
    g = torch.Generator().manual_seed(C)
    y = (torch.randn(N, C, hw, hw, generator=g) * 2 + 0.5).to(torch.bfloat16).float().cuda().requires_grad_(True)
    r = torch.randn(N, C, hw, hw, generator=g).to(torch.bfloat16).float().cuda().requires_grad_(res)
    gamma = (torch.rand(C, generator=g) + 0.5).cuda().requires_grad_(True)
    beta = torch.randn(C, generator=g).cuda().requires_grad_(True)
    rm, rv = torch.zeros(C).cuda(), torch.ones(C).cuda()
    rm2, rv2 = rm.clone(), rv.clone()
    z = F.batch_norm(y, rm, rv, gamma, beta, True, 0.1, 1e-5)
    if res:
        z = z + r
    out_ref = F.relu(z) if relu else z
    dout = torch.randn(out_ref.shape, generator=g).to(torch.bfloat16).float().cuda()
    grads = torch.autograd.grad(out_ref, [y, gamma, beta] + ([r] if res else []), dout)

    yh = K.nchw_to_nhwc(y.detach(), Cp)
    # statistics the conv epilogue would have produced
    stats = torch.zeros(2, Cp, dtype=torch.float64, device="cuda")
    stats[0, :C] = y.detach().double().sum(dim=(0, 2, 3))
    stats[1, :C] = (y.detach().double() ** 2).sum(dim=(0, 2, 3))
    scale, shift, mean, invstd = K.bn_finalize((stats[0], stats[1]), N * hw * hw, gamma.detach(), beta.detach(), 1e-5,
                                               0.1, rm2, rv2, Cp)
    assert stats.abs().max() == 0  # reset
    assert rel_l2(rm2, rm) < 1e-5 and rel_l2(rv2, rv) < 1e-5
    rh = K.nchw_to_nhwc(r.detach(), Cp) if res else None
    out = K.bn_apply(yh, scale, shift, relu, residual=rh)
    assert rel_l2(K.nhwc_to_nchw(out, C), out_ref) < 4e-3
    if Cp > C:
        assert out[..., C:].abs().max() == 0
    # backward
    dgamma, dbeta = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dres = torch.empty_like(yh) if res else None
    dy = K.bn_backward(K.nchw_to_nhwc(dout, Cp), out, yh, mean, invstd, scale, relu, dres=dres, dgamma=dgamma,
                       dbeta=dbeta)
    # the mask comes from the bf16-rounded forward output, the reference's from fp32: compare with bf16 slack
    assert rel_l2(K.nhwc_to_nchw(dy, C), grads[0]) < 1e-2
    assert rel_l2(dgamma, grads[1]) < 1e-2
    assert rel_l2(dbeta, grads[2]) < 1e-2
    if res:
        assert rel_l2(K.nhwc_to_nchw(dres, C), grads[3]) < 1e-2


def test_bn_eval_and_dropout():
    from zs3_b200 import kernels as K
    N, C, hw = 2, 256, 16
    g = torch.Generator().manual_seed(1)
    y = torch.randn(N, C, hw, hw, generator=g).to(torch.bfloat16).float().cuda()
    gamma, beta = (torch.rand(C, generator=g) + 0.5).cuda(), torch.randn(C, generator=g).cuda()
    rm, rv = torch.randn(C, generator=g).cuda(), (torch.rand(C, generator=g) + 0.5).cuda()
    ref = F.relu(F.batch_norm(y, rm, rv, gamma, beta, False, 0.1, 1e-5))
    scale, shift, mean, invstd = K.bn_eval_coeffs(gamma, beta, rm, rv, 1e-5, C)
    yh = K.nchw_to_nhwc(y, C)
    out = K.bn_apply(yh, scale, shift, True)
    assert rel_l2(K.nhwc_to_nchw(out, C), ref) < 4e-3
    # explicit keep-mask
    mask = (torch.rand(N, hw, hw, C, generator=g) > 0.5).to(torch.uint8).cuda()
    out_m = K.bn_apply(yh, scale, shift, True, drop_p=0.5, keep_mask=mask)
    ref_m = _nhwc(ref) * mask.float() * 2.0
    assert rel_l2(out_m.float(), ref_m) < 4e-3
    # counter-based RNG: right keep rate, deterministic in (seed, offset), scaled by 1/(1-p)
    for p in (0.5, 0.1):
        o1 = K.bn_apply(yh, scale, shift, True, drop_p=p, seed=123, offset=77)
        o2 = K.bn_apply(yh, scale, shift, True, drop_p=p, seed=123, offset=77)
        o3 = K.bn_apply(yh, scale, shift, True, drop_p=p, seed=124, offset=77)
        assert torch.equal(o1, o2) and not torch.equal(o1, o3)
        pos = out.float() > 0
        kept = (o1.float() > 0) & pos
        rate = kept.sum().item() / pos.sum().item()
        assert abs(rate - (1 - p)) < 0.01, rate
        assert rel_l2(o1.float()[kept], out.float()[kept] / (1 - p)) < 4e-3


def test_bn_stats_kernel():
    from zs3_b200 import kernels as K
    g = torch.Generator().manual_seed(3)
    for C, hw in ((64, 33), (1024, 9), (2048, 5)):
        y = (torch.randn(3, C, hw, hw, generator=g) * 2 + 1).to(torch.bfloat16).float().cuda()
        stats = torch.zeros(2, C, dtype=torch.float64, device="cuda")
        K.bn_stats(K.nchw_to_nhwc(y, C), (stats[0], stats[1]))
        assert rel_l2(stats[0], y.double().sum(dim=(0, 2, 3))) < 1e-5
        assert rel_l2(stats[1], (y.double() ** 2).sum(dim=(0, 2, 3))) < 1e-5


def test_bn_apply_fused_finalize_matches_separate_finalize():
    from zs3_b200 import kernels as K
    g = torch.Generator().manual_seed(11)
    N, C, hw = 3, 48, 14
    Cp = K.cpad(C)
    y = (torch.randn(N, C, hw, hw, generator=g) * 1.5 - 0.3).to(torch.bfloat16).float().cuda()
    gamma, beta = (torch.rand(C, generator=g) + 0.5).cuda(), torch.randn(C, generator=g).cuda()
    yh = K.nchw_to_nhwc(y, Cp)
    def fresh_stats():
        st = torch.zeros(2, 2048, dtype=torch.float64, device="cuda")
        K.bn_stats(yh, (st[0, :Cp], st[1, :Cp]))
        return st
    rm1, rv1 = torch.zeros(C).cuda(), torch.ones(C).cuda()
    st1 = fresh_stats()
    sc, sh, mu, istd = K.bn_finalize((st1[0, :Cp], st1[1, :Cp]), N * hw * hw, gamma, beta, 1e-5, 0.1, rm1, rv1, Cp)
    ref = K.bn_apply(yh, sc, sh, True)
    rm2, rv2 = torch.zeros(C).cuda(), torch.ones(C).cuda()
    st2 = fresh_stats()
    other = torch.full((2, 2048), 7.0, dtype=torch.float64, device="cuda")
    coef = torch.empty(4, Cp, device="cuda")
    out = K.bn_apply(yh, None, None, True, finalize=dict(stats=(st2[0, :Cp], st2[1, :Cp]), count=N * hw * hw, gamma=gamma,
                                                          beta=beta, eps=1e-5, momentum=0.1, running_mean=rm2,
                                                          running_var=rv2, coef=coef, c_real=C,
                                                          reset=(other[0], other[1], 300)))
    assert torch.equal(out, ref)
    assert torch.allclose(coef[0], sc) and torch.allclose(coef[1], sh) and torch.allclose(coef[2], mu) and torch.allclose(coef[3], istd)
    assert torch.allclose(rm1, rm2) and torch.allclose(rv1, rv2)
    assert other[:, :300].abs().max() == 0 and (other[:, 300:] == 7.0).all()
    ref_bn = F.batch_norm(y, torch.zeros(C).cuda(), torch.ones(C).cuda(), gamma, beta, True, 0.1, 1e-5)
    assert rel_l2(K.nhwc_to_nchw(out, C), F.relu(ref_bn)) < 4e-3
