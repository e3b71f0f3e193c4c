"""Split-precision TRAINING mode (zs3_b200/parity_train.py, csrc/parity_train.cu): kernel-level parity against torch
fp32/fp64 and full-model forward + backward parity against the oracle (oracle/zs3_oracle.py, pinned to the real
reference by tests/golden).

Tolerances (written next to each assert): BASELINE.json's north star asks for logits within 1e-3 of the reference;
gradients are held to the same relative-L2 bound in the regimes where the reference's own arithmetic (fp32) is itself
within 1e-3 of fp64 -- eval/frozen-BN statistics -- and, in the chaotic train-mode-BN regime at random init
(fp32 vs fp64 already differs by ~1e-3), to 5x the error of the reference's fp32 arithmetic."""
import ctypes as C
import os
import sys

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))

pytestmark = pytest.mark.gpu


def _nhwc(x, cs=None):
    """NCHW fp32 -> NHWC fp32 with zero-padded channel stride"""
    n, c, h, w = x.shape
    cs = cs or (c + 63) // 64 * 64
    out = torch.zeros(n, h, w, cs, device=x.device, dtype=torch.float32)
    out[..., :c] = x.permute(0, 2, 3, 1)
    return out


def _nchw(x, c):
    return x[..., :c].permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("pieces,tol", [(1, 2.0 ** -8), (2, 2.0 ** -16), (3, 2.0 ** -23)])
def test_split_pieces_sum_to_the_value(pieces, tol):
    from zs3_b200 import parity_train as PT
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(4096 * 8, generator=g) * torch.logspace(-6, 6, 4096 * 8)).cuda()
    ps = PT.split(x, pieces)
    s = sum(p.double() for p in ps)
    assert ((s - x.double()).abs() <= tol * x.abs().double() + 1e-40).all()
    assert (torch.sign(ps[0].float()) == torch.sign(x)).all()   # the backward takes the ReLU mask from piece 0


@pytest.mark.parametrize("relu,res,drop", [(True, False, False), (True, True, False), (False, False, False), (True, False, True)])
def test_bn_act_and_backward_match_torch(relu, res, drop):
    """y -> BN(train) -> (+res) -> ReLU -> Dropout(mask) forward (fp32 and pieces) and its backward against torch
    autograd in fp64: <= 2e-6 forward (fp32 rounding), <= 1e-5 backward sums/dy."""
    from zs3_b200 import _lib as L
    from zs3_b200 import kernels as K
    from zs3_b200 import parity_train as PT
    g = torch.Generator().manual_seed(1)
    n, c, h, w = 3, 128, 9, 11
    y = torch.randn(n, c, h, w, generator=g).cuda() * 2 + 0.3
    r = torch.randn(n, c, h, w, generator=g).cuda() if res else None
    gamma, beta = (torch.rand(c, generator=g) + 0.5).cuda(), (torch.randn(c, generator=g) * 0.1).cuda()
    keep = (torch.rand(n, h, w, c, generator=g) > 0.5).to(torch.uint8).cuda() if drop else None
    dout = torch.randn(n, c, h, w, generator=g).cuda()
    # torch reference in fp64
    yd = y.double().requires_grad_(True)
    rd = r.double().requires_grad_(True) if res else None
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    z = F.batch_norm(yd, None, None, gd, bd, True, 0.1, 1e-5)
    if res:
        z = z + rd
    if relu:
        z = F.relu(z)
    if drop:
        z = z * keep.permute(0, 3, 1, 2).double() / 0.5
    z.backward(dout.double())
    # ours
    yh = _nhwc(y)
    stats = torch.zeros(2, c, dtype=torch.float64, device="cuda")
    stats[0] = yh.double().sum((0, 1, 2))
    stats[1] = (yh.double() ** 2).sum((0, 1, 2))
    scale, shift, mean, invstd = K.bn_finalize((stats[0], stats[1]), n * h * w, gamma, beta, 1e-5, 0.1, None, None, c)
    eng = PT.SplitPrecisionTrainer.__new__(PT.SplitPrecisionTrainer)
    eng.P, eng.rng_calls = 2, 0
    out = eng._bn_act(yh, scale, shift, relu, residual=_nhwc(r) if res else None, drop_p=0.5 if drop else 0.0,
                      keep_mask=keep)
    assert rel_l2(_nchw(out.f32, c), z.detach()) < 2e-6
    assert rel_l2(_nchw(out.pieces[0].float() + out.pieces[1].float(), c), z.detach()) < 2e-5   # 16 mantissa bits
    a = L.BnBwdF32Args()
    dh = _nhwc(dout)
    a.dout, a.dout_cstride = dh.data_ptr(), c
    if relu or drop:
        a.relu = 1
        a.act_hi, a.act_hi_cstride = out.pieces[0].data_ptr(), c
    a.y, a.y_cstride = yh.data_ptr(), c
    a.mean, a.invstd, a.scale = mean.data_ptr(), invstd.data_ptr(), scale.data_ptr()
    a.M, a.C, a.grad_scale, a.training = n * h * w, c, 2.0 if drop else 1.0, 1
    sums = torch.full((2, c), 7.0, dtype=torch.float64, device="cuda")  # dirty on purpose: the call zeroes them
    a.sum_dz, a.sum_dzx = sums[0].data_ptr(), sums[1].data_ptr()
    dy = torch.empty_like(yh)
    a.dy, a.dy_cstride = dy.data_ptr(), c
    pcs = [torch.empty(yh.shape, dtype=torch.bfloat16, device="cuda") for _ in range(2)]
    for i, t in enumerate(pcs):
        a.dy_pieces[i] = t.data_ptr()
    a.n_pieces, a.piece_cstride = 2, c
    dres = torch.empty_like(yh) if res else None
    if res:
        a.dres, a.dres_cstride = dres.data_ptr(), c
    dgam, dbet = torch.ones(c, device="cuda"), torch.ones(c, device="cuda")
    a.dgamma, a.dbeta, a.C_real, a.param_accumulate = dgam.data_ptr(), dbet.data_ptr(), c, 1
    L.check(L.lib().zs3_bn_bwd_f32(C.byref(a), L.stream_ptr()), "zs3_bn_bwd_f32")
    assert rel_l2(_nchw(dy, c), yd.grad) < 1e-5
    assert rel_l2(_nchw(pcs[0].float() + pcs[1].float(), c), yd.grad) < 2e-5
    assert rel_l2(dgam - 1, gd.grad) < 1e-5 and rel_l2(dbet - 1, bd.grad) < 1e-5   # accumulated onto the ones
    if res:
        assert rel_l2(_nchw(dres, c), rd.grad) < 1e-6


def test_maxpool_and_bilinear_backward_match_torch():
    from zs3_b200 import _lib as L
    g = torch.Generator().manual_seed(2)
    st = L.stream_ptr()
    # max-pool 3x3/2 pad 1 (resnet.py:82), ties included (ReLU zeros)
    x = torch.relu(torch.randn(2, 64, 17, 19, generator=g)).cuda().requires_grad_(True)
    ref = F.max_pool2d(x, 3, 2, 1)
    dy = torch.randn(ref.shape, generator=g).cuda()
    ref.backward(dy)
    xh = _nhwc(x.detach())
    n, h, w, cs = xh.shape
    ho, wo = ref.shape[2], ref.shape[3]
    y = torch.empty(n, ho, wo, cs, device="cuda")
    arg = torch.empty(n, ho, wo, cs, dtype=torch.uint8, device="cuda")
    L.check(L.lib().zs3_maxpool_arg_f32(L.ptr(xh), L.ptr(y), L.ptr(arg), n, h, w, cs, ho, wo, 3, 2, 1, st), "maxpool")
    assert torch.equal(_nchw(y, 64), ref.detach())
    dx = torch.empty_like(xh)
    L.check(L.lib().zs3_maxpool_bwd_f32(L.ptr(_nhwc(dy)), L.ptr(arg), L.ptr(dx), n, h, w, cs, ho, wo, 3, 2, 1, st), "maxpool_bwd")
    assert rel_l2(_nchw(dx, 64), x.grad) < 1e-6
    # bilinear align_corners x4 (decoder.py:33-35), NHWC
    a = torch.randn(2, 64, 9, 9, generator=g).cuda().requires_grad_(True)
    up = F.interpolate(a, size=(33, 33), mode="bilinear", align_corners=True)
    du = torch.randn(up.shape, generator=g).cuda()
    up.backward(du)
    da = torch.empty(2, 9, 9, 64, device="cuda")
    L.check(L.lib().zs3_bilinear_bwd_f32(L.ptr(_nhwc(du)), L.ptr(da), 2, 9, 9, 33, 33, 64, 64, 64, 0, st), "bilinear_bwd")
    assert rel_l2(_nchw(da, 64), a.grad) < 1e-5
    # from NCHW logits gradients (deeplab.py:44), 21 real channels in a 64-wide NHWC gradient, odd sizes
    s = torch.randn(2, 21, 17, 17, generator=g).cuda().requires_grad_(True)
    lg = F.interpolate(s, size=(65, 65), mode="bilinear", align_corners=True)
    dl = torch.randn(lg.shape, generator=g).cuda()
    lg.backward(dl)
    ds = torch.zeros(2, 17, 17, 64, device="cuda")
    L.check(L.lib().zs3_bilinear_bwd_f32(L.ptr(dl.contiguous()), L.ptr(ds), 2, 17, 17, 65, 65, 21, 0, 64, 1, st), "bilinear_bwd")
    assert rel_l2(_nchw(ds, 21), s.grad) < 1e-5 and ds[..., 21:].abs().max() == 0
    # channel sums (bias gradient) and accumulating broadcast (global-pool backward)
    t = torch.randn(5, 7, 3, 64, generator=g).cuda()
    sums = torch.empty(64, dtype=torch.float64, device="cuda")
    L.check(L.lib().zs3_channel_sums_f32(L.ptr(t), 64, 5 * 7 * 3, 64, L.ptr(sums), st), "channel_sums")
    assert rel_l2(sums, t.double().sum((0, 1, 2))) < 1e-6
    v = torch.randn(5, 64, generator=g).cuda()
    t2 = t.clone()
    L.check(L.lib().zs3_spatial_broadcast_acc_f32(L.ptr(v), L.ptr(t2), 5, 21, 64, 0.25, 1, st), "broadcast_acc")
    assert rel_l2(t2, t + 0.25 * v[:, None, None, :]) < 1e-6


def _calibrated_state(x, seed=1, num_classes=21, output_stride=16):
    """random-init weights whose BatchNorm running statistics are this batch's statistics (one fp64 train-mode pass with
    momentum 1): eval-mode BN then normalises like a trained network does -- activations and logits O(1) -- instead of
    the identity-BN / logits ~1e5 regime of untouched running stats.  The frozen-BN fine-tuning regime of the reference
    (train_pascal.py --freeze-bn) with a well-conditioned network."""
    import zs3_oracle as O
    st = O.init_deeplab_state(seed=seed, num_classes=num_classes, output_stride=output_stride)
    s64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in st.items()}
    mom, O.BN_MOMENTUM = O.BN_MOMENTUM, 1.0
    try:
        with torch.no_grad():
            O.deeplab_forward(s64, x.double(), training=True, output_stride=output_stride, drop_p=(0.0, 0.0, 0.0))
    finally:
        O.BN_MOMENTUM = mom
    return {k: (v.float() if v.is_floating_point() else v) for k, v in s64.items()}


def _cudnn_step(st, x, target, training, tf32):
    """the reference's GPU arithmetic (stock torch on cuda, cuDNN TF32 allowed = torch's default) as a yardstick"""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = tf32
    try:
        return _oracle_step(st, x, target, training, torch.float32, "cuda")
    finally:
        torch.backends.cudnn.allow_tf32 = old


def _oracle_step(st, x, target, training, dtype, device):
    """loss, logits and every parameter gradient of the oracle in `dtype` on `device` (Dropout off)"""
    import zs3_oracle as O
    s = {k: (v.to(device=device, dtype=dtype if v.is_floating_point() else v.dtype)).clone() for k, v in st.items()}
    for k, v in s.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    logits = O.deeplab_forward(s, x.to(device=device, dtype=dtype), training=training, drop_p=(0.0, 0.0, 0.0))
    loss = O.cross_entropy(logits, target.to(device))
    loss.backward()
    grads = {k: v.grad for k, v in s.items() if v.requires_grad}
    return loss.detach(), logits.detach(), grads, s


def _our_step(st, x, target, training, pieces, freeze=False):
    from zs3_b200 import parity_train as PT
    from zs3_b200.modeling.deeplab import DeepLab
    ncls = st["decoder.pred_conv.weight"].shape[0]
    model = DeepLab(num_classes=ncls, sync_bn=True, pretrained=False)
    model.load_state_dict({k: v.clone() for k, v in st.items()})
    model = model.cuda()
    model.train() if training else model.eval()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    eng = PT.SplitPrecisionTrainer(model, pieces=pieces, optimizer=False)
    loss, logits = eng.loss_and_grads(x.cuda(), target.cuda(), return_logits=True)
    torch.cuda.synchronize()
    return loss, logits, {k: p.grad for k, p in model.named_parameters()}, model


def _global_rel(grads, ref):
    num = sum(float(((grads[k].double() - ref[k].double().to(grads[k].device)) ** 2).sum()) for k in ref)
    den = sum(float((ref[k].double() ** 2).sum()) for k in ref)
    return (num / den) ** 0.5


def _state_for(regime, x, **kw):
    """'identity': untouched-looking running statistics (randomised around mean 0 / var 1): BatchNorm is close to the
    identity, the network's error gain is ~1e2 and the north star's 1e-3 is reachable -- BASELINE configs[0]'s regime.
    'calibrated': running statistics = this batch's statistics: the normalised random-init network has an error gain of
    ~1e3 (measured: cuDNN-TF32 logits are 0.6 away from fp64, fp32 2e-3), so bounds are relative to those yardsticks."""
    import zs3_oracle as O
    if regime == "identity":
        return O.init_deeplab_state(seed=kw.get("seed", 1), num_classes=kw.get("num_classes", 21),
                                    output_stride=kw.get("output_stride", 16), randomize_bn=True)
    return _calibrated_state(x, **kw)


def _check_against_yardsticks(pieces, regime, e_log, e_grad, y_log, y_grad, per_tensor):
    """pieces=3 is held to the reference's fp32 arithmetic (within 2x of ITS distance to fp64, or 1e-3); pieces=2 to the
    reference's GPU arithmetic (cuDNN TF32): at least 10x closer on logits, 2x on gradients, or 1e-3.
    y_* = {"f32": .., "tf32": ..}; per_tensor = [(name, ours, f32, tf32)]"""
    if regime == "identity":
        assert e_log < 1e-3                                   # BASELINE.json north star
    if pieces == 3:
        assert e_log < max(1e-3, 2 * y_log["f32"]) and e_grad < max(1e-3, 2 * y_grad["f32"])
    else:
        assert e_log < max(1e-3, 0.1 * y_log["tf32"]) and e_grad < max(1e-3, 0.5 * y_grad["tf32"])
    for name, e, f32, tf32 in per_tensor:
        assert e < max(1e-2, 3 * f32 if pieces == 3 else tf32), (name, e, f32, tf32)   # small cancellation-heavy tensors (BN biases): fp32 TMEM accumulation


@pytest.mark.parametrize("regime", ["identity", "calibrated"])
@pytest.mark.parametrize("pieces", [3, 2])
def test_full_model_eval_bn_forward_and_all_gradients(pieces, regime):
    """Frozen-BN training step (eval statistics, every conv and affine parameter trainable): logits AND every one of the
    312 parameter gradients against the fp64 oracle, each tensor asserted, with the reference's own arithmetic (stock
    torch on this GPU: fp32, and cuDNN with TF32 allowed = the reference's GPU path) as yardsticks."""
    x = torch.randn(2, 3, 65, 65, generator=torch.Generator().manual_seed(11))
    target = torch.randint(0, 21, (2, 65, 65), generator=torch.Generator().manual_seed(12)).float()
    target[:, :3] = 255
    st = _state_for(regime, x)
    loss_ref, logits_ref, g_ref, _ = _oracle_step(st, x, target, False, torch.float64, "cpu")
    _, logits_tf32, g_tf32, _ = _cudnn_step(st, x, target, False, True)
    _, logits_f32, g_f32, _ = _cudnn_step(st, x, target, False, False)
    loss, logits, g, _ = _our_step(st, x, target, False, pieces)
    cpu = lambda d: {k: v.cpu() for k, v in d.items()}  # noqa: E731
    g, g_tf32, g_f32 = cpu(g), cpu(g_tf32), cpu(g_f32)
    e_log, e_all = rel_l2(logits.cpu(), logits_ref), _global_rel(g, g_ref)
    y_log = {"f32": rel_l2(logits_f32.cpu(), logits_ref), "tf32": rel_l2(logits_tf32.cpu(), logits_ref)}
    y_grad = {"f32": _global_rel(g_f32, g_ref), "tf32": _global_rel(g_tf32, g_ref)}
    per = [(k, rel_l2(g[k], g_ref[k]), rel_l2(g_f32[k], g_ref[k]), rel_l2(g_tf32[k], g_ref[k])) for k in g_ref]
    worst = max(per, key=lambda t: t[1])
    print(f"pieces={pieces} eval-BN/{regime}: logits {e_log:.2e} (torch fp32 {y_log['f32']:.2e}, cuDNN tf32 {y_log['tf32']:.2e})  "
          f"loss {loss.item():.7f} vs {loss_ref.item():.7f}  grads global {e_all:.2e} (fp32 {y_grad['f32']:.2e}, tf32 "
          f"{y_grad['tf32']:.2e}) worst {worst[1]:.2e} ({worst[0]}; fp32 {worst[2]:.2e}, tf32 {worst[3]:.2e})")
    assert set(g) == set(g_ref) and all(v is not None for v in g.values())
    _check_against_yardsticks(pieces, regime, e_log, e_all, y_log, y_grad, per)


def test_full_model_train_bn_step_vs_fp64_oracle_with_fp32_yardstick():
    """Train-mode BatchNorm at random init (the chaotic regime): pieces=3 against the fp64 oracle, with the fp32 oracle's
    own distance to fp64 as the yardstick; running statistics and num_batches_tracked follow F.batch_norm."""
    import zs3_oracle as O
    x = torch.randn(2, 3, 65, 65, generator=torch.Generator().manual_seed(11))
    target = torch.randint(0, 21, (2, 65, 65), generator=torch.Generator().manual_seed(12)).float()
    target[:, :3] = 255
    st = O.init_deeplab_state(seed=1)
    _, logits64, g64, s64 = _oracle_step(st, x, target, True, torch.float64, "cpu")
    _, logits32, g32, _ = _oracle_step(st, x, target, True, torch.float32, "cpu")
    y_log, y_grad = rel_l2(logits32, logits64), _global_rel(g32, g64)
    loss, logits, g, model = _our_step(st, x, target, True, 3)
    e_log, e_grad = rel_l2(logits.cpu(), logits64), _global_rel({k: v.cpu() for k, v in g.items()}, g64)
    print(f"train-BN: logits {e_log:.2e} (fp32 oracle {y_log:.2e})  grads {e_grad:.2e} (fp32 oracle {y_grad:.2e})")
    assert e_log < max(1e-3, 5 * y_log)
    assert e_grad < max(1e-3, 5 * y_grad)
    assert rel_l2(model.backbone.bn1.running_mean.cpu(), s64["backbone.bn1.running_mean"]) < 1e-5
    assert rel_l2(model.decoder.last_conv[5].running_var.cpu(), s64["decoder.last_conv.5.running_var"]) < 1e-3
    assert int(model.backbone.layer3[5].bn2.num_batches_tracked) == 1


@pytest.mark.parametrize("ncls,os_", [(60, 16), (21, 8)])
def test_context_classes_and_output_stride_8(ncls, os_):
    """DeepLab(num_classes=60) (Pascal-Context) and output_stride=8 (resnet.py:69-76, aspp.py:47-52) at model level"""
    import zs3_oracle as O
    from zs3_b200 import parity_train as PT
    from zs3_b200.modeling.deeplab import DeepLab
    x = torch.randn(2, 3, 65, 65, generator=torch.Generator().manual_seed(21))
    target = torch.randint(0, ncls, (2, 65, 65), generator=torch.Generator().manual_seed(22)).float()
    st = _state_for("identity", x, seed=2, num_classes=ncls, output_stride=os_)
    s64 = {k: (v.double().requires_grad_("running" not in k) if v.is_floating_point() else v) for k, v in st.items()}
    logits_ref = O.deeplab_forward(s64, x.double(), training=False, output_stride=os_)
    O.cross_entropy(logits_ref, target).backward()
    model = DeepLab(num_classes=ncls, output_stride=os_, sync_bn=True, pretrained=False)
    model.load_state_dict(st)
    model = model.cuda().eval()
    eng = PT.SplitPrecisionTrainer(model, pieces=2, optimizer=False)
    loss, logits = eng.loss_and_grads(x.cuda(), target.cuda(), return_logits=True)
    e = rel_l2(logits.cpu(), logits_ref.detach())
    g = {k: p.grad.cpu() for k, p in model.named_parameters()}
    g_ref = {k: v.grad for k, v in s64.items() if v.is_floating_point() and v.requires_grad}
    e_g = _global_rel(g, g_ref)
    _, _, g_tf32, _ = _cudnn_step(st, x, target, False, True)
    y_g = _global_rel({k: v.cpu() for k, v in g_tf32.items()}, g_ref)
    print(f"C={ncls} OS={os_}: logits {e:.2e} grads {e_g:.2e} (cuDNN tf32 {y_g:.2e})")
    assert e < 1e-3 and e_g < max(1e-3, y_g)
    # and the bf16 throughput path on the same model (module API): bf16 tolerance
    with torch.no_grad():
        out = model(x.cuda())
    assert tuple(out.shape) == (2, ncls, 65, 65) and rel_l2(out.cpu(), logits_ref.detach()) < 3e-2


@pytest.mark.parametrize("regime", ["identity", "calibrated"])
def test_fwd_bwd_parity_at_513_against_fp64_on_the_device(regime):
    """2 x 3 x 513 x 513 (the benchmark resolution; the wgrad pixel splits and TMA reduce-adds run at M = 33 282 here and
    the decoder at 129^2): pieces=2 logits and decoder / ASPP / layer4 / layer1 / stem gradients against the oracle
    evaluated in fp64 ON THE GPU (the oracle is plain torch; cuDNN/cuBLAS fp64 is the checker here, never the product),
    next to the reference's GPU arithmetic (cuDNN TF32) on the same tensors."""
    x = torch.randn(2, 3, 513, 513, generator=torch.Generator().manual_seed(31))
    target = torch.randint(0, 21, (2, 17, 17), generator=torch.Generator().manual_seed(32)).float()
    target = F.interpolate(target[:, None], size=(513, 513), mode="nearest")[:, 0].contiguous()
    target[:, :5] = 255
    st = _state_for(regime, x)
    loss_ref, logits_ref, g_ref, _ = _oracle_step(st, x, target, False, torch.float64, "cuda")
    _, logits_tf32, g_tf32, _ = _cudnn_step(st, x, target, False, True)
    loss, logits, g, _ = _our_step(st, x, target, False, 2)
    e_log, y_log = rel_l2(logits, logits_ref), rel_l2(logits_tf32, logits_ref)
    e_g, y_g = _global_rel(g, g_ref), _global_rel(g_tf32, g_ref)
    print(f"513^2 pieces=2/{regime}: logits {e_log:.2e} (cuDNN tf32 {y_log:.2e}) loss {loss.item():.7f} vs "
          f"{loss_ref.item():.7f}; grads global {e_g:.2e} (cuDNN tf32 {y_g:.2e})")
    if regime == "identity":
        assert e_log < 1e-3
    assert e_log < max(1e-3, 0.1 * y_log) and e_g < max(1e-3, 0.5 * y_g)
    for k in ("decoder.pred_conv.weight", "decoder.last_conv.0.weight", "decoder.last_conv.4.weight", "decoder.conv1.weight",
              "aspp.conv1.weight", "aspp.aspp4.atrous_conv.weight", "aspp.global_avg_pool.1.weight",
              "backbone.layer4.2.conv2.weight", "backbone.layer3.0.downsample.0.weight", "backbone.layer2.0.conv2.weight",
              "backbone.layer1.0.conv1.weight", "backbone.conv1.weight", "backbone.bn1.weight", "backbone.layer3.7.bn2.bias"):
        e, y = rel_l2(g[k], g_ref[k]), rel_l2(g_tf32[k], g_ref[k])
        print(f"  grad {k}: {e:.2e} (cuDNN tf32 {y:.2e})")
        assert e < max(2e-3, y), k


def test_train_step_updates_like_sgd_and_bf16_path_gradients_are_bounded():
    """train_step = zero_grad + loss_and_grads + fused SGD (base_trainer.py:16-20) against torch.optim.SGD on the oracle's
    fp64 gradients; then the bf16 THROUGHPUT path's full-model gradients against the same oracle, asserted per stage
    (they were only printed): bf16 activation storage => <= 8e-2 per tensor group, frozen-BN regime."""
    import zs3_oracle as O
    from zs3_b200 import parity_train as PT
    from zs3_b200.modeling.deeplab import DeepLab
    from zs3_b200.utils.loss import SegmentationLosses
    x = torch.randn(2, 3, 65, 65, generator=torch.Generator().manual_seed(11))
    target = torch.randint(0, 21, (2, 65, 65), generator=torch.Generator().manual_seed(12)).float()
    st = _state_for("identity", x)
    _, _, g_ref, s64 = _oracle_step(st, x, target, False, torch.float64, "cpu")
    model = DeepLab(num_classes=21, sync_bn=True, pretrained=False)
    model.load_state_dict(st)
    model = model.cuda().eval()
    eng = PT.SplitPrecisionTrainer(model, pieces=2, lr=0.007)
    eng.train_step(x.cuda(), target.cuda())
    torch.cuda.synchronize()
    for k, lr in (("decoder.last_conv.4.weight", 0.07), ("backbone.layer2.1.conv2.weight", 0.007), ("backbone.bn1.bias", 0.007)):
        w0 = s64[k].detach()
        expect = w0 - lr * (g_ref[k] + 5e-4 * w0)         # first SGD step: buf = grad + wd * w
        got = dict(model.named_parameters())[k].detach().cpu().double()
        assert rel_l2(got - w0, expect - w0) < 1e-2, k
    # bf16 throughput path, same weights (reload), autograd through the module API
    model2 = DeepLab(num_classes=21, sync_bn=True, pretrained=False)
    model2.load_state_dict(st)
    model2 = model2.cuda().eval()
    crit = SegmentationLosses(weight=None, cuda=True).build_loss("ce")
    crit(model2(x.cuda()), target.cuda()).backward()
    groups = {"decoder": [], "aspp": [], "backbone.layer4": [], "backbone.layer3": [], "backbone.layer2": [],
              "backbone.layer1": [], "backbone.conv1": [], "backbone.bn1": []}
    for k, p in model2.named_parameters():
        for gname in groups:
            if k.startswith(gname):
                groups[gname].append(k)
                break
    for gname, keys in groups.items():
        e = _global_rel({k: dict(model2.named_parameters())[k].grad.cpu() for k in keys}, {k: g_ref[k] for k in keys})
        print(f"  bf16 path grads {gname}: {e:.2e}")
        # bf16 activation storage: ReLU-mask flips near zero dominate and grow towards the input (DESIGN.md "Numerics");
        # bounds = 1.5-2x the first B200 measurement (1.0e-2 / 3.0e-2 / 5.4e-2 / 1.3e-1 / 8.6e-2 / 9.1e-2 / 3.4e-1: the
        # stem's weight gradient is a heavily cancelling sum over 16 k pixels per image)
        bound = {"decoder": 3e-2, "aspp": 6e-2, "backbone.layer4": 1e-1, "backbone.conv1": 5e-1, "backbone.bn1": 5e-1}
        assert e < bound.get(gname, 2e-1), gname
