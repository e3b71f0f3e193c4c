"""Per-kernel parity of the tcgen05 implicit-GEMM conv (fprop, dgrad-as-fprop, wgrad) against
torch.nn.functional.conv2d in fp32 on the SAME bf16-rounded operands (so the only differences are
fp32 accumulation order and, for bf16 outputs, the final rounding)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TOL_F32_OUT = 2e-5   # fp32 accumulate, different summation order
TOL_BF16_OUT = 4e-3  # one bf16 rounding of the output (2^-9 relative per element)


def _mk(N, H, W, Cin, Cout, R, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, Cin, H, W, generator=g).to(torch.bfloat16).float()
    w = (torch.randn(Cout, Cin, R, R, generator=g) / (Cin * R * R) ** 0.5).to(torch.bfloat16).float()
    return x.cuda(), w.cuda()


CONV_CASES = [
    # N, H, W, Cin, Cout, R, stride, pad, dil
    (2, 17, 17, 64, 64, 1, 1, 0, 1),
    (2, 33, 33, 256, 128, 1, 1, 0, 1),
    (1, 33, 33, 128, 256, 1, 1, 0, 1),
    (2, 33, 33, 64, 64, 3, 1, 1, 1),
    (2, 33, 33, 128, 256, 3, 1, 2, 2),
    (2, 33, 33, 64, 512, 3, 1, 12, 12),
    (2, 65, 65, 64, 128, 3, 2, 1, 1),
    (2, 65, 65, 128, 256, 1, 2, 0, 1),
    (3, 20, 20, 48, 21, 3, 1, 1, 1),      # channel padding on both sides
    (16, 33, 33, 256, 256, 3, 1, 1, 1),   # layer3 conv2 shape at full batch
    # filter rows that lie in the zero padding for whole tiles are skipped (valid_filter_rows): ASPP dilations at 33x33,
    # a dilation larger than the map (OS8 ASPP: only the centre row ever touches a pixel), tiles straddling two images,
    # maps so small that one 128-pixel tile spans several images, a strided 3x3
    (5, 33, 33, 64, 256, 3, 1, 18, 18),
    (3, 33, 33, 128, 64, 3, 1, 6, 6),
    (2, 17, 17, 64, 64, 3, 1, 24, 24),
    (9, 5, 5, 64, 64, 3, 1, 4, 4),
    (7, 9, 9, 64, 128, 3, 1, 3, 3),
    (3, 33, 33, 64, 64, 3, 2, 8, 8),
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("out_f32", [True, False])
def test_fprop(case, out_f32):
    from zs3_b200 import kernels as K
    N, H, W, Cin, Cout, R, stride, pad, dil = case
    x, w = _mk(N, H, W, Cin, Cout, R)
    ref = F.conv2d(x, w, stride=stride, padding=pad, dilation=dil)
    cin_p, cout_p = K.cpad(Cin), K.cpad(Cout)
    xh = K.nchw_to_nhwc(x, cin_p)
    wp = K.pack_weight(w, cout_p, cin_p)
    stats = torch.zeros(2, cout_p, dtype=torch.float64, device="cuda")
    y = K.conv_fprop([(xh, wp)], R, R, stride, pad, dil, cout_p, out_f32=out_f32, stats=(stats[0], stats[1]))
    torch.cuda.synchronize()
    got = y[..., :Cout].permute(0, 3, 1, 2).float()
    err = rel_l2(got, ref)
    assert err < (TOL_F32_OUT if out_f32 else TOL_BF16_OUT), f"rel_l2={err}"
    if Cout < cout_p:
        assert y[..., Cout:].abs().max().item() == 0.0
    # fused BatchNorm statistics.  fp32 outputs: statistics of the fp32 accumulators.  bf16 outputs: statistics of the
    # STORED (bf16-rounded) tensor, i.e. exactly what the normalisation pass will read (tight), which differ from the
    # fp32 ones only by bf16 rounding noise (loose: 2^-9 per element, averaging out over the batch)
    src = ref if out_f32 else got
    s_ref = src.double().sum(dim=(0, 2, 3))
    q_ref = (src.double() ** 2).sum(dim=(0, 2, 3))
    assert rel_l2(stats[0, :Cout], s_ref) < 1e-4 or (stats[0, :Cout] - s_ref).abs().max() < 1e-2
    assert rel_l2(stats[1, :Cout], q_ref) < 1e-5
    if not out_f32:
        assert rel_l2(stats[0, :Cout], ref.double().sum(dim=(0, 2, 3))) < 5e-3
        assert rel_l2(stats[1, :Cout], (ref.double() ** 2).sum(dim=(0, 2, 3))) < 1e-3


def test_fprop_bias_accumulate_segments():
    from zs3_b200 import kernels as K
    N, H, W = 2, 33, 33
    xa, wa = _mk(N, H, W, 256, 256, 3, seed=1)
    xb, wb = _mk(N, H, W, 48, 256, 3, seed=2)
    bias = torch.randn(256, device="cuda")
    ref = F.conv2d(torch.cat([xa, xb], 1), torch.cat([wa, wb], 1), bias=bias, padding=1)
    w_full = torch.cat([wa, wb], 1)
    segs = [(K.nchw_to_nhwc(xa, 256), K.pack_weight(w_full, 256, 256, 0, 256)),
            (K.nchw_to_nhwc(xb, 64), K.pack_weight(w_full, 256, 64, 256, 48))]
    y = K.conv_fprop(segs, 3, 3, 1, 1, 1, 256, out_f32=True, bias=bias)
    err = rel_l2(y.permute(0, 3, 1, 2), ref)
    assert err < TOL_F32_OUT, err
    # accumulate on top of an existing tensor
    y2 = K.conv_fprop(segs, 3, 3, 1, 1, 1, 256, out=y.clone(), accumulate=True, bias=bias)
    assert rel_l2(y2, 2 * y) < TOL_F32_OUT
    yb = y.to(torch.bfloat16)
    y3 = K.conv_fprop(segs, 3, 3, 1, 1, 1, 256, out=yb.clone(), accumulate=True)  # bf16 read-modify-write, no bias
    assert rel_l2(y3.float(), yb.float() + y - bias) < TOL_BF16_OUT


@pytest.mark.parametrize("case", [
    (2, 33, 33, 64, 64, 3, 1, 1, 1),
    (2, 33, 33, 128, 256, 3, 1, 4, 4),
    (2, 33, 33, 256, 64, 1, 1, 0, 1),
])
def test_dgrad_as_fprop(case):
    """dX = conv(dY, flipped/transposed W): the data gradient autograd computes for stride-1 convs."""
    from zs3_b200 import kernels as K
    N, H, W, Cin, Cout, R, stride, pad, dil = case
    x, w = _mk(N, H, W, Cin, Cout, R)
    x.requires_grad_(True)
    y = F.conv2d(x, w, stride=stride, padding=pad, dilation=dil)
    g = torch.Generator().manual_seed(3)
    dy = torch.randn(y.shape, generator=g).to(torch.bfloat16).float().cuda()
    (dx_ref,) = torch.autograd.grad(y, x, dy)
    cin_p, cout_p = K.cpad(Cin), K.cpad(Cout)
    wT = K.pack_weight(w, cout_p, cin_p, mode=1)  # [cin_p][taps][cout_p]
    dyh = K.nchw_to_nhwc(dy, cout_p)
    dx = K.conv_fprop([(dyh, wT)], R, R, 1, dil * (R - 1) - pad, dil, cin_p, out_f32=True)
    err = rel_l2(dx[..., :Cin].permute(0, 3, 1, 2), dx_ref)
    assert err < TOL_F32_OUT, err


def test_dgrad_strided_1x1_scatter():
    from zs3_b200 import kernels as K
    N, H, W, Cin, Cout = 2, 65, 65, 128, 256
    x, w = _mk(N, H, W, Cin, Cout, 1)
    x.requires_grad_(True)
    y = F.conv2d(x, w, stride=2)
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(5)).to(torch.bfloat16).float().cuda()
    (dx_ref,) = torch.autograd.grad(y, x, dy)
    wT = K.pack_weight(w, Cout, Cin, mode=1)
    dx = torch.zeros(N, H, W, Cin, dtype=torch.float32, device="cuda")
    K.conv_fprop([(K.nchw_to_nhwc(dy, Cout), wT)], 1, 1, 1, 0, 1, Cin, out=dx, scatter=(2, H, W))
    assert rel_l2(dx.permute(0, 3, 1, 2), dx_ref) < TOL_F32_OUT


WGRAD_CASES = [
    (2, 17, 17, 64, 64, 1, 1, 0, 1),
    (2, 33, 33, 128, 256, 1, 1, 0, 1),
    (2, 33, 33, 64, 128, 3, 1, 1, 1),
    (2, 33, 33, 256, 64, 3, 1, 2, 2),
    (2, 65, 65, 64, 128, 3, 2, 1, 1),
    (3, 20, 20, 48, 21, 3, 1, 1, 1),
    (16, 33, 33, 256, 256, 3, 1, 1, 1),
]


@pytest.mark.parametrize("case", WGRAD_CASES)
def test_wgrad(case):
    from zs3_b200 import kernels as K
    N, H, W, Cin, Cout, R, stride, pad, dil = case
    x, w = _mk(N, H, W, Cin, Cout, R)
    w.requires_grad_(True)
    y = F.conv2d(x, w, stride=stride, padding=pad, dilation=dil)
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(9)).to(torch.bfloat16).float().cuda()
    (dw_ref,) = torch.autograd.grad(y, w, dy)
    cin_p, cout_p = K.cpad(Cin), K.cpad(Cout)
    dw = K.conv_wgrad(K.nchw_to_nhwc(x, cin_p), K.nchw_to_nhwc(dy, cout_p), R, R, stride, pad, dil, cin_p, cout_p)
    g = torch.zeros_like(w)
    K.unpack_wgrad(dw, g)
    torch.cuda.synchronize()
    err = rel_l2(g, dw_ref)
    assert err < 5e-5, err


def test_pack_roundtrip_and_layout():
    from zs3_b200 import kernels as K
    w = torch.randn(21, 48, 3, 3, device="cuda").to(torch.bfloat16).float()
    wp = K.pack_weight(w, 64, 64)
    assert torch.equal(wp[:21, :, :48].float(), w.permute(0, 2, 3, 1).reshape(21, 9, 48))
    assert wp[21:].abs().max() == 0 and wp[:, :, 48:].abs().max() == 0
    wt = K.pack_weight(w, 64, 64, mode=1)
    assert torch.equal(wt[:48, :, :21].float(), w.flip(2, 3).permute(1, 2, 3, 0).reshape(48, 9, 21))
    x = torch.randn(2, 5, 7, 9, device="cuda")
    xh = K.nchw_to_nhwc(x, 64)
    assert torch.equal(xh[..., :5].float(), x.to(torch.bfloat16).float().permute(0, 2, 3, 1))
    assert torch.equal(K.nhwc_to_nchw(xh, 5), x.to(torch.bfloat16).float())


@pytest.mark.parametrize("case", [(2, 33, 33, 64, 128, 3, 1, 1, 1), (3, 20, 20, 48, 21, 3, 1, 1, 1),
                                  (2, 17, 17, 256, 256, 1, 1, 0, 1)])
def test_wgrad_direct_into_krsc_grad_and_krsc_pack(case):
    """parameters stored KRSC (torch channels_last): packing reads them as is, wgrad accumulates into .grad in place"""
    from zs3_b200 import kernels as K
    N, H, W, Cin, Cout, R, stride, pad, dil = case
    x, w = _mk(N, H, W, Cin, Cout, R)
    w_cl = w.contiguous(memory_format=torch.channels_last)
    cin_p, cout_p = K.cpad(Cin), K.cpad(Cout)
    assert torch.equal(K.pack_weight(w_cl, cout_p, cin_p), K.pack_weight(w, cout_p, cin_p))
    assert torch.equal(K.pack_weight(w_cl, cout_p, cin_p, mode=1), K.pack_weight(w, cout_p, cin_p, mode=1))
    w.requires_grad_(True)
    y = F.conv2d(x, w, stride=stride, padding=pad, dilation=dil)
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(9)).to(torch.bfloat16).float().cuda()
    (dw_ref,) = torch.autograd.grad(y, w, dy)
    # two K-segments accumulate into the same KRSC gradient tensor, twice (gradient accumulation semantics)
    g = torch.zeros_like(w_cl)
    half = (Cin // 2) // 8 * 8
    xh = K.nchw_to_nhwc(x, cin_p)
    xa, xb = K.nchw_to_nhwc(x[:, :half], K.cpad(half)), K.nchw_to_nhwc(x[:, half:], K.cpad(Cin - half))
    dyh = K.nchw_to_nhwc(dy, cout_p)
    for _ in range(2):
        K.conv_wgrad(xa, dyh, R, R, stride, pad, dil, K.cpad(half), cout_p, dw=g, dw_view=(Cin, 0, Cout, half))
        K.conv_wgrad(xb, dyh, R, R, stride, pad, dil, K.cpad(Cin - half), cout_p, dw=g,
                     dw_view=(Cin, half, Cout, Cin - half))
    assert g.is_contiguous(memory_format=torch.channels_last)
    assert rel_l2(g, 2 * dw_ref) < 5e-5


@pytest.mark.parametrize("case", [
    (2, 33, 33, 64, 64, 3, 1, 1, 1),
    (2, 33, 33, 128, 256, 3, 1, 4, 4),
    (2, 33, 33, 256, 64, 1, 1, 0, 1),
    (3, 20, 20, 48, 21, 3, 1, 1, 1),
    (2, 17, 17, 1024, 256, 1, 1, 0, 1),
    (5, 33, 33, 64, 256, 3, 1, 18, 18),   # culled filter rows in the data gradient
    (2, 17, 17, 64, 64, 3, 1, 24, 24),
    (9, 5, 5, 64, 64, 3, 1, 4, 4),
])
def test_dgrad_from_forward_packed_weights(case):
    """data gradient straight from the forward-packed weights (MN-major B tiles, taps flipped in the kernel)"""
    from zs3_b200 import kernels as K
    N, H, W, Cin, Cout, R, stride, pad, dil = case
    x, w = _mk(N, H, W, Cin, Cout, R)
    x.requires_grad_(True)
    y = F.conv2d(x, w, stride=stride, padding=pad, dilation=dil)
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16).float().cuda()
    (dx_ref,) = torch.autograd.grad(y, x, dy)
    cin_p, cout_p = K.cpad(Cin), K.cpad(Cout)
    wp = K.pack_weight(w, cout_p, cin_p)  # forward layout [cout_p][taps][cin_p]
    dx = K.conv_fprop([(K.nchw_to_nhwc(dy, cout_p), wp)], R, R, 1, dil * (R - 1) - pad, dil, cin_p, out_f32=True,
                      w_forward_layout=True)
    assert rel_l2(dx[..., :Cin].permute(0, 3, 1, 2), dx_ref) < TOL_F32_OUT
    if cin_p > Cin:
        assert dx[..., Cin:].abs().max() == 0


def test_inference_epilogue_folds_bn_residual_relu():
    from zs3_b200 import kernels as K
    N, H, W, Cin, Cout = 2, 33, 33, 128, 256
    x, w = _mk(N, H, W, Cin, Cout, 3)
    g = torch.Generator().manual_seed(4)
    scale, shift = (torch.rand(Cout, generator=g) + 0.5).cuda(), torch.randn(Cout, generator=g).cuda()
    res = torch.randn(N, Cout, H, W, generator=g).to(torch.bfloat16).float().cuda()
    ref = F.relu(F.conv2d(x, w, padding=2, dilation=2) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) + res)
    y = K.conv_fprop([(K.nchw_to_nhwc(x, Cin), K.pack_weight(w, Cout, Cin))], 3, 3, 1, 2, 2, Cout,
                     epilogue=(scale, shift, K.nchw_to_nhwc(res, Cout), True))
    assert rel_l2(y.permute(0, 3, 1, 2).float(), ref) < TOL_BF16_OUT
    y2 = K.conv_fprop([(K.nchw_to_nhwc(x, Cin), K.pack_weight(w, Cout, Cin))], 3, 3, 1, 2, 2, Cout, out_f32=True,
                      epilogue=(scale, shift, None, False))
    ref2 = F.conv2d(x, w, padding=2, dilation=2) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    assert rel_l2(y2.permute(0, 3, 1, 2), ref2) < TOL_F32_OUT


@pytest.mark.parametrize("case", [
    # N, H, Cin, Cout, R, dil, residual, relu
    (2, 33, 128, 256, 3, 2, True, True),
    (3, 17, 64, 64, 1, 1, True, True),       # 64-column tiles, rows past M in the last tile
    (2, 33, 256, 128, 1, 1, False, True),
    (1, 20, 64, 512, 1, 1, True, False),     # two n-tiles, residual without ReLU
    (2, 33, 64, 48, 3, 1, False, False),     # channel padding: the padded output channels stay zero
    (16, 33, 256, 1024, 1, 1, True, True),   # layer3 conv3 at full batch: several tiles per CTA
])
def test_inference_epilogue_hot_variant(case):
    """frozen / eval-mode BatchNorm (+ residual, + ReLU) folded into the pipelined TMA-store epilogue (MODE 3)"""
    from zs3_b200 import kernels as K
    N, H, Cin, Cout, R, dil, with_res, relu = case
    pad = dil * (R - 1) // 2
    x, w = _mk(N, H, H, Cin, Cout, R)
    cin_p, cout_p = K.cpad(Cin), K.cpad(Cout)
    g = torch.Generator().manual_seed(4)
    scale = torch.zeros(cout_p).cuda()
    shift = torch.zeros(cout_p).cuda()
    scale[:Cout] = (torch.rand(Cout, generator=g) + 0.5).cuda()
    shift[:Cout] = torch.randn(Cout, generator=g).cuda()
    res = torch.randn(N, Cout, H, H, generator=g).to(torch.bfloat16).float().cuda() if with_res else None
    ref = F.conv2d(x, w, padding=pad, dilation=dil) * scale[:Cout].view(1, -1, 1, 1) + shift[:Cout].view(1, -1, 1, 1)
    if with_res:
        ref = ref + res
    if relu:
        ref = F.relu(ref)
    y = K.conv_fprop([(K.nchw_to_nhwc(x, cin_p), K.pack_weight(w, cout_p, cin_p))], R, R, 1, pad, dil, cout_p,
                     epilogue=(scale, shift, K.nchw_to_nhwc(res, cout_p) if with_res else None, relu))
    assert rel_l2(y[..., :Cout].permute(0, 3, 1, 2).float(), ref) < TOL_BF16_OUT
    if cout_p > Cout:
        assert y[..., Cout:].abs().max().item() == 0.0
