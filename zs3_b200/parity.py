"""fp32-grade forward of the DeepLab path on the SAME tcgen05 conv kernel ("fp32x3" parity mode).

The bf16 tensor-core path stores activations in bf16, which bounds end-to-end agreement with the fp32 reference
at ~1e-2 (DESIGN.md "Numerics").  BASELINE.json's north star asks for logits within 1e-3, so this module runs
the network with fp32 activations and emulates fp32 convolutions on the bf16 tensor cores: both operands are
split into three bf16 pieces (24 mantissa bits) and the six significant cross products are reduced as six
K-segments of one fp32 TMEM accumulator by zs3_conv_fprop.  Forward only (no autograd); about 6x the tensor work
of the bf16 path, so it is a verification mode, not the throughput path.
"""
import ctypes as C

import torch

from . import _lib as L
from . import kernels as K

# (activation piece, weight piece) pairs, smallest products first so the fp32 accumulator adds them in
# increasing magnitude: lo=2, mid=1, hi=0
_PAIRS = [(1, 1), (0, 2), (2, 0), (0, 1), (1, 0), (0, 0)]


def _st():
    return L.stream_ptr()


def split3(x):
    """fp32 tensor -> three bf16 tensors with x == hi + mid + lo up to 2^-24 |x|"""
    x = x.contiguous()
    outs = [torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) for _ in range(3)]
    L.check(L.lib().zs3_split3_f32(L.ptr(x), L.ptr(outs[0]), L.ptr(outs[1]), L.ptr(outs[2]), x.numel(), _st()),
            "zs3_split3_f32")
    return outs


def _pack_components(w, cout_p, cin_p, ci0, ci_count):
    """three packed bf16 pieces [cout_p][taps][cin_p] of an fp32 conv weight (OIHW-contiguous or KRSC)"""
    cout, cin, r, s = w.shape
    krsc = K.is_krsc(w) and not (w.is_contiguous() and r * s > 1)
    src = w.detach() if krsc else w.detach().contiguous()
    outs = []
    for comp in range(3):
        dst = torch.empty((cout_p, r * s, cin_p), dtype=torch.bfloat16, device=w.device)
        L.check(L.lib().zs3_pack_weight_component(L.ptr(src), cout, cin, r, s, ci0, ci_count, L.ptr(dst), cout_p, cin_p,
                                                  int(krsc), comp, _st()), "zs3_pack_weight_component")
        outs.append(dst)
    return outs


def conv_fp32(xs, channels, weight, R, S, stride, pad, dil, cout, bias=None, stats=None):
    """fp32-grade conv over the virtual concat of fp32 NHWC tensors `xs`; returns fp32 [N,Ho,Wo,cpad(cout)]"""
    cout_p = K.cpad(cout)
    y = None
    ci = 0
    for i, (x, c_real) in enumerate(zip(xs, channels)):
        cin_p = x.shape[3]
        x3 = split3(x)
        w3 = _pack_components(weight, cout_p, cin_p, ci, c_real)
        segs = [(x3[a], w3[b]) for a, b in _PAIRS]
        last = i == len(xs) - 1
        y = K.conv_fprop(segs, R, S, stride, pad, dil, cout_p, out=y, out_f32=True, accumulate=i > 0,
                         bias=bias if i == 0 else None, stats=stats if last else None)
        ci += c_real
    return y


def _bn_coeffs(bn, y, training, stats):
    n, h, w, cs = y.shape
    if bn is None:
        ones = torch.ones(cs, dtype=torch.float32, device=y.device)
        return ones, torch.zeros_like(ones)
    wgt = bn.weight.detach() if bn.weight is not None else None
    b = bn.bias.detach() if bn.bias is not None else None
    if training:
        sc, sh, _, _ = K.bn_finalize(stats, n * h * w, wgt, b, bn.eps, 0.1, None, None, cs)  # running stats untouched
    else:
        sc, sh, _, _ = K.bn_eval_coeffs(wgt, b, bn.running_mean, bn.running_var, bn.eps, cs)
    return sc, sh


def conv_bn_act_fp32(xs, channels, conv, bn, relu=True, residual=None, weight=None, geom=None):
    weight = conv.weight if weight is None else weight
    R, S, stride, pad, dil = geom or (conv.kernel_size[0], conv.kernel_size[1], conv.stride[0], conv.padding[0],
                                      conv.dilation[0])
    cout = weight.shape[0]
    training = bn is not None and bn.training
    stats = None
    if training:
        z = torch.zeros(2, K.cpad(cout), dtype=torch.float64, device=xs[0].device)
        stats = (z[0], z[1])
    y = conv_fp32(xs, channels, weight, R, S, stride, pad, dil, cout, stats=stats)
    scale, shift = _bn_coeffs(bn, y, training, stats)
    n, h, w, cs = y.shape
    out = torch.empty_like(y)
    L.check(L.lib().zs3_bn_apply_f32(L.ptr(y), cs, L.ptr(residual), residual.shape[3] if residual is not None else 0,
                                     L.ptr(out), cs, L.ptr(scale), L.ptr(shift), n * h * w, cs, int(relu), _st()),
            "zs3_bn_apply_f32")
    return out


def _bilinear(x, ho, wo, c_out=None, to_nchw=False):
    n, hi, wi, cs = x.shape
    c_out = cs if c_out is None else c_out
    y = torch.empty((n, c_out, ho, wo) if to_nchw else (n, ho, wo, cs), dtype=torch.float32, device=x.device)
    L.check(L.lib().zs3_bilinear_f32(L.ptr(x), L.ptr(y), n, hi, wi, ho, wo, cs, c_out, int(to_nchw), _st()),
            "zs3_bilinear_f32")
    return y


def _check_dropout_inactive(model):
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout) and m.training and m.p > 0:
            raise RuntimeError("parity forward needs Dropout inactive (eval mode or p=0): masks are RNG dependent")


@torch.no_grad()
def deeplab_forward_fp32x3(model, input, return_features=False):
    """DeepLab.forward (zs3/modeling/deeplab.py:40-45) with fp32-grade arithmetic.  BatchNorm follows each module's
    .training flag (batch statistics in train mode; running statistics are NOT updated)."""
    _check_dropout_inactive(model)
    if not input.is_cuda:
        raise RuntimeError("zs3_b200 runs on CUDA (sm_100a) tensors only")
    bb = model.backbone
    x = input.contiguous().float()
    n, c, h, w = x.shape
    # stem: im2col (fp32) -> GEMM -> BN -> ReLU -> maxpool
    conv = bb.conv1
    R, stride, pad = conv.kernel_size[0], conv.stride[0], conv.padding[0]
    ho, wo = K.conv_out_size(h, R, stride, pad, 1), K.conv_out_size(w, R, stride, pad, 1)
    kreal = c * R * R
    kpad = K.cpad(kreal)
    cols = torch.empty((n, ho, wo, kpad), dtype=torch.float32, device=x.device)
    L.check(L.lib().zs3_stem_im2col_f32(L.ptr(x), L.ptr(cols), n, c, h, w, R, stride, pad, ho, wo, kpad, 0, _st()),
            "zs3_stem_im2col_f32")
    w2d = conv.weight.detach().reshape(conv.out_channels, kreal, 1, 1).contiguous()
    a = conv_bn_act_fp32([cols], [kreal], conv, bb.bn1, relu=True, weight=w2d, geom=(1, 1, 1, 0, 1))
    k, ps, pp = bb.maxpool.kernel_size, bb.maxpool.stride, bb.maxpool.padding
    hp, wp = K.conv_out_size(ho, k, ps, pp, 1), K.conv_out_size(wo, k, ps, pp, 1)
    cs = a.shape[3]
    xcur = torch.empty((n, hp, wp, cs), dtype=torch.float32, device=x.device)
    L.check(L.lib().zs3_maxpool_f32(L.ptr(a), L.ptr(xcur), n, ho, wo, cs, hp, wp, k, ps, pp, _st()), "zs3_maxpool_f32")
    low = None
    for li, layer in enumerate([bb.layer1, bb.layer2, bb.layer3, bb.layer4]):
        for blk in layer:
            out = conv_bn_act_fp32([xcur], [blk.inplanes], blk.conv1, blk.bn1)
            out = conv_bn_act_fp32([out], [blk.planes], blk.conv2, blk.bn2)
            res = xcur
            if blk.downsample is not None:
                res = conv_bn_act_fp32([xcur], [blk.inplanes], blk.downsample[0], blk.downsample[1], relu=False)
            xcur = conv_bn_act_fp32([out], [blk.planes], blk.conv3, blk.bn3, relu=True, residual=res)
        if li == 0:
            low = xcur
    # ASPP
    asp = model.aspp
    n, h33, w33, cs = xcur.shape
    branches = [conv_bn_act_fp32([xcur], [asp.inplanes], m.atrous_conv, m.bn) for m in (asp.aspp1, asp.aspp2, asp.aspp3, asp.aspp4)]
    g = torch.empty((n, 1, 1, cs), dtype=torch.float32, device=x.device)
    L.check(L.lib().zs3_spatial_sum_f32(L.ptr(xcur), L.ptr(g), n, h33 * w33, cs, 1.0 / (h33 * w33), _st()),
            "zs3_spatial_sum_f32")
    gbn = asp.global_avg_pool[2] if asp.global_avg_pool_bn else None
    g = conv_bn_act_fp32([g], [asp.inplanes], asp.global_avg_pool[1], gbn)
    gb = torch.empty((n, h33, w33, g.shape[3]), dtype=torch.float32, device=x.device)
    L.check(L.lib().zs3_spatial_broadcast_f32(L.ptr(g), L.ptr(gb), n, h33 * w33, g.shape[3], _st()),
            "zs3_spatial_broadcast_f32")
    a = conv_bn_act_fp32(branches + [gb], [256] * 5, asp.conv1, asp.bn1)
    # decoder
    dec = model.decoder
    lowf = conv_bn_act_fp32([low], [256], dec.conv1, dec.bn1)
    up = _bilinear(a, lowf.shape[1], lowf.shape[2])
    f = conv_bn_act_fp32([up, lowf], [256, 48], dec.last_conv[0], dec.last_conv[1])
    f = conv_bn_act_fp32([f], [256], dec.last_conv[4], dec.last_conv[5])
    if return_features:
        return f[..., :256].permute(0, 3, 1, 2).contiguous()
    pc = dec.pred_conv
    bias_p = torch.zeros(K.cpad(pc.out_channels), dtype=torch.float32, device=x.device)
    if pc.bias is not None:
        bias_p[:pc.out_channels] = pc.bias.detach()
    logits_small = conv_fp32([f], [256], pc.weight, 1, 1, 1, 0, 1, pc.out_channels, bias=bias_p)
    return _bilinear(logits_small, h, w, c_out=pc.out_channels, to_nchw=True)
