"""Split-precision TRAINING path of the DeepLab step (forward + backward + SGD) on the same tcgen05 conv kernels.

Why it exists.  BASELINE.json's north star asks for logits within 1e-3 of the fp32 reference.  The throughput path
(zs3_b200/functional.py) stores activations in bf16 and lands at ~1e-2; the reference's own GPU path (cuDNN, TF32
allowed) carries 10 mantissa bits per operand.  This module trains the network with fp32 activations and
gradients and feeds every conv operand to the tensor cores as P bf16 pieces (x = p0 + p1 [+ p2]):

    pieces = 1   bf16 operands, fp32 storage                       1 product  per conv
    pieces = 2   16 mantissa bits per operand (> TF32's 10)        3 products per conv   <- default
    pieces = 3   24 mantissa bits ("fp32x3", as zs3_b200/parity.py) 6 products per conv

fprop and dgrad reduce the products as K-segments of ONE zs3_conv_fprop launch (fp32 TMEM accumulator, fp32
output); wgrad issues one zs3_conv_wgrad launch per product, all reduce-adding into the same KRSC gradient.
BatchNorm / ReLU / residual / Dropout / pooling / bilinear run as fp32 kernels (csrc/parity_train.cu) that write
the NEXT conv's operand pieces in the same pass.  No autograd: the backward is written out explicitly, mirroring
zs3/modeling/deeplab.py:40-45, backbone/resnet.py:33-53,186-197, aspp.py:103-116, decoder.py:29-68 and the step of
zs3/base_trainer.py:16-20 (zero_grad, forward, CE, backward, SGD).
"""
import ctypes as C
import os

import torch

from . import _lib as L
from . import kernels as K


def _st():
    return L.stream_ptr()


def _pairs(p):
    """(activation piece, weight piece) products kept: index sum < p; small products first (fp32 accumulation order)"""
    out = [(a, b) for a in range(p) for b in range(p) if a + b < p]
    return sorted(out, key=lambda ab: -(ab[0] + ab[1]))


class Act:
    """an fp32 NHWC activation [N,H,W,Cs] and/or its bf16 operand pieces"""
    __slots__ = ("f32", "pieces", "shape")

    def __init__(self, f32=None, pieces=None, shape=None):
        self.f32, self.pieces = f32, pieces
        self.shape = tuple(shape if shape is not None else (f32.shape if f32 is not None else pieces[0].shape))


def _ptr_array(tensors):
    arr = (C.c_void_p * 3)()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def split(x, pieces):
    """fp32 tensor (numel % 8 == 0) -> `pieces` bf16 tensors of the same shape"""
    x = x.contiguous()
    outs = [torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) for _ in range(pieces)]
    L.check(L.lib().zs3_split_f32(L.ptr(x), _ptr_array(outs), pieces, x.numel(), _st()), "zs3_split_f32")
    return outs


class _Rec:
    """what one conv -> BN -> act layer keeps for its backward"""
    __slots__ = ("conv", "bn", "xs", "ranges", "geom", "y", "out", "coef", "relu", "p", "training", "has_res", "weight",
                 "wgeom", "grad_view")


class SplitPrecisionTrainer:
    """One optimisation step of zs3/base_trainer.py:16-20 with fp32 activations and split bf16 operands.

    model: zs3_b200.modeling.deeplab.DeepLab (conv weights KRSC, as its constructor leaves them).
    After `loss_and_grads(image, target)` every parameter's `.grad` holds the gradient of
    criterion(model(image), target) (weighted CE with ignore_index, / batch); `train_step` adds the fused SGD update."""

    def __init__(self, model, pieces=2, class_weight=None, ignore_index=255, batch_average=True, lr=0.007, momentum=0.9,
                 weight_decay=5e-4, nesterov=False, optimizer=True):
        if pieces not in (1, 2, 3):
            raise ValueError("pieces must be 1, 2 or 3")
        self.model, self.P = model, pieces
        self.pairs = _pairs(pieces)
        self.class_weight = None if class_weight is None else class_weight.float().contiguous()
        self.ignore_index, self.batch_average = ignore_index, batch_average
        self._wcache, self._epoch = {}, 0
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("zs3_b200 runs on CUDA (sm_100a) tensors only; there is no CPU path")
        self.dev = dev
        self.stats = torch.zeros(2, 2048, dtype=torch.float64, device=dev)     # conv-epilogue BN statistics
        self.bsums = torch.zeros(2, 2048, dtype=torch.float64, device=dev)     # BN-backward sums
        self.rng_calls = 0
        self._counters = []
        self.flat = self.opt = None
        if optimizer:
            from .parallel import FlatParams, FusedSGD
            groups = [list(model.get_1x_lr_params()), list(model.get_10x_lr_params())]
            self.flat = FlatParams(groups)
            self.opt = FusedSGD(self.flat, [lr, lr * 10], momentum, weight_decay, nesterov)

    # ------------------------------------------------------------------------------------------ weights
    def _weight_pieces(self, w4, cout_p, cin_p, ci0, c_real, tag):
        """P packed bf16 pieces [cout_p][taps][cin_p] of the fp32 weight view w4 [O,I,R,S] (cached per step)"""
        key = (tag, ci0, c_real, cin_p, cout_p)
        ent = self._wcache.get(key)
        if ent is not None and ent[0] == self._epoch:
            return ent[1]
        cout, cin, r, s = w4.shape
        krsc = K.is_krsc(w4) and not (w4.is_contiguous() and r * s > 1)
        src = w4.detach() if krsc else w4.detach().contiguous()
        outs = []
        for comp in range(self.P):
            dst = torch.empty((cout_p, r * s, cin_p), dtype=torch.bfloat16, device=w4.device)
            L.check(L.lib().zs3_pack_weight_component(L.ptr(src), cout, cin, r, s, ci0, c_real, L.ptr(dst), cout_p, cin_p,
                                                      int(krsc), comp, _st()), "zs3_pack_weight_component")
            outs.append(dst)
        self._wcache[key] = (self._epoch, outs)
        return outs

    # ------------------------------------------------------------------------------------------ conv primitives
    def _conv(self, xs, wsets, geom, cout_p, bias=None, stats=None, out=None, accumulate=False, dgrad=False,
              scatter=None, flops=0.0):
        """sum over inputs i and kept (a, b) products of conv(xs[i].pieces[a], wsets[i][b]); <= 6 K-segments per launch"""
        R, S, stride, pad, dil = geom
        segs = [(x.pieces[a], w[b]) for x, w in zip(xs, wsets) for a, b in self.pairs]
        y = out
        for i in range(0, len(segs), L.ZS3_MAX_SEGMENTS):
            chunk = segs[i:i + L.ZS3_MAX_SEGMENTS]
            last = i + L.ZS3_MAX_SEGMENTS >= len(segs)
            y = K.conv_fprop(chunk, R, S, stride, pad, dil, cout_p, out=y, out_f32=True, accumulate=accumulate or i > 0,
                             bias=bias if i == 0 else None, stats=stats if last else None, scatter=scatter,
                             flops=flops * len(chunk) / max(1, len(segs)), kind="conv_dgrad" if dgrad else "conv_fprop",
                             w_forward_layout=dgrad)
        return y

    def _bn_act(self, y, scale, shift, relu, residual=None, drop_p=0.0, keep_mask=None, want_f32=True, want_pieces=True):
        n, h, w, cs = y.shape
        a = L.BnActF32Args()
        a.y, a.y_cstride = y.data_ptr(), cs
        if residual is not None:
            a.residual, a.res_cstride = residual.data_ptr(), residual.shape[3]
        a.scale, a.shift = scale.data_ptr(), shift.data_ptr()
        out = torch.empty_like(y) if want_f32 else None
        if out is not None:
            a.out, a.out_cstride = out.data_ptr(), cs
        pieces = None
        if want_pieces:
            pieces = [torch.empty(y.shape, dtype=torch.bfloat16, device=y.device) for _ in range(self.P)]
            for i, t in enumerate(pieces):
                a.pieces[i] = t.data_ptr()
            a.n_pieces, a.piece_cstride = self.P, cs
        a.M, a.C, a.relu = n * h * w, cs, int(relu)
        if drop_p > 0:
            a.drop_p = float(drop_p)
            if keep_mask is not None:
                a.drop_mode, a.keep_mask = 2, keep_mask.data_ptr()
            else:
                self.rng_calls += 1
                a.drop_mode = 1
                a.seed = (torch.initial_seed() + 0x9E3779B97F4A7C15 * int(os.environ.get("RANK", "0"))) & 0xFFFFFFFFFFFFFFFF
                a.offset = (self.rng_calls << 36) & 0xFFFFFFFFFFFFFFFF
        L.check(L.lib().zs3_bn_act_f32(C.byref(a), _st()), "zs3_bn_act_f32")
        return Act(out, pieces, y.shape)

    # ------------------------------------------------------------------------------------------ conv -> BN -> act
    def cba(self, conv, bn, xs, channels, relu=True, residual=None, drop=None, keep_mask=None, weight=None, geom=None,
            want_f32=True, tag=None, grad_view=None):
        """[concat ->] conv -> BatchNorm -> (+residual) -> (ReLU) -> (Dropout); returns (Act, record)"""
        w4 = conv.weight if weight is None else weight
        R, S, stride, pad, dil = geom or (conv.kernel_size[0], conv.kernel_size[1], conv.stride[0], conv.padding[0],
                                          conv.dilation[0])
        cout = w4.shape[0]
        cout_p = K.cpad(cout)
        tag = tag or id(conv)
        wsets, ranges, ci = [], [], 0
        for x, c_real in zip(xs, channels):
            cin_p = x.shape[3]
            wsets.append(self._weight_pieces(w4, cout_p, cin_p, ci, c_real, tag))
            ranges.append((ci, c_real, cin_p))
            ci += c_real
        if ci != w4.shape[1]:
            raise ValueError(f"segments provide {ci} channels, conv expects {w4.shape[1]}")
        n, h, w_, _ = xs[0].shape
        ho, wo = K.conv_out_size(h, R, stride, pad, dil), K.conv_out_size(w_, S, stride, pad, dil)
        training = bn is not None and bn.training
        stats = (self.stats[0, :cout_p], self.stats[1, :cout_p]) if training else None
        fl = 2.0 * n * ho * wo * cout * R * S * w4.shape[1] * len(self.pairs)
        y = self._conv(xs, wsets, (R, S, stride, pad, dil), cout_p, stats=stats, flops=fl)
        if bn is None:
            coef = torch.zeros((4, cout_p), dtype=torch.float32, device=y.device)
            coef[0, :cout] = 1.0
            coef[3, :cout] = 1.0
            coef = (coef[0], coef[1], coef[2], coef[3])
        else:
            g = bn.weight.detach() if bn.weight is not None else None
            b = bn.bias.detach() if bn.bias is not None else None
            if training:
                mom = bn.momentum if bn.momentum is not None else 0.1
                coef = K.bn_finalize(stats, n * ho * wo, g, b, bn.eps, mom, bn.running_mean, bn.running_var, cout_p)
                if bn.num_batches_tracked is not None:
                    self._counters.append(bn.num_batches_tracked)
            else:
                coef = K.bn_eval_coeffs(g, b, bn.running_mean, bn.running_var, bn.eps, cout_p)
        p = float(drop.p) if (drop is not None and drop.training and drop.p > 0) else 0.0
        out = self._bn_act(y, coef[0], coef[1], relu, residual=None if residual is None else residual.f32, drop_p=p,
                           keep_mask=keep_mask if p > 0 else None, want_f32=want_f32)
        rec = _Rec()
        rec.conv, rec.bn, rec.xs, rec.ranges = conv, bn, list(xs), ranges
        rec.geom = (R, S, stride, pad, dil, cout, cout_p)
        rec.y, rec.out, rec.coef, rec.relu, rec.p, rec.training = y, out, coef, relu, p, training
        rec.has_res, rec.weight, rec.wgeom, rec.grad_view = residual is not None, w4, tag, grad_view
        return out, rec

    def cba_backward(self, rec, dout, need_dx, dx_into=None, need_w=True):
        """backward of cba.  dout: fp32 [N,Ho,Wo,cout_p].  need_dx: per-input flags.  dx_into: per-input (tensor,
        accumulate) targets or None.  Weight / affine gradients are ACCUMULATED into the parameters' .grad.
        Returns (dxs, dres)."""
        R, S, stride, pad, dil, cout, cout_p = rec.geom
        y, bn = rec.y, rec.bn
        n, ho, wo, _ = y.shape
        dev = y.device
        a = L.BnBwdF32Args()
        a.dout, a.dout_cstride = dout.data_ptr(), dout.shape[3]
        if rec.relu or rec.p > 0:
            a.relu = 1
            if rec.out.f32 is not None:
                a.act, a.act_cstride = rec.out.f32.data_ptr(), cout_p
            else:
                a.act_hi, a.act_hi_cstride = rec.out.pieces[0].data_ptr(), cout_p
        a.y, a.y_cstride = y.data_ptr(), cout_p
        scale, shift, mean, invstd = rec.coef
        a.mean, a.invstd, a.scale = mean.data_ptr(), invstd.data_ptr(), scale.data_ptr()
        a.M, a.C = n * ho * wo, cout_p
        a.grad_scale, a.training = 1.0 / (1.0 - rec.p), int(rec.training)
        a.sum_dz, a.sum_dzx = self.bsums[0].data_ptr(), self.bsums[1].data_ptr()
        dy_pieces = [torch.empty(y.shape, dtype=torch.bfloat16, device=dev) for _ in range(self.P)]
        for i, t in enumerate(dy_pieces):
            a.dy_pieces[i] = t.data_ptr()
        a.n_pieces, a.piece_cstride = self.P, cout_p
        dres = None
        if rec.has_res:
            dres = torch.empty_like(y)
            a.dres, a.dres_cstride = dres.data_ptr(), cout_p
        if bn is not None and bn.weight is not None and bn.weight.requires_grad and need_w:
            for prm in (bn.weight, bn.bias):
                if prm.grad is None:
                    prm.grad = torch.zeros_like(prm)
            a.dgamma, a.dbeta, a.C_real, a.param_accumulate = bn.weight.grad.data_ptr(), bn.bias.grad.data_ptr(), cout, 1
        L.check(L.lib().zs3_bn_bwd_f32(C.byref(a), _st()), "zs3_bn_bwd_f32")
        dy = Act(None, dy_pieces, y.shape)
        # ---- data gradients
        zero_insert = stride > 1 and R > 1
        dy_z = dy
        if zero_insert and any(need_dx):
            hz, wz = (ho - 1) * stride + 1, (wo - 1) * stride + 1
            zs = []
            for t in dy_pieces:
                z = torch.zeros((n, hz, wz, cout_p), dtype=torch.bfloat16, device=dev)
                z[:, ::stride, ::stride] = t
                zs.append(z)
            dy_z = Act(None, zs, zs[0].shape)
        dxs = []
        w4 = rec.weight
        for i, x in enumerate(rec.xs):
            ci0, c_real, cin_p = rec.ranges[i]
            if not need_dx[i]:
                dxs.append(None)
                continue
            wt = self._weight_pieces(w4, cout_p, cin_p, ci0, c_real, rec.wgeom)   # forward-packed, read MN-major
            h_in, w_in = x.shape[1], x.shape[2]
            tgt, acc = (dx_into[i] if dx_into is not None and dx_into[i] is not None else (None, False))
            fl = 2.0 * n * ho * wo * cout * R * S * c_real * len(self.pairs)
            if stride == 1:
                dx = self._conv([dy], [wt], (R, S, 1, dil * (R - 1) - pad, dil), cin_p, out=tgt, accumulate=acc,
                                dgrad=True, flops=fl)
            elif R == 1:
                dx = tgt if tgt is not None else torch.zeros(x.shape, dtype=torch.float32, device=dev)
                self._conv([dy], [wt], (1, 1, 1, 0, 1), cin_p, out=dx, accumulate=acc, dgrad=True,
                           scatter=(stride, h_in, w_in), flops=fl)
            else:
                hp, wp = h_in + 2 * pad - dil * (R - 1), w_in + 2 * pad - dil * (S - 1)
                padded = Act(None, [_pad_to(t, hp, wp) for t in dy_z.pieces])
                dx = self._conv([padded], [wt], (R, S, 1, dil * (R - 1) - pad, dil), cin_p, out=tgt, accumulate=acc,
                                dgrad=True, flops=fl)
            dxs.append(dx)
        # ---- weight gradient: one launch per kept product, reduce-added into the (KRSC) .grad
        if need_w and w4.requires_grad:
            base = rec.conv.weight
            if base.grad is None:
                base.grad = torch.zeros_like(base)
            gview = base.grad if rec.grad_view is None else rec.grad_view(base.grad)
            if gview.data_ptr() != base.grad.data_ptr() or gview.shape != w4.shape:
                raise RuntimeError("weight-gradient view must alias the parameter's .grad")
            direct = K.is_krsc(gview) or (gview.shape[2] * gview.shape[3] == 1 and gview.is_contiguous())
            for i, x in enumerate(rec.xs):
                ci0, c_real, cin_p = rec.ranges[i]
                fl = 2.0 * n * ho * wo * cout * R * S * c_real
                if direct:
                    for pa, pb in self.pairs:
                        K.conv_wgrad(x.pieces[pa], dy_pieces[pb], R, S, stride, pad, dil, cin_p, cout_p, dw=gview,
                                     flops=fl, dw_view=(w4.shape[1], ci0, cout, c_real))
                else:
                    dw = torch.zeros((cout_p, R * S, cin_p), dtype=torch.float32, device=dev)
                    for pa, pb in self.pairs:
                        K.conv_wgrad(x.pieces[pa], dy_pieces[pb], R, S, stride, pad, dil, cin_p, cout_p, dw=dw, flops=fl)
                    K.unpack_wgrad(dw, gview, ci0, c_real, accumulate=True)
        return dxs, dres

    # ------------------------------------------------------------------------------------------ network pieces
    def _stem(self, x):
        bb = self.model.backbone
        conv = bb.conv1
        n, c, h, w = x.shape
        R, stride, pad = conv.kernel_size[0], conv.stride[0], conv.padding[0]
        ho, wo = K.conv_out_size(h, R, stride, pad, 1), K.conv_out_size(w, R, stride, pad, 1)
        kreal, kpad = c * R * R, K.cpad(c * R * R)
        krsc = K.is_krsc(conv.weight) and not conv.weight.is_contiguous()
        cols = torch.empty((n, ho, wo, kpad), dtype=torch.float32, device=x.device)
        L.check(L.lib().zs3_stem_im2col_f32(L.ptr(x), L.ptr(cols), n, c, h, w, R, stride, pad, ho, wo, kpad, int(krsc),
                                            _st()), "zs3_stem_im2col_f32")
        cols_act = Act(None, split(cols, self.P), cols.shape)
        del cols
        wd = conv.weight.detach()
        w2d = wd.permute(0, 2, 3, 1).reshape(conv.out_channels, kreal, 1, 1) if krsc else \
            wd.reshape(conv.out_channels, kreal, 1, 1)
        w2d.requires_grad_(conv.weight.requires_grad)
        shape2d = tuple(w2d.shape)
        # the [cout][147] GEMM gradient IS the parameter's memory order: (r, s, c) for KRSC storage, (c, r, s) for OIHW
        gv = (lambda g: g.permute(0, 2, 3, 1).reshape(shape2d)) if krsc else (lambda g: g.reshape(shape2d))
        a, rec = self.cba(conv, bb.bn1, [cols_act], [kreal], relu=True, weight=w2d, geom=(1, 1, 1, 0, 1),
                          tag=("stem", id(conv)), grad_view=gv)
        k, ps, pp = bb.maxpool.kernel_size, bb.maxpool.stride, bb.maxpool.padding
        hp, wp = K.conv_out_size(ho, k, ps, pp, 1), K.conv_out_size(wo, k, ps, pp, 1)
        cs = a.shape[3]
        y = torch.empty((n, hp, wp, cs), dtype=torch.float32, device=x.device)
        arg = torch.empty((n, hp, wp, cs), dtype=torch.uint8, device=x.device)
        L.check(L.lib().zs3_maxpool_arg_f32(L.ptr(a.f32), L.ptr(y), L.ptr(arg), n, ho, wo, cs, hp, wp, k, ps, pp, _st()),
                "zs3_maxpool_arg_f32")
        out = Act(y, split(y, self.P))
        return out, (rec, arg, (n, ho, wo, cs), (k, ps, pp), krsc)

    def _stem_backward(self, saved, dout):
        rec, arg, in_shape, (k, ps, pp), krsc = saved
        n, ho, wo, cs = in_shape
        da = torch.empty(in_shape, dtype=torch.float32, device=dout.device)
        L.check(L.lib().zs3_maxpool_bwd_f32(L.ptr(dout), L.ptr(arg), L.ptr(da), n, ho, wo, cs, dout.shape[1], dout.shape[2],
                                            k, ps, pp, _st()), "zs3_maxpool_bwd_f32")
        self.cba_backward(rec, da, [False])   # the image needs no gradient

    def _bottleneck(self, blk, x):
        o1, r1 = self.cba(blk.conv1, blk.bn1, [x], [blk.inplanes], want_f32=False)
        o2, r2 = self.cba(blk.conv2, blk.bn2, [o1], [blk.planes], want_f32=False)
        rd, res = None, x
        if blk.downsample is not None:
            res, rd = self.cba(blk.downsample[0], blk.downsample[1], [x], [blk.inplanes], relu=False)
        o3, r3 = self.cba(blk.conv3, blk.bn3, [o2], [blk.planes], relu=True, residual=res)
        return o3, (r1, r2, r3, rd)

    def _bottleneck_backward(self, saved, dout, need_x=True):
        r1, r2, r3, rd = saved
        (d2,), dres = self.cba_backward(r3, dout, [True])
        (d1,), _ = self.cba_backward(r2, d2, [True])
        dx = None
        if rd is not None:
            (dx,), _ = self.cba_backward(rd, dres, [need_x])
        elif need_x:
            dx = dres
        into = [(dx, True)] if (need_x and dx is not None) else None
        (dx1,), _ = self.cba_backward(r1, d1, [need_x], dx_into=into)
        return dx if dx is not None else dx1

    # ------------------------------------------------------------------------------------------ forward / backward
    def forward(self, image, keep_masks=None):
        """returns (logits [N,C,H,W] fp32, tape)"""
        m = self.model
        km = keep_masks or {}
        self.stats.zero_()
        self._counters = []
        x = image.contiguous().float()
        tape = {}
        cur, tape["stem"] = self._stem(x)
        tape["blocks"] = []
        low = None
        for li, layer in enumerate([m.backbone.layer1, m.backbone.layer2, m.backbone.layer3, m.backbone.layer4]):
            for blk in layer:
                cur, sv = self._bottleneck(blk, cur)
                tape["blocks"].append(sv)
            if li == 0:
                low = cur
        asp = m.aspp
        n, h, w, cs = cur.shape
        branches, recs = [], []
        for i in range(1, 5):
            mod = getattr(asp, f"aspp{i}")
            o, r = self.cba(mod.atrous_conv, mod.bn, [cur], [asp.inplanes], want_f32=False)
            branches.append(o)
            recs.append(r)
        pooled = torch.empty((n, 1, 1, cs), dtype=torch.float32, device=x.device)
        L.check(L.lib().zs3_spatial_sum_f32(L.ptr(cur.f32), L.ptr(pooled), n, h * w, cs, 1.0 / (h * w), _st()),
                "zs3_spatial_sum_f32")
        pooled_act = Act(pooled, split(pooled, self.P))
        gbn = asp.global_avg_pool[2] if asp.global_avg_pool_bn else None
        pg, rg = self.cba(asp.global_avg_pool[1], gbn, [pooled_act], [asp.inplanes])
        width = asp.conv1.out_channels
        gb = torch.empty((n, h, w, pg.shape[3]), dtype=torch.float32, device=x.device)
        L.check(L.lib().zs3_spatial_broadcast_f32(L.ptr(pg.f32), L.ptr(gb), n, h * w, pg.shape[3], _st()),
                "zs3_spatial_broadcast_f32")
        gb_act = Act(None, split(gb, self.P), gb.shape)
        del gb
        a, rp = self.cba(asp.conv1, asp.bn1, branches + [gb_act], [width] * 5, drop=asp.dropout,
                         keep_mask=km.get("aspp.dropout"))
        tape["aspp"] = (recs, rg, rp, (n, h, w, cs))
        dec = m.decoder
        lowf, rl = self.cba(dec.conv1, dec.bn1, [low], [dec.LOW_LEVEL_WIDTH], want_f32=False)
        hl, wl = lowf.shape[1], lowf.shape[2]
        up = torch.empty((n, hl, wl, a.shape[3]), dtype=torch.float32, device=x.device)
        L.check(L.lib().zs3_bilinear_f32(L.ptr(a.f32), L.ptr(up), n, a.shape[1], a.shape[2], hl, wl, a.shape[3], a.shape[3],
                                         0, _st()), "zs3_bilinear_f32")
        up_act = Act(None, split(up, self.P), up.shape)
        del up
        lc = dec.last_conv
        f1, rf1 = self.cba(lc[0], lc[1], [up_act, lowf], [dec.WIDTH, dec.REDUCED_WIDTH], drop=lc[3],
                           keep_mask=km.get("decoder.dropout0"), want_f32=False)
        f2, rf2 = self.cba(lc[4], lc[5], [f1], [dec.WIDTH], drop=lc[7], keep_mask=km.get("decoder.dropout1"), want_f32=False)
        pc = dec.pred_conv
        ncls, cp = pc.out_channels, K.cpad(pc.out_channels)
        bias_p = torch.zeros(cp, dtype=torch.float32, device=x.device)
        if pc.bias is not None:
            bias_p[:ncls] = pc.bias.detach()
        wp = self._weight_pieces(pc.weight, cp, f2.shape[3], 0, pc.in_channels, id(pc))
        scores = self._conv([f2], [wp], (1, 1, 1, 0, 1), cp, bias=bias_p,
                            flops=2.0 * n * hl * wl * ncls * pc.in_channels * len(self.pairs))
        hh, ww = image.shape[2], image.shape[3]
        logits = torch.empty((n, ncls, hh, ww), dtype=torch.float32, device=x.device)
        L.check(L.lib().zs3_bilinear_f32(L.ptr(scores), L.ptr(logits), n, hl, wl, hh, ww, cp, ncls, 1, _st()),
                "zs3_bilinear_f32")
        tape["decoder"] = (rl, rf1, rf2, f2, (n, hl, wl, cp), (a.shape[1], a.shape[2]))
        if self._counters:
            torch._foreach_add_(self._counters, 1)     # num_batches_tracked of every training-mode BatchNorm
            self._counters = []
        return logits, tape

    def backward(self, tape, dlogits):
        """dlogits: fp32 [N,C,H,W].  Accumulates every parameter gradient into .grad."""
        m = self.model
        dec, asp = m.decoder, m.aspp
        rl, rf1, rf2, f2, (n, hl, wl, cp), (ha, wa) = tape["decoder"]
        dev = dlogits.device
        ncls = dlogits.shape[1]
        dscores = torch.zeros((n, hl, wl, cp), dtype=torch.float32, device=dev)
        L.check(L.lib().zs3_bilinear_bwd_f32(L.ptr(dlogits), L.ptr(dscores), n, hl, wl, dlogits.shape[2], dlogits.shape[3],
                                             ncls, 0, cp, 1, _st()), "zs3_bilinear_bwd_f32")
        # pred_conv (decoder.py:26): bias gradient, weight gradient, data gradient
        pc = dec.pred_conv
        ds = Act(dscores, split(dscores, self.P))
        if pc.bias is not None and pc.bias.requires_grad:
            sums = torch.empty(cp, dtype=torch.float64, device=dev)
            L.check(L.lib().zs3_channel_sums_f32(L.ptr(dscores), cp, n * hl * wl, cp, L.ptr(sums), _st()),
                    "zs3_channel_sums_f32")
            g = sums[:ncls].float()
            pc.bias.grad = g if pc.bias.grad is None else pc.bias.grad.add_(g)
        if pc.weight.requires_grad:
            if pc.weight.grad is None:
                pc.weight.grad = torch.zeros_like(pc.weight)
            for pa, pb in self.pairs:
                K.conv_wgrad(f2.pieces[pa], ds.pieces[pb], 1, 1, 1, 0, 1, f2.shape[3], cp, dw=pc.weight.grad,
                             dw_view=(pc.in_channels, 0, ncls, pc.in_channels))
        wp = self._weight_pieces(pc.weight, cp, f2.shape[3], 0, pc.in_channels, id(pc))
        df2 = self._conv([ds], [wp], (1, 1, 1, 0, 1), f2.shape[3], dgrad=True)
        (df1,), _ = self.cba_backward(rf2, df2, [True])
        (dup, dlowf), _ = self.cba_backward(rf1, df1, [True, True])
        # low-level branch: its data gradient meets layer2's at layer1's output
        (dlow,), _ = self.cba_backward(rl, dlowf, [True])
        da = torch.empty((n, ha, wa, dup.shape[3]), dtype=torch.float32, device=dev)
        L.check(L.lib().zs3_bilinear_bwd_f32(L.ptr(dup), L.ptr(da), n, ha, wa, hl, wl, dup.shape[3], dup.shape[3],
                                             dup.shape[3], 0, _st()), "zs3_bilinear_bwd_f32")
        # ASPP
        recs, rg, rp, (n, h, w, cs) = tape["aspp"]
        douts, _ = self.cba_backward(rp, da, [True] * 5)
        dpg = torch.empty((n, 1, 1, douts[4].shape[3]), dtype=torch.float32, device=dev)
        L.check(L.lib().zs3_spatial_sum_f32(L.ptr(douts[4]), L.ptr(dpg), n, h * w, douts[4].shape[3], 1.0, _st()),
                "zs3_spatial_sum_f32")
        (dpooled,), _ = self.cba_backward(rg, dpg, [True])
        dx = None
        for r, d in zip(recs, douts[:4]):
            into = [(dx, True)] if dx is not None else None
            (dxi,), _ = self.cba_backward(r, d, [True], dx_into=into)
            dx = dx if dx is not None else dxi
        L.check(L.lib().zs3_spatial_broadcast_acc_f32(L.ptr(dpooled), L.ptr(dx), n, h * w, cs, 1.0 / (h * w), 1, _st()),
                "zs3_spatial_broadcast_acc_f32")
        # backbone, last block first; the decoder's low-level gradient joins after layer2's first block has run
        blocks = tape["blocks"]
        n_l1 = len(m.backbone.layer1)
        for bi in range(len(blocks) - 1, -1, -1):
            if bi == n_l1 - 1:
                dx.add_(dlow)   # layer1's output feeds layer2 AND the decoder (resnet.py:192-197)
            dx = self._bottleneck_backward(blocks[bi], dx, need_x=True)
        self._stem_backward(tape["stem"], dx)

    def loss_and_grads(self, image, target, keep_masks=None, return_logits=False):
        """criterion(model(image), target) of zs3/utils/loss.py:31-46 and its gradients (accumulated into .grad)"""
        logits, tape = self.forward(image, keep_masks)
        n, c, h, w = logits.shape
        target = target.contiguous().float()
        accum = torch.empty(2, dtype=torch.float64, device=logits.device)
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        div = float(n) if self.batch_average else 1.0
        L.check(L.lib().zs3_ce_fwd(L.ptr(logits), L.ptr(target), L.ptr(self.class_weight), n, c, h * w, self.ignore_index,
                                   div, L.ptr(accum), L.ptr(loss), _st()), "zs3_ce_fwd")
        dlogits = torch.empty_like(logits)
        one = torch.ones((), dtype=torch.float32, device=logits.device)
        L.check(L.lib().zs3_ce_bwd(L.ptr(logits), L.ptr(target), L.ptr(self.class_weight), n, c, h * w, self.ignore_index,
                                   div, L.ptr(accum), L.ptr(one), L.ptr(dlogits), _st()), "zs3_ce_bwd")
        keep = logits if return_logits else None
        del logits
        self.backward(tape, dlogits)
        return (loss, keep) if return_logits else loss

    def invalidate(self):
        """call after parameters were changed by anyone else than train_step (the packed weight pieces are cached)"""
        self._epoch += 1

    def zero_grad(self):
        if self.flat is not None:
            self.flat.zero_grad()
        else:
            for p in self.model.parameters():
                if p.grad is not None:
                    p.grad.zero_()

    def train_step(self, image, target):
        """zero_grad -> forward -> CE -> backward -> SGD (base_trainer.py:16-20); returns the device loss"""
        if self.opt is None:
            raise RuntimeError("constructed with optimizer=False")
        self.zero_grad()
        loss = self.loss_and_grads(image, target)
        self.opt.step()
        self._epoch += 1          # packed weight pieces are stale now
        return loss


def _pad_to(t, h, w):
    n, hh, ww, c = t.shape
    if hh == h and ww == w:
        return t
    out = torch.zeros((n, h, w, c), dtype=t.dtype, device=t.device)
    out[:, :hh, :ww] = t
    return out
