"""Thin torch-tensor wrappers over the C ABI (include/zs3b200.h).

Everything here works on CUDA tensors only and launches on torch's current stream.  Activations are
NHWC bf16 tensors of shape [N, H, W, Cs] with Cs (the channel stride) a multiple of 64; channels beyond
the logical channel count are zero.
"""
import ctypes as C

import torch

from . import _lib as L


# Optional per-launch device timing (bench.py's roofline leg).  When PROFILE is a dict, every conv launch is
# bracketed by CUDA events on the launching stream: PROFILE[kind] = [(start_event, end_event, nominal_flops)].
PROFILE = None


class _Timed:
    def __init__(self, kind, flops, tag=""):
        self.kind, self.flops, self.tag = kind, flops, tag

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if PROFILE is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.setdefault(self.kind, []).append((self.e0, e1, self.flops))
            PROFILE.setdefault("_tags", []).append((self.kind, self.tag, self.e0, e1, self.flops))


def cpad(c, to=64):
    return (c + to - 1) // to * to


def _chk_act(t, name):
    if not (t.is_cuda and t.dtype == torch.bfloat16 and t.dim() == 4 and t.is_contiguous()):
        raise ValueError(f"{name}: expected a contiguous CUDA bf16 NHWC tensor, got {t.dtype} {tuple(t.shape)}")


def conv_out_size(h, k, stride, pad, dil):
    return (h + 2 * pad - dil * (k - 1) - 1) // stride + 1


def is_krsc(w):
    """True if the 4-D weight is stored KRSC (torch channels_last), the layout the kernels pack from cheaply."""
    return w.dim() == 4 and w.is_contiguous(memory_format=torch.channels_last)


def pack_weight(w, cout_pad, cin_pad, ci_begin=0, ci_count=None, mode=0):
    """fp32 conv weight [O,I,R,S] (OIHW-contiguous or channels_last) -> packed bf16.
    mode 0: [cout_pad][R*S][cin_pad]; mode 1 (dgrad): [cin_pad][R*S][cout_pad] with spatially flipped taps."""
    cout, cin, r, s = w.shape
    if ci_count is None:
        ci_count = cin - ci_begin
    w = w.detach().float()
    if is_krsc(w) and not (w.is_contiguous() and r * s > 1):
        mode += 2  # KRSC source
    else:
        w = w.contiguous()
    shape = (cout_pad, r * s, cin_pad) if (mode & 1) == 0 else (cin_pad, r * s, cout_pad)
    dst = torch.empty(shape, dtype=torch.bfloat16, device=w.device)
    L.check(L.lib().zs3_pack_weight(L.ptr(w), cout, cin, r, s, ci_begin, ci_count, L.ptr(dst), cout_pad, cin_pad, mode,
                                    L.stream_ptr()), "zs3_pack_weight")
    return dst


def unpack_wgrad(dw, grad_oihw, ci_begin=0, ci_count=None, accumulate=False):
    cout, cin, r, s = grad_oihw.shape
    if ci_count is None:
        ci_count = cin - ci_begin
    cout_pad, taps, cin_pad = dw.shape
    L.check(L.lib().zs3_unpack_wgrad(L.ptr(dw), cout_pad, cin_pad, L.ptr(grad_oihw), cout, cin, r, s, ci_begin,
                                     ci_count, int(accumulate), L.stream_ptr()), "zs3_unpack_wgrad")


def conv_fprop(segments, R, S, stride, pad, dil, cout_pad, out=None, out_f32=False, accumulate=False, bias=None,
               stats=None, scatter=None, flops=0.0, kind="conv_fprop", w_forward_layout=False, epilogue=None):
    # epilogue = (scale, shift, residual or None, relu): eval-mode BatchNorm folded into the conv (inference)
    """segments: list of (x [N,H,W,Cs] bf16, w_packed [cout_pad, R*S, cin_pad] bf16).
    stats: optional (sum, sqsum) fp64 [cout_pad] accumulators.  scatter: optional (sp_stride, y_H, y_W).
    Returns y [N,Ho,Wo,cout_pad] (or the provided `out`)."""
    x0 = segments[0][0]
    _chk_act(x0, "conv_fprop x")
    n, h, w_, _ = x0.shape
    ho = conv_out_size(h, R, stride, pad, dil)
    wo = conv_out_size(w_, S, stride, pad, dil)
    a = L.ConvArgs()
    a.N, a.H, a.W, a.Ho, a.Wo = n, h, w_, ho, wo
    a.R, a.S, a.stride, a.pad, a.dil = R, S, stride, pad, dil
    a.cout_pad = cout_pad
    a.num_segments = len(segments)
    for i, (x, wp) in enumerate(segments):
        _chk_act(x, f"conv_fprop seg{i}")
        if tuple(x.shape[:3]) != (n, h, w_):
            raise ValueError("conv_fprop: segments disagree on N/H/W")
        kdim, ndim = (0, 2) if w_forward_layout else (2, 0)  # forward layout: [k][taps][n]
        if wp.dtype != torch.bfloat16 or wp.shape[ndim] != cout_pad or wp.shape[1] != R * S:
            raise ValueError(f"conv_fprop: packed weight shape {tuple(wp.shape)} does not match")
        a.seg[i].x = x.data_ptr()
        a.seg[i].x_cstride = x.shape[3]
        a.seg[i].w = wp.data_ptr()
        a.seg[i].cin_pad = wp.shape[kdim]
    if out is None:
        if scatter is not None:
            raise ValueError("scatter needs an explicit output tensor")
        out = torch.empty((n, ho, wo, cout_pad), dtype=torch.float32 if out_f32 else torch.bfloat16, device=x0.device)
    a.y = out.data_ptr()
    a.y_cstride = out.shape[-1]
    a.y_is_f32 = int(out.dtype == torch.float32)
    if scatter is not None:
        a.y_sp_stride, a.y_H, a.y_W = scatter
    a.accumulate = int(accumulate)
    a.w_forward_layout = int(w_forward_layout)
    if epilogue is not None:
        sc, sh, res, relu = epilogue
        a.ep_scale, a.ep_shift, a.ep_relu = sc.data_ptr(), sh.data_ptr(), int(relu)
        if res is not None:
            _chk_act(res, "conv_fprop epilogue residual")
            a.ep_residual, a.ep_res_cstride = res.data_ptr(), res.shape[3]
    a.bias = None if bias is None else bias.data_ptr()
    if stats is not None:
        a.stat_sum = stats[0].data_ptr()
        a.stat_sqsum = stats[1].data_ptr()
    tag = ""
    if PROFILE is not None:
        tag = (f"N{n} {h}x{w_}->{ho}x{wo} k{R} s{stride} d{dil} cin{sum(s_.cin_pad for s_ in a.seg[:len(segments)])} "
               f"cout{cout_pad} segs{len(segments)}{' stats' if stats is not None else ''}")
    with _Timed(kind, flops, tag):
        L.check(L.lib().zs3_conv_fprop(C.byref(a), L.stream_ptr()), "zs3_conv_fprop")
    return out


def conv_wgrad(x, dy, R, S, stride, pad, dil, cin_pad, cout_pad, dw=None, k_splits=0, flops=0.0, dw_view=None):
    """dw[cout_pad][R*S][cin_pad] fp32 (+)= sum_p dy[p] (x) x[p@tap].  Returns dw.
    dw_view = (ld, ci_offset, cout_valid, cin_valid): accumulate straight into a KRSC gradient tensor `dw`."""
    _chk_act(x, "conv_wgrad x")
    _chk_act(dy, "conv_wgrad dy")
    n, h, w_, _ = x.shape
    ho, wo = dy.shape[1], dy.shape[2]
    if dw is None:
        dw = torch.zeros((cout_pad, R * S, cin_pad), dtype=torch.float32, device=x.device)
    a = L.WgradArgs()
    a.N, a.H, a.W, a.Ho, a.Wo = n, h, w_, ho, wo
    a.R, a.S, a.stride, a.pad, a.dil = R, S, stride, pad, dil
    a.x = x.data_ptr()
    a.x_cstride = x.shape[3]
    a.cin_pad = cin_pad
    a.dy = dy.data_ptr()
    a.dy_cstride = dy.shape[3]
    a.cout_pad = cout_pad
    a.dw = dw.data_ptr()
    a.k_splits = k_splits
    if dw_view is not None:
        a.dw_ld, a.dw_ci_offset, a.cout_valid, a.cin_valid = dw_view
    tag = f"N{n} {h}x{w_}->{ho}x{wo} k{R} s{stride} d{dil} cin{cin_pad} cout{cout_pad}" if PROFILE is not None else ""
    with _Timed("conv_wgrad", flops, tag):
        L.check(L.lib().zs3_conv_wgrad(C.byref(a), L.stream_ptr()), "zs3_conv_wgrad")
    return dw


def nchw_to_nhwc(x, cs=None):
    """fp32 NCHW -> bf16 NHWC with channel stride cs (zero padded)."""
    n, c, h, w = x.shape
    cs = cpad(c) if cs is None else cs
    x = x.contiguous().float()
    out = torch.empty((n, h, w, cs), dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().zs3_nchw_f32_to_nhwc_bf16(L.ptr(x), L.ptr(out), n, c, h * w, cs, L.stream_ptr()),
            "zs3_nchw_f32_to_nhwc_bf16")
    return out


def nhwc_to_nchw(x, c):
    """bf16 NHWC [N,H,W,Cs] -> fp32 NCHW [N,c,H,W]."""
    _chk_act(x, "nhwc_to_nchw")
    n, h, w, cs = x.shape
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    L.check(L.lib().zs3_nhwc_bf16_to_nchw_f32(L.ptr(x), L.ptr(out), n, c, h * w, cs, L.stream_ptr()),
            "zs3_nhwc_bf16_to_nchw_f32")
    return out


def bn_finalize(stats, count, gamma, beta, eps, momentum, running_mean, running_var, cpad_, reset=True):
    """Returns (scale, shift, mean, invstd), each fp32 [cpad_]."""
    dev = stats[0].device
    c = gamma.numel()
    coef = torch.empty((4, cpad_), dtype=torch.float32, device=dev)
    L.check(L.lib().zs3_bn_finalize(L.ptr(stats[0]), L.ptr(stats[1]), int(count), L.ptr(gamma), L.ptr(beta),
                                    float(eps), float(momentum), L.ptr(running_mean), L.ptr(running_var),
                                    L.ptr(coef[0]), L.ptr(coef[1]), L.ptr(coef[2]), L.ptr(coef[3]), c, cpad_,
                                    int(reset), L.stream_ptr()), "zs3_bn_finalize")
    return coef[0], coef[1], coef[2], coef[3]


def bn_stats(y, stats):
    """stats[0] += per-channel sum of y, stats[1] += per-channel sum of squares (y: NHWC bf16)"""
    _chk_act(y, "bn_stats")
    n, h, w, cs = y.shape
    L.check(L.lib().zs3_bn_stats(L.ptr(y), cs, n * h * w, cs, L.ptr(stats[0]), L.ptr(stats[1]), L.stream_ptr()),
            "zs3_bn_stats")


def bn_eval_coeffs(gamma, beta, running_mean, running_var, eps, cpad_):
    dev = running_mean.device
    c = running_mean.numel()
    coef = torch.empty((4, cpad_), dtype=torch.float32, device=dev)
    L.check(L.lib().zs3_bn_eval_coeffs(L.ptr(gamma), L.ptr(beta), L.ptr(running_mean), L.ptr(running_var), float(eps),
                                       L.ptr(coef[0]), L.ptr(coef[1]), L.ptr(coef[2]), L.ptr(coef[3]), c, cpad_,
                                       L.stream_ptr()), "zs3_bn_eval_coeffs")
    return coef[0], coef[1], coef[2], coef[3]


def bn_apply(y, scale, shift, relu, residual=None, out=None, drop_p=0.0, seed=0, offset=0, keep_mask=None,
             offset_dev=None, finalize=None, relu_mask=None):
    """finalize = dict(stats=(sum, sqsum), count, gamma, beta, eps, momentum, running_mean, running_var, coef [4][C],
    reset=(sum, sqsum) or None): derive the affine from raw batch statistics inside the kernel (fused bn_finalize)."""
    _chk_act(y, "bn_apply y")
    n, h, w, cs = y.shape
    if out is None:
        out = torch.empty_like(y)
    a = L.BnApplyArgs()
    a.y, a.y_cstride = y.data_ptr(), cs
    if residual is not None:
        _chk_act(residual, "bn_apply residual")
        a.residual, a.res_cstride = residual.data_ptr(), residual.shape[3]
    a.out, a.out_cstride = out.data_ptr(), out.shape[3]
    if finalize is not None:
        f = finalize
        a.stat_sum, a.stat_sqsum, a.count = f["stats"][0].data_ptr(), f["stats"][1].data_ptr(), int(f["count"])
        a.gamma = None if f["gamma"] is None else f["gamma"].data_ptr()
        a.beta = None if f["beta"] is None else f["beta"].data_ptr()
        a.eps, a.momentum = float(f["eps"]), float(f["momentum"])
        if f["running_mean"] is not None:
            a.running_mean, a.running_var = f["running_mean"].data_ptr(), f["running_var"].data_ptr()
        a.C_real = int(f["c_real"])
        coef = f["coef"]
        a.scale_out, a.shift_out = coef[0].data_ptr(), coef[1].data_ptr()
        a.mean_out, a.invstd_out = coef[2].data_ptr(), coef[3].data_ptr()
        if f.get("reset") is not None:
            a.reset_sum, a.reset_sqsum = f["reset"][0].data_ptr(), f["reset"][1].data_ptr()
            a.reset_count = int(f["reset"][2])
        a.sync_clamp = int(bool(f.get("sync_clamp", False)))
    else:
        a.scale, a.shift = scale.data_ptr(), shift.data_ptr()
    a.M, a.C = n * h * w, cs
    a.relu = int(relu)
    a.drop_p = float(drop_p)
    if keep_mask is not None:
        a.drop_mode = 2
        a.keep_mask = keep_mask.data_ptr()
    elif drop_p > 0:
        a.drop_mode = 1
    a.seed, a.offset = int(seed), int(offset)
    if offset_dev is not None:
        a.offset_dev = offset_dev.data_ptr()
    if relu_mask is not None:  # uint8 [M * C/8]: one ReLU bit per element for the backward (relu_mask= of bn_backward)
        assert relu_mask.dtype == torch.uint8 and relu_mask.numel() * 8 >= n * h * w * cs
        a.relu_mask_out = relu_mask.data_ptr()
    L.check(L.lib().zs3_bn_apply(C.byref(a), L.stream_ptr()), "zs3_bn_apply")
    return out


def bn_backward(dout, out, y, mean, invstd, scale, relu, grad_scale=1.0, training=True, dres=None,
                dres_accumulate=False, dgamma=None, dbeta=None, param_accumulate=False, scatter=None, dy=None,
                scratch=None, shift=None, sums=None, reset=None, relu_mask=None, sync=None):
    """Two-phase BatchNorm(+ReLU/+Dropout) backward.  Returns dy (bf16, same layout as y unless scatter).
    With `shift` given (plain conv->BN->ReLU layers) the ReLU mask is recomputed from y instead of read from `out`."""
    _chk_act(dout, "bn_backward dout")
    n, h, w, cs = y.shape
    if sums is not None:
        scratch = sums  # (sum_dz, sum_dzx): pre-zeroed; the caller alternates two buffers and passes `reset`
    elif scratch is None:
        scratch = torch.zeros((2, cs), dtype=torch.float64, device=y.device)
    else:
        scratch.zero_()
    a = L.BnBwdArgs()
    a.dout, a.dout_cstride = dout.data_ptr(), dout.shape[3]
    relu_mode = 0
    if relu and shift is not None:
        relu_mode = 2
        a.shift = shift.data_ptr()
    elif relu and relu_mask is not None:
        relu_mode = 3  # bit mask written by bn_apply(relu_mask=...): 1/16 of the bytes of re-reading `out`
        a.relu_mask = relu_mask.data_ptr()
    elif relu:
        relu_mode = 1
        a.out, a.out_cstride = out.data_ptr(), out.shape[3]
    a.y, a.y_cstride = y.data_ptr(), cs
    a.mean, a.invstd, a.scale = mean.data_ptr(), invstd.data_ptr(), scale.data_ptr()
    a.M, a.C = n * h * w, cs
    a.relu, a.grad_scale, a.training = relu_mode, float(grad_scale), int(training)
    a.sum_dz, a.sum_dzx = scratch[0].data_ptr(), scratch[1].data_ptr()
    if scatter is not None:
        sp, hy, wy = scatter
        if dy is None:
            dy = torch.zeros((n, hy, wy, cs), dtype=torch.bfloat16, device=y.device)
        a.dy_sp_stride, a.sp_Ho, a.sp_Wo, a.dy_H, a.dy_W = sp, h, w, hy, wy
    elif dy is None:
        dy = torch.empty_like(y)
    a.dy, a.dy_cstride = dy.data_ptr(), dy.shape[3]
    if dres is not None:
        a.dres, a.dres_cstride, a.dres_accumulate = dres.data_ptr(), dres.shape[3], int(dres_accumulate)
    if dgamma is not None:
        a.dgamma, a.dbeta, a.C_real = dgamma.data_ptr(), dbeta.data_ptr(), dgamma.numel()
        a.param_accumulate = int(param_accumulate)
    if reset is not None:
        a.reset_sum_dz, a.reset_sum_dzx, a.reset_count = reset[0].data_ptr(), reset[1].data_ptr(), int(reset[2])
    st = L.stream_ptr()
    L.check(L.lib().zs3_bn_bwd_reduce(C.byref(a), st), "zs3_bn_bwd_reduce")
    if sync is not None:
        # Synchronised BatchNorm: sync = (all_reduce, buffer holding sum_dz / sum_dzx, world size).  dy needs the sums
        # over ALL ranks (divided by the global count); dgamma / dbeta keep the rank-local sums (they join the step's
        # gradient all-reduce), so they are taken from the buffer before it is reduced.
        all_reduce, buf, world = sync
        if dgamma is not None:
            c_real = dgamma.numel()
            if param_accumulate:
                dgamma.add_(scratch[1][:c_real].float())
                dbeta.add_(scratch[0][:c_real].float())
            else:
                dgamma.copy_(scratch[1][:c_real])
                dbeta.copy_(scratch[0][:c_real])
            a.dgamma, a.dbeta = None, None
        all_reduce(buf)
        a.stat_count = n * h * w * world
    L.check(L.lib().zs3_bn_bwd_apply(C.byref(a), st), "zs3_bn_bwd_apply")
    return dy


def im2col_probe(x, pad, upper, stride, cpp, ppc, c, w, h, n, off_w, off_h):
    """Raw (swizzled) shared-memory image of one im2col TMA load: uint8 [ppc*cpp*2]."""
    _chk_act(x, "im2col_probe x")
    nn, hh, ww, cc = x.shape
    out = torch.empty(ppc * cpp * 2, dtype=torch.uint8, device=x.device)
    L.check(L.lib().zs3_debug_im2col_probe(L.ptr(x), nn, hh, ww, cc, pad, upper, stride, cpp, ppc, c, w, h, n, off_w,
                                           off_h, L.ptr(out), L.stream_ptr()), "zs3_debug_im2col_probe")
    return out


def stem_im2col(x, R, stride, pad, ho, wo, kpad, krsc=False):
    """fp32 NCHW image -> bf16 im2col matrix viewed as NHWC [N, ho, wo, kpad]
    (k = c*R*R + r*R + s, or (r*R + s)*C + c with krsc=True)."""
    n, c, h, w = x.shape
    x = x.contiguous().float()
    cols = torch.empty((n, ho, wo, kpad), dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().zs3_stem_im2col(L.ptr(x), L.ptr(cols), n, c, h, w, R, stride, pad, ho, wo, kpad, int(krsc),
                                    L.stream_ptr()), "zs3_stem_im2col")
    return cols


def maxpool_fwd(x, k, stride, pad):
    _chk_act(x, "maxpool_fwd")
    n, h, w, c = x.shape
    ho, wo = conv_out_size(h, k, stride, pad, 1), conv_out_size(w, k, stride, pad, 1)
    y = torch.empty((n, ho, wo, c), dtype=torch.bfloat16, device=x.device)
    arg = torch.empty((n, ho, wo, c), dtype=torch.uint8, device=x.device)
    L.check(L.lib().zs3_maxpool_fwd(L.ptr(x), L.ptr(y), L.ptr(arg), n, h, w, c, ho, wo, k, stride, pad,
                                    L.stream_ptr()), "zs3_maxpool_fwd")
    return y, arg


def maxpool_bwd(dy, arg, in_shape, k, stride, pad):
    n, h, w, c = in_shape
    dx = torch.empty(in_shape, dtype=torch.bfloat16, device=dy.device)
    L.check(L.lib().zs3_maxpool_bwd(L.ptr(dy), L.ptr(arg), L.ptr(dx), n, h, w, c, dy.shape[1], dy.shape[2], k, stride,
                                    pad, L.stream_ptr()), "zs3_maxpool_bwd")
    return dx


def bilinear_fwd(x, ho, wo):
    _chk_act(x, "bilinear_fwd")
    n, hi, wi, c = x.shape
    y = torch.empty((n, ho, wo, c), dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().zs3_bilinear_fwd(L.ptr(x), L.ptr(y), n, hi, wi, ho, wo, c, c, c, L.stream_ptr()),
            "zs3_bilinear_fwd")
    return y


def bilinear_bwd(dy, hi, wi):
    _chk_act(dy, "bilinear_bwd")
    n, ho, wo, c = dy.shape
    dx = torch.empty((n, hi, wi, c), dtype=torch.bfloat16, device=dy.device)
    L.check(L.lib().zs3_bilinear_bwd(L.ptr(dy), L.ptr(dx), n, hi, wi, ho, wo, c, c, c, 0, L.stream_ptr()),
            "zs3_bilinear_bwd")
    return dx


def upsample_logits_fwd(x, c, ho, wo):
    _chk_act(x, "upsample_logits_fwd")
    n, hi, wi, cs = x.shape
    y = torch.empty((n, c, ho, wo), dtype=torch.float32, device=x.device)
    L.check(L.lib().zs3_upsample_logits_fwd(L.ptr(x), L.ptr(y), n, c, hi, wi, cs, ho, wo, L.stream_ptr()),
            "zs3_upsample_logits_fwd")
    return y


def upsample_logits_bwd(dy, in_shape, c):
    n, hi, wi, cs = in_shape
    dx = torch.empty(in_shape, dtype=torch.bfloat16, device=dy.device)
    L.check(L.lib().zs3_upsample_logits_bwd(L.ptr(dy), L.ptr(dx), n, c, hi, wi, cs, dy.shape[2], dy.shape[3],
                                            L.stream_ptr()), "zs3_upsample_logits_bwd")
    return dx


def spatial_sum(x, scale):
    """[N,H,W,C] -> [N,1,1,C], y = scale * sum over H,W."""
    _chk_act(x, "spatial_sum")
    n, h, w, c = x.shape
    y = torch.empty((n, 1, 1, c), dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().zs3_spatial_sum(L.ptr(x), L.ptr(y), n, h * w, c, c, c, float(scale), L.stream_ptr()),
            "zs3_spatial_sum")
    return y


def spatial_broadcast(x, h, w, scale, out=None, accumulate=False):
    """[N,1,1,C] -> [N,h,w,C], y = scale * x  (accumulate: out += scale * x)."""
    n, _, _, c = x.shape
    y = out if out is not None else torch.empty((n, h, w, c), dtype=torch.bfloat16, device=x.device)
    _chk_act(y, "spatial_broadcast out")
    L.check(L.lib().zs3_spatial_broadcast(L.ptr(x), L.ptr(y), n, h * w, c, x.shape[3], y.shape[3], float(scale),
                                          int(bool(accumulate)), L.stream_ptr()), "zs3_spatial_broadcast")
    return y


def channel_sums(t):
    """fp64 per-channel sums over all pixels of an NHWC bf16 tensor (bias gradients)."""
    _chk_act(t, "channel_sums")
    n, h, w, cs = t.shape
    scratch = torch.zeros((2, cs), dtype=torch.float64, device=t.device)
    ones = torch.ones(cs, dtype=torch.float32, device=t.device)
    zeros = torch.zeros(cs, dtype=torch.float32, device=t.device)
    a = L.BnBwdArgs()
    a.dout, a.dout_cstride = t.data_ptr(), cs
    a.y, a.y_cstride = t.data_ptr(), cs
    a.mean, a.invstd, a.scale = zeros.data_ptr(), ones.data_ptr(), ones.data_ptr()
    a.M, a.C = n * h * w, cs
    a.relu, a.grad_scale, a.training = 0, 1.0, 1
    a.sum_dz, a.sum_dzx = scratch[0].data_ptr(), scratch[1].data_ptr()
    L.check(L.lib().zs3_bn_bwd_reduce(C.byref(a), L.stream_ptr()), "zs3_bn_bwd_reduce")
    return scratch[0]


def cast_f32_to_bf16(src, dst):
    L.check(L.lib().zs3_cast_f32_to_bf16(L.ptr(src), L.ptr(dst), src.numel(), L.stream_ptr()), "zs3_cast_f32_to_bf16")


def sgd_step(p, g, buf, lr, momentum, weight_decay, nesterov, first_step, grad_scale=1.0):
    """lr: a Python float, or a 1-element fp32 CUDA tensor read by the kernel at run time (CUDA-graph safe)"""
    if isinstance(lr, torch.Tensor):
        L.check(L.lib().zs3_sgd_step_lrdev(L.ptr(p), L.ptr(g), L.ptr(buf), p.numel(), L.ptr(lr), float(momentum),
                                           float(weight_decay), int(nesterov), int(first_step), float(grad_scale),
                                           L.stream_ptr()), "zs3_sgd_step_lrdev")
        return
    L.check(L.lib().zs3_sgd_step(L.ptr(p), L.ptr(g), L.ptr(buf), p.numel(), float(lr), float(momentum),
                                 float(weight_decay), int(nesterov), int(first_step), float(grad_scale),
                                 L.stream_ptr()), "zs3_sgd_step")


def adam_step(p, g, m, v, lr, beta1, beta2, eps, step, grad_scale=1.0):
    L.check(L.lib().zs3_adam_step(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), float(lr), float(beta1),
                                  float(beta2), float(eps), int(step), float(grad_scale), L.stream_ptr()),
            "zs3_adam_step")
