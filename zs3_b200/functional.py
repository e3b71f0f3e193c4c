"""autograd Functions that stitch the sm_100a kernels (zs3_b200/kernels.py) into the DeepLab graph.

Internal activations are NHWC bf16 tensors [N, H, W, Cs] (Cs = channels padded to a multiple of 64).
Parameters stay fp32 tensors in the reference's layout (OIHW conv weights, [C] BatchNorm affine), so
state_dict()s are interchangeable with the reference's; the Functions return parameter gradients in
that layout.
"""
import ctypes as C

import os

import torch

from . import _lib as L
from . import kernels as K
from .modeling.sync_batchnorm import batchnorm as SBN

_SCRATCH = {}
PENDING_BATCH_COUNTERS = []  # num_batches_tracked buffers to bump (one fused foreach add per model forward)


def flush_batch_counters():
    if PENDING_BATCH_COUNTERS:
        torch._foreach_add_(PENDING_BATCH_COUNTERS, 1)
        PENDING_BATCH_COUNTERS.clear()


class UpsampleCrossEntropy(torch.autograd.Function):
    """CrossEntropy(F.interpolate(x, size, bilinear, align_corners=True), target) computed from the LOW-resolution
    NHWC bf16 class scores x (deeplab.py:44 + loss.py:31-46 in one pass each way): the [N,C,H,W] fp32 logits and
    their gradient never touch HBM.  Used by the training runtime; the module API still returns real logits."""

    @staticmethod
    def forward(ctx, x, num_classes, target, weight, ignore_index, div):
        if not (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4):
            raise ValueError("UpsampleCrossEntropy: x must be a CUDA NHWC bf16 tensor")
        target = target.contiguous().float()
        n, hi, wi, cs = x.shape
        ho, wo = int(target.shape[-2]), int(target.shape[-1])
        if target.numel() != n * ho * wo:
            raise ValueError("UpsampleCrossEntropy: target must be [N, H, W]")
        x = x.contiguous()
        accum = torch.empty(2, dtype=torch.float64, device=x.device)
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        wt = None if weight is None else weight.contiguous().float()
        L.check(L.lib().zs3_upsample_ce_fwd(L.ptr(x), L.ptr(target), L.ptr(wt), n, int(num_classes), hi, wi, cs, ho, wo,
                                            int(ignore_index), float(div), L.ptr(accum), L.ptr(loss), L.stream_ptr()),
                "zs3_upsample_ce_fwd")
        ctx.save_for_backward(x, target, accum)
        ctx.info = (int(num_classes), wt, int(ignore_index), float(div), ho, wo)
        return loss

    @staticmethod
    def backward(ctx, gout):
        x, target, accum = ctx.saved_tensors
        c, wt, ignore, div, ho, wo = ctx.info
        n, hi, wi, cs = x.shape
        dx = torch.empty_like(x)
        gout = gout.contiguous().float()
        if (ho == 4 * (hi - 1) + 1 and wo == 4 * (wi - 1) + 1 and wo <= 544 and hi >= 2 and wi >= 2
                and 7 * c * (wi + 1) * 4 <= 224 * 1024 and os.environ.get("ZS3_CE_BWD_X4", "1") == "1"):
            # DeepLab's exact x4 geometry: one softmax per output pixel, deterministic two-partial combine
            ws = torch.empty(L.lib().zs3_upsample4_ce_bwd_workspace_size(n, c, hi, wi), dtype=torch.uint8, device=x.device)
            L.check(L.lib().zs3_upsample4_ce_bwd(L.ptr(x), L.ptr(target), L.ptr(wt), n, c, hi, wi, cs, ho, wo, ignore, div,
                                                 L.ptr(accum), L.ptr(gout), L.ptr(dx), L.ptr(ws), ws.numel(),
                                                 L.stream_ptr()), "zs3_upsample4_ce_bwd")
            return dx, None, None, None, None, None
        L.check(L.lib().zs3_upsample_ce_bwd(L.ptr(x), L.ptr(target), L.ptr(wt), n, c, hi, wi, cs, ho, wo, ignore, div,
                                            L.ptr(accum), L.ptr(gout), L.ptr(dx), L.stream_ptr()), "zs3_upsample_ce_bwd")
        return dx, None, None, None, None, None


class IdentityBN:
    """stand-in for 'no normalisation' (aspp.global_avg_pool without BN, zs3/modeling/aspp.py:90-95)"""
    training = False
    weight = None
    bias = None
    eps = 0.0
    momentum = 0.1
    num_batches_tracked = None

    def __init__(self, c, device):
        self.running_mean = torch.zeros(c, dtype=torch.float32, device=device)
        self.running_var = torch.ones(c, dtype=torch.float32, device=device)



class _StatBuffers:
    """Two alternating fp64 (sum, sqsum) accumulators per device for the training-mode BatchNorm statistics.
    Layer L accumulates into buffer `cur` (conv epilogue or bn_stats), its bn_apply derives the affine from it
    (fused finalize) and zeroes the OTHER buffer, which layer L+1 then uses: no separate finalize/reset launches."""
    per_device = {}

    def __init__(self, dev):
        self.buf = [torch.zeros(2, 2048, dtype=torch.float64, device=dev) for _ in range(2)]
        self.cur = 0
        self.dirty = [0, 0]

    @classmethod
    def get(cls, dev):
        key = str(dev)
        if key not in cls.per_device:
            cls.per_device[key] = cls(dev)
        return cls.per_device[key]

    def current(self, c):
        self.dirty[self.cur] = max(self.dirty[self.cur], c)
        b = self.buf[self.cur]
        return b[0, :c], b[1, :c]

    def current_buffer(self):
        """the whole [2, 2048] fp64 buffer `current()` hands out views of (one contiguous 32 KB all-reduce message)"""
        return self.buf[self.cur]

    def flip(self):
        """returns (sum, sqsum, count) of the other buffer to be zeroed by this layer's bn_apply, then switches"""
        o = 1 - self.cur
        res = (self.buf[o][0], self.buf[o][1], self.dirty[o])
        self.dirty[o] = 0
        self.cur = o
        return res

    def reset(self):
        """both buffers zero and a fixed starting side: called once per model forward so that a captured CUDA graph
        replays a self-consistent sequence"""
        if self.dirty[0] or self.dirty[1] or self.cur != 0:
            for b in self.buf:
                b.zero_()
        self.cur, self.dirty = 0, [0, 0]


class _BwdSumBuffers(_StatBuffers):
    """same alternation for the BatchNorm-backward sums (sum dz, sum dz*xhat)"""
    per_device = {}


def reset_stat_buffers(dev):
    _StatBuffers.get(dev).reset()
    _BwdSumBuffers.get(dev).reset()


def _scratch64(dev, tag="fwd", n=2 * 2048):
    """fp64 scratch: "fwd" holds the conv-epilogue BN statistics (kept zero between uses by bn_finalize's
    reset), "bwd" the BN-backward sums (zeroed before each use)."""
    key = (dev, tag, n)
    t = _SCRATCH.get(key)
    if t is None:
        t = torch.zeros(n, dtype=torch.float64, device=dev)
        _SCRATCH[key] = t
    return t


class _RngState:
    """Counter-based dropout RNG: (seed, running offset) on the host, in the spirit of torch's Philox state.
    `device_counter` (an int64 CUDA tensor, optional) is added on the device so that a captured CUDA graph
    draws fresh masks on every replay (the trainer bumps it once per step)."""
    offset = 0
    device_counter = None

    @classmethod
    def next(cls, numel):
        # the data-parallel rank is mixed into the seed: replicas must not draw identical Dropout masks
        rank = int(os.environ.get("RANK", "0"))
        seed = (torch.initial_seed() + 0x9E3779B97F4A7C15 * rank) & 0xFFFFFFFFFFFFFFFF
        off = cls.offset
        cls.offset += numel
        return seed, off


# BatchNorm statistics are taken in the conv epilogue (from the staged bf16 tile) for every layer with at least this
# many 64-wide k-blocks per tile; 0 = always.  (The first version, a shuffle-tree reduction, only paid off >= 18.)
FUSE_STATS_MIN_KB = int(os.environ.get("ZS3_FUSE_STATS_MIN_KB", "0"))

_WEIGHT_EPOCH = [0]


def invalidate_weight_caches():
    """Call after parameters were modified behind autograd's back (raw-pointer optimizer kernels, NCCL
    broadcasts into the flat buffer): torch's version counters do not see those writes."""
    _WEIGHT_EPOCH[0] += 1


def _packed_weight(conv, mode, ci_begin, ci_count, cin_pad, cout_pad):
    """bf16 kernel-layout copy of conv.weight, cached until the parameter changes (version counter)."""
    w = conv.weight
    shadow = conv.__dict__.get("_zs3_bf16_shadow")
    if (shadow is not None and mode == 0 and ci_begin == 0 and ci_count == w.shape[1] == cin_pad
            and cout_pad == w.shape[0] and shadow[1] == w.data_ptr()):
        # parameter lives in a trainer's flat KRSC buffer whose bf16 shadow is refreshed once per step
        return shadow[0]
    cache = conv.__dict__.setdefault("_zs3_pack_cache", {})
    key = (mode, ci_begin, ci_count, cin_pad, cout_pad)
    ent = cache.get(key)
    ver = (w._version, w.data_ptr(), _WEIGHT_EPOCH[0])
    if ent is None or ent[0] != ver:
        ent = (ver, K.pack_weight(w.detach(), cout_pad, cin_pad, ci_begin, ci_count, mode))
        cache[key] = ent
    return ent[1]


def to_krsc_(module):
    """Store every conv weight of `module` in KRSC order (torch channels_last): shapes, names and values are
    unchanged (state_dict compatible), but the kernels' packed bf16 copy becomes a plain cast and the weight
    gradient kernel can accumulate straight into `.grad`."""
    for m in module.modules():
        if isinstance(m, torch.nn.Conv2d):
            m.weight.data = m.weight.data.contiguous(memory_format=torch.channels_last)
            if m.weight.grad is not None:
                m.weight.grad = m.weight.grad.contiguous(memory_format=torch.channels_last)
    return module


def _direct_grad_buffer(w):
    """fp32 KRSC gradient tensor of parameter `w` to accumulate into in place, or None if the layout does not allow it"""
    if not (K.is_krsc(w) and w.dtype == torch.float32):
        return None
    if w.grad is None:
        w.grad = torch.zeros_like(w)  # preserve_format keeps the KRSC strides
    g = w.grad
    if K.is_krsc(g) and g.dtype == torch.float32 and g.stride() == w.stride():
        return g
    return None


def _direct_small_grad(p):
    """1-D fp32 parameter gradient buffer to accumulate into in place (created on first use)"""
    if p is None:
        return None
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    g = p.grad
    return g if (g.dtype == torch.float32 and g.is_contiguous()) else None


def _bn_forward_coeffs(bn, stats, count, cs):
    """(scale, shift, mean, invstd) for this BN: batch statistics in training mode (updates the running
    buffers like F.batch_norm), running statistics otherwise."""
    w = bn.weight.detach() if bn.weight is not None else None
    b = bn.bias.detach() if bn.bias is not None else None
    if stats is not None:
        mom = bn.momentum if bn.momentum is not None else 0.1
        if bn.num_batches_tracked is not None:
            PENDING_BATCH_COUNTERS.append(bn.num_batches_tracked)
        return K.bn_finalize(stats, count, w, b, bn.eps, mom, bn.running_mean, bn.running_var, cs)
    return K.bn_eval_coeffs(w, b, bn.running_mean, bn.running_var, bn.eps, cs)


class _Saved:
    """what a conv->BN->act forward keeps for its backward"""
    __slots__ = ("conv", "bn", "relu", "p", "training", "ranges", "has_res", "geom", "flops_per_cin", "mask_from_y",
                 "y", "out", "mean", "invstd", "scale", "shift", "xs", "relu_bits", "sync_w")


def cba_forward(conv, bn, relu, drop_p, drop_training, keep_mask, residual, xs, channels, need_backward=True):
    """[concat ->] conv -> BatchNorm -> (+residual) -> (ReLU) -> (Dropout).  Returns (out, saved).
    need_backward: whether any input of the enclosing autograd node requires grad (torch.is_grad_enabled() is always
    False inside Function.forward, so the node passes any(ctx.needs_input_grad))."""
    R, S = conv.kernel_size
    stride, pad, dil = conv.stride[0], conv.padding[0], conv.dilation[0]
    cout = conv.out_channels
    cout_p = K.cpad(cout)
    training = bn.training
    if len(channels) != len(xs):
        raise ValueError("conv_bn_act: segment count mismatch")
    # input channel ranges of the (virtually concatenated) segments
    segs, ranges, ci = [], [], 0
    for x, c_real in zip(xs, channels):
        cin_p = x.shape[3]
        segs.append((x, _packed_weight(conv, 0, ci, c_real, cin_p, cout_p)))
        ranges.append((ci, c_real, cin_p))
        ci += c_real
    if ci != conv.in_channels:
        raise ValueError(f"segments provide {ci} channels, conv expects {conv.in_channels}")
    dev = xs[0].device
    n_, h_, w_, _ = xs[0].shape
    if not training and not need_backward and not (drop_p and drop_training):
        # inference: eval-mode BatchNorm, residual add and ReLU are folded into the conv epilogue -- the layer is
        # ONE kernel and neither the pre-BN tensor nor any statistic is ever written
        ho_, wo_ = K.conv_out_size(h_, R, stride, pad, dil), K.conv_out_size(w_, S, stride, pad, dil)
        scale, shift, _, _ = _bn_forward_coeffs(bn, None, 0, cout_p)
        out = K.conv_fprop(segs, R, S, stride, pad, dil, cout_p, epilogue=(scale, shift, residual, relu),
                           flops=2.0 * n_ * ho_ * wo_ * cout * R * S * conv.in_channels)
        return out, None
    stats = None
    if training:
        sb = _StatBuffers.get(dev)
        stats = sb.current(cout_p)
    ho_, wo_ = K.conv_out_size(h_, R, stride, pad, dil), K.conv_out_size(w_, S, stride, pad, dil)
    macs_per_cin = 2.0 * n_ * ho_ * wo_ * cout * R * S  # nominal FLOPs per input channel
    # BatchNorm statistics: fused into the conv epilogue when the main loop is long enough to hide it
    # (>= FUSE_STATS_MIN_KB k-blocks of 64 per tile), otherwise one extra streaming pass over y
    kblocks = R * S * sum(x.shape[3] for x in xs) // 64
    fuse_stats = stats is not None and kblocks >= FUSE_STATS_MIN_KB
    y = K.conv_fprop(segs, R, S, stride, pad, dil, cout_p, stats=stats if fuse_stats else None,
                     flops=macs_per_cin * conv.in_channels)
    if stats is not None and not fuse_stats:
        K.bn_stats(y, stats)
    n, ho, wo, _ = y.shape
    sync_w = SBN.sync_world(bn) if training else 1
    if sync_w > 1:
        SBN.all_reduce_stats(sb.current_buffer())   # (sum, sum of squares) over all ranks, before the fused finalize
    seed = off = 0
    p = float(drop_p) if (drop_p and drop_training) else 0.0
    if p > 0 and keep_mask is None:
        seed, off = _RngState.next(y.numel())
    fin = None
    if training:
        # fused finalize: bn_apply derives scale/shift from the raw sums, publishes the coefficients for the
        # backward, updates the running statistics and zeroes the other statistics buffer
        coef = torch.empty((4, cout_p), dtype=torch.float32, device=dev)
        scale, shift, mean, invstd = coef[0], coef[1], coef[2], coef[3]
        if bn.num_batches_tracked is not None:
            PENDING_BATCH_COUNTERS.append(bn.num_batches_tracked)
        fin = dict(stats=stats, count=n * ho * wo * sync_w, gamma=bn.weight.detach() if bn.weight is not None else None,
                   beta=bn.bias.detach() if bn.bias is not None else None, eps=bn.eps,
                   momentum=bn.momentum if bn.momentum is not None else 0.1, running_mean=bn.running_mean,
                   running_var=bn.running_var, coef=coef, c_real=cout, reset=sb.flip(), sync_clamp=sync_w > 1)
    else:
        scale, shift, mean, invstd = _bn_forward_coeffs(bn, None, n * ho * wo, cout_p)
    # residual joins (out = relu(bn(y) + x)): the backward needs the ReLU mask twice; keep it as one bit per element
    # instead of re-reading the bf16 output in both backward passes
    relu_bits = None
    if relu and residual is not None and p == 0.0 and need_backward:
        relu_bits = torch.empty(y.numel() // 8, dtype=torch.uint8, device=dev)
    out = K.bn_apply(y, scale, shift, relu, residual=residual, drop_p=p, seed=seed, offset=off,
                     keep_mask=keep_mask if p > 0 else None,
                     offset_dev=_RngState.device_counter if (p > 0 and keep_mask is None) else None, finalize=fin,
                     relu_mask=relu_bits)
    sv = _Saved()
    sv.sync_w = sync_w
    sv.relu_bits = relu_bits
    sv.conv, sv.bn, sv.relu, sv.p, sv.training = conv, bn, relu, p, training
    sv.ranges, sv.has_res = ranges, residual is not None
    sv.geom = (R, S, stride, pad, dil, cout, cout_p)
    sv.flops_per_cin = macs_per_cin
    # plain conv->BN->ReLU: the backward recomputes the ReLU mask from y, `out` need not be re-read
    sv.mask_from_y = bool(relu) and residual is None and p == 0.0
    sv.y, sv.out, sv.mean, sv.invstd, sv.scale, sv.shift, sv.xs = y, out, mean, invstd, scale, shift, list(xs)
    return out, sv


_DEBUG_FINITE = os.environ.get("ZS3_DEBUG_FINITE", "0") == "1"

# Weight-gradient kernels on a side stream.  In the backward of a layer the data gradient is on the critical path (the
# next layer's BatchNorm backward needs it), the weight gradient is not: it only has to be done before the optimizer.
# The BatchNorm-backward kernels that follow are HBM-bound, use no shared memory and leave the tensor cores idle, so a
# weight-gradient kernel (tensor-bound, one smem-heavy CTA per SM) can run under them.  ZS3_WGRAD_STREAM=1 forks every
# in-place weight-gradient launch to one side stream (ordered after everything enqueued so far) and joins it at the end
# of the backward pass (autograd engine callback), also inside CUDA-graph capture.
_WGRAD_SIDE = {"on": os.environ.get("ZS3_WGRAD_STREAM", "0") == "1", "stream": None, "pending": False}


def join_wgrad_stream():
    """make the current stream wait for the weight-gradient kernels forked by _wgrad_async (no-op when none pending)"""
    st = _WGRAD_SIDE
    if st["pending"]:
        torch.cuda.current_stream().wait_stream(st["stream"])
        st["pending"] = False


def _wgrad_async(tensors, fn):
    """run fn() -- launches of weight-gradient kernels that READ `tensors` and reduce into persistent .grad buffers --
    on the side stream when enabled, else inline"""
    st = _WGRAD_SIDE
    if not st["on"] or not tensors[0].is_cuda:
        fn()
        return
    if st["stream"] is None:
        st["stream"] = torch.cuda.Stream()
    side = st["stream"]
    side.wait_stream(torch.cuda.current_stream())      # the operands (and the data gradient launched before) are enqueued
    with torch.cuda.stream(side):
        fn()
    for t in tensors:
        t.record_stream(side)                          # the allocator must not hand these blocks out before the kernel ran
    if not st["pending"]:
        st["pending"] = True
        try:
            torch.autograd.Variable._execution_engine.queue_callback(join_wgrad_stream)
        except RuntimeError:                            # not inside a backward pass: the caller joins explicitly
            pass


def _check_finite(what, t, sv):
    """debug aid (ZS3_DEBUG_FINITE=1): name the first tensor of a backward that holds a non-finite value"""
    if _DEBUG_FINITE and t is not None and not bool(torch.isfinite(t.float()).all()):
        bad = (~torch.isfinite(t.float())).nonzero()
        raise FloatingPointError(f"{what} of {sv.conv} holds {bad.shape[0]} non-finite values, first at {bad[0].tolist()}, "
                                 f"shape {tuple(t.shape)}")


def cba_backward(sv, dout, need_dx, need_w=True, need_affine=True, dx_into=None):
    """Backward of cba_forward.  need_dx: per-segment flags.  dx_into: optional per-segment (tensor, accumulate)
    targets the data gradient is written (or added) into -- how a residual join is summed without an extra kernel.
    Returns (dxs, dweight, dgamma, dbeta, dres); parameter gradients that were accumulated in place come back None."""
    conv, bn = sv.conv, sv.bn
    R, S, stride, pad, dil, cout, cout_p = sv.geom
    y, xs = sv.y, sv.xs
    dout = dout.contiguous()
    dev = dout.device
    dgamma = dbeta = None
    direct_affine = False
    if bn.weight is not None and need_affine:
        dgamma, dbeta = _direct_small_grad(bn.weight), _direct_small_grad(bn.bias)
        direct_affine = dgamma is not None and dbeta is not None
        if not direct_affine:
            dgamma = torch.empty(cout, dtype=torch.float32, device=dev)
            dbeta = torch.empty(cout, dtype=torch.float32, device=dev)
    dres = torch.empty_like(dout) if sv.has_res else None
    bb = _BwdSumBuffers.get(dev)
    sums = bb.current(cout_p)
    sync = (SBN.all_reduce_stats, bb.current_buffer(), sv.sync_w) if (sv.training and sv.sync_w > 1) else None
    # strided 3x3: write dy zero-inserted so that the data gradient is a stride-1 conv (see conv_igemm.cu)
    zero_insert = stride > 1 and R > 1
    n, ho, wo, _ = y.shape
    h_in, w_in = xs[0].shape[1], xs[0].shape[2]
    scatter = None
    if zero_insert:
        scatter = (stride, (ho - 1) * stride + 1, (wo - 1) * stride + 1)
    dy_dense_needed = need_w and zero_insert
    dy = K.bn_backward(dout, sv.out, y, sv.mean, sv.invstd, sv.scale, sv.relu, grad_scale=1.0 / (1.0 - sv.p),
                       training=sv.training, dres=dres, dgamma=dgamma, dbeta=dbeta, sums=sums, reset=bb.flip(),
                       param_accumulate=direct_affine, scatter=None if dy_dense_needed else scatter,
                       shift=sv.shift if sv.mask_from_y else None, relu_mask=getattr(sv, "relu_bits", None), sync=sync)
    if direct_affine:
        dgamma = dbeta = None  # already added to bn.weight.grad / bn.bias.grad
    if _DEBUG_FINITE:
        for nm, t in (("dout", dout), ("saved y", y), ("saved out", sv.out), ("mean", sv.mean), ("invstd", sv.invstd),
                      ("scale", sv.scale), ("bn-backward sums", sums[0]), ("dy", dy), ("dres", dres)):
            _check_finite(nm, t, sv)
    dy_z = dy
    if dy_dense_needed:
        # both layouts are needed: dense for wgrad, zero-inserted for dgrad
        dy_z = torch.zeros((n, scatter[1], scatter[2], cout_p), dtype=torch.bfloat16, device=dev)
        dy_z[:, ::stride, ::stride] = dy
    dxs = []
    # weight gradient: accumulated in place into conv.weight.grad when the parameter is stored KRSC (then
    # nothing is returned to autograd for it), else produced as a fresh OIHW tensor
    gbuf = _direct_grad_buffer(conv.weight) if need_w else None
    dweight = torch.empty_like(conv.weight) if (need_w and gbuf is None) else None
    for i, x in enumerate(xs):
        ci0, c_real, cin_p = sv.ranges[i]
        if need_dx[i]:
            # the data gradient reads the FORWARD-packed weights (MN-major B tiles, taps flipped in the kernel):
            # no transposed weight copy is ever materialised
            wt = _packed_weight(conv, 0, ci0, c_real, cin_p, cout_p)  # [cout_p][taps][cin_p]
            fl = sv.flops_per_cin * c_real
            tgt, acc = (dx_into[i] if dx_into is not None and dx_into[i] is not None else (None, False))
            if stride == 1:
                dx = K.conv_fprop([(dy, wt)], R, S, 1, dil * (R - 1) - pad, dil, cin_p, out=tgt, accumulate=acc, flops=fl,
                                  kind="conv_dgrad", w_forward_layout=True)
            elif R == 1:
                dx = tgt if tgt is not None else torch.zeros_like(x)
                K.conv_fprop([(dy, wt)], 1, 1, 1, 0, 1, cin_p, out=dx, scatter=(stride, h_in, w_in), accumulate=acc,
                             flops=fl, kind="conv_dgrad", w_forward_layout=True)
            else:
                # zero-inserted dy has (ho-1)*s+1 rows; pad so that the output covers the full input extent
                dx = K.conv_fprop([(_pad_to(dy_z, h_in + 2 * pad - dil * (R - 1), w_in + 2 * pad - dil * (S - 1)), wt)],
                                  R, S, 1, dil * (R - 1) - pad, dil, cin_p, out=tgt, accumulate=acc, flops=fl,
                                  kind="conv_dgrad", w_forward_layout=True)
            _check_finite(f"dx[{i}] (accumulate={acc})", dx, sv)
            dxs.append(dx)
        else:
            dxs.append(None)
        if need_w and gbuf is not None:
            _wgrad_async([x, dy], lambda x=x, ci0=ci0, c_real=c_real, cin_p=cin_p: K.conv_wgrad(
                x, dy, R, S, stride, pad, dil, cin_p, cout_p, dw=gbuf, flops=sv.flops_per_cin * c_real,
                dw_view=(conv.in_channels, ci0, cout, c_real)))
        elif need_w:
            dw = K.conv_wgrad(x, dy, R, S, stride, pad, dil, cin_p, cout_p, flops=sv.flops_per_cin * c_real)
            K.unpack_wgrad(dw, dweight, ci0, c_real, accumulate=False)
    return dxs, dweight, dgamma, dbeta, dres


class ConvBnAct(torch.autograd.Function):
    """[concat ->] conv -> BatchNorm -> (+residual) -> (ReLU) -> (Dropout), one fused forward/backward.

    Replaces the nn.Conv2d/BatchNorm/ReLU/Dropout runs of zs3/modeling/backbone/resnet.py:33-53,
    zs3/modeling/aspp.py:25-29,111-116 and zs3/modeling/decoder.py:30-38."""

    @staticmethod
    def forward(ctx, conv, bn, relu, drop_p, drop_training, keep_mask, channels, weight, gamma, beta, residual, *xs):
        out, sv = cba_forward(conv, bn, relu, drop_p, drop_training, keep_mask, residual, xs, channels,
                              need_backward=any(ctx.needs_input_grad))
        ctx.sv = sv
        return out

    @staticmethod
    def backward(ctx, dout):
        need = ctx.needs_input_grad
        dxs, dweight, dgamma, dbeta, dres = cba_backward(ctx.sv, dout, list(need[11:]), need_w=need[7],
                                                         need_affine=need[8])
        ctx.sv = None
        return (None, None, None, None, None, None, None, dweight, dgamma, dbeta, dres, *dxs)


class BottleneckFn(torch.autograd.Function):
    """A whole residual bottleneck (zs3/modeling/backbone/resnet.py:33-53) as ONE autograd node: conv1-bn1-relu,
    conv2-bn2-relu, conv3-bn3 (+ downsample conv-bn), residual add, relu.  Fusing the block lets the backward sum
    the two gradients that meet at the block input inside the data-gradient conv's epilogue (accumulate) instead of
    a separate elementwise add."""

    @staticmethod
    def forward(ctx, blk, x, *params):
        nb = any(ctx.needs_input_grad)
        out1, s1 = cba_forward(blk.conv1, blk.bn1, True, 0.0, False, None, None, [x], [blk.inplanes], nb)
        out2, s2 = cba_forward(blk.conv2, blk.bn2, True, 0.0, False, None, None, [out1], [blk.planes], nb)
        sd = None
        res = x
        if blk.downsample is not None:
            res, sd = cba_forward(blk.downsample[0], blk.downsample[1], False, 0.0, False, None, None, [x],
                                  [blk.inplanes], nb)
        out3, s3 = cba_forward(blk.conv3, blk.bn3, True, 0.0, False, None, res, [out2], [blk.planes], nb)
        ctx.saved = (s1, s2, s3, sd)
        ctx.nparams = len(params)
        return out3

    @staticmethod
    def backward(ctx, dout):
        s1, s2, s3, sd = ctx.saved
        ctx.saved = None
        need_x = ctx.needs_input_grad[1]
        need_p = any(ctx.needs_input_grad[2:])
        grads = {}
        (d2,), grads["w3"], grads["g3"], grads["b3"], dres = cba_backward(s3, dout, [True], need_p, need_p)
        (d1,), grads["w2"], grads["g2"], grads["b2"], _ = cba_backward(s2, d2, [True], need_p, need_p)
        dx = None
        if sd is not None:
            (dx,), grads["wd"], grads["gd"], grads["bd"], _ = cba_backward(sd, dres, [need_x], need_p, need_p)
        elif need_x:
            dx = dres
        # the gradient through conv1 is ADDED onto the residual-path gradient in the dgrad epilogue
        into = [(dx, True)] if (need_x and dx is not None) else None
        (dx1,), grads["w1"], grads["g1"], grads["b1"], _ = cba_backward(s1, d1, [need_x], need_p, need_p, dx_into=into)
        if need_x and dx is None:
            dx = dx1
        order = ["w1", "g1", "b1", "w2", "g2", "b2", "w3", "g3", "b3"] + (["wd", "gd", "bd"] if sd is not None else [])
        return (None, dx, *[grads[k] for k in order])


class ASPPFn(torch.autograd.Function):
    """The whole ASPP head (zs3/modeling/aspp.py:103-116) as ONE autograd node.  Forward = the five branches and the
    1280->256 projection (five K-segments, concat never materialised).  The point is the backward: the five branch
    gradients meet at the backbone output and are summed by the data-gradient convs' TMA reduce-add epilogue
    (and the pooling branch by an accumulating broadcast) instead of four elementwise adds by the autograd engine."""

    @staticmethod
    def forward(ctx, aspp, keep_mask, x, *params):
        nb = any(ctx.needs_input_grad)
        n, h, w, _ = x.shape
        cin = aspp.inplanes
        branches, outs = [], []
        for i in range(1, 5):
            m = getattr(aspp, f"aspp{i}")
            o, sv = cba_forward(m.atrous_conv, m.bn, True, 0.0, False, None, None, [x], [cin], nb)
            branches.append(sv)
            outs.append(o)
        width = aspp.conv1.out_channels
        norm = aspp.global_avg_pool[2] if aspp.global_avg_pool_bn else IdentityBN(width, x.device)
        pooled = K.spatial_sum(x, 1.0 / (h * w))
        pg, sg = cba_forward(aspp.global_avg_pool[1], norm, True, 0.0, False, None, None, [pooled], [cin], nb)
        outs.append(K.spatial_broadcast(pg, h, w, 1.0))
        out, sp = cba_forward(aspp.conv1, aspp.bn1, True, aspp.dropout.p, aspp.dropout.training, keep_mask, None, outs,
                              [width] * 5, nb)
        ctx.saved = (branches, sg, sp, (h, w), aspp.global_avg_pool_bn)
        return out

    @staticmethod
    def backward(ctx, dout):
        branches, sg, sp, (h, w), gap_bn = ctx.saved
        ctx.saved = None
        need_x = ctx.needs_input_grad[2]
        need_p = any(ctx.needs_input_grad[3:])
        douts, gw, gg, gb, _ = cba_backward(sp, dout, [True] * 5, need_p, need_p)
        proj = [gw, gg, gb]
        dpg = K.spatial_sum(douts[4].contiguous(), 1.0)
        (dpooled,), pw, pgm, pbt, _ = cba_backward(sg, dpg, [need_x], need_p, need_p and gap_bn)
        dx, grads = None, []
        for sv, d in zip(branches, douts[:4]):
            into = [(dx, True)] if dx is not None else None
            (dxi,), bw, bg, bb, _ = cba_backward(sv, d, [need_x], need_p, need_p, dx_into=into)
            dx = dx if dx is not None else dxi
            grads += [bw, bg, bb]
        if need_x:
            K.spatial_broadcast(dpooled, h, w, 1.0 / (h * w), out=dx, accumulate=True)
        grads += [pw] + ([pgm, pbt] if gap_bn else []) + proj
        return (None, None, dx, *grads)


def aspp_head(aspp, x, keep_mask=None):
    params = []
    for i in range(1, 5):
        m = getattr(aspp, f"aspp{i}")
        params += [m.atrous_conv.weight, m.bn.weight, m.bn.bias]
    params.append(aspp.global_avg_pool[1].weight)
    if aspp.global_avg_pool_bn:
        params += [aspp.global_avg_pool[2].weight, aspp.global_avg_pool[2].bias]
    params += [aspp.conv1.weight, aspp.bn1.weight, aspp.bn1.bias]
    return ASPPFn.apply(aspp, keep_mask, x, *params)


def bottleneck(blk, x):
    params = [blk.conv1.weight, blk.bn1.weight, blk.bn1.bias, blk.conv2.weight, blk.bn2.weight, blk.bn2.bias,
              blk.conv3.weight, blk.bn3.weight, blk.bn3.bias]
    if blk.downsample is not None:
        params += [blk.downsample[0].weight, blk.downsample[1].weight, blk.downsample[1].bias]
    return BottleneckFn.apply(blk, x, *params)


def _pad_to(t, h, w):
    """zero-pad a [N,H,W,C] tensor at the bottom/right to (h, w) (no-op if already that size)."""
    n, hh, ww, c = t.shape
    if hh == h and ww == w:
        return t
    out = torch.zeros((n, h, w, c), dtype=t.dtype, device=t.device)
    out[:, :hh, :ww] = t
    return out


def conv_bn_act(xs, channels, conv, bn, relu=True, residual=None, drop_p=0.0, drop_training=False, keep_mask=None):
    """xs: list of NHWC bf16 tensors (a virtual channel concat), channels: their logical channel counts."""
    return ConvBnAct.apply(conv, bn, relu, drop_p, bool(drop_training), keep_mask, list(channels), conv.weight,
                           bn.weight, bn.bias, residual, *xs)


class ConvBias(torch.autograd.Function):
    """1x1 conv with bias and no normalisation: decoder.pred_conv (zs3/modeling/decoder.py:26,66-68)."""

    @staticmethod
    def forward(ctx, conv, weight, bias, x):
        cout, cin = conv.out_channels, conv.in_channels
        cout_p, cin_p = K.cpad(cout), x.shape[3]
        wp = _packed_weight(conv, 0, 0, cin, cin_p, cout_p)
        bias_p = torch.zeros(cout_p, dtype=torch.float32, device=x.device)
        if bias is not None:
            bias_p[:cout] = bias.detach()
        fl = 2.0 * x.shape[0] * x.shape[1] * x.shape[2] * cout * cin
        y = K.conv_fprop([(x, wp)], 1, 1, 1, 0, 1, cout_p, bias=bias_p, flops=fl)
        ctx.conv, ctx.flops = conv, fl
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        conv = ctx.conv
        cout, cin = conv.out_channels, conv.in_channels
        cout_p, cin_p = K.cpad(cout), x.shape[3]
        dy = dy.contiguous()
        dx = dw_oihw = dbias = None
        if ctx.needs_input_grad[3]:
            wt = _packed_weight(conv, 0, 0, cin, cin_p, cout_p)
            dx = K.conv_fprop([(dy, wt)], 1, 1, 1, 0, 1, cin_p, flops=ctx.flops, kind="conv_dgrad",
                              w_forward_layout=True)
        if ctx.needs_input_grad[1]:
            gbuf = _direct_grad_buffer(conv.weight)
            if gbuf is not None:
                K.conv_wgrad(x, dy, 1, 1, 1, 0, 1, cin_p, cout_p, dw=gbuf, flops=ctx.flops,
                             dw_view=(cin, 0, cout, cin))
            else:
                dw = K.conv_wgrad(x, dy, 1, 1, 1, 0, 1, cin_p, cout_p, flops=ctx.flops)
                dw_oihw = torch.empty_like(conv.weight)
                K.unpack_wgrad(dw, dw_oihw)
        if conv.bias is not None and ctx.needs_input_grad[2]:
            dbias = K.channel_sums(dy)[:cout].float()
        return None, dw_oihw, dbias, dx


class Stem(torch.autograd.Function):
    """conv7x7/s2 -> BN -> ReLU -> maxpool3x3/s2 (zs3/modeling/backbone/resnet.py:186-190) from the NCHW fp32 image."""

    @staticmethod
    def forward(ctx, conv, bn, pool, weight, gamma, beta, x):
        n, c, h, w = x.shape
        R = conv.kernel_size[0]
        stride, pad = conv.stride[0], conv.padding[0]
        cout = conv.out_channels
        cout_p = K.cpad(cout)
        ho, wo = K.conv_out_size(h, R, stride, pad, 1), K.conv_out_size(w, R, stride, pad, 1)
        kreal = c * R * R
        kpad = K.cpad(kreal)
        krsc = K.is_krsc(conv.weight) and not conv.weight.is_contiguous()
        cols = K.stem_im2col(x, R, stride, pad, ho, wo, kpad, krsc=krsc)          # [n, ho, wo, kpad]
        wp = _packed_weight_2d(conv, kreal, kpad, cout_p, krsc)
        training = bn.training
        fl = 2.0 * n * ho * wo * cout * kreal
        if training:
            sb = _StatBuffers.get(x.device)
            stats = sb.current(cout_p)
            y = K.conv_fprop([(cols, wp)], 1, 1, 1, 0, 1, cout_p, flops=fl)
            K.bn_stats(y, stats)  # K = 192: far too short to hide the statistics in the conv epilogue
            sync_w = SBN.sync_world(bn)
            if sync_w > 1:
                SBN.all_reduce_stats(sb.current_buffer())
            coef = torch.empty((4, cout_p), dtype=torch.float32, device=x.device)
            scale, shift, mean, invstd = coef[0], coef[1], coef[2], coef[3]
            if bn.num_batches_tracked is not None:
                PENDING_BATCH_COUNTERS.append(bn.num_batches_tracked)
            fin = dict(stats=stats, count=n * ho * wo * sync_w, gamma=bn.weight.detach(), beta=bn.bias.detach(), eps=bn.eps,
                       momentum=bn.momentum if bn.momentum is not None else 0.1, running_mean=bn.running_mean,
                       running_var=bn.running_var, coef=coef, c_real=cout, reset=sb.flip(), sync_clamp=sync_w > 1)
            a = K.bn_apply(y, scale, shift, True, finalize=fin)
        else:
            y = K.conv_fprop([(cols, wp)], 1, 1, 1, 0, 1, cout_p, flops=fl)
            scale, shift, mean, invstd = _bn_forward_coeffs(bn, None, n * ho * wo, cout_p)
            a = K.bn_apply(y, scale, shift, True)
        k, ps, pp = pool.kernel_size, pool.stride, pool.padding
        out, arg = K.maxpool_fwd(a, k, ps, pp)
        ctx.conv, ctx.bn, ctx.training, ctx.krsc = conv, bn, training, krsc
        ctx.sync_w = SBN.sync_world(bn) if training else 1
        ctx.pool = (k, ps, pp)
        ctx.dims = (kreal, kpad, cout, cout_p)
        ctx.flops = fl
        ctx.save_for_backward(cols, y, a, arg, mean, invstd, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        cols, y, a, arg, mean, invstd, scale = ctx.saved_tensors
        conv, bn = ctx.conv, ctx.bn
        kreal, kpad, cout, cout_p = ctx.dims
        k, ps, pp = ctx.pool
        da = K.maxpool_bwd(dout.contiguous(), arg, a.shape, k, ps, pp)
        dev = dout.device
        dgamma = torch.empty(cout, dtype=torch.float32, device=dev)
        dbeta = torch.empty(cout, dtype=torch.float32, device=dev)
        sc = _scratch64(dev, "bwd")
        sync = (SBN.all_reduce_stats, sc, ctx.sync_w) if ctx.sync_w > 1 else None
        dy = K.bn_backward(da, a, y, mean, invstd, scale, True, training=ctx.training, dgamma=dgamma, dbeta=dbeta,
                           scratch=sc[:2 * cout_p].view(2, cout_p), sync=sync)
        gbuf = _direct_grad_buffer(conv.weight) if ctx.krsc else None
        if gbuf is not None:
            # KRSC memory of the [cout,3,7,7] gradient is exactly the [cout][147] GEMM weight gradient
            K.conv_wgrad(cols, dy, 1, 1, 1, 0, 1, kpad, cout_p, dw=gbuf, flops=ctx.flops, dw_view=(kreal, 0, cout, kreal))
            dweight = None
        else:
            dw = K.conv_wgrad(cols, dy, 1, 1, 1, 0, 1, kpad, cout_p, flops=ctx.flops)
            flat = torch.empty((cout, kreal, 1, 1), dtype=torch.float32, device=dev)
            K.unpack_wgrad(dw, flat)
            if ctx.krsc:  # k ran over (r, s, c)
                dweight = flat.view(cout, conv.kernel_size[0], conv.kernel_size[1], -1).permute(0, 3, 1, 2)
            else:
                dweight = flat.view_as(conv.weight) if conv.weight.is_contiguous() else flat.reshape(conv.weight.shape)
        return None, None, None, dweight, dgamma, dbeta, None  # the image needs no gradient (stem dgrad is never used)


def _packed_weight_2d(conv, kreal, kpad, cout_p, krsc=False):
    w = conv.weight
    cache = conv.__dict__.setdefault("_zs3_pack_cache", {})
    key = ("stem", kpad, cout_p, krsc)
    ent = cache.get(key)
    ver = (w._version, w.data_ptr(), _WEIGHT_EPOCH[0])
    if ent is None or ent[0] != ver:
        w2 = w.detach().permute(0, 2, 3, 1).reshape(w.shape[0], kreal, 1, 1) if krsc else \
            w.detach().reshape(w.shape[0], kreal, 1, 1)
        ent = (ver, K.pack_weight(w2, cout_p, kpad))
        cache[key] = ent
    return ent[1]


class Bilinear(torch.autograd.Function):
    """F.interpolate(bilinear, align_corners=True) NHWC bf16 (zs3/modeling/decoder.py:33-35)."""

    @staticmethod
    def forward(ctx, x, ho, wo):
        ctx.in_hw = (x.shape[1], x.shape[2])
        return K.bilinear_fwd(x, ho, wo)

    @staticmethod
    def backward(ctx, dy):
        return K.bilinear_bwd(dy.contiguous(), ctx.in_hw[0], ctx.in_hw[1]), None, None


class SpatialMean(torch.autograd.Function):
    """nn.AdaptiveAvgPool2d((1,1)) (zs3/modeling/aspp.py:84): [N,H,W,C] -> [N,1,1,C]."""

    @staticmethod
    def forward(ctx, x):
        ctx.shape = x.shape
        n, h, w, c = x.shape
        return K.spatial_sum(x, 1.0 / (h * w))

    @staticmethod
    def backward(ctx, dy):
        n, h, w, c = ctx.shape
        return K.spatial_broadcast(dy.contiguous(), h, w, 1.0 / (h * w))


class SpatialBroadcast(torch.autograd.Function):
    """bilinear 'upsampling' of a 1x1 map = broadcast (zs3/modeling/aspp.py:109): [N,1,1,C] -> [N,H,W,C]."""

    @staticmethod
    def forward(ctx, x, h, w):
        return K.spatial_broadcast(x, h, w, 1.0)

    @staticmethod
    def backward(ctx, dy):
        return K.spatial_sum(dy.contiguous(), 1.0), None, None


class UpsampleLogits(torch.autograd.Function):
    """F.interpolate(..., size=input, bilinear, align_corners=True) of the class scores (deeplab.py:44,55):
    NHWC bf16 in, NCHW fp32 out (what the reference API returns)."""

    @staticmethod
    def forward(ctx, x, c, ho, wo):
        ctx.info = (x.shape, c)
        return K.upsample_logits_fwd(x, c, ho, wo)

    @staticmethod
    def backward(ctx, dy):
        shape, c = ctx.info
        return K.upsample_logits_bwd(dy.contiguous().float(), shape, c), None, None, None


class FromNCHW(torch.autograd.Function):
    """fp32 NCHW (reference API) -> NHWC bf16 (internal)."""

    @staticmethod
    def forward(ctx, x):
        ctx.c = x.shape[1]
        return K.nchw_to_nhwc(x)

    @staticmethod
    def backward(ctx, dy):
        return K.nhwc_to_nchw(dy.contiguous(), ctx.c)


class ToNCHW(torch.autograd.Function):
    """NHWC bf16 (internal) -> fp32 NCHW (reference API)."""

    @staticmethod
    def forward(ctx, x, c):
        ctx.cs = x.shape[3]
        return K.nhwc_to_nchw(x, c)

    @staticmethod
    def backward(ctx, dy):
        return K.nchw_to_nhwc(dy.contiguous(), ctx.cs), None


class CrossEntropy(torch.autograd.Function):
    """nn.CrossEntropyLoss(weight, ignore_index, mean) then /batch (zs3/utils/loss.py:31-46), fused."""

    @staticmethod
    def forward(ctx, logit, target, weight, ignore_index, div):
        logit = logit.contiguous().float()
        target = target.contiguous().float()
        n, c, h, w = logit.shape
        accum = torch.empty(2, dtype=torch.float64, device=logit.device)
        loss = torch.empty((), dtype=torch.float32, device=logit.device)
        wt = None if weight is None else weight.contiguous().float()
        L.check(L.lib().zs3_ce_fwd(L.ptr(logit), L.ptr(target), L.ptr(wt), n, c, h * w, int(ignore_index), float(div),
                                   L.ptr(accum), L.ptr(loss), L.stream_ptr()), "zs3_ce_fwd")
        ctx.save_for_backward(logit, target, accum)
        ctx.wt, ctx.ignore, ctx.div = wt, int(ignore_index), float(div)
        return loss

    @staticmethod
    def backward(ctx, gout):
        logit, target, accum = ctx.saved_tensors
        n, c, h, w = logit.shape
        dlogit = torch.empty_like(logit)
        gout = gout.contiguous().float()
        L.check(L.lib().zs3_ce_bwd(L.ptr(logit), L.ptr(target), L.ptr(ctx.wt), n, c, h * w, ctx.ignore, ctx.div,
                                   L.ptr(accum), L.ptr(gout), L.ptr(dlogit), L.stream_ptr()), "zs3_ce_bwd")
        return dlogit, None, None, None, None
