"""autograd Functions of the GMMN generator and the MMD loss on the fp32 CUDA kernels (csrc/gmmn.cu)."""
import ctypes as C

import torch

from . import _lib as L
from .functional import _RngState


def _f32(t):
    return t.contiguous().float()


def sgemm(A, B, Cout, M, N, K, transA=False, transB=False, bias=None, accumulate=False, idxA=None, idxB=None,
          dyn_count=None, dyn_dim=0):
    a = L.SgemmArgs()
    a.A, a.lda, a.transA = A.data_ptr(), A.stride(0), int(transA)
    a.B, a.ldb, a.transB = B.data_ptr(), B.stride(0), int(transB)
    a.C, a.ldc = Cout.data_ptr(), Cout.stride(0)
    a.M, a.N, a.K = M, N, K
    a.idxA = None if idxA is None else idxA.data_ptr()
    a.idxB = None if idxB is None else idxB.data_ptr()
    a.bias = None if bias is None else bias.data_ptr()
    a.accumulate = int(accumulate)
    a.dyn_count = None if dyn_count is None else dyn_count.data_ptr()
    a.dyn_dim = dyn_dim
    L.check(L.lib().zs3_sgemm(C.byref(a), L.stream_ptr()), "zs3_sgemm")
    return Cout


class Linear(torch.autograd.Function):
    """y = x W^T + b (nn.Linear, zs3/modeling/gmmn.py:18,31,34).  The backward first compacts the rows of dy
    that are non-zero: in the ZS3 step only the 128 sampled rows carry gradient (train_pascal_GMMN.py:229-237)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x, w = _f32(x), _f32(weight)
        n, k = x.shape
        o = w.shape[0]
        y = torch.empty((n, o), dtype=torch.float32, device=x.device)
        sgemm(x, w, y, n, o, k, transB=True, bias=None if bias is None else _f32(bias))
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = _f32(dy)
        n, k = x.shape
        o = w.shape[0]
        dev = x.device
        rows = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        count = torch.empty(1, dtype=torch.int32, device=dev)
        st = L.stream_ptr()
        L.check(L.lib().zs3_find_active_rows(L.ptr(dy), n, o, L.ptr(rows), L.ptr(count), st), "zs3_find_active_rows")
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            # dx = dy W; rows without gradient stay exactly zero
            dxc = torch.empty((n, k), dtype=torch.float32, device=dev)
            sgemm(dy, w, dxc, n, k, o, idxA=rows, dyn_count=count, dyn_dim=1)
            dx = RowScatter.scatter(dxc, rows, count, n)
        if ctx.needs_input_grad[1]:
            dw = torch.empty((o, k), dtype=torch.float32, device=dev)
            sgemm(dy, x, dw, o, k, n, transA=True, idxA=rows, idxB=rows, dyn_count=count, dyn_dim=2)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.empty(o, dtype=torch.float32, device=dev)
            L.check(L.lib().zs3_col_sum(L.ptr(dy), o, L.ptr(rows), L.ptr(count), n, o, L.ptr(db), 0, st), "zs3_col_sum")
        return dx, dw, db


class RowScatter:
    """dense[rows[r]] = compact[r] for r < *count, zeros elsewhere (keeps the sparsity for the next layer)."""

    @staticmethod
    def scatter(compact, rows, count, n):
        out = torch.zeros_like(compact)
        # C[rows[i]] is not expressible as a gather on the output side; use the transposed identity:
        # out = E^T compact with E the selection matrix -> implemented as an index copy on device without host sync
        idx = rows.long()
        valid = (torch.arange(n, device=rows.device) < count.long()).unsqueeze(1)
        out.index_add_(0, torch.where(valid[:, 0], idx, torch.zeros_like(idx)),
                       torch.where(valid, compact, torch.zeros_like(compact)))
        return out


class LeakyDropout(torch.autograd.Function):
    """nn.LeakyReLU(0.2) -> nn.Dropout(p) (zs3/modeling/gmmn.py:19-20) in one pass."""

    @staticmethod
    def forward(ctx, x, slope, p, training, keep_mask):
        x = _f32(x)
        y = torch.empty_like(x)
        p_eff = float(p) if training else 0.0
        mode, seed, off = 0, 0, 0
        if p_eff > 0:
            if keep_mask is not None:
                mode = 2
                keep_mask = keep_mask.contiguous().to(torch.uint8)
            else:
                mode = 1
                seed, off = _RngState.next((x.numel() + 3) // 4)
        L.check(L.lib().zs3_leaky_dropout_fwd(L.ptr(x), L.ptr(y), x.numel(), float(slope), mode, p_eff, seed, off,
                                              L.ptr(keep_mask) if mode == 2 else None, L.stream_ptr()),
                "zs3_leaky_dropout_fwd")
        ctx.save_for_backward(y)
        ctx.slope, ctx.p = float(slope), p_eff
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = _f32(dy)
        dx = torch.empty_like(dy)
        rows, cols = y.shape
        L.check(L.lib().zs3_leaky_dropout_bwd(L.ptr(dy), L.ptr(y), L.ptr(dx), rows, cols, None, None, ctx.slope, ctx.p,
                                              L.stream_ptr()), "zs3_leaky_dropout_bwd")
        return dx, None, None, None, None


class Concat2(torch.autograd.Function):
    """torch.cat((embd, noise), 1) (gmmn.py:44)"""

    @staticmethod
    def forward(ctx, a, b):
        a, b = _f32(a), _f32(b)
        n = a.shape[0]
        y = torch.empty((n, a.shape[1] + b.shape[1]), dtype=torch.float32, device=a.device)
        L.check(L.lib().zs3_concat2(L.ptr(a), a.shape[1], L.ptr(b), b.shape[1], L.ptr(y), n, L.stream_ptr()),
                "zs3_concat2")
        ctx.c1 = a.shape[1]
        return y

    @staticmethod
    def backward(ctx, dy):
        return dy[:, :ctx.c1], dy[:, ctx.c1:]


class MomentLoss(torch.autograd.Function):
    """GMMNLoss.moment_loss (zs3/utils/loss.py:99-115): forward + analytic backward, fp32."""

    @staticmethod
    def forward(ctx, gen, real, sigma):
        gen, real = _f32(gen), _f32(real)
        M, D = gen.shape
        N = real.shape[0]
        Lr = M + N
        dev = gen.device
        P = torch.empty((Lr, Lr), dtype=torch.float32, device=dev)
        loss2 = torch.empty(1, dtype=torch.float64, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        sig = (C.c_float * len(sigma))(*[float(s) for s in sigma])
        L.check(L.lib().zs3_mmd_fwd(L.ptr(gen), L.ptr(real), M, N, D, sig, len(sigma), L.ptr(P), L.ptr(loss2),
                                    L.ptr(loss), L.stream_ptr()), "zs3_mmd_fwd")
        ctx.save_for_backward(gen, real, P, loss)
        return loss

    @staticmethod
    def backward(ctx, gout):
        gen, real, P, loss = ctx.saved_tensors
        M, D = gen.shape
        N = real.shape[0]
        gout = _f32(gout).reshape(1)
        dgen = torch.empty_like(gen) if ctx.needs_input_grad[0] else None
        dreal = torch.empty_like(real) if ctx.needs_input_grad[1] else None
        L.check(L.lib().zs3_mmd_bwd(L.ptr(gen), L.ptr(real), M, N, D, L.ptr(P), L.ptr(loss.reshape(1)), L.ptr(gout),
                                    L.ptr(dgen), L.ptr(dreal), L.stream_ptr()), "zs3_mmd_bwd")
        return dgen, dreal, None
