"""Host side of the fused generator update (csrc/gmmn_fused.cu, `zs3_gmmn_train_fused`).

One launch = a work list of sequential (image, class) generator iterations of zs3/train_pascal_GMMN.py:211-240
(gather of the sampled rows -> GMMNnetwork.forward -> GMMNLoss.moment_loss -> backward -> Adam.step).  The packing
helpers only put pointers and strides into the C structs of include/zs3b200.h; `FusedGeneratorUpdater.run`
launches on the current CUDA stream and refuses anything that is not a CUDA tensor (there is no CPU path).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L

MAX_ROWS = 128


def row_source(t, rows=None, row_stride=None, col_stride=None):
    """zs3_row_source over tensor `t`: element (r, k) = t.data[row(r) * row_stride + k * col_stride], row(r) = rows[r]
    (int32 tensor) or r.  Default strides are those of a 2-D tensor."""
    if t.dtype != torch.float32:
        raise TypeError("row sources are fp32 tensors")
    if rows is not None and (rows.dtype != torch.int32 or not rows.is_contiguous()):
        raise TypeError("row gathers are contiguous int32 tensors")
    s = L.RowSource()
    s.base = t.data_ptr()
    s.rows = None if rows is None else rows.data_ptr()
    s.row_stride = t.stride(0) if row_stride is None else int(row_stride)
    s.col_stride = (t.stride(1) if t.dim() > 1 else 1) if col_stride is None else int(col_stride)
    return s


FORWARD_ONLY = 1   # ZS3_GMMN_FORWARD_ONLY


def pack_item(emb, noise, real, rows, keep_mask=None, keep_rows=None, adj=None, out=None, forward_only=False):
    """emb / noise / real: zs3_row_source; keep_mask: uint8 [*, hidden] contiguous or None; keep_rows: int32 or None;
    adj: fp32 [rows, rows] contiguous adjacency (graph generator) or None; out: fp32 [rows, feat] contiguous tensor that
    receives the generated features, or None; forward_only: generate `out` and stop (no loss / backward / Adam)."""
    if not 1 <= int(rows) <= MAX_ROWS:
        raise ValueError(f"a generator update samples 1..{MAX_ROWS} rows, got {rows}")
    it = L.GmmnItem()
    it.emb, it.noise, it.real = emb, noise, real
    if keep_mask is not None and (keep_mask.dtype != torch.uint8 or not keep_mask.is_contiguous()):
        raise TypeError("keep_mask is a contiguous uint8 tensor")
    it.keep_mask = None if keep_mask is None else keep_mask.data_ptr()
    it.keep_rows = None if keep_rows is None else keep_rows.data_ptr()
    if adj is not None:
        if adj.dtype != torch.float32 or not adj.is_contiguous() or tuple(adj.shape) != (int(rows), int(rows)):
            raise TypeError("adj is a contiguous fp32 [rows, rows] tensor")
        it.adj = adj.data_ptr()
    if out is not None:
        if out.dtype != torch.float32 or not out.is_contiguous() or out.shape[0] != int(rows):
            raise TypeError("out is a contiguous fp32 [rows, feat] tensor")
        it.out = out.data_ptr()
    if forward_only and out is None:
        raise ValueError("a forward-only item needs an `out` tensor")
    it.rows = int(rows)
    it.flags = FORWARD_ONLY if forward_only else 0
    return it


def items_to_bytes(items):
    """list of zs3_gmmn_item, or the numpy structured array of pack_items_vectorized"""
    if isinstance(items, np.ndarray):
        return items.tobytes()
    arr = (L.GmmnItem * len(items))(*items)
    return bytes(memoryview(arr).cast("B"))


ITEM_DTYPE = np.dtype(L.GmmnItem)   # numpy mirror of zs3_gmmn_item (same offsets: derived from the ctypes struct)


def pack_items_vectorized(images, rows, emb, emb_rows, noise, real, real_rows, keep_rows):
    """A whole work list at once.  Update k belongs to image images[k]; `emb` / `real` = (base pointer, bytes per
    image, column stride in elements) of an NCHW-like map whose rows are gathered through emb_rows[k] / real_rows[k]
    (int32 [n, rows] CUDA tensors, row stride 1; an optional 4th tuple entry overrides the row stride, e.g. a
    [C, E] class-embedding table gathered by class id); noise [n, rows, Z] fp32 (dense); keep_rows int32 [n, rows]."""
    n = len(images)
    a = np.zeros(n, dtype=ITEM_DTYPE)
    k = np.arange(n, dtype=np.int64)
    for name, src, idx in (("emb", emb, emb_rows), ("real", real, real_rows)):
        base, img_bytes, cs = src[:3]
        rs = src[3] if len(src) > 3 else 1     # row stride in elements (a [C, E] class-embedding table: E, with cs = 1)
        if idx.dtype != torch.int32 or not idx.is_contiguous():
            raise TypeError("row gathers are contiguous int32 tensors")
        a[name]["base"] = (base + images * img_bytes).astype(np.uint64)
        a[name]["rows"] = (idx.data_ptr() + k * (rows * 4)).astype(np.uint64)
        a[name]["row_stride"], a[name]["col_stride"] = rs, cs
    if noise.dtype != torch.float32 or not noise.is_contiguous():
        raise TypeError("noise is a contiguous fp32 tensor [n, rows, Z]")
    a["noise"]["base"] = (noise.data_ptr() + k * (rows * noise.shape[2] * 4)).astype(np.uint64)
    a["noise"]["rows"] = 0
    a["noise"]["row_stride"], a["noise"]["col_stride"] = noise.shape[2], 1
    a["keep_mask"] = 0
    a["keep_rows"] = (keep_rows.data_ptr() + k * (rows * 4)).astype(np.uint64)
    if not 1 <= rows <= MAX_ROWS:
        raise ValueError(f"a generator update samples 1..{MAX_ROWS} rows, got {rows}")
    a["rows"] = rows
    return a


def pack_args(items_ptr, n_items, dims, params, sigma, losses, workspace, *, adam=None, grads=None, lr=2e-4,
              betas=(0.9, 0.999), eps=1e-8, step0=0, slope=0.2, drop_p=0.5, seed=0, offset=0, max_rows=MAX_ROWS,
              phase_stamps=None, weights_in_out=False):
    """dims = (embed_dim, noise_dim, hidden, feat); params = (w1, b1, w2, b2) fp32 contiguous tensors;
    adam = ([exp_avg x4], [exp_avg_sq x4]) for in-place Adam, or grads = [g x4] for gradient output."""
    a = L.GmmnTrainArgs()
    a.items, a.n_items, a.max_rows = items_ptr, n_items, max_rows
    a.embed_dim, a.noise_dim, a.hidden, a.feat = dims
    for t in params:
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise TypeError("generator parameters must be contiguous fp32 tensors")
    a.w1, a.b1, a.w2, a.b2 = (t.data_ptr() for t in params)
    a.apply_adam = 1 if adam is not None else 0
    if adam is not None:
        for i in range(4):
            a.adam_m[i], a.adam_v[i] = adam[0][i].data_ptr(), adam[1][i].data_ptr()
    if grads is not None:
        for i in range(4):
            a.grad[i] = grads[i].data_ptr()
    a.lr, a.beta1, a.beta2, a.eps, a.step0 = lr, betas[0], betas[1], eps, int(step0)
    if not 1 <= len(sigma) <= 8:
        raise ValueError("1..8 MMD bandwidths")
    for i, s in enumerate(sigma):
        a.sigma[i] = float(s)
    a.nsigma = len(sigma)
    a.slope, a.drop_p, a.seed, a.offset = slope, drop_p, int(seed), int(offset)
    a.losses = losses.data_ptr()
    a.workspace, a.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    a.phase_stamps = None if phase_stamps is None else phase_stamps.data_ptr()   # int64 [n_items, 8]
    a.weights_in_out = 1 if weights_in_out else 0   # pygcn GraphConvolution weights are [in, out]
    return a


def _launch_seed(calls):
    """64-bit Dropout seed of one launch: splitmix64 of (torch seed, launch counter).  The kernel adds
    (item << 40) + element/4 to the offset, so the launch counter must not share those bits (round 1 passed it as
    offset = calls << 44: item 16 of launch c drew the masks of item 0 of launch c+1, and it wrapped after 2^20
    launches)."""
    m = 0xFFFFFFFFFFFFFFFF
    z = (torch.initial_seed() + 0x9E3779B97F4A7C15 * (calls + 1)) & m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    return z ^ (z >> 31)


class FusedGeneratorUpdater:
    """Runs work lists of generator updates for a `GMMNnetwork` (hidden_size > 0) or a `GMMNnetwork_GCN` (items then
    carry the image's adjacency matrix) and its torch.optim.Adam.
    The Adam state (`exp_avg`, `exp_avg_sq`, `step`) stays in `optimizer.state`, so checkpoints and a later
    `optimizer.step()` keep working."""

    def __init__(self, generator, optimizer, sigma=(2, 5, 10, 20, 40, 80)):
        self.generator, self.optimizer, self.sigma = generator, optimizer, tuple(float(s) for s in sigma)
        if hasattr(generator, "gcn1") and hasattr(generator, "gcn2"):
            # GMMNnetwork_GCN (gmmn.py:52-67): two GraphConvolutions, weights [in, out]; items carry the adjacency
            self.graph = True
            self.lin1, self.act, self.drop, self.lin2 = generator.gcn1, generator.relu, generator.dropout, generator.gcn2
            self.params = (self.lin1.weight, self.lin1.bias, self.lin2.weight, self.lin2.bias)
            self.in_dim, self.hidden = self.lin1.weight.shape
            self.feat = self.lin2.weight.shape[1]
        else:
            self.graph = False
            model = generator.model
            if not isinstance(model, torch.nn.Sequential) or len(model) != 4:
                raise NotImplementedError("the fused update covers the one-hidden-layer generator (gmmn.py:17-21)")
            if getattr(generator, "semantic_reconstruction", False):
                raise NotImplementedError("semantic_reconstruction has a second output; use the unfused modules")
            self.lin1, self.act, self.drop, self.lin2 = model[0], model[1], model[2], model[3]
            self.params = (self.lin1.weight, self.lin1.bias, self.lin2.weight, self.lin2.bias)
            self.hidden, self.in_dim = self.lin1.weight.shape
            self.feat = self.lin2.weight.shape[0]
        if not isinstance(optimizer, torch.optim.Adam):
            raise NotImplementedError("the fused update implements torch.optim.Adam (train_pascal_GMMN.py:65-67)")
        if len(optimizer.param_groups) != 1:
            raise NotImplementedError("one Adam parameter group expected")
        g = optimizer.param_groups[0]
        if g.get("weight_decay", 0) or g.get("amsgrad", False) or g.get("maximize", False):
            raise NotImplementedError("plain Adam only (no weight decay / amsgrad / maximize)")
        ids = {id(p) for p in g["params"]}
        if ids != {id(p) for p in self.params}:
            raise ValueError("the optimizer must hold exactly the generator's four parameters")
        self._workspace = None
        self._calls = 0

    def _adam_state(self):
        ms, vs, step = [], [], None
        for p in self.params:
            st = self.optimizer.state[p]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            ms.append(st["exp_avg"])
            vs.append(st["exp_avg_sq"])
            s = int(st["step"].item()) if torch.is_tensor(st["step"]) else int(st["step"])
            step = s if step is None else step
            if s != step:
                raise RuntimeError("Adam step counters of the generator parameters disagree")
        return ms, vs, step

    def run(self, items, embed_dim, noise_dim, keepalive=(), phase_stamps=None):
        """items: list of zs3_gmmn_item (pack_item); returns the device tensor of their moment losses.
        phase_stamps: optional int64 CUDA tensor [len(items), 8] receiving %globaltimer at the phase boundaries."""
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("zs3_b200 runs on CUDA (sm_100a) tensors only; there is no CPU path")
        if embed_dim + noise_dim != self.in_dim:
            raise ValueError("embed_dim + noise_dim must equal the generator's input width")
        n = len(items)
        losses = torch.empty(n, dtype=torch.float32, device=dev)
        if n == 0:
            return losses
        lib = L.lib()
        if self._workspace is None:
            nbytes = lib.zs3_gmmn_train_workspace_size(embed_dim, noise_dim, self.hidden, self.feat)
            self._workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        ms, vs, step = self._adam_state()
        g = self.optimizer.param_groups[0]
        host = torch.frombuffer(bytearray(items_to_bytes(items)), dtype=torch.uint8)
        dev_items = host.to(dev, non_blocking=False)
        training = self.generator.training
        self._calls += 1
        a = pack_args(dev_items.data_ptr(), n, (embed_dim, noise_dim, self.hidden, self.feat),
                      tuple(p.data for p in self.params), self.sigma, losses, self._workspace, adam=(ms, vs),
                      lr=float(g["lr"]), betas=tuple(g["betas"]), eps=float(g["eps"]), step0=step,
                      slope=float(self.act.negative_slope), drop_p=float(self.drop.p) if training else 0.0,
                      seed=_launch_seed(self._calls), offset=0, phase_stamps=phase_stamps, weights_in_out=self.graph)
        L.check(lib.zs3_gmmn_train_fused(C.byref(a), L.stream_ptr()), "zs3_gmmn_train_fused")
        n_upd = n if isinstance(items, np.ndarray) else sum(1 for it in items if not (it.flags & FORWARD_ONLY))
        for p in self.params:
            self.optimizer.state[p]["step"] += n_upd      # forward-only items take no Adam step
        del keepalive  # the caller's tensors were only needed until the launch was enqueued (stream-ordered allocator)
        return losses

    def gradients(self, item, embed_dim, noise_dim):
        """(loss, [dW1, db1, dW2, db2]) of ONE update without applying it (tests / drop-in autograd use)."""
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("zs3_b200 runs on CUDA (sm_100a) tensors only; there is no CPU path")
        lib = L.lib()
        if self._workspace is None:
            nbytes = lib.zs3_gmmn_train_workspace_size(embed_dim, noise_dim, self.hidden, self.feat)
            self._workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        losses = torch.empty(1, dtype=torch.float32, device=dev)
        grads = [torch.empty_like(p) for p in self.params]
        dev_items = torch.frombuffer(bytearray(items_to_bytes([item])), dtype=torch.uint8).to(dev)
        self._calls += 1
        a = pack_args(dev_items.data_ptr(), 1, (embed_dim, noise_dim, self.hidden, self.feat),
                      tuple(p.data for p in self.params), self.sigma, losses, self._workspace, grads=grads,
                      slope=float(self.act.negative_slope),
                      drop_p=float(self.drop.p) if self.generator.training else 0.0,
                      seed=_launch_seed(self._calls), offset=0, weights_in_out=self.graph)
        L.check(lib.zs3_gmmn_train_fused(C.byref(a), L.stream_ptr()), "zs3_gmmn_train_fused")
        return losses[0], grads
