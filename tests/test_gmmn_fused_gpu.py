"""Fused generator-update kernel (zs3_gmmn_train_fused, csrc/gmmn_fused.cu) on the GPU: against the reference's
golden gradients (tests/golden/gmmn.npz), the autograd oracle with an injected Dropout mask, sequential Adam
updates, and bit-reproducibility.  Tolerance 1e-3 on the GMMN loss as in BASELINE.json's north_star (fp32
arithmetic: observed ~1e-6)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import rel_l2

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))

pytestmark = pytest.mark.gpu
TOL = 1e-3
DIMS = (300, 300, 256, 256)


def _inputs():  # same draws as tests/test_gmmn_gpu.py / tests/golden/make_golden.py
    g = torch.Generator().manual_seed(21)
    emb = torch.randn(200, 300, generator=g) * 0.06
    z = torch.rand(200, 300, generator=g)
    real = torch.relu(torch.randn(200, 256, generator=g))
    idx = torch.randint(0, 200, (128,), generator=g)
    return emb, z, real, idx


def _generator(seed=3, train=True):
    import zs3_oracle as O
    from zs3_b200.modeling.gmmn import GMMNnetwork
    gst = O.init_gmmn_state(seed=seed)
    gen = GMMNnetwork(300, 300, 256, 256)
    gen.load_state_dict(gst)
    gen = gen.cuda()
    gen.train(train)
    return gst, gen


def _item(emb, z, real, ridx, mask=None, nchw=False):
    """device tensors; real features optionally laid out like an NCHW map ([256][pixels])"""
    from zs3_b200 import gmmn_fused as GF
    keep = [emb, z, ridx]
    if nchw:
        real_t = real.t().contiguous()
        rs = GF.row_source(real_t, ridx, row_stride=1, col_stride=real.shape[0])
        keep.append(real_t)
    else:
        rs = GF.row_source(real, ridx)
        keep.append(real)
    m8 = None if mask is None else mask.to(torch.uint8).contiguous()
    keep.append(m8)
    return GF.pack_item(GF.row_source(emb, ridx), GF.row_source(z, ridx), rs, ridx.numel(), keep_mask=m8,
                        keep_rows=ridx), keep


def test_fused_gradients_vs_reference_golden():
    from zs3_b200.gmmn_fused import FusedGeneratorUpdater
    gold = np.load(os.path.join(HERE, "golden", "gmmn.npz"))
    emb, z, real, idx = (t.cuda() for t in _inputs())
    _, gen = _generator(train=False)
    upd = FusedGeneratorUpdater(gen, torch.optim.Adam(gen.parameters(), lr=2e-4))
    item, keep = _item(emb, z, real, idx.to(torch.int32))
    loss, grads = upd.gradients(item, 300, 300)
    torch.cuda.synchronize()
    assert abs(loss.item() - float(gold["mmd_loss"])) < TOL * float(gold["mmd_loss"])
    for (k, _), g in zip(gen.named_parameters(), grads):
        g = g.detach().cpu().reshape(-1)
        assert abs(g.double().norm().item() - float(gold["gen_gradnorm/" + k])) < TOL * float(gold["gen_gradnorm/" + k])
        assert rel_l2(g[:: max(1, g.numel() // 2048)][:2048], torch.from_numpy(gold["gen_grad/" + k])) < TOL


@pytest.mark.parametrize("rows,nchw", [(128, True), (77, False)])
def test_fused_gradients_with_dropout_mask_vs_oracle(rows, nchw):
    import zs3_oracle as O
    from zs3_b200.gmmn_fused import FusedGeneratorUpdater
    emb, z, real, _ = _inputs()
    g = torch.Generator().manual_seed(5)
    mask = torch.rand(200, 256, generator=g) > 0.5
    ridx = torch.randint(0, 200, (rows,), generator=g)
    gst, gen = _generator()
    st = {k: v.clone().requires_grad_(True) for k, v in gst.items()}
    ref_loss = O.moment_loss(O.gmmn_forward(st, emb, z, training=True, keep_mask=mask)[ridx], real[ridx])
    ref = torch.autograd.grad(ref_loss, list(st.values()))
    upd = FusedGeneratorUpdater(gen, torch.optim.Adam(gen.parameters(), lr=2e-4))
    item, keep = _item(emb.cuda(), z.cuda(), real.cuda(), ridx.to(torch.int32).cuda(), mask.cuda(), nchw=nchw)
    loss, grads = upd.gradients(item, 300, 300)
    torch.cuda.synchronize()
    assert abs(loss.item() - ref_loss.item()) < TOL * abs(ref_loss.item())
    for gg, r, k in zip(grads, ref, st):
        assert rel_l2(gg.cpu(), r) < TOL, k


def _work_list(n_items, seed0=30):
    cases = []
    for i in range(n_items):
        g = torch.Generator().manual_seed(seed0 + i)
        n = 150 + 17 * i
        cases.append((torch.randn(n, 300, generator=g) * 0.06, torch.rand(n, 300, generator=g),
                      torch.relu(torch.randn(n, 256, generator=g)), torch.rand(n, 256, generator=g) > 0.5,
                      torch.randint(0, n, (128,), generator=g)))
    return cases


def _run_list(cases, step_offset=0):
    from zs3_b200.gmmn_fused import FusedGeneratorUpdater
    _, gen = _generator()
    opt = torch.optim.Adam(gen.parameters(), lr=2e-4)
    upd = FusedGeneratorUpdater(gen, opt)
    items, keep = [], []
    for emb, z, real, mask, ridx in cases:
        it, k = _item(emb.cuda(), z.cuda(), real.cuda(), ridx.to(torch.int32).cuda(), mask.cuda(), nchw=True)
        items.append(it)
        keep.append(k)
    losses = upd.run(items, 300, 300, keepalive=keep)
    torch.cuda.synchronize()
    return gen, opt, losses.cpu()


def test_fused_work_list_matches_sequential_oracle_updates():
    """five dependent updates in ONE launch == five oracle iterations of train_pascal_GMMN.py:211-240"""
    import zs3_oracle as O
    cases = _work_list(5)
    gst = O.init_gmmn_state(seed=3)
    ref = {k: v.clone().requires_grad_(True) for k, v in gst.items()}
    adam = {k: [torch.zeros_like(v), torch.zeros_like(v)] for k, v in ref.items()}
    ref_losses = []
    for t, (emb, z, real, mask, ridx) in enumerate(cases):
        loss = O.moment_loss(O.gmmn_forward(ref, emb, z, training=True, keep_mask=mask)[ridx], real[ridx])
        ref_losses.append(loss.item())
        grads = torch.autograd.grad(loss, list(ref.values()))
        with torch.no_grad():
            for (k, p), g in zip(ref.items(), grads):
                O.adam_step(p, g, adam[k][0], adam[k][1], t + 1)
    gen, opt, losses = _run_list(cases)
    assert np.allclose(losses.numpy(), ref_losses, rtol=TOL)
    for k, p in gen.state_dict().items():
        assert rel_l2(p.cpu(), ref[k].detach()) < 1e-5, k
        assert rel_l2(p.cpu() - gst[k], ref[k].detach() - gst[k]) < 5e-3, k      # the accumulated update itself
    # the Adam state lives in the torch optimizer: step counters advanced, moments match, a stock step still works
    for p, k in zip(gen.parameters(), ref):
        st = opt.state[p]
        assert int(st["step"]) == 5
        assert rel_l2(st["exp_avg"].cpu(), adam[k][0]) < 1e-3, k
        assert rel_l2(st["exp_avg_sq"].cpu(), adam[k][1]) < 1e-3, k
    for p in gen.parameters():
        p.grad = torch.zeros_like(p)
    opt.step()
    assert int(opt.state[next(gen.parameters())]["step"]) == 6


def test_fused_work_list_is_bit_reproducible():
    cases = _work_list(3, seed0=50)
    gen_a, _, la = _run_list(cases)
    gen_b, _, lb = _run_list(cases)
    assert torch.equal(la, lb)
    for pa, pb in zip(gen_a.parameters(), gen_b.parameters()):
        assert torch.equal(pa, pb)


def test_fused_counter_rng_dropout_and_argument_checks():
    from zs3_b200 import _lib as L
    from zs3_b200 import gmmn_fused as GF
    emb, z, real, idx = (t.cuda() for t in _inputs())
    _, gen = _generator()
    upd = GF.FusedGeneratorUpdater(gen, torch.optim.Adam(gen.parameters(), lr=2e-4))
    ridx = idx.to(torch.int32)
    item = GF.pack_item(GF.row_source(emb, ridx), GF.row_source(z, ridx), GF.row_source(real, ridx), 128, keep_rows=ridx)
    l1, g1 = upd.gradients(item, 300, 300)
    l2, g2 = upd.gradients(item, 300, 300)      # a new call draws a new mask
    torch.cuda.synchronize()
    assert torch.isfinite(l1) and torch.isfinite(l2) and l1.item() != l2.item()
    # about half of the hidden units are dropped: db1 = column sums of dH is exactly zero only for a column whose
    # 128 rows were all dropped (p = 2^-128); the bias gradient of the second layer never is
    assert all(torch.isfinite(g).all() for g in g1)
    with pytest.raises(ValueError):
        GF.pack_item(GF.row_source(emb), GF.row_source(z), GF.row_source(real), 129)
    with pytest.raises(L.Zs3NativeError):
        a = L.GmmnTrainArgs()
        L.check(L.lib().zs3_gmmn_train_fused(a, None), "zs3_gmmn_train_fused")
    with pytest.raises(RuntimeError):
        _, cpu_gen = _generator()
        cpu_gen = cpu_gen.cpu()
        GF.FusedGeneratorUpdater(cpu_gen, torch.optim.Adam(cpu_gen.parameters())).run([item], 300, 300)
