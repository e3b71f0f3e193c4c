"""CPU (gloo, world_size 2): host logic of the data-parallel runtime -- batch sharding, flat parameter/gradient
buffers, one all-reduce per step, averaging equivalent to the single-process full batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_batch_partitions_exactly():
    from zs3_b200.parallel import shard_batch
    for n in (16, 17, 3, 128):
        for world in (1, 2, 3, 8):
            spans = [shard_batch(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_flat_params_views_and_grad_accumulation():
    from zs3_b200.parallel import FlatParams
    torch.manual_seed(0)
    m1, m2 = torch.nn.Linear(5, 3), torch.nn.Linear(3, 2)
    ref = [p.detach().clone() for p in list(m1.parameters()) + list(m2.parameters())]
    fp = FlatParams([list(m1.parameters()), list(m2.parameters())])
    assert fp.group_ranges == [(0, 18), (18, 26)]
    for p, r in zip(fp.params, ref):
        assert torch.equal(p.detach(), r)
        assert p.data_ptr() >= fp.flat.data_ptr()
    fp.zero_grad()
    out = m2(torch.relu(m1(torch.randn(4, 5)))).sum()
    out.backward()
    assert fp.grad.abs().sum() > 0
    assert all(p.grad.data_ptr() >= fp.grad.data_ptr() for p in fp.params)  # autograd accumulated in place
    g1 = fp.grad.clone()
    out2 = m2(torch.relu(m1(torch.randn(4, 5)))).sum()
    out2.backward()
    assert not torch.equal(fp.grad, g1)
    fp.zero_grad()
    assert fp.grad.abs().sum() == 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from zs3_b200.parallel import FlatParams, init_distributed, shard_batch
    r, _, w = init_distributed()
    assert (r, w) == (rank, world) and dist.get_backend() == "gloo"
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.ReLU(), torch.nn.Linear(4, 2))
    x, y = torch.randn(8, 6), torch.randn(8, 2)
    fp = FlatParams([list(model.parameters())])
    dist.broadcast(fp.flat, 0)
    a, b = shard_batch(8, rank, world)
    fp.zero_grad()
    # per-rank mean loss over its shard, as the CE kernel divides by the local batch
    loss = ((model(x[a:b]) - y[a:b]) ** 2).mean()
    loss.backward()
    dist.all_reduce(fp.grad)          # the one collective of the step
    fp.grad.mul_(1.0 / world)         # folded into the fused SGD kernel on the GPU path (grad_scale)
    q.put((rank, fp.grad.clone()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_allreduce_matches_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=90) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.ReLU(), torch.nn.Linear(4, 2))
    x, y = torch.randn(8, 6), torch.randn(8, 2)
    loss = ((model(x) - y) ** 2).mean()
    loss.backward()
    full = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.allclose(got[0], got[1])
    assert torch.allclose(got[0], full, atol=1e-6)


def _worker_step2(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from zs3_b200.parallel import exchange_step2, init_distributed
    init_distributed()
    torch.manual_seed(0)
    gen = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.LeakyReLU(0.2), torch.nn.Linear(4, 3))
    head = torch.nn.Linear(3, 2)
    gp = list(gen.parameters())
    snap = torch.cat([p.detach().reshape(-1) for p in gp])
    # each rank's own chain of sequential generator updates (different data per rank), then its classifier backward
    opt = torch.optim.Adam(gp, lr=1e-2)
    g = torch.Generator().manual_seed(100 + rank)
    for _ in range(3):
        opt.zero_grad()
        (gen(torch.randn(5, 6, generator=g)) ** 2).mean().backward()
        opt.step()
    head(torch.randn(4, 3, generator=g)).pow(2).mean().backward()
    local_delta = torch.cat([p.detach().reshape(-1) for p in gp]) - snap
    local_grad = torch.cat([p.grad.reshape(-1) for p in head.parameters()])
    exchange_step2(gp, snap, list(head.parameters()), world)
    q.put((rank,) + tuple(t.numpy().copy() for t in (       # by value: the worker may exit before the parent reads
        local_delta, local_grad, torch.cat([p.detach().reshape(-1) for p in gp]),
        torch.cat([p.grad.reshape(-1) for p in head.parameters()]), snap)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_step2_exchange_averages_generator_delta_and_classifier_gradients():
    """step 2 shards by image; ONE message [generator delta | pred_conv gradients] per iteration.  After the
    exchange both ranks hold snapshot + mean(delta) and the mean classifier gradient (gloo, world_size 2)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_step2, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {r[0]: tuple(torch.from_numpy(a) for a in r[1:]) for r in (q.get(timeout=90) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (d0, g0, w0, cg0, snap), (d1, g1, w1, cg1, _) = got[0], got[1]
    assert not torch.allclose(d0, d1)                         # the ranks really diverged before the exchange
    assert torch.allclose(w0, w1) and torch.allclose(cg0, cg1)
    assert torch.allclose(w0, snap + (d0 + d1) / 2, atol=1e-7)
    assert torch.allclose(cg0, (g0 + g1) / 2, atol=1e-7)


def test_step2_exchange_single_rank_is_identity():
    from zs3_b200.parallel import exchange_step2
    torch.manual_seed(1)
    gen, head = torch.nn.Linear(4, 4), torch.nn.Linear(4, 2)
    snap = torch.cat([p.detach().reshape(-1) for p in gen.parameters()]) - 0.5
    head(torch.randn(3, 4)).sum().backward()
    before = [p.detach().clone() for p in gen.parameters()], [p.grad.clone() for p in head.parameters()]
    exchange_step2(list(gen.parameters()), snap, list(head.parameters()), 1)
    for a, b in zip(gen.parameters(), before[0]):
        assert torch.allclose(a, b, atol=1e-7)
    for a, b in zip(head.parameters(), before[1]):
        assert torch.equal(a.grad, b)


class _FreeingBlock(torch.autograd.Function):
    """y = tanh(x * w); drops its saved state after backward like the fused network nodes (functional.BottleneckFn),
    so a node that is run twice fails loudly.  in_place=True mimics the nodes that add the weight gradient into
    w.grad themselves and return None for it (KRSC conv weights, functional.py)."""

    @staticmethod
    def forward(ctx, x, w, in_place):
        ctx.saved = (x, w, in_place)
        return torch.tanh(x * w)

    @staticmethod
    def backward(ctx, g):
        x, w, in_place = ctx.saved
        ctx.saved = None
        d = g * (1 - torch.tanh(x * w) ** 2)
        gw = (d * x).sum().reshape(w.shape)
        if in_place:
            w.grad = gw if w.grad is None else w.grad + gw
            gw = None
        return d * w, gw, None


def _toy_forward(ws, inp, alias):
    """stem -> layer1 (= low-level feature) -> layer2 (= x) | layer3(x) + decoder(low) -> loss: the shape of the
    DeepLab graph around the backbone cut (zs3_b200/modeling/backbone/resnet.py)"""
    h = _FreeingBlock.apply(inp, ws[0], False)
    low = _FreeingBlock.apply(h, ws[1], True)
    x = _FreeingBlock.apply(low, ws[2], False)
    low_dec = low.view_as(low) if alias else low
    out = _FreeingBlock.apply(x, ws[3], True) + _FreeingBlock.apply(low_dec, ws[4], False)
    return (out ** 2).mean(), (x, low_dec)


def test_cut_backward_two_stage_equals_single_pass():
    """parallel.cut_backward: gradients above the cut first (so their all-reduce can start), the rest in tail();
    the cut has to be an antichain -- naming layer1's output itself (upstream of layer2's output) would hand the
    decoder's gradient AND layer2's gradient to the same tensor, i.e. stage 1 would have to run layer2"""
    from zs3_b200.parallel import cut_backward
    torch.manual_seed(0)
    inp = torch.randn(4, 5)
    w0 = [torch.randn(1).requires_grad_(True) for _ in range(5)]
    loss, _ = _toy_forward(w0, inp, alias=True)
    loss.backward()
    ref = [w.grad.clone() for w in w0]

    w1 = [w.detach().clone().requires_grad_(True) for w in w0]
    loss, cut = _toy_forward(w1, inp, alias=True)
    tail = cut_backward(loss, list(cut), [w1[3], w1[4]])
    assert w1[3].grad is not None and w1[4].grad is not None           # above the cut: ready before the tail
    assert all(w.grad is None for w in w1[:3])                        # below the cut: untouched so far
    assert torch.allclose(w1[3].grad, ref[3]) and torch.allclose(w1[4].grad, ref[4])
    tail()
    for a, b in zip(w1, ref):
        assert torch.allclose(a.grad, b, atol=1e-7)
    # a second step accumulates onto existing .grad buffers (the flat gradient buffer is zeroed, not dropped)
    for w in w1:
        w.grad.zero_()
    loss, cut = _toy_forward(w1, inp, alias=True)
    cut_backward(loss, list(cut), [w1[3], w1[4]])()
    for a, b in zip(w1, ref):
        assert torch.allclose(a.grad, b, atol=1e-7)
