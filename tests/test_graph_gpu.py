"""zs3_label_components (csrc/graph.cu) on the GPU: bit-exact against the reference construct_adj_mat golden vectors
(tests/golden/graph.npz) and, at the full config-5 size (batch of 129x129 maps), against the oracle."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))

pytestmark = pytest.mark.gpu


def test_components_match_reference_golden():
    from zs3_b200.graph import label_components
    gold = np.load(os.path.join(HERE, "golden", "graph.npz"))
    for name in sorted({k.split("/")[0] for k in gold.files}):
        seg = torch.from_numpy(gold[name + "/seg"]).cuda()
        h, w = seg.shape
        n_nodes, node_label, node_seed, adj, node_map = label_components(seg.reshape(1, -1), h, w, max_nodes=64,
                                                                         want_node_map=True)
        n = int(n_nodes[0])
        assert n == len(gold[name + "/node_label"]), name
        assert np.array_equal(node_label[0, :n].cpu().numpy(), gold[name + "/node_label"].astype(np.int32)), name
        assert np.array_equal(node_seed[0, :n].cpu().numpy(), gold[name + "/node_seed"]), name
        assert np.array_equal(node_map[0].cpu().numpy().reshape(h, w), gold[name + "/node_map"]), name
        a = adj[0].cpu().numpy()
        assert np.array_equal(a[:n, :n], gold[name + "/adj"]), name
        assert a[n:].sum() == 0 and a[:, n:].sum() == 0


def _voronoi_batch(B, hw, seed):
    rng = np.random.RandomState(seed)
    out = np.zeros((B, hw, hw), dtype=np.float32)
    yy, xx = np.mgrid[0:hw, 0:hw]
    for b in range(B):
        k = rng.randint(5, 40)
        sites = rng.randint(0, hw, size=(k, 2))
        cls = rng.randint(0, rng.randint(4, 11), size=k)
        d = (yy[..., None] - sites[:, 0]) ** 2 + (xx[..., None] - sites[:, 1]) ** 2
        out[b] = cls[d.argmin(-1)]
        out[b][rng.rand(hw, hw) < 0.002] = 255
    return out


def test_components_full_size_batch_vs_oracle_and_wrapper():
    """config 5: batch of 8 label maps at the 129x129 feature resolution, 5-40 Voronoi cells over 4-10 classes
     + isolated ignore pixels; every output bit-exact; construct_adj_mat wrapper gathers seeds"""
    import zs3_graph_oracle as GO
    from zs3_b200.graph import construct_adj_mat, label_components
    B, hw = 8, 129
    seg = _voronoi_batch(B, hw, seed=7)
    segc = torch.from_numpy(seg).cuda()
    n_nodes, node_label, node_seed, adj, node_map = label_components(segc.reshape(B, -1), hw, hw, max_nodes=256,
                                                                     want_node_map=True)
    torch.cuda.synchronize()
    for b in range(B):
        nm, nl, ns, ad = GO.cluster_graph(seg[b])
        n = len(nl)
        assert int(n_nodes[b]) == n
        assert np.array_equal(node_map[b].cpu().numpy().reshape(hw, hw), nm)
        assert np.array_equal(node_label[b, :n].cpu().numpy(), nl.astype(np.int32))
        assert np.array_equal(node_seed[b, :n].cpu().numpy(), ns)
        assert np.array_equal(adj[b, :n, :n].cpu().numpy(), ad)
    g = torch.Generator().manual_seed(1)
    emb = torch.randn(B, 12, hw, hw, generator=g).cuda()
    feat = torch.randn(B, 9, hw, hw, generator=g).cuda()
    res = construct_adj_mat(segc, emb, feat)
    for b, (a, lbl, e, f) in enumerate(res):
        nm, nl, ns, ad = GO.cluster_graph(seg[b])
        assert np.array_equal(a.cpu().numpy(), ad)
        assert torch.equal(e.cpu(), emb[b].reshape(12, -1)[:, torch.from_numpy(ns).long().cuda()].t().cpu())
        assert torch.equal(f.cpu(), feat[b].reshape(9, -1)[:, torch.from_numpy(ns).long().cuda()].t().cpu())
    # idempotence-style property at full size: relabelling the node map reproduces the same graph
    again = label_components(node_map.float(), hw, hw, max_nodes=256, want_node_map=True)
    assert torch.equal(again[0], n_nodes) and torch.equal(again[4], node_map) and torch.equal(again[3], adj)


def test_components_nearest_gather_and_gcn_generator():
    """labels gathered from the full-resolution map through the nearest source index; the dense adjacency feeds
    GMMNnetwork_GCN (zs3/modeling/gmmn.py:52-67) and matches the oracle's graph convolution"""
    import zs3_graph_oracle as GO
    import zs3_oracle as O
    from zs3_b200.graph import label_components
    from zs3_b200.modeling.gmmn import GMMNnetwork_GCN
    full = torch.from_numpy(_voronoi_batch(2, 513, seed=11))
    idx = torch.arange(513 * 513, dtype=torch.float32).view(1, 1, 513, 513)
    src = torch.nn.functional.interpolate(idx, size=(129, 129), mode="nearest").view(-1).to(torch.int32)
    small = torch.nn.functional.interpolate(full[:, None], size=(129, 129), mode="nearest")[:, 0]
    n_nodes, node_label, node_seed, adj, _ = label_components(full.cuda().reshape(2, -1), 129, 129, src_index=src.cuda())
    for b in range(2):
        nm, nl, ns, ad = GO.cluster_graph(small[b].numpy())
        n = len(nl)
        assert int(n_nodes[b]) == n and np.array_equal(adj[b, :n, :n].cpu().numpy(), ad)
        assert np.array_equal(node_seed[b, :n].cpu().numpy(), ns)
    n = int(n_nodes[0])
    gen = GMMNnetwork_GCN().cuda().eval()
    g = torch.Generator().manual_seed(5)
    emb, z = torch.randn(n, 300, generator=g) * 0.06, torch.rand(n, 300, generator=g)
    out = gen(emb.cuda(), z.cuda(), adj[0, :n, :n].contiguous())
    st = {k: v.detach().cpu() for k, v in gen.state_dict().items()}
    ref = O.gmmn_gcn_forward(st, emb, z, adj[0, :n, :n].cpu())
    assert torch.allclose(out.cpu(), ref, rtol=1e-4, atol=1e-5)
