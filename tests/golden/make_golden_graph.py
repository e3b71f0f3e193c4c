"""Generates tests/golden/graph.npz by running the REAL construct_adj_mat of the reference
(/root/reference/zs3/train_context_GMMN_GCNcontext.py:33-102) on synthetic label maps.

Run in the build container only:  python tests/golden/make_golden_graph.py
The trainer module cannot be imported here (tensorboardX / matplotlib are not installed), so the two function
definitions it needs are compiled from the reference file's syntax tree -- executed from where they lie, not copied.
"""
import ast
import itertools
import os

import numpy as np
import scipy.sparse as sp
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_FILE = "/root/reference/zs3/train_context_GMMN_GCNcontext.py"


def load_reference_function():
    tree = ast.parse(open(REF_FILE).read())
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef)
              and n.name in ("construct_adj_mat", "sparse_mx_to_torch_sparse_tensor")]
    assert len(wanted) == 2
    ns = {"np": np, "itertools": itertools, "sp": sp, "torch": torch}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), REF_FILE, "exec"), ns)
    return ns["construct_adj_mat"]


def label_maps():
    rng = np.random.RandomState(1)
    maps = {}
    # Voronoi cells of 14 sites over 6 classes at the feature resolution of a 513x513 crop, 255 on cell borders
    h = w = 129
    sites = rng.randint(0, h, size=(14, 2))
    cls = rng.randint(0, 6, size=14)
    yy, xx = np.mgrid[0:h, 0:w]
    d = (yy[..., None] - sites[:, 0]) ** 2 + (xx[..., None] - sites[:, 1]) ** 2
    near = d.argsort(-1)
    m = cls[near[..., 0]].astype(np.float32)
    border = np.abs(np.sqrt(np.take_along_axis(d, near[..., :1], -1)) - np.sqrt(np.take_along_axis(d, near[..., 1:2], -1)))[..., 0] < 0.8
    m[border] = 255
    maps["voronoi129"] = m
    # blocky 33x33 map
    g = rng.randint(0, 4, size=(5, 5))
    maps["blocks33"] = np.kron(g, np.ones((7, 7)))[:33, :33].astype(np.float32)
    # a single cluster: adj_mat is None in the reference
    maps["single9"] = np.full((9, 9), 3, dtype=np.float32)
    # diagonal contacts only: 8-connectivity merges the diagonal runs of a checkerboard into two clusters
    maps["checker8"] = ((np.add.outer(np.arange(8), np.arange(8)) % 2) * 7).astype(np.float32)
    # salt noise: many single-pixel clusters
    m = np.zeros((21, 17), dtype=np.float32)
    m[rng.rand(21, 17) < 0.15] = 255
    m[10:, 9:] += 1
    maps["noise21x17"] = m
    return maps


def main():
    fn = load_reference_function()
    out = {}
    for name, seg in label_maps().items():
        h, w = seg.shape
        emb = np.random.RandomState(2).randn(5, h, w).astype(np.float32)
        feat = np.random.RandomState(3).randn(7, h, w).astype(np.float32)
        adj, pix, lbl, emb_gcn, feat_gcn = fn(seg, emb, feat)
        n = len(lbl)
        seeds = np.array([pix[k][0][0] * w + pix[k][0][1] for k in range(n)], dtype=np.int32)
        node_map = np.zeros((h, w), dtype=np.int32)
        for k in range(n):
            for (i, j) in pix[k]:
                node_map[i, j] = k
        assert np.array_equal(emb_gcn, emb.reshape(5, -1)[:, seeds].T)       # node embedding = seed pixel's
        assert np.array_equal(feat_gcn, feat.reshape(7, -1)[:, seeds].T)     # node feature = seed pixel's
        dense = np.zeros((n, n), dtype=np.float32) if adj is None else adj.to_dense().numpy()
        assert (adj is None) == (n == 1)
        out[name + "/seg"] = seg
        out[name + "/node_label"] = np.array(lbl, dtype=np.float32)
        out[name + "/node_seed"] = seeds
        out[name + "/node_map"] = node_map
        out[name + "/adj"] = dense
        print(name, seg.shape, "clusters", n, "edges", int(dense.sum()) // 2)
    np.savez_compressed(os.path.join(HERE, "graph.npz"), **out)


if __name__ == "__main__":
    main()
