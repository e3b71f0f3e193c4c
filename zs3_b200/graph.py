"""Semantic-cluster graphs of label maps on the device (csrc/graph.cu, `zs3_label_components`): the GPU counterpart
of construct_adj_mat (zs3/train_context_GMMN_GCNcontext.py:33-102) for the GCN-context generator (config 5)."""
import ctypes as C

import torch

from . import _lib as L


def label_components(labels, h, w, src_index=None, max_nodes=256, want_node_map=False):
    """labels: float CUDA tensor [B, *] (class ids as floats, as the reference keeps them); pixel q of the h x w
    graph grid reads labels[b, src_index[q]] (int32 [h*w]) or labels[b, q].
    Returns (n_nodes [B] int32, node_label [B, max_nodes] int32, node_seed [B, max_nodes] int32,
    adj [B, max_nodes, max_nodes] float32, node_map [B, h*w] int32 or None) -- device tensors, no host sync."""
    if not labels.is_cuda:
        raise RuntimeError("zs3_b200 runs on CUDA (sm_100a) tensors only; there is no CPU path")
    labels = labels.float()
    if labels.dim() > 2:
        labels = labels.reshape(labels.shape[0], -1)
    if labels.stride(1) != 1:
        labels = labels.contiguous()
    if src_index is None and labels.shape[1] != h * w:
        raise ValueError("labels must be [B, h*w] when no src_index is given")
    B, dev = labels.shape[0], labels.device
    n_nodes = torch.empty(B, dtype=torch.int32, device=dev)
    node_label = torch.zeros((B, max_nodes), dtype=torch.int32, device=dev)
    node_seed = torch.zeros((B, max_nodes), dtype=torch.int32, device=dev)
    adj = torch.empty((B, max_nodes, max_nodes), dtype=torch.float32, device=dev)
    node_map = torch.empty((B, h * w), dtype=torch.int32, device=dev) if want_node_map else None
    a = L.ComponentsArgs()
    a.labels, a.image_stride = labels.data_ptr(), labels.stride(0)
    a.src_index = None if src_index is None else src_index.contiguous().data_ptr()
    a.B, a.h, a.w, a.max_nodes = B, h, w, max_nodes
    a.n_nodes, a.node_label, a.node_seed = n_nodes.data_ptr(), node_label.data_ptr(), node_seed.data_ptr()
    a.node_map = None if node_map is None else node_map.data_ptr()
    a.adj = adj.data_ptr()
    L.check(L.lib().zs3_label_components(C.byref(a), L.stream_ptr()), "zs3_label_components")
    return n_nodes, node_label, node_seed, adj, node_map


def construct_adj_mat(segmap, embeddingmap, featmap, max_nodes=256):
    """Batched, device-side construct_adj_mat: segmap [B, h, w] float labels, embeddingmap [B, E, h, w],
    featmap [B, F, h, w] or None (CUDA tensors).  Returns a list with one tuple per image,
    (adj [n, n] dense float32 or None when n == 1, node labels [n], embedding_GCN [n, E], feat_GCN [n, F] or None):
    the same quantities as train_context_GMMN_GCNcontext.py:93-102 (clsidx_2_pixidx is not consumed by the trainer
    and is available as `label_components(..., want_node_map=True)`).  One host sync (the node counts)."""
    B, h, w = segmap.shape
    n_nodes, node_label, node_seed, adj, _ = label_components(segmap.reshape(B, -1), h, w, max_nodes=max_nodes)
    counts = n_nodes.tolist()
    out = []
    for b, n in enumerate(counts):
        if n > max_nodes:
            raise RuntimeError(f"image {b}: {n} clusters exceed max_nodes={max_nodes}")
        seeds = node_seed[b, :n].long()
        emb = embeddingmap[b].reshape(embeddingmap.shape[1], -1)[:, seeds].t().contiguous()
        feat = None if featmap is None else featmap[b].reshape(featmap.shape[1], -1)[:, seeds].t().contiguous()
        out.append((adj[b, :n, :n].contiguous() if n > 1 else None, node_label[b, :n], emb, feat))
    return out
