"""GMMN generators with the reference's module surface (zs3/modeling/gmmn.py) on the fp32 CUDA kernels."""
import torch
from torch import nn

from .. import gmmn_ops as G


def _require_cuda(t):
    if not t.is_cuda:
        raise RuntimeError("zs3_b200 runs on CUDA (sm_100a) tensors only; there is no CPU path")


def _run_sequential(model, x, keep_mask=None):
    """Executes an nn.Sequential of Linear / LeakyReLU / Dropout holders with the fused kernels."""
    mods = list(model) if isinstance(model, nn.Sequential) else [model]
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.Linear):
            x = G.Linear.apply(x, m.weight, m.bias)
            i += 1
        elif isinstance(m, nn.LeakyReLU):
            p, training = 0.0, False
            if i + 1 < len(mods) and isinstance(mods[i + 1], nn.Dropout):
                p, training = mods[i + 1].p, mods[i + 1].training
                i += 1
            x = G.LeakyDropout.apply(x, m.negative_slope, p, training, keep_mask)
            i += 1
        else:
            raise NotImplementedError(type(m))
    return x


def _xavier_(root, kind=nn.Linear):
    """xavier-uniform weights, bias 0.01 (gmmn.py:24-27, 62-65); the reconstruction layer keeps torch's default"""
    for m in root.modules():
        if type(m) is kind:
            nn.init.xavier_uniform_(m.weight)
            nn.init.constant_(m.bias, 0.01)


class GMMNnetwork(nn.Module):
    """zs3/modeling/gmmn.py:6-49; parameters live under model.0 / model.3 like the reference's nn.Sequential."""

    def __init__(self, noise_dim, embed_dim, hidden_size, feature_dim, semantic_reconstruction=False):
        super().__init__()

        width_in = noise_dim + embed_dim
        if hidden_size:   # one hidden layer: Linear -> LeakyReLU(0.2) -> Dropout(0.5) -> Linear  (model.0 / model.3)
            layers = [nn.Linear(width_in, hidden_size), nn.LeakyReLU(0.2, inplace=True), nn.Dropout(p=0.5),
                      nn.Linear(hidden_size, feature_dim)]
            self.model = nn.Sequential(*layers)
        else:             # linear generator
            self.model = nn.Linear(width_in, feature_dim)
        _xavier_(self.model)
        self.semantic_reconstruction = semantic_reconstruction
        if self.semantic_reconstruction:
            self.semantic_reconstruction_layer = nn.Linear(feature_dim, noise_dim + embed_dim)

    def forward(self, embd, noise, keep_mask=None):
        """keep_mask: optional uint8 [n, hidden] Dropout keep mask (parity tests); None = counter-based RNG"""
        _require_cuda(embd)
        x = G.Concat2.apply(embd, noise)
        features = _run_sequential(self.model, x, keep_mask)
        if self.semantic_reconstruction:
            m = self.semantic_reconstruction_layer
            return features, G.Linear.apply(features, m.weight, m.bias)
        return features


class GraphConvolution(nn.Module):
    """Dense restatement of pygcn.layers.GraphConvolution (tkipf/pygcn; not vendored by the reference, see
    SURVEY.md 8c "parity unpinned"): output = adj @ (x @ W) + b with W [in, out]."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(in_features, out_features))
        self.bias = nn.Parameter(torch.empty(out_features)) if bias else None
        stdv = 1.0 / (out_features ** 0.5)
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    def forward(self, x, adj):
        if adj.is_sparse:
            adj = adj.to_dense()
        support = G.Linear.apply(x, self.weight.t(), None)            # x @ W
        return G.Linear.apply(adj, support.t(), self.bias)            # adj @ support + b


class GMMNnetwork_GCN(nn.Module):
    """zs3/modeling/gmmn.py:52-67"""

    def __init__(self, noise_dim=300, embed_dim=300, hidden_size=256, feature_dim=256):
        super().__init__()
        self.gcn1 = GraphConvolution(noise_dim + embed_dim, hidden_size)
        self.relu = nn.LeakyReLU(0.2)
        self.dropout = nn.Dropout(p=0.5)
        self.gcn2 = GraphConvolution(hidden_size, feature_dim)
        _xavier_(self, GraphConvolution)

    def forward(self, embd, noise, adj_mat, keep_mask=None):
        _require_cuda(embd)
        x = self.gcn1(G.Concat2.apply(embd, noise), adj_mat)
        x = G.LeakyDropout.apply(x, self.relu.negative_slope, self.dropout.p, self.dropout.training, keep_mask)
        return self.gcn2(x, adj_mat)
