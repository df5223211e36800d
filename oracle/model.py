"""ORACLE (test infrastructure, never shipped): the reference models restated in plain torch.

Follows ``/root/reference/model.py`` (SAGE :292-333, GCN :386-439, GATv2 :115-234,
custom_GATv2Conv :13-112).  The DGL layers they wrap (``dglnn.SAGEConv('mean')``,
``dglnn.GraphConv(norm='both')``, ``dglnn.GATv2Conv``) are NOT under /root/reference; their
forward is restated from the DGL 2.2.1 semantics in SURVEY.md §8(a12-a14).  PARITY UNPINNED
(see ``oracle/dglops.py``).  Aggregations are ``index_add_`` in the working dtype; autograd
differentiates them, which is the fp32 reference the CUDA kernels' backward is checked against.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import dglops as ops


def _spmm(block, x, w=None):
    """out[i] = Σ_{e→i} w_e · x[src_e]  (g-SpMM u_mul_e / copy_u, sum)."""
    m = x[block.src]
    if w is not None:
        m = m * w.view(-1, *([1] * (x.dim() - 1))).to(x.dtype)
    out = torch.zeros((block.num_dst_nodes(),) + tuple(x.shape[1:]), dtype=x.dtype)
    return out.index_add(0, block.dst, m)


class SAGEConv(nn.Module):
    """dglnn.SAGEConv(in, out, 'mean') with ``edge_weight``."""

    def __init__(self, in_feats, out_feats, aggregator_type="mean"):
        super().__init__()
        assert aggregator_type == "mean"
        self._in, self._out = in_feats, out_feats
        self.fc_neigh = nn.Linear(in_feats, out_feats, bias=False)
        self.fc_self = nn.Linear(in_feats, out_feats, bias=True)
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_uniform_(self.fc_self.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=gain)

    def forward(self, block, feat, edge_weight=None):
        feat_src = feat
        feat_dst = feat[: block.num_dst_nodes()]
        lin_before_mp = self._in > self._out
        h = self.fc_neigh(feat_src) if lin_before_mp else feat_src
        deg = block.in_degrees().clamp(min=1).to(h.dtype)
        neigh = _spmm(block, h, edge_weight) / deg.unsqueeze(-1)
        if not lin_before_mp:
            neigh = self.fc_neigh(neigh)
        return self.fc_self(feat_dst) + neigh


class GraphConv(nn.Module):
    """dglnn.GraphConv(norm='both', allow_zero_in_degree=True) with ``edge_weight`` (a13)."""

    def __init__(self, in_feats, out_feats, activation=None, allow_zero_in_degree=True):
        super().__init__()
        self._in, self._out = in_feats, out_feats
        self.weight = nn.Parameter(torch.empty(in_feats, out_feats))
        self.bias = nn.Parameter(torch.zeros(out_feats))
        nn.init.xavier_uniform_(self.weight)
        self._activation = activation

    def forward(self, block, feat, edge_weight=None):
        feat_src = feat
        norm = block.out_degrees().to(feat.dtype).clamp(min=1).pow(-0.5)
        feat_src = feat_src * norm.unsqueeze(-1)
        if self._in > self._out:
            rst = _spmm(block, feat_src @ self.weight, edge_weight)
        else:
            rst = _spmm(block, feat_src, edge_weight) @ self.weight
        norm = block.in_degrees().to(feat.dtype).clamp(min=1).pow(-0.5)
        rst = rst * norm.unsqueeze(-1) + self.bias
        if self._activation is not None:
            rst = self._activation(rst)
        return rst


class GATv2Conv(nn.Module):
    """custom_GATv2Conv (model.py:13-112) over dglnn.GATv2Conv(share_weights=True, bias=False)."""

    def __init__(self, in_feats, out_feats, num_heads, feat_drop=0.0, attn_drop=0.0, negative_slope=0.2,
                 residual=False, activation=None, allow_zero_in_degree=True, bias=False, share_weights=True):
        super().__init__()
        assert share_weights and not bias
        self._num_heads, self._out_feats = num_heads, out_feats
        self.fc_src = nn.Linear(in_feats, out_feats * num_heads, bias=False)
        self.attn = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.feat_drop = nn.Dropout(feat_drop)
        self.attn_drop = nn.Dropout(attn_drop)
        self.negative_slope = negative_slope
        if residual:
            if in_feats != out_feats * num_heads:
                self.res_fc = nn.Linear(in_feats, num_heads * out_feats, bias=False)
            else:
                self.res_fc = nn.Identity()
        else:
            self.res_fc = None
        self.activation = activation
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_normal_(self.fc_src.weight, gain=gain)
        nn.init.xavier_normal_(self.attn, gain=gain)
        if isinstance(self.res_fc, nn.Linear):
            nn.init.xavier_normal_(self.res_fc.weight, gain=gain)

    def forward(self, block, feat, edge_weight=None, get_attention=False):
        h_src = h_dst = self.feat_drop(feat)                                    # model.py:69
        feat_src = self.fc_src(h_src).view(-1, self._num_heads, self._out_feats)   # :70
        feat_dst = feat_src[: block.number_of_dst_nodes()]                      # :72,78
        h_dst = h_dst[: block.number_of_dst_nodes()]                            # :79
        e = feat_src[block.src] + feat_dst[block.dst]                           # :82 u_add_v
        e = F.leaky_relu(e, self.negative_slope)                                # :83
        e = (e * self.attn).sum(dim=-1).unsqueeze(dim=2)                        # :86
        a = self.attn_drop(ops.edge_softmax(block, e))                          # :88-90
        rst = torch.zeros((block.num_dst_nodes(), self._num_heads, self._out_feats), dtype=feat_src.dtype)
        rst = rst.index_add(0, block.dst, feat_src[block.src] * a)              # :98
        if self.res_fc is not None:
            rst = rst + self.res_fc(h_dst).view(h_dst.shape[0], -1, self._out_feats)   # :101-103
        if self.activation:
            rst = self.activation(rst)                                          # :105-106
        return (rst, e) if get_attention else rst                               # :108-112


class SAGE(nn.Module):
    """model.py:292-333."""

    def __init__(self, in_feats, n_hidden, n_classes, n_layers, activation, dropout):
        super().__init__()
        self.layers = nn.ModuleList()
        if n_layers > 1:
            self.layers.append(SAGEConv(in_feats, n_hidden, "mean"))
            for _ in range(1, n_layers - 1):
                self.layers.append(SAGEConv(n_hidden, n_hidden, "mean"))
            self.layers.append(SAGEConv(n_hidden, n_classes, "mean"))
        else:
            self.layers.append(SAGEConv(in_feats, n_classes, "mean"))
        self.dropout = nn.Dropout(dropout)
        self.activation = activation

    def forward(self, blocks, x):
        h = x
        for l, (layer, block) in enumerate(zip(self.layers, blocks)):
            block.srcdata["embed_norm"] = torch.reshape(torch.norm(h, dim=1, keepdim=True), (-1,))   # :318
            h = layer(block, h, edge_weight=block.edata.get("edge_weights"))    # :321-329
            if l < len(self.layers) - 1:
                h = self.dropout(self.activation(h))                            # :330-332
        return h


class GCN(nn.Module):
    """model.py:386-439."""

    def __init__(self, in_feats, n_hidden, n_classes, n_layers, activation, dropout):
        super().__init__()
        self.layers = nn.ModuleList()
        if n_layers > 1:
            self.layers.append(GraphConv(in_feats, n_hidden, activation=activation))
            for _ in range(1, n_layers - 1):
                self.layers.append(GraphConv(n_hidden, n_hidden, activation=activation))
            self.layers.append(GraphConv(n_hidden, n_classes))
        else:
            self.layers.append(GraphConv(in_feats, n_classes))
        self.dropout = nn.Dropout(dropout)

    def forward(self, blocks, x):
        h = x
        for l, (layer, block) in enumerate(zip(self.layers, blocks)):
            block.srcdata["embed_norm"] = torch.reshape(torch.norm(h, dim=1, keepdim=True), (-1,))   # :425
            h = layer(block, h, edge_weight=block.edata.get("edge_weights"))
            if l < len(self.layers) - 1:
                h = self.dropout(h)                                             # :437-438
        return h


class GATv2(nn.Module):
    """model.py:115-234."""

    def __init__(self, num_layers, in_dim, num_hidden, num_classes, heads, activation, feat_drop,
                 attn_drop, negative_slope, residual):
        super().__init__()
        self.gatv2_layers = nn.ModuleList()
        mk = lambda i, o, h, res, act: GATv2Conv(i, o, h, feat_drop, attn_drop, negative_slope, res, act,
                                                 bias=False, share_weights=True)
        if num_layers > 1:
            self.gatv2_layers.append(mk(in_dim, num_hidden, heads[0], False, activation))
            for l in range(1, num_layers - 1):
                self.gatv2_layers.append(mk(num_hidden * heads[l - 1], num_hidden, heads[l], residual, activation))
            self.gatv2_layers.append(mk(num_hidden * heads[-2], num_classes, heads[-1], residual, None))
        else:
            self.gatv2_layers.append(mk(in_dim, num_classes, heads[-1], residual, None))

    def forward(self, blocks, inputs):
        h = inputs
        for l, block in enumerate(blocks):
            block.srcdata["embed_norm"] = torch.reshape(torch.norm(h, dim=1, keepdim=True), (-1,))   # :211
            h, a = self.gatv2_layers[l](block, h, edge_weight=block.edata.get("edge_weights"),
                                        get_attention=True)                     # :214-223
            block.edata["a_ij"] = torch.mean(a.squeeze(dim=-1), dim=1)          # :224-227
            h = h.flatten(1) if l < len(blocks) - 1 else h.mean(1)              # :228-232
        return h
