"""Drop-in models: ``SAGE`` / ``GCN`` / ``GATv2`` with the reference constructor signatures
(``model.py:292-310,386-419,115-205``) over the sm_100a aggregation kernels.

The DGL layers the reference instantiates — ``dglnn.SAGEConv(in, out, 'mean')``,
``dglnn.GraphConv(norm='both', allow_zero_in_degree=True)`` and ``dglnn.GATv2Conv`` subclassed as
``custom_GATv2Conv`` (``model.py:13-112``) — are re-implemented here with DGL-compatible
parameter names (``fc_neigh`` / ``fc_self`` / ``weight`` / ``bias`` / ``fc_src`` / ``attn`` /
``res_fc``) so reference checkpoints' state dicts line up.  The dense projections stay torch
GEMMs; message passing, edge softmax and ``embed_norm`` are the custom kernels (``ops.py``).
"""
from __future__ import annotations

import contextlib
import os

import torch
import torch.nn as nn

from . import ops
from .graph import Graph


class _LinearSplitK(torch.autograd.Function):
    """``F.linear(x, pad(weight), bias)`` whose weight gradient ``G^T X`` (an [out, in] result reduced over
    thousands of rows: a handful of output tiles for 148 SMs) is computed as a batched GEMM over row chunks;
    the chunk partials are added in chunk order (deterministic) by ``bliss_splitk_accumulate`` straight into
    ``weight.grad`` when that buffer exists (the flat gradient buffer of ``parallel.FlatGrads``), else summed
    and returned.  ``weight`` is the un-padded parameter; ``x`` may carry zero-padded extra columns."""

    SPLIT, MIN_ROWS = int(os.environ.get("BLISS_SPLITK", "4")), 2048
    #: the self projection's weight gradient on its side stream, beside the neighbour projection's (ablation switch)
    WGRAD_OVERLAP = os.environ.get("BLISS_WGRAD_OVERLAP", "1") == "1"

    @staticmethod
    def forward(ctx, x, weight, bias, side=None):
        # ``side``: run the forward GEMM on that stream (the caller has made it wait for the inputs and joins it
        # before using the result).  Autograd records the CALLER's stream for this node, so the backward GEMMs run
        # on the main stream: beside the backward SpMM they would only fight it for SMs (measured: 79 us instead
        # of 50 for the aggregation, 82 instead of 14 for the GEMM).
        extra = x.shape[1] - weight.shape[1]
        stored = getattr(weight, "_bliss_padded", None)      # the parameter's own storage, padded (parallel.flat_layout)
        if extra and (stored is None or stored.shape[1] != x.shape[1]):
            stored = None
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
            w = weight if extra == 0 else (stored if stored is not None else torch.nn.functional.pad(weight, (0, extra)))
            y = torch.nn.functional.linear(x, w, bias)
        if side is not None and w is not weight and w is not stored:
            w.record_stream(torch.cuda.current_stream())     # allocated on the side stream, read by backward on this one
        ctx.save_for_backward(x, w)
        ctx.weight, ctx.has_bias, ctx.side = weight, bias is not None, side
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        weight = ctx.weight
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = gy @ w
        side = ctx.side if (_LinearSplitK.WGRAD_OVERLAP and ctx.needs_input_grad[1]) else None
        ev = getattr(ops, "LAST_SPMM_BWD", None) if side is not None else None
        if side is not None and ev is not None:
            # this node runs last in its layer's backward pass (its forward was issued first): its weight gradient only
            # needs the epilogue's gradient, so it goes to the side stream behind the layer's backward aggregation and
            # runs beside the neighbour projection's weight gradient; the backward pass joins the stream when it ends
            side.wait_event(ev)
            torch.autograd.Variable._execution_engine.queue_callback(lambda: torch.cuda.current_stream().wait_stream(side))
        else:
            side = None
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
            gw = _LinearSplitK._wgrad(ctx, x, gy, weight)
        if side is not None:
            gy.record_stream(side)
            x.record_stream(side)
            if gw is not None:
                torch.cuda.current_stream().wait_stream(side)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gy.sum(0)
        return gx, gw, gb, None

    @staticmethod
    def _wgrad(ctx, x, gy, weight):
        gw = None
        if ctx.needs_input_grad[1]:
            S = _LinearSplitK.SPLIT
            k = (x.shape[0] // S) * S
            gyc, xc = gy.contiguous(), x.contiguous()
            n_out, n_in_pad, n_in = gyc.shape[1], xc.shape[1], weight.shape[1]
            rem = k < x.shape[0]
            part = torch.empty((S + int(rem), n_out, n_in_pad), dtype=gyc.dtype, device=gyc.device)
            torch.bmm(gyc[:k].view(S, k // S, -1).transpose(1, 2), xc[:k].view(S, k // S, -1), out=part[:S])
            if rem:
                torch.mm(gyc[k:].t(), xc[k:], out=part[S])
            g = weight.grad
            gp = getattr(weight, "_bliss_padded_grad", None)
            if g is not None and gp is not None and gp.shape[1] == n_in_pad and g.data_ptr() == gp.data_ptr():
                # the gradient lives in padded storage too: accumulate whole padded rows (the extra columns add zeros)
                ops.N.call("bliss_splitk_accumulate", ops.N.ptr(part), part.shape[0], n_out, n_in_pad, n_in_pad,
                           ops.N.ptr(gp), ops.N.stream())
            elif g is not None and g.is_contiguous() and g.dtype == torch.float32 and g.is_cuda:
                ops.N.call("bliss_splitk_accumulate", ops.N.ptr(part), part.shape[0], n_out, n_in_pad, n_in,
                           ops.N.ptr(g), ops.N.stream())       # accumulated in place: nothing to return
            else:
                gw = part.sum(0)[:, :n_in]
        return gw


def _linear(x, lin: nn.Linear, use_bias: bool = True, side=None):
    """``lin(x)``; when the feature table's rows were zero-padded to a 16-byte multiple
    (``train.DataModule(pad_features=True)``: 602 -> 604 columns keeps cuBLAS off its unaligned
    kernels, ~2-3x on the input-layer GEMMs) the weight is zero-padded to match — same result."""
    extra = x.shape[-1] - lin.in_features
    bias = lin.bias if use_bias else None
    if x.is_cuda and x.dim() == 2 and x.shape[0] >= _LinearSplitK.MIN_ROWS and torch.is_grad_enabled():
        return _LinearSplitK.apply(x, lin.weight, bias, side)
    with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
        w = lin.weight if extra == 0 else torch.nn.functional.pad(lin.weight, (0, extra))
        return torch.nn.functional.linear(x, w, bias)


_SIDE_STREAMS = {}


def _side_stream(device):
    s = _SIDE_STREAMS.get(device)
    if s is None:
        s = _SIDE_STREAMS[device] = torch.cuda.Stream(device)
    return s


class SAGEConv(nn.Module):
    """``dglnn.SAGEConv(in_feats, out_feats, 'mean')`` with ``edge_weight``."""

    def __init__(self, in_feats, out_feats, aggregator_type="mean", feat_drop=0.0, bias=True):
        super().__init__()
        if aggregator_type != "mean":
            raise NotImplementedError("the reference only builds SAGEConv(..., 'mean') (model.py:303-308)")
        self._in_src_feats = self._in_dst_feats = in_feats
        self._out_feats = out_feats
        self.feat_drop = nn.Dropout(feat_drop)
        self.fc_neigh = nn.Linear(in_feats, out_feats, bias=False)
        self.fc_self = nn.Linear(in_feats, out_feats, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_uniform_(self.fc_self.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=gain)

    def forward_parts(self, graph, feat, edge_weight=None, use_bias=False):
        """The two summands of the layer output: ``(fc_self(h_dst) [without its bias], h_neigh)``."""
        feat_src = self.feat_drop(feat)
        feat_dst = feat_src[: graph.number_of_dst_nodes()]
        lin_before_mp = self._in_src_feats > self._out_feats
        # the self projection does not depend on the aggregation: on CUDA it runs on a side stream beside the
        # (L2-bandwidth-bound) SpMM, and autograd replays its weight-gradient GEMM on that stream too
        side = _side_stream(feat.device) if feat.is_cuda else None
        if side is not None:
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            h_self = _linear(feat_dst, self.fc_self, use_bias=use_bias, side=side)
        h = _linear(feat_src, self.fc_neigh) if lin_before_mp else feat_src
        h_neigh = ops.spmm(graph, h, edge_weight, dst_scale=ops.mean_scale(graph))   # u_mul_e + fn.mean
        if not lin_before_mp:
            h_neigh = _linear(h_neigh, self.fc_neigh)
        if side is not None:
            main.wait_stream(side)
            h_self.record_stream(main)
        else:
            h_self = _linear(feat_dst, self.fc_self, use_bias=use_bias)
        return h_self, h_neigh

    def forward(self, graph, feat, edge_weight=None):
        h_self, h_neigh = self.forward_parts(graph, feat, edge_weight, use_bias=True)
        return h_self + h_neigh


class GraphConv(nn.Module):
    """``dglnn.GraphConv(in, out, norm='both', activation=..., allow_zero_in_degree=True)`` (a13)."""

    def __init__(self, in_feats, out_feats, norm="both", weight=True, bias=True, activation=None,
                 allow_zero_in_degree=False):
        super().__init__()
        if norm != "both" or not weight:
            raise NotImplementedError("the reference only builds GraphConv(norm='both') with a weight")
        self._in_feats, self._out_feats = in_feats, out_feats
        self.weight = nn.Parameter(torch.empty(in_feats, out_feats))
        self.bias = nn.Parameter(torch.empty(out_feats)) if bias else None
        self._activation = activation
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, graph, feat, weight=None, edge_weight=None):
        src_norm = getattr(graph, "_gcn_src_norm", None)
        if src_norm is None:
            if getattr(graph, "csc_pos", None) is not None:   # sampled block: out-degrees from the transpose (no host sync)
                t_indptr = ops.block_transpose(graph)[0]
                out_deg = t_indptr[1:] - t_indptr[:-1]
            else:                                             # whole-graph inference block
                out_deg = graph.out_degrees()
            src_norm = out_deg.to(torch.float32).clamp(min=1).pow(-0.5)
            dst_norm = graph.in_degrees().to(torch.float32).clamp(min=1).pow(-0.5)
            if graph.num_src_nodes() == src_norm.numel() and not getattr(graph, "_static_padded", False):
                graph._gcn_src_norm, graph._gcn_dst_norm = src_norm, dst_norm
        else:
            dst_norm = graph._gcn_dst_norm
        w = self.weight
        if feat.shape[-1] > self._in_feats:       # zero-padded feature rows (see _linear)
            w = torch.nn.functional.pad(w, (0, 0, 0, feat.shape[-1] - self._in_feats))
        if self._in_feats > self._out_feats:
            rst = ops.spmm(graph, torch.matmul(feat, w), edge_weight, src_scale=src_norm, dst_scale=dst_norm)
        else:
            rst = torch.matmul(ops.spmm(graph, feat, edge_weight, src_scale=src_norm, dst_scale=dst_norm), w)
        if self.bias is not None:
            rst = rst + self.bias
        if self._activation is not None:
            rst = self._activation(rst)
        return rst


class custom_GATv2Conv(nn.Module):
    """``custom_GATv2Conv`` (``model.py:13-112``) — GATv2 layer with shared projection, no bias,
    returning the **pre-softmax logits** as attention (``:108-110``); ``edge_weight`` accepted and
    ignored like the reference (``:91-96`` commented out)."""

    def __init__(self, in_feats, out_feats, num_heads, feat_drop=0.0, attn_drop=0.0, negative_slope=0.2,
                 residual=False, activation=None, allow_zero_in_degree=False, bias=True, share_weights=False):
        super().__init__()
        if not share_weights or bias:
            raise NotImplementedError("the reference only builds GATv2Conv(bias=False, share_weights=True)")
        self._num_heads, self._in_src_feats, self._out_feats = num_heads, in_feats, out_feats
        self._allow_zero_in_degree = allow_zero_in_degree
        self.fc_src = nn.Linear(in_feats, out_feats * num_heads, bias=False)
        self.fc_dst = self.fc_src
        self.attn = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.feat_drop = nn.Dropout(feat_drop)
        self.attn_drop = nn.Dropout(attn_drop)
        self.negative_slope = negative_slope
        if residual:
            self.res_fc = (nn.Linear(in_feats, num_heads * out_feats, bias=False)
                           if in_feats != out_feats * num_heads else nn.Identity())
        else:
            self.res_fc = None
        self.activation = activation
        self.share_weights = share_weights
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_normal_(self.fc_src.weight, gain=gain)
        nn.init.xavier_normal_(self.attn, gain=gain)
        if isinstance(self.res_fc, nn.Linear):
            nn.init.xavier_normal_(self.res_fc.weight, gain=gain)

    def forward(self, graph, feat, edge_weight=None, get_attention=False):
        if not self._allow_zero_in_degree and bool((graph.in_degrees() == 0).any()):
            raise RuntimeError("There are 0-in-degree nodes in the graph (model.py:49-61)")
        h_src = h_dst = self.feat_drop(feat)                                              # :69
        feat_src = _linear(h_src, self.fc_src).view(-1, self._num_heads, self._out_feats)  # :70
        h_dst = h_dst[: graph.number_of_dst_nodes()]                                      # :79
        mask = None
        if self.training and self.attn_drop.p > 0:                                        # :88 attn_drop
            keep = 1.0 - self.attn_drop.p
            mask = (torch.rand(graph.num_edges(), self._num_heads, device=feat.device) < keep).float() / keep
        rst, e = ops.gatv2_attention(graph, feat_src, self.attn, self.negative_slope, mask)   # :80-98
        if self.res_fc is not None:
            res = _linear(h_dst, self.res_fc) if isinstance(self.res_fc, nn.Linear) else self.res_fc(h_dst)
            rst = rst + res.view(h_dst.shape[0], -1, self._out_feats)                     # :101-103
        if self.activation:
            rst = self.activation(rst)                                                    # :105-106
        return (rst, e.unsqueeze(-1)) if get_attention else rst                           # :108-112


def _edge_weights(block):
    return block.edata["edge_weights"] if "edge_weights" in block.edata else None


class SAGE(nn.Module):
    """``model.py:292-333``."""

    def __init__(self, in_feats, n_hidden, n_classes, n_layers, activation, dropout):
        super().__init__()
        self.init(in_feats, n_hidden, n_classes, n_layers, activation, dropout)

    def init(self, in_feats, n_hidden, n_classes, n_layers, activation, dropout):
        self.n_layers, self.n_hidden, self.n_classes = n_layers, n_hidden, n_classes
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
        h, norm = x, None
        fuse = self.activation is torch.nn.functional.relu and x.is_cuda
        if fuse and self.training and self.dropout.p > 0 and not getattr(self, "_external_drop_step", False):
            self._drop_step(x.device).add_(1)           # one Philox step per forward pass (device scalar: replayable)
        for l, (layer, block) in enumerate(zip(self.layers, blocks)):
            if l == 0 and norm is None:
                norm = getattr(x, "_bliss_row_norm", None)       # came with the feature gather (train._padded_fwd_bwd)
            block.srcdata["embed_norm"] = ops.row_norm(h) if norm is None else norm       # :318
            norm = None
            last = l == len(self.layers) - 1
            if not last and fuse and layer._out_feats % 4 == 0 and layer._out_feats <= 1024:
                # bias + relu + dropout (:330-332) and the next layer's embed_norm in one launch
                h_self, h_neigh = layer.forward_parts(block, h, edge_weight=_edge_weights(block))   # :321-329
                p = self.dropout.p if self.training else 0.0
                h, norm = ops.sage_epilogue(h_self, h_neigh, layer.fc_self.bias, True, p, self._drop_seed,
                                            self._drop_step(x.device) if p > 0 else None, l, True)
                continue
            h = layer(block, h, edge_weight=_edge_weights(block))                         # :321-329
            if not last:
                h = self.dropout(self.activation(h))                                      # :330-332
        return h

    _drop_seed = 0x5EED

    def _drop_step(self, device):
        t = getattr(self, "_drop_step_t", None)
        if t is None or t.device != device:
            t = torch.zeros(1, dtype=torch.int64, device=device)
            self._drop_step_t = t
        return t

    def inference(self, g, device, batch_size, use_uva=False, num_workers=0):
        """``model.py:335-383``: layer-wise full-neighbour inference (no sampling, no edge weights).
        The whole graph's CSC is one destination-major CSR, so each layer is a single SpMM."""
        return _full_inference(self, g, lambda l, h: self.dropout(self.activation(h)))


class GCN(nn.Module):
    """``model.py:386-439``."""

    def __init__(self, in_feats, n_hidden, n_classes, n_layers, activation, dropout):
        super().__init__()
        self.init(in_feats, n_hidden, n_classes, n_layers, activation, dropout)

    def init(self, in_feats, n_hidden, n_classes, n_layers, activation, dropout):
        self.n_layers, self.n_hidden, self.n_classes = n_layers, n_hidden, n_classes
        self.layers = nn.ModuleList()
        if n_layers > 1:
            self.layers.append(GraphConv(in_feats, n_hidden, activation=activation, allow_zero_in_degree=True))
            for _ in range(1, n_layers - 1):
                self.layers.append(GraphConv(n_hidden, n_hidden, activation=activation, allow_zero_in_degree=True))
            self.layers.append(GraphConv(n_hidden, n_classes, allow_zero_in_degree=True))
        else:
            self.layers.append(GraphConv(in_feats, n_classes, allow_zero_in_degree=True))
        self.dropout = nn.Dropout(dropout)
        self.activation = activation

    def forward(self, blocks, x):
        h = x
        for l, (layer, block) in enumerate(zip(self.layers, blocks)):
            block.srcdata["embed_norm"] = ops.row_norm(h)                             # :425
            h = layer(block, h, edge_weight=_edge_weights(block))                         # :428-436
            if l < len(self.layers) - 1:
                h = self.dropout(h)                                                       # :437-438
        return h

    def inference(self, g, device, batch_size, use_uva=False, num_workers=0):
        """``model.py:441-488``."""
        return _full_inference(self, g, lambda l, h: self.dropout(h))


class GATv2(nn.Module):
    """``model.py:115-234``."""

    def __init__(self, num_layers, in_dim, num_hidden, num_classes, heads, activation, feat_drop, attn_drop,
                 negative_slope, residual):
        super().__init__()
        self.num_layers, self.num_hidden, self.num_classes, self.heads = num_layers, num_hidden, num_classes, heads
        self.activation = activation
        self.gatv2_layers = nn.ModuleList()
        mk = lambda i, o, h, res, act: custom_GATv2Conv(                                  # noqa: E731
            i, o, h, feat_drop, attn_drop, negative_slope, res, act, bias=False, share_weights=True,
            allow_zero_in_degree=True)
        if num_layers > 1:
            self.gatv2_layers.append(mk(in_dim, num_hidden, heads[0], False, self.activation))
            for l in range(1, num_layers - 1):
                self.gatv2_layers.append(mk(num_hidden * heads[l - 1], num_hidden, heads[l], residual, self.activation))
            self.gatv2_layers.append(mk(num_hidden * heads[-2], num_classes, heads[-1], residual, None))
        else:
            self.gatv2_layers.append(mk(in_dim, num_classes, heads[-1], residual, None))

    def forward(self, blocks, inputs):
        h = inputs
        for l, block in enumerate(blocks):
            block.srcdata["embed_norm"] = ops.row_norm(h)                             # :211
            h, a = self.gatv2_layers[l](block, h, edge_weight=_edge_weights(block), get_attention=True)
            block.edata["a_ij"] = torch.mean(a.squeeze(dim=-1), dim=1)                    # :224-227
            h = h.flatten(1) if l < len(blocks) - 1 else h.mean(1)                        # :228-232
        return h

    def inference(self, g, device, batch_size, use_uva=False, num_workers=0):
        """``model.py:236-289``."""
        blk = _whole_graph_block(g)
        was_training = self.training
        self.eval()
        h = g.ndata["features"].float()
        with torch.no_grad():
            for l, layer in enumerate(self.gatv2_layers):
                h = layer(blk, h)
                h = h.flatten(1) if l < len(self.gatv2_layers) - 1 else h.mean(1)
        self.train(was_training)
        return h


class _WholeGraphBlock:
    """The full graph seen as one block (every node is source and destination) — what
    ``MultiLayerFullNeighborSampler(1)`` yields batch by batch in the reference's ``inference``."""

    is_block = True

    def __init__(self, g: Graph):
        if g.num_edges() >= 2 ** 31:
            raise NotImplementedError("whole-graph inference needs |E| < 2^31")
        self.indptr = g.indptr.to(torch.int32)
        self.edge_src = g.indices
        n = g.num_nodes()
        self.edge_dst = torch.repeat_interleave(torch.arange(n, dtype=torch.int32, device=g.device), g.in_degrees())
        self._n = n
        self.srcdata, self.dstdata, self.edata = {}, {}, {}
        self._transpose = None
        # 32-edge row segments for the balanced SpMM (hub columns of a power-law graph hold 10^4 edges)
        segs = ((g.in_degrees().long() + ops.SPMM_SEG - 1) // ops.SPMM_SEG).clamp(min=1)
        self.seg_ptr = torch.cat([segs.new_zeros(1), segs.cumsum(0)]).to(torch.int32)

    def num_src_nodes(self):
        return self._n

    num_dst_nodes = number_of_dst_nodes = number_of_src_nodes = num_src_nodes

    def num_edges(self):
        return int(self.edge_src.numel())

    @property
    def device(self):
        return self.indptr.device

    def in_degrees(self):
        return self.indptr[1:] - self.indptr[:-1]

    def out_degrees(self):
        return torch.bincount(self.edge_src.long(), minlength=self._n)


def _whole_graph_block(g: Graph):
    blk = getattr(g, "_whole_block", None)
    if blk is None:
        blk = _WholeGraphBlock(g)
        g._whole_block = blk
    return blk


def _full_inference(model, g: Graph, post):
    blk = _whole_graph_block(g)
    was_training = model.training
    model.eval()
    h = g.ndata["features"].float()
    with torch.no_grad():
        for l, layer in enumerate(model.layers):
            h = layer(blk, h)
            if l < len(model.layers) - 1:
                h = post(l, h)
    model.train(was_training)
    return h
