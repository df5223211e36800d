"""Autograd-aware wrappers over the C ABI (``include/bliss_b200.h``): feature gather, row norms,
SpMM (SAGE / GCN aggregation) and the fused GATv2 attention.  Every function launches the
hand-written sm_100a kernels on torch's current stream; there is no PyTorch fallback — a tensor
that is not a contiguous fp32 CUDA tensor raises.

Reference call sites replaced: ``train_lightning.py:138`` (feature fetch), ``model.py:318,425,211``
(``th.norm``), ``dglnn.SAGEConv`` / ``GraphConv`` message passing (``model.py:321-329,428-436``) and
``custom_GATv2Conv.forward`` (``model.py:80-99``).
"""
from __future__ import annotations

import os

import torch

from . import _native as N


def _req(t: torch.Tensor, dtype=torch.float32, name="tensor"):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (the BLISS hot path has no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def gather_rows(table: torch.Tensor, nid: torch.Tensor, with_norm: bool = False, out=None, norm_out=None):
    """``table[nid]`` (lazy DGL frame gather, ``train_lightning.py:138-139``); optionally also the
    row L2 norms of the gathered rows (``embed_norm`` of layer 0, ``model.py:318``).  ``out`` / ``norm_out``: persistent
    destination buffers (the pipelined step gathers the next batch's inputs ahead of time)."""
    if table.dtype != torch.float32 or table.dim() != 2:
        # labels / masks / non-fp32 frames: not the hot path, plain indexing
        if out is not None:
            torch.index_select(table, 0, nid.long(), out=out)
        else:
            out = table[nid.long()]
        return (out, None) if with_norm else out
    table = _req(table, name="table")
    nid = _req(nid, torch.int32, "nid")
    n, d = nid.numel(), table.shape[1]
    if out is None:
        out = torch.empty((n, d), dtype=torch.float32, device=table.device)
    elif out.shape != (n, d) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("gather_rows: out must be a contiguous float32 [len(nid), dim] tensor")
    norm = norm_out if norm_out is not None else (
        torch.empty(n, dtype=torch.float32, device=table.device) if with_norm else None)
    N.call("bliss_gather_rows", N.ptr(table), N.ptr(nid), n, d, N.ptr(out), N.ptr(norm), N.stream())
    return (out, norm) if with_norm else out


def row_norm(x: torch.Tensor) -> torch.Tensor:
    """``th.norm(h, dim=1)`` (``model.py:318``); detached — the bandit only reads it."""
    x = _req(x.detach(), name="x")
    x2 = x.reshape(x.shape[0], -1)
    out = torch.empty(x2.shape[0], dtype=torch.float32, device=x.device)
    N.call("bliss_row_norm", N.ptr(x2), x2.shape[0], x2.shape[1], N.ptr(out), N.stream())
    return out


def block_transpose_into(block, pool):
    """Build the transpose of ``block`` (exact sizes) into a LayerPool's capacity buffers: running the
    scan over ``cap_src`` rows leaves the padded source rows empty (t_indptr tail = E_b)."""
    E, n_dst = block.num_edges(), block.num_dst_nodes()
    N.call("bliss_block_transpose", N.ptr(block.edge_src), N.ptr(block.edge_dst), E, pool.cap_src, n_dst,
           N.ptr(pool.t_indptr), N.ptr(pool.t_cursor), N.ptr(pool.t_bits), N.ptr(pool.t_pre), pool.t_words,
           N.ptr(pool.t_dst), N.ptr(pool.t_perm),
           N.ptr(pool.t_seg_ptr), 1, None, None, None, N.stream())    # counts were accumulated by the fill kernel (out_deg)
    block._transpose = (pool.t_indptr[:block.num_src_nodes() + 1], pool.t_dst[:E], pool.t_perm[:E],
                        pool.t_seg_ptr[:block.num_src_nodes() + 1])


def block_transpose(block):
    """Source-major CSR of a block (cached on the block) for the backward aggregation."""
    _wait_ready(block, "_t_ready")
    if block._transpose is None:
        E, n_src, n_dst = block.num_edges(), block.num_src_nodes(), block.num_dst_nodes()
        dev = block.device
        t_indptr = torch.empty(n_src + 1, dtype=torch.int32, device=dev)
        t_cursor = torch.empty(max(n_src, 1), dtype=torch.int32, device=dev)
        t_words = (n_dst + 31) // 32
        t_bits = torch.empty(max(n_src * t_words, 1), dtype=torch.int32, device=dev)     # cleared by the call
        t_pre = torch.empty(max(n_src * t_words, 1), dtype=torch.int32, device=dev)
        t_dst = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
        t_perm = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
        t_seg = torch.empty(n_src + 1, dtype=torch.int32, device=dev)
        N.call("bliss_block_transpose", N.ptr(block.edge_src), N.ptr(block.edge_dst), E, n_src, n_dst,
               N.ptr(t_indptr), N.ptr(t_cursor), N.ptr(t_bits), N.ptr(t_pre), t_words, N.ptr(t_dst), N.ptr(t_perm),
               N.ptr(t_seg),
               0, None, None, None, N.stream())
        block._transpose = (t_indptr, t_dst[:E], t_perm[:E], t_seg)
    return block._transpose


SPMM_SEG = 32   # csrc/common.cuh BLISS_SPMM_SEG


def _spmm_tiling(d: int):
    """(padded row width of the partial-sum scratch, number of column tiles) that covers whichever vector
    width ``bliss_spmm`` picks (``launch_spmm`` in csrc/aggregate.cu: tile = 32 * VEC * NCH, NCH <= 8)."""
    width, tiles = 0, 0
    for vec in (4, 2, 1):
        if d % vec:
            continue
        per_lane = -(-d // (32 * vec))
        nch = 1
        while nch < per_lane and nch < 8:
            nch *= 2
        tile = 32 * vec * nch
        n = -(-d // tile)
        width, tiles = max(width, n * tile), max(tiles, n)
    return width, tiles


def _spmm_raw(indptr, col, perm, w, sscale, dscale, agg, x, n_rows, seg_ptr=None):
    """One ``bliss_spmm`` call.  With ``seg_ptr`` (sampled blocks) rows are cut into 32-edge segments, one
    warp each, 8 consecutive segments per CTA; rows that span several CTAs combine their partials through
    a scratch buffer sized from the host-side bound edges/32 + rows (no sync)."""
    d = x.shape[1]
    y = torch.empty((n_rows, d), dtype=torch.float32, device=x.device)
    partial, item_cap = None, 0
    if seg_ptr is not None:
        width, _ = _spmm_tiling(d)
        item_cap = int(col.numel()) // SPMM_SEG + n_rows + 1
        partial = torch.empty(item_cap * width, dtype=torch.float32, device=x.device)
    N.call("bliss_spmm", N.ptr(indptr), N.ptr(col), N.ptr(perm), N.ptr(w), N.ptr(sscale), N.ptr(dscale),
           agg, N.ptr(x), n_rows, d, N.ptr(seg_ptr), N.ptr(partial), item_cap, N.ptr(y), N.stream())
    return y


def _wait_ready(block, what="_ready"):
    """A block of the whole-step graph is filled (``_ready``) and transposed (``_t_ready``) on a side stream
    (``sampler.enqueue_static``): the first kernel that reads its edges / its transpose waits for the event."""
    ev = getattr(block, what, None)
    if ev is not None:
        torch.cuda.current_stream().wait_event(ev)


#: recorded behind the latest backward aggregation (model._LinearSplitK's overlapped weight gradient waits for it)
LAST_SPMM_BWD = None


def _wgrad_overlap_on(t):
    return t.is_cuda and os.environ.get("BLISS_WGRAD_OVERLAP", "1") == "1"


class _SpMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, block, w, sscale, dscale):
        x = _req(x, name="x")
        _wait_ready(block)
        ctx.block, ctx.w, ctx.sscale, ctx.dscale = block, w, sscale, dscale
        return _spmm_raw(block.indptr, block.edge_src, None, w, sscale, dscale, N.AGG_SUM, x,
                         block.num_dst_nodes(), getattr(block, "seg_ptr", None))

    @staticmethod
    def backward(ctx, gy):
        block = ctx.block
        t_indptr, t_dst, t_perm, t_seg = block_transpose(block)
        gy = _req(gy, name="grad")
        w, perm = ctx.w, t_perm
        t_w = getattr(block, "_t_w", None)
        if t_w is not None and w is not None and w.data_ptr() == t_w[0]:
            w, perm = t_w[1], None            # the transpose carries these weights in its own order: no gather through perm
        # dx_c = sscale_c * Σ_{e: src_e = c} w_e * dscale_{dst_e} * dy_{dst_e}
        gx = _spmm_raw(t_indptr, t_dst, perm, w, ctx.dscale, ctx.sscale, N.AGG_SUM, gy,
                       block.num_src_nodes(), t_seg)
        global LAST_SPMM_BWD
        if _wgrad_overlap_on(gy):
            LAST_SPMM_BWD = torch.cuda.Event()
            LAST_SPMM_BWD.record()
        return gx, None, None, None, None


def spmm(block, x, edge_weight=None, src_scale=None, dst_scale=None):
    """``y_i = dst_scale_i · Σ_{e→i} w_e · src_scale_{src(e)} · x_{src(e)}`` (g-SpMM ``u_mul_e``/sum).
    ``x`` is [n_src, D] fp32; differentiable w.r.t. ``x``."""
    if x.dim() != 2:
        raise ValueError("spmm expects [n_src, D]")
    if edge_weight is not None:
        edge_weight = _req(edge_weight.detach(), name="edge_weight")
    return _SpMM.apply(x, block, edge_weight, src_scale, dst_scale)


def mean_scale(block) -> torch.Tensor:
    """1 / max(in_degree, 1) per destination (``fn.mean``), cached on the block."""
    s = getattr(block, "_mean_scale", None)
    if s is None:
        s = 1.0 / block.in_degrees().clamp(min=1).to(torch.float32)
        block._mean_scale = s
    return s


class _GATv2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, attn, block, drop_mask, slope):
        feat = _req(feat, name="feat")          # [n_src, H, D]
        _wait_ready(block)
        attn_c = _req(attn, name="attn").reshape(attn.shape[-2], attn.shape[-1])
        n_src, H, D = feat.shape
        n_dst, E = block.num_dst_nodes(), block.num_edges()
        dev = feat.device
        out = torch.empty((n_dst, H, D), dtype=torch.float32, device=dev)
        logits = torch.empty((E, H), dtype=torch.float32, device=dev)
        rmax = torch.empty((n_dst, H), dtype=torch.float32, device=dev)
        rsum = torch.empty((n_dst, H), dtype=torch.float32, device=dev)
        N.call("bliss_gatv2_fwd", N.ptr(block.indptr), N.ptr(block.edge_src), N.ptr(feat), N.ptr(attn_c),
                                        N.ptr(drop_mask), slope, n_dst, H, D, N.ptr(out), N.ptr(logits),
                                        N.ptr(rmax), N.ptr(rsum), N.stream())
        ctx.save_for_backward(feat, attn_c, logits, rmax, rsum, out)
        ctx.block, ctx.drop_mask, ctx.slope, ctx.attn_shape = block, drop_mask, slope, attn.shape
        ctx.mark_non_differentiable(logits)
        return out, logits

    @staticmethod
    def backward(ctx, gout, _glogits):
        feat, attn_c, logits, rmax, rsum, out = ctx.saved_tensors
        block, mask, slope = ctx.block, ctx.drop_mask, ctx.slope
        n_src, H, D = feat.shape
        n_dst, E = block.num_dst_nodes(), block.num_edges()
        gout = _req(gout, name="grad")
        dev = feat.device
        gfeat = torch.empty_like(feat)
        gattn = torch.zeros_like(attn_c)
        glogit = torch.empty((max(E, 1), H), dtype=torch.float32, device=dev)
        L = N.lib()
        N.call("bliss_gatv2_bwd_dst", N.ptr(block.indptr), N.ptr(block.edge_src), N.ptr(feat), N.ptr(attn_c),
                                      N.ptr(mask), N.ptr(logits), N.ptr(rmax), N.ptr(rsum), N.ptr(out),
                                      N.ptr(gout), slope, n_dst, H, D, N.ptr(glogit), N.ptr(gfeat),
                                      N.ptr(gattn), N.stream())
        t_indptr, t_dst, t_perm, _ = block_transpose(block)
        N.call("bliss_gatv2_bwd_src", N.ptr(t_indptr), N.ptr(t_dst), N.ptr(t_perm), N.ptr(feat), N.ptr(attn_c),
                                      N.ptr(mask), N.ptr(logits), N.ptr(rmax), N.ptr(rsum), N.ptr(gout),
                                      N.ptr(glogit), slope, n_src, n_dst, H, D, N.ptr(gfeat), N.stream())
        return gfeat, gattn.reshape(ctx.attn_shape), None, None, None


def gatv2_attention(block, feat, attn, negative_slope, drop_mask=None):
    """Fused SDDMM(u_add_v) + leaky-relu + ·attn + edge-softmax + SpMM(u_mul_e) of
    ``custom_GATv2Conv.forward`` (``model.py:80-99``) with shared source/destination projection.

    feat [n_src, H, D], attn [1, H, D] → (ft [n_dst, H, D], logits [E, H]) where ``logits`` are the
    pre-softmax scores the reference returns as attention (``model.py:108-110``).  ``drop_mask``
    [E, H] (already scaled by 1/(1-p)) multiplies the softmax output (``attn_drop``, :88-90)."""
    if drop_mask is not None:
        drop_mask = _req(drop_mask, name="drop_mask")
    return _GATv2.apply(feat, attn, block, drop_mask, float(negative_slope))


class _SageEpilogue(torch.autograd.Function):
    """``dropout(relu(a + b + bias))`` + row norms, one launch each way (``csrc/epilogue.cu``)."""

    @staticmethod
    def forward(ctx, a, b, bias, relu, p_drop, seed, step_dev, layer, want_norm):
        a, b = _req(a, name="a"), _req(b, name="b")
        n, d = a.shape
        y = torch.empty_like(a)
        norm = torch.empty(n, dtype=torch.float32, device=a.device) if want_norm else None
        N.call("bliss_sage_epilogue_fwd", N.ptr(a), N.ptr(b), N.ptr(bias), n, d, int(relu), float(p_drop), int(seed),
               N.ptr(step_dev), int(layer), N.ptr(y), N.ptr(norm), N.stream())
        ctx.save_for_backward(y)
        ctx.gate, ctx.p_drop, ctx.has_bias = bool(relu), float(p_drop), bias is not None
        if norm is None:
            norm = y.new_empty(0)
        ctx.mark_non_differentiable(norm)
        return y, norm

    @staticmethod
    def backward(ctx, gy, _gnorm):
        (y,) = ctx.saved_tensors
        gy = _req(gy, name="grad")
        n, d = y.shape
        gz = torch.empty_like(y)
        gbias = partial = None
        if ctx.has_bias:
            partial = torch.empty((N.lib().bliss_sage_epilogue_parts(), d), dtype=torch.float32, device=y.device)
            gbias = torch.empty(d, dtype=torch.float32, device=y.device)
        if not ctx.gate and ctx.p_drop > 0:
            raise RuntimeError("sage_epilogue: dropout without relu is not supported (the gate is read off y > 0)")
        N.call("bliss_sage_epilogue_bwd", N.ptr(gy), N.ptr(y), n, d, int(ctx.gate), ctx.p_drop, N.ptr(gz),
               N.ptr(partial), N.ptr(gbias), N.stream())
        return gz, gz, gbias, None, None, None, None, None, None


def sage_epilogue(a, b, bias, relu=True, p_drop=0.0, seed=0, step_dev=None, layer=0, want_norm=True):
    """``y = dropout(relu(a + b + bias), p_drop)`` and ``||y_r||_2`` per row — the tail of a hidden SAGE layer
    (``model.py:321-332``) and the next layer's ``embed_norm`` (``model.py:318``).  Differentiable w.r.t.
    ``a``, ``b`` and ``bias``.  Dropout draws are Philox(seed; element, layer, ``*step_dev``)."""
    y, norm = _SageEpilogue.apply(a, b, bias, relu, p_drop, seed, step_dev, layer, want_norm)
    return (y, norm) if want_norm else (y, None)


class _XentMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        logits = _req(logits, name="logits")
        labels = _req(labels, torch.int64, "labels")
        n, c = logits.shape
        row = torch.empty(n, dtype=torch.float32, device=logits.device)
        loss = torch.empty(1, dtype=torch.float32, device=logits.device)
        grad = torch.empty_like(logits)
        N.call("bliss_xent_mean", N.ptr(logits), N.ptr(labels), n, c, N.ptr(row), N.ptr(loss), N.ptr(grad), N.stream())
        ctx.save_for_backward(grad)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None


def cross_entropy_mean(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """``nn.CrossEntropyLoss()(logits, labels)`` (``train_lightning.py:77-79,142``) with the gradient produced in
    the forward launch."""
    return _XentMean.apply(logits, labels)
