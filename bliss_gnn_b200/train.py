"""The caller of the hot path: data module, training step and CLI with the reference's flags.

Replaces ``train_lightning.py`` of the reference without Lightning / DGL:
* :class:`DataModule`  — ``train_lightning.py:307-422``: dataset, self-loops, ``--undirected``, static
  edge weights ``w``, sampler factory by ``--sampler``, seed-batch iterators (shuffle + drop_last for
  training, in-order for validation).
* :class:`Trainer.training_step` — ``train_lightning.py:100-168`` + ``BatchSizeCallback.on_train_batch_end``
  (``:463-471``): sample → lazy feature / label gather → forward → loss → backward → Adam → ``exp3``.
* data-parallel over one process per GPU (``torchrun``): every rank samples and aggregates its own
  seed batches over the replicated graph; NCCL all-reduces one flat gradient buffer and all-gathers
  the sparse bandit updates (``sampler._update_distributed``).  The reference is single-device.
"""
from __future__ import annotations

import argparse
import math
import os
import time
from typing import Iterator, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .graph import NID, Graph, add_self_loops_and_build, load_dataset, normalized_edata
from .model import GCN, SAGE, GATv2, SAGEConv
from .parallel import FlatAdam, FlatGrads, shard_batches
from .sampler import (BanditLadiesSampler, LadiesSampler, MultiLayerFullNeighborSampler, NeighborSampler,
                      PoissonBanditLadiesSampler, PoissonLadiesSampler)


def make_sampler(name: str, fanouts, importance_sampling=1, eta=0.1, num_steps=-1, model="sage", rng_seed=0,
                 normalize="lazy"):
    """Sampler factory with the reference's flag semantics (``train_lightning.py:349-370``)."""
    if name == "full":                                                           # :349-350
        return MultiLayerFullNeighborSampler(len(fanouts), rng_seed=rng_seed)
    if name == "neighbor":                                                       # :351-357
        return NeighborSampler(fanouts, rng_seed=rng_seed)
    if "ladies" in name:
        return (PoissonLadiesSampler if "poisson" in name else LadiesSampler)(fanouts, rng_seed=rng_seed)
    if "bandit" in name:
        cls = PoissonBanditLadiesSampler if "poisson" in name else BanditLadiesSampler
        return cls(fanouts, importance_sampling=importance_sampling, node_embedding="features",
                   num_steps=num_steps, eta=eta, model=model, rng_seed=rng_seed, normalize=normalize)
    raise ValueError(f"unknown sampler {name}")


class DataModule:
    """``train_lightning.py:307-422`` over the plain :class:`Graph` container."""

    def __init__(self, dataset_name, undirected=False, data_cpu=False, use_uva=False, fan_out=(128, 256), eta=0.4,
                 device=torch.device("cpu"), batch_size=64, num_workers=0, sampler="bandit",
                 importance_sampling=1, cache_size=0, num_steps=500, model="sage", seed=0, rank=0, world_size=1,
                 graph: Optional[Graph] = None, normalize="lazy", pad_features: bool = True, graph_seed: int = 0):
        self.sampler_name, self.num_steps, self.eta = sampler, num_steps, eta
        if graph is None:     # (a synthetic graph is the same for every run of --k-runs: runs differ by their RNG streams only)
            g, n_classes, multilabel = load_dataset(dataset_name, device=device, seed=graph_seed)
        else:
            g, n_classes, multilabel = graph, graph.n_classes, graph.multilabel
        if undirected:                                                           # :337-339
            src, dst = g.coo()
            feats = dict(g.ndata)
            g = add_self_loops_and_build(torch.cat([src, dst]), torch.cat([dst, src]), g.num_nodes())
            g.ndata.update(feats)
        g = g.to(device)
        self.train_nid = torch.nonzero(g.ndata["train_mask"], as_tuple=True)[0].to(torch.int32)   # :344-346
        self.val_nid = torch.nonzero(g.ndata["val_mask"], as_tuple=True)[0].to(torch.int32)
        self.test_nid = torch.nonzero(g.ndata["test_mask"], as_tuple=True)[0].to(torch.int32)
        g.edata["w"] = normalized_edata(g)                                       # :359,362
        fanouts = [int(_) for _ in fan_out]
        # every rank draws its own stream: the Philox key carries the rank
        self.sampler = make_sampler(sampler, fanouts, importance_sampling, eta, num_steps, model,
                                    rng_seed=(seed + 2) + (rank << 32), normalize=normalize)
        self.g = g
        self.device = device
        self.batch_size = batch_size
        self.in_feats = g.ndata["features"].shape[1]
        if pad_features and g.device.type == "cuda" and self.in_feats % 4:
            # rows of the feature table start 16-byte aligned: 128-bit gather loads and aligned GEMM
            # operands for the input layer (the models zero-pad their first weight view to match)
            g.ndata["features"] = F.pad(g.ndata["features"], (0, 4 - self.in_feats % 4)).contiguous()
        self.n_classes, self.multilabel = n_classes, multilabel
        self.rank, self.world_size = rank, world_size
        self._epoch = 0
        self._seed = seed + 1

    def train_batches(self) -> Iterator[torch.Tensor]:
        """shuffle=True, drop_last=True (``:396-408``); rank r takes batches r, r+R, … of the shared
        epoch permutation."""
        gen = torch.Generator().manual_seed(self._seed + self._epoch)
        self._epoch += 1
        perm = self.train_nid[torch.randperm(self.train_nid.numel(), generator=gen).to(self.train_nid.device)]
        n_batches = perm.numel() // self.batch_size
        for b in shard_batches(n_batches, self.rank, self.world_size):
            yield perm[b * self.batch_size:(b + 1) * self.batch_size]

    def val_batches(self) -> Iterator[torch.Tensor]:
        """shuffle=False, drop_last=False (``:410-422``)."""
        for b in range(0, self.val_nid.numel(), self.batch_size):
            yield self.val_nid[b:b + self.batch_size]


def micro_f1(pred: torch.Tensor, labels: torch.Tensor, multilabel: bool) -> float:
    """torchmetrics Multiclass/MultilabelF1Score(average='micro') (``train_lightning.py:68-72``)."""
    if multilabel:
        p = (torch.sigmoid(pred) > 0.5)
        y = labels > 0.5
        tp = (p & y).sum().item()
        return 2 * tp / max(p.sum().item() + y.sum().item(), 1)
    return (pred.argmax(1) == labels).float().mean().item()


def build_model(name: str, in_feats, n_hidden, n_classes, n_layers, dropout=0.1, num_in_heads=4, num_out_heads=1,
                attn_dropout=0.1, negative_slope=0.2, residual=False, faithful_gcn_quirk=True):
    """Model factory of ``train_lightning.py:581-618``.  The reference builds a **SAGE** module for
    ``--model gcn`` (``:597-607``); ``faithful_gcn_quirk=False`` builds the real GCN (``GCNLightning``)."""
    name = name.lower()
    if "gat" in name:
        heads = ([num_in_heads] * (n_layers - 1)) + [num_out_heads]              # :247
        return GATv2(n_layers, in_feats, n_hidden, n_classes, heads, F.elu, dropout, attn_dropout,
                     negative_slope, residual)
    if "gcn" in name and not faithful_gcn_quirk:
        return GCN(in_feats, n_hidden, n_classes, n_layers, F.relu, dropout)
    return SAGE(in_feats, n_hidden, n_classes, n_layers, F.relu, dropout)


class Trainer:
    """One training step = the hot path end to end (``train_lightning.py:100-168,463-471``)."""

    def __init__(self, datamodule: DataModule, model: nn.Module, lr=0.002, process_group=None,
                 static_graph: bool = False, eager_warmup: int = 4, pipeline: bool = True):
        """``static_graph=True``: after ``eager_warmup`` ordinary steps the blocks are built into
        capacity-padded persistent buffers (``sampler.LayerPool``) and the whole step (sampling, model
        forward / backward, Adam, bandit update) runs as ONE replayed CUDA graph — the step stops being
        bound by Python dispatch.  ``pipeline=True``: the step's device counters (sizes, capacity flags)
        are copied to pinned memory stream-ordered and consumed while the NEXT step runs, so the host
        never waits inside a step; ``flush()`` consumes the outstanding read."""
        self.dm, self.model, self.pg = datamodule, model, process_group
        self.static_graph, self.eager_warmup = bool(static_graph), int(eager_warmup)
        self.pipeline, self._pending = bool(pipeline), None
        self._ctr_pins, self._call_done, self._set_free = {}, None, [None, None]
        self._loss_hist, self._loss_pins = [], None
        self._seed_stage, self._seed_staged, self._copy_stream = [None, None], [None, None], None
        self.total_sampled_edges = 0        # Σ block edges over all consumed steps (bench.py's edges/s)
        self.pool_resizes = 0               # capacity re-sizings (each one re-captures the step graph)
        self._dev_step_mirror = None        # host's view of the device-side Philox step counter
        self._sizing_steps = 0              # ordinary (eager) steps seen so far: they size the capacity pools
        # the data-parallel code path (two graphs with the collectives in between) can be forced on a
        # single rank, so it is testable on one GPU
        self._force_dp = bool(os.environ.get("BLISS_FORCE_DP_PATH")) and process_group is not None
        self._graph, self._pools, self._padded, self._exchange, self._gradx = None, None, None, None, None
        self._sets, self._graphs, self._graph_kernel_counts = None, {}, {}
        self._cur, self._next_ready, self._prefetched_seeds = 0, False, None
        self._max_src, self._max_edges = None, None
        self.graph_replays, self.graph_kernels = 0, 0
        self.graph_kernel_launches = 0      # hand-written kernels launched by graph replays so far (bench.py)
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.loss_fn = nn.BCEWithLogitsLoss() if datamodule.multilabel else nn.CrossEntropyLoss()   # :77-79
        if not datamodule.multilabel and next(model.parameters()).is_cuda:
            self.loss_fn = ops.cross_entropy_mean       # same loss, forward + gradient in one launch
        # The feature rows are padded to a 16-byte multiple (DataModule): the input layer's weights get the matching zero
        # columns inside their flat storage (parallel.flat_layout), so no padded copy is built per step.
        feats = datamodule.g.ndata["features"] if hasattr(datamodule, "g") else None
        extra = int(feats.shape[1] - datamodule.in_feats) if feats is not None and feats.dim() == 2 else 0
        if extra > 0 and next(model.parameters()).is_cuda:
            for name, p_ in model.named_parameters():
                if name.startswith("layers.0.") and p_.dim() == 2 and p_.shape[1] == datamodule.in_feats \
                        and "fc_" in name and isinstance(getattr(model, "layers", [None])[0], SAGEConv):
                    p_._bliss_pad_cols = extra
        # one flat gradient buffer: a single all-reduce per step (~0.46 M parameters for SAGE/Reddit)
        self.grads = FlatGrads(model.parameters())
        params = self.grads.params
        self._flat_grad = self.grads.flat
        self._grads_clean = False
        if params[0].is_cuda:     # one flat Adam launch per step (csrc/optim.cu); it also clears the gradients
            self.optimizer = FlatAdam(self.grads, lr=lr)
        else:                     # host-logic tests (gloo): the product path refuses CPU graphs anyway
            self.optimizer = torch.optim.Adam(params, lr=lr)                         # :206
        self.scheduler = torch.optim.lr_scheduler.StepLR(self.optimizer, gamma=0.01, step_size=5)   # :208 (per epoch)
        if process_group is not None and hasattr(datamodule.sampler, "process_group"):
            datamodule.sampler.process_group = process_group
        self.num_steps = 0
        self.w = 0.99
        n_layers = len(datamodule.sampler.nodes_per_layer)
        self.cum_sampled_nodes = [0.0] * (n_layers + 1)
        self.cum_sampled_edges = [0.0] * n_layers
        self.last_blocks = None

    def _ema(self, mfgs):
        """EMA sampled nodes / edges per layer (``train_lightning.py:104-136``)."""
        self.num_steps += 1
        self.total_sampled_edges += sum(m.num_edges() for m in mfgs)
        for i, mfg in enumerate(mfgs):
            self.cum_sampled_nodes[i] = self.cum_sampled_nodes[i] * self.w + mfg.num_src_nodes()
            self.cum_sampled_edges[i] = self.cum_sampled_edges[i] * self.w + mfg.num_edges()
        self.cum_sampled_nodes[len(mfgs)] = self.cum_sampled_nodes[len(mfgs)] * self.w + mfgs[-1].num_dst_nodes()

    def num_sampled_edges(self, i):
        return self.cum_sampled_edges[i] * (1 - self.w) / (1 - self.w ** self.num_steps)

    def num_sampled_nodes(self, i):
        return self.cum_sampled_nodes[i] * (1 - self.w) / (1 - self.w ** self.num_steps)

    def training_step(self, seeds: torch.Tensor, next_seeds: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One training step on ``seeds``.  ``next_seeds`` (optional): the batch the NEXT call will train on — the
        look-ahead of a data loader; the static-graph path then samples its blocks in the shadow of this step's
        backward pass (after this step's ``exp3``, so the trajectory is the reference's)."""
        self._graph_loss = None
        if self.static_graph:
            if self._full_graph_ok():
                loss = self._training_step_full_graph(seeds, next_seeds)
            else:
                loss = self._training_step_static_partial(seeds)
        else:
            dm, g = self.dm, self.dm.g
            input_nodes, output_nodes, mfgs = dm.sampler.sample_blocks(g, seeds)
            loss = self._eager_rest(mfgs)
        # (a step that was not a graph replay keeps the tensor itself)
        self._loss_hist = (self._loss_hist + [self._graph_loss or ("eager", loss)])[-2:]
        return loss

    def host_loss(self, back: int = 0) -> float:
        """The loss of the last ``training_step`` call (``back=0``: waits for that step) or of the call before it
        (``back=1``: in a loop that runs one step ahead of the device this does not stall).  A replayed step graph
        copies its loss to pinned host memory itself, so no copy has to be enqueued between two step graphs."""
        kind, ref = self._loss_hist[-1 - back]
        if kind == "eager":
            return float(ref)
        slot, p = ref
        self._call_done[slot].synchronize()
        return float(self._loss_pins[p])

    def _eager_rest(self, mfgs):
        dm, g = self.dm, self.dm.g
        self._ema(mfgs)
        batch_inputs = mfgs[0].srcdata["features"]                               # :138  (gather kernel)
        batch_labels = mfgs[-1].dstdata["labels"]                                # :139
        batch_pred = self.model(mfgs, batch_inputs)                              # :141
        loss = self.loss_fn(batch_pred, batch_labels)                            # :142
        self._zero_grads()
        ops.LAST_SPMM_BWD = None          # (an event of an earlier pass / capture must not be waited for)
        loss.backward()
        self.grads.all_reduce_mean_(self.pg)
        self._optimizer_step()
        if "bandit" in dm.sampler_name:                                          # :469-471
            dm.sampler.exp3(mfgs, g)
        # detached: nothing may keep this step's autograd graph (and its AccumulateGrad nodes) alive,
        # or a later CUDA-graph capture would see nodes bound to another stream
        self.last_blocks, self.last_pred, self.last_labels = mfgs, batch_pred.detach(), batch_labels
        return loss.detach()

    def _zero_grads(self):
        """The flat Adam launch leaves the gradient buffer cleared; zero it only when something else
        (a backward pass without a step) has written to it since."""
        if not self._grads_clean:
            self._flat_grad.zero_()
        self._grads_clean = False

    def _sync_lr(self):
        """A captured step reads lr from a device scalar: refresh it when the scheduler changed it."""
        if isinstance(self.optimizer, FlatAdam):
            self.optimizer.sync_lr()

    def _optimizer_step(self):
        self.optimizer.step()
        self._grads_clean = isinstance(self.optimizer, FlatAdam)

    # ---- static-shape path: capacity-padded blocks in persistent pools + CUDA graphs ------------------------
    #: edge capacity of a pool = FACTOR x the largest block seen in the sizing steps + SLACK
    POOL_EDGE_FACTOR, POOL_EDGE_SLACK = 1.45, 4096

    def _alloc_pools(self):
        """Two pool sets (double buffering): while the backward pass of step t still reads set p, the blocks of
        step t+1 are sampled into set 1-p (``_training_step_full_graph``).  The eager-sampling variant
        (``_training_step_static_partial``) uses set 0 only."""
        from .sampler import LayerPool
        from .graph import Block
        dm, g = self.dm, self.dm.g
        fan, L = dm.sampler.nodes_per_layer, len(dm.sampler.nodes_per_layer)
        dev = g.device
        bandit = "bandit" in dm.sampler_name
        if getattr(self, "_next_ready", False):             # an outstanding prefetch is dropped with its pool set
            dm.sampler.step -= 1
        # edges of a block vary by +-17 % around their median from batch to batch at the Reddit shape (measured over
        # 400 steps): 1.45x the largest count seen so far; the high-water mark in _consume_counters re-sizes at 92 %
        cap_e = [int(self.POOL_EDGE_FACTOR * self._max_edges[l]) + self.POOL_EDGE_SLACK for l in range(L)]
        self._exchange = None
        if (self.world > 1 or self._force_dp) and bandit:   # ranks must agree on the exchange layout
            from .parallel import BanditExchange
            t = torch.tensor(cap_e, dtype=torch.int64, device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX, group=self.pg)
            cap_e = [int(v) for v in t.tolist()]
            self._exchange = BanditExchange(cap_e, self.world, dev, self.pg)
        if (self.world > 1 or self._force_dp) and self._gradx is None and isinstance(self.optimizer, FlatAdam) \
                and os.environ.get("BLISS_P2P", "1") != "0":
            from .parallel import GradExchange
            try:       # gradient all-reduce through peer memory, fused into the Adam launch (csrc/optim.cu)
                self._gradx = GradExchange(self._flat_grad.numel(), self.world, dev, self.pg)
            except Exception as e:
                import warnings
                warnings.warn(f"peer-memory gradient exchange unavailable ({type(e).__name__}: {e}); using NCCL all-reduce")
                self._gradx = False
        sets = []
        for p in range(2):
            seeds_static = torch.zeros(dm.batch_size, dtype=torch.int32, device=dev)
            pools, cd = [None] * L, dm.batch_size
            for l in reversed(range(L)):                       # output layer first: cap_dst[l] = cap_src[l+1]
                # sources = destinations + Poisson-selected nodes (mean <= fan-out, sigma <= sqrt(fan-out)); the
                # high-water mark in _consume_counters re-sizes long before a capacity can be hit
                if isinstance(dm.sampler, NeighborSampler):    # per-seed fan-outs: sized from the eager steps, bounded by |V|
                    cap_src = min(g.num_nodes(), int(1.5 * self._max_src[l]) + 256)
                else:
                    cap_src = max(cd + fan[l] + int(6 * math.sqrt(fan[l])) + 64, int(1.12 * self._max_src[l]) + 64)
                cap_src = (cap_src + 63) // 64 * 64            # row counts the split-K weight gradients divide evenly
                cap_src = min(cap_src, g.num_nodes())          # (a frontier never holds more than |V| nodes)
                pools[l] = LayerPool(dev, cd, cap_src, cap_e[l], bandit=bandit)
                cd = cap_src
            padded = []
            for l in range(L):
                pool = pools[l]
                dst_nid = pools[l + 1].src_nid if l < L - 1 else seeds_static
                pb = Block(pool.indptr, pool.e32[0], pool.e32[1], pool.src_nid, dst_nid, graph=g, csc_pos=pool.csc_pos)
                pb.seg_ptr, pb._mean_scale, pb._static_padded = pool.seg_ptr, pool.inv_deg, True
                if getattr(dm.sampler, "attach_weights", True):
                    pb.edata["edge_weights"] = pool.e32[3].view(torch.float32)
                pb._transpose = (pool.t_indptr, pool.t_dst, pool.t_perm, pool.t_seg_ptr)
                pool.padded = pb
                padded.append(pb)
            pset = _PoolSet(pools, padded, seeds_static, ctr_base=8 * p)
            feats, labels = g.ndata["features"], g.ndata["labels"]
            if feats.dtype == torch.float32 and feats.dim() == 2:      # inputs gathered ahead with the blocks
                pset.x = torch.zeros((pools[0].cap_src, feats.shape[1]), dtype=torch.float32, device=dev)
                pset.x_norm = torch.zeros(pools[0].cap_src, dtype=torch.float32, device=dev)
                pset.y = torch.zeros((dm.batch_size,) + tuple(labels.shape[1:]), dtype=labels.dtype, device=dev)
            sets.append(pset)
        self._sets = sets
        self._pools, self._padded, self._seeds_static = sets[0].pools, sets[0].padded, sets[0].seeds
        self._graph, self._graphs = None, {}
        self._cur, self._next_ready, self._prefetched_seeds = 0, False, None

    def _padded_fwd_bwd(self, step_optimizer: bool, after_forward=None, before_backward=None, pset=None, inputs=None):
        """``after_forward`` / ``before_backward``: hooks of the whole-step graph — work that only needs the
        forward pass is forked onto side streams there, work the backward pass needs is joined."""
        loss, pred, y = self._padded_fwd(pset, inputs)
        self._fwd_loss = loss.detach()
        if after_forward is not None:
            after_forward()
        self._zero_grads()
        if before_backward is not None:
            before_backward()
        ops.LAST_SPMM_BWD = None          # (an event of an earlier pass / capture must not be waited for)
        loss.backward()
        if step_optimizer:
            self._optimizer_step()
        return loss.detach(), pred, y

    @staticmethod
    def _gather_labels(labels, nid32, out=None):
        """``labels[nid]`` in one launch: int64 class ids ride through the row-gather kernel as 2-float rows
        (torch's index / index_select paths cost 2-3 launches or a slow generic gather here)."""
        if labels.dim() == 1 and labels.dtype == torch.int64 and labels.is_contiguous():
            buf = None if out is None else out.view(torch.float32).view(-1, 2)
            res = ops.gather_rows(labels.view(torch.float32).view(-1, 2), nid32, out=buf)
            return res.view(torch.int64).view(-1)
        if out is not None:
            return torch.index_select(labels, 0, nid32.long(), out=out)
        return labels[nid32.long()]

    def _gather_inputs(self, pset, ahead: bool):
        """Input features (with their row norms: layer 0's embed_norm) and labels of the batch sampled into ``pset``.
        ``ahead``: into the set's persistent buffers, as the tail of the look-ahead sampling."""
        g = self.dm.g
        if ahead and pset.x is not None:
            ops.gather_rows(g.ndata["features"], pset.pools[0].src_nid, with_norm=True, out=pset.x, norm_out=pset.x_norm)
            self._gather_labels(g.ndata["labels"], pset.seeds_in, out=pset.y)
            return pset.x, pset.x_norm, pset.y
        x, norm = ops.gather_rows(g.ndata["features"], pset.pools[0].src_nid, with_norm=True)
        return x, norm, self._gather_labels(g.ndata["labels"], pset.seeds)

    def _padded_fwd(self, pset=None, inputs=None):
        pset = pset or self._sets[0]
        x, norm, y = inputs if inputs is not None else self._gather_inputs(pset, False)
        x._bliss_row_norm = norm                       # layer 0's embed_norm comes with the gather (model.SAGE/GCN/GATv2)
        pred = self.model(pset.padded, x)
        if pred.shape[0] != self.dm.batch_size:          # (the top layer's capacity is the batch size: usually a no-op)
            pred = pred[: self.dm.batch_size]
        return self.loss_fn(pred, y), pred.detach(), y

    def _capture(self):
        self.last_pred = None
        for pb in self._padded:                            # drop tensors of earlier forward passes
            pb._ready = pb._t_ready = pb._t_w = None       # (and side-branch events of a whole-step capture)
            for k in ("embed_norm",):
                dict.pop(pb.srcdata, k, None)
            dict.pop(pb.edata, "a_ij", None)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up off the default stream (no optimizer step)
            for _ in range(2):
                self._padded_fwd_bwd(False)
        torch.cuda.current_stream().wait_stream(side)
        from . import _native
        before = _native.STATS.launches
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._static_loss, self._static_pred, self._static_y = self._padded_fwd_bwd(self.world == 1)
        self.graph_kernels = _native.STATS.launches - before    # hand-written kernels inside one replay
        _native.STATS.launches = before

    def _training_step_static_partial(self, seeds: torch.Tensor) -> torch.Tensor:
        """Eager sampling into the pools + replayed forward/backward(/Adam): the variant for samplers whose stage
        methods are overridden or whose draws are injected (the collectives sit between the sampler and the
        optimizer)."""
        dm, g, smp = self.dm, self.dm.g, self.dm.sampler
        L = len(smp.nodes_per_layer)
        if self._sets is None:
            if self._sizing_steps < self.eager_warmup or seeds.numel() != dm.batch_size:
                return self._sizing_step(seeds)
            self._alloc_pools()
        if seeds.numel() != dm.batch_size:                  # ragged last batch: ordinary path
            _, _, mfgs = smp.sample_blocks(g, seeds)
            return self._eager_rest(mfgs)
        self._drop_prefetch()
        self._seeds_static.copy_(seeds, non_blocking=True)
        _, _, mfgs = smp.sample_blocks(g, self._seeds_static, pools=self._pools)
        if not smp.pool_used:                               # per-stage sampler path (overridden stage / profiling)
            return self._eager_rest(mfgs)
        if smp.pool_overflow:                               # rare: grow the capacities, re-capture next step
            for l, b in enumerate(mfgs):
                self._max_src[l] = max(self._max_src[l], b.num_src_nodes())
                self._max_edges[l] = max(self._max_edges[l], b.num_edges())
            loss = self._eager_rest(mfgs)
            self._alloc_pools()
            return loss
        for l, b in enumerate(mfgs):
            ops.block_transpose_into(b, self._pools[l])
        if self._graph is None:
            self._capture()
        self._ema(mfgs)
        self._sync_lr()
        self._graph.replay()
        self.graph_replays += 1
        self.graph_kernel_launches += self.graph_kernels
        if self.world > 1:
            self.grads.all_reduce_mean_(self.pg)
            self._optimizer_step()
        for l, b in enumerate(mfgs):                        # what exp3 reads from the forward pass
            pb = self._padded[l]
            b.srcdata["embed_norm"] = pb.srcdata["embed_norm"][: b.num_src_nodes()]
            if "a_ij" in dict.keys(pb.edata):
                b.edata["a_ij"] = pb.edata["a_ij"][: b.num_edges()]
        if "bandit" in dm.sampler_name:
            smp.exp3(mfgs, g, exchange=self._exchange)
        self.last_blocks, self.last_pred, self.last_labels = mfgs, self._static_pred, self._static_y
        return self._static_loss

    def _sizing_step(self, seeds):
        """An ordinary (eager) step whose block sizes feed the capacity of the pools."""
        dm, g, smp = self.dm, self.dm.g, self.dm.sampler
        L = len(smp.nodes_per_layer)
        self._drop_prefetch()
        _, _, mfgs = smp.sample_blocks(g, seeds)
        self._sizing_steps += 1
        if self._max_src is None:
            self._max_src, self._max_edges = [0] * L, [0] * L
        for l, b in enumerate(mfgs):
            self._max_src[l] = max(self._max_src[l], b.num_src_nodes())
            self._max_edges[l] = max(self._max_edges[l], b.num_edges())
        return self._eager_rest(mfgs)

    def set_batch_size(self, batch_size: int):
        """``--vertex-limit`` changes the batch size between epochs (``train_lightning.py:480-485``): the capacity pools
        are sized by the batch, so they are re-sized from fresh eager steps and the step graphs re-captured."""
        self.flush()
        self._drop_prefetch()
        self.dm.batch_size = int(batch_size)
        self._sets, self._graphs, self._graph = None, {}, None
        self._pools = self._padded = None
        self._sizing_steps, self._max_src, self._max_edges = 0, None, None

    def _drop_prefetch(self):
        """Forget the blocks sampled ahead for a batch that is not going to be trained on next (eager step in
        between, re-sized pools, checkpoint): their Philox step is given back, so the next sampling draws exactly
        what an un-pipelined run would have drawn."""
        if getattr(self, "_next_ready", False):
            self._next_ready, self._prefetched_seeds = False, None
            self.dm.sampler.step -= 1

    # ---- whole step in CUDA graphs, sampling of step t+1 in the shadow of step t's backward pass --------------
    def _full_graph_ok(self) -> bool:
        smp = self.dm.sampler
        return smp._stages_not_overridden() and smp.inject_uniforms is None

    def _training_step_full_graph(self, seeds: torch.Tensor, next_seeds: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One step on pool set p = replay of the step graph of that set:

            forward(t) ─┬─ backward(t) ─ Adam(t) ───────────────────────────────┬─ (end of step t)
                        └─ exp3(t) ─ sample_blocks(t+1) into set 1-p (prefetch) ─┘

        ``exp3(t)`` needs only the forward pass (``embed_norm``, ``a_ij``) and ``sample_blocks(t+1)`` only the weights
        after ``exp3(t)`` — the reference's order (``bandit_sampler.py:251-267`` after ``training_step``,
        ``train_lightning.py:463-471``) is kept, the sampling just no longer waits for the backward pass it does
        not depend on.  ``next_seeds`` is the batch of the next call (the data loader's look-ahead); without it the
        step samples its own blocks first (prologue graph) and nothing is sampled ahead."""
        dm, g, smp = self.dm, self.dm.g, self.dm.sampler
        L = len(smp.nodes_per_layer)
        if self._sets is None or seeds.numel() != dm.batch_size:
            if self._sizing_steps < self.eager_warmup or self._max_src is None or seeds.numel() != dm.batch_size:
                return self._sizing_step(seeds)                # ordinary steps: size the pools / ragged batch
            self._alloc_pools()
        if not self._graphs:
            self._capture_full()
        if self._call_done is None:
            self._call_done = [torch.cuda.Event(), torch.cuda.Event()]
            self._copy_stream = torch.cuda.Stream()
        self._sync_lr()
        if self._next_ready and seeds is not self._prefetched_seeds:
            self._drop_prefetch()                              # the caller trains on another batch than announced
        if smp.step != self._dev_step_mirror:     # eager sampling in between (validation, ragged batch) advanced the
            self._step_dev.fill_(smp.step)        # host's Philox step: the device counter follows, draws never repeat
            self._dev_step_mirror = smp.step
        p = self._cur
        cur, nxt = self._sets[p], self._sets[1 - p]
        fresh = []                      # (kind, set): pinned counter buffers this call's graphs write (see _ctr_pin)
        if not self._next_ready:
            cur.seeds_in.copy_(seeds, non_blocking=True)
            self._replay(("S", p))
            smp.step += 1
            fresh.append(("S", p))
        prefetch = next_seeds is not None and next_seeds.numel() == dm.batch_size
        if prefetch:
            self._stage_seeds(nxt, next_seeds)
        if getattr(self, "_dp_one_graph", False) or (self._exchange is None and not (self.world > 1 or self._force_dp)):
            self._replay(("G" if prefetch else "GN", p))
        else:                                                  # data parallel: see _capture_full
            main = torch.cuda.current_stream()
            if self._exchange is not None and self._exchange.p2p:
                # peer-memory exchange: no collective call.  The apply + look-ahead sampling branch is launched at the
                # START of the step: its first kernel polls the flags the ranks' reward kernels raise during A1
                start = torch.cuda.Event()
                start.record(main)
                with torch.cuda.stream(self._side_apply):
                    self._side_apply.wait_event(start)
                    self._replay(("B1" if prefetch else "B1N", p))
                self._replay(("A1", p))
                self._replay(("A2", p))
            else:
                self._replay(("A1", p))
                a1_done = torch.cuda.Event()
                a1_done.record(main)
                work = None
                if self._exchange is not None:    # the bandit all-gather runs on NCCL's stream beside the backward pass
                    work = torch.distributed.all_gather_into_tensor(self._exchange.recv, self._exchange.send,
                                                                    group=self.pg, async_op=True)
                self._replay(("A2", p))
                with torch.cuda.stream(self._side_apply):      # … and so do the apply pass over all ranks' updates
                    self._side_apply.wait_event(a1_done)       # and the sampling of the next batch
                    if work is not None:
                        work.wait()
                    self._replay(("B1" if prefetch else "B1N", p))
            if not self._gradx:
                self.grads.all_reduce_mean_(self.pg)
            self._replay(("B2", 0))
            main.wait_stream(self._side_apply)
        if prefetch:
            smp.step += 1
            fresh.append(("G", 1 - p))
        self._dev_step_mirror = smp.step
        self._next_ready, self._prefetched_seeds = prefetch, (next_seeds if prefetch else None)
        self._cur = 1 - p
        self._static_loss, self._static_pred, self._static_y = self._static_out[p if prefetch or ("N", p) not in self._static_out
                                                                                else ("N", p)]
        bandit_mirror = getattr(smp, "_w_csc", None) is not None
        slot = self.graph_replays & 1
        self.graph_replays += 1
        self._call_done[slot].record()            # (counters and loss were copied to pinned memory inside the graphs)
        self._graph_loss = ("graph", (slot, p))
        for _, q in fresh:                        # … and the seed buffers its sampling read may be overwritten after it
            self._set_free[q] = self._call_done[slot]
        smp.tick_renorm(L, mirrored=bandit_mirror)
        if self.pipeline:                                      # consume the PREVIOUS call's counters
            prev, self._pending = self._pending, (slot, fresh)
            if prev is None:
                return self._static_loss
            slot, fresh = prev
        self._consume_counters(slot, fresh)
        return self._static_loss

    def _ctr_pin(self, kind, pset):
        """Pinned host copy of a pool set's counter blocks, one per (graph kind that samples into the set, set):
        S[p] and G[1-p] both sample into set p and may be in flight one call apart, so each has its own buffer; a
        buffer is re-written two calls after the call that consumes it (``_consume_counters``, one call late)."""
        smp = self.dm.sampler
        key = (kind, pset.ctr_base // 8)
        if key not in self._ctr_pins:
            self._ctr_pins[key] = torch.zeros((8, smp._wsp.ctr_all.shape[1]), dtype=torch.uint8).pin_memory()
        return self._ctr_pins[key]

    def _stage_seeds(self, pset, seeds):
        """The announced batch into the set's seed buffer.  Host seeds (the data loader's case) go through a pinned
        staging buffer on a copy stream that only waits for the last step that read the buffer — two calls back — so
        the copy runs beside the previous step instead of between two step graphs; a device tensor may have been
        produced on the current stream just now and is copied in stream order."""
        if seeds.is_cuda:
            pset.seeds_in.copy_(seeds, non_blocking=True)
            return
        main = torch.cuda.current_stream()
        i = pset.ctr_base // 8
        if self._seed_stage[i] is None:
            self._seed_stage[i] = torch.zeros(self.dm.batch_size, dtype=torch.int32).pin_memory()
            self._seed_staged[i] = torch.cuda.Event()
        elif self._seed_stage[i].numel() != seeds.numel():
            self._seed_staged[i].synchronize()
            self._seed_stage[i] = torch.zeros(seeds.numel(), dtype=torch.int32).pin_memory()
        else:
            self._seed_staged[i].synchronize()    # (the previous copy out of the staging buffer: two calls ago)
        self._seed_stage[i].copy_(seeds)
        cs = self._copy_stream
        if self._set_free[i] is not None:
            cs.wait_event(self._set_free[i])
        with torch.cuda.stream(cs):
            pset.seeds_in.copy_(self._seed_stage[i], non_blocking=True)
            self._seed_staged[i].record(cs)
        main.wait_event(self._seed_staged[i])

    def _replay(self, key):
        self._graphs[key].replay()
        self.graph_kernel_launches += self._graph_kernel_counts.get(key, 0)

    def flush(self):
        """Consume the counters of the last enqueued step (pipelined mode)."""
        if self._pending is not None:
            (slot, fresh), self._pending = self._pending, None
            self._consume_counters(slot, fresh)

    def _consume_counters(self, slot, fresh):
        """Sizes and capacity flags of the blocks sampled by one call (``fresh``: the pool sets it sampled into)."""
        dm, g, smp = self.dm, self.dm.g, self.dm.sampler
        L = len(smp.nodes_per_layer)
        grow = False
        from . import _native
        self._call_done[slot].synchronize()
        for kind, p in fresh:
            pset = self._sets[p] if self._sets is not None else None
            raw = self._ctr_pins[(kind, p)].numpy()
            ctrs = [_native.Counters.from_buffer_copy(raw[l].tobytes()) for l in range(L)]
            for l, c in enumerate(ctrs):
                if c.error:
                    cap = (pset.pools[l].cap_src, pset.pools[l].cap_edges) if pset is not None else ("?", "?")
                    msg = (f"static step: capacity of layer {l} exceeded (n_src {c.n_src}/{cap[0]}, "
                           f"edges {c.n_edges}/{cap[1]}); the step is invalid — "
                           "raise the pool margins or use static_graph=False")
                    if self.world > 1:            # abort on every rank together (at the next vote), not on this one alone
                        self._cap_error = getattr(self, "_cap_error", None) or msg
                        continue
                    raise RuntimeError(msg)
                self._max_src[l] = max(self._max_src[l], c.n_src)
                self._max_edges[l] = max(self._max_edges[l], c.n_edges)
                if pset is not None:
                    grow |= c.n_src > 0.92 * pset.pools[l].cap_src or c.n_edges > 0.92 * pset.pools[l].cap_edges
            smp.last_counters = ctrs
            self.num_steps += 1
            for i, c in enumerate(ctrs):
                self.cum_sampled_nodes[i] = self.cum_sampled_nodes[i] * self.w + c.n_src
                self.cum_sampled_edges[i] = self.cum_sampled_edges[i] * self.w + c.n_edges
            self.cum_sampled_nodes[L] = self.cum_sampled_nodes[L] * self.w + dm.batch_size
            self.total_sampled_edges += sum(int(c.n_edges) for c in ctrs)
            self.last_blocks = _CounterBlocks(ctrs)
        self.last_pred, self.last_labels = self._static_pred, self._static_y
        if self.world > 1:        # re-sizing allocates collectively: agree on it, every 32 steps
            self._grow_pending = getattr(self, "_grow_pending", False) or grow
            grow = False
            if self.num_steps % 32 == 0:
                grow = self._dp_vote()
        if grow:                                               # high-water mark: re-size before it can overflow
            self.pool_resizes += 1
            self._alloc_pools()

    def _dp_vote(self) -> bool:
        """Data parallel, every 32 steps: do the pools have to grow (any rank over its high-water mark), did a capacity
        overflow or an exchange time-out happen anywhere?  One 4-float MAX all-reduce whose result is read at the NEXT
        vote (pinned copy behind an event), so the host never waits for the device here; every rank acts on the same
        result at the same step (re-sizing and aborting are collective)."""
        dev = self.dm.g.device
        res = None
        if getattr(self, "_vote", None) is not None:
            self._vote.synchronize()
            res = self._vote_pin.clone()
        else:
            self._vote_dev = torch.zeros(4, dtype=torch.float32, device=dev)
            self._vote_stage = torch.zeros(4, dtype=torch.float32).pin_memory()
            self._vote_pin = torch.zeros(4, dtype=torch.float32).pin_memory()
            self._vote = torch.cuda.Event()
        self._vote_stage[0] = 1.0 if self._grow_pending else 0.0
        self._vote_stage[1] = 1.0 if getattr(self, "_cap_error", None) else 0.0
        self._vote_stage[2:] = 0.0
        self._grow_pending = False
        v = self._vote_dev
        v.copy_(self._vote_stage, non_blocking=True)
        if self._exchange is not None and self._exchange.p2p:
            v[2:3].copy_((self._exchange.err != 0).float())
        if self._gradx:
            v[3:4].copy_((self._gradx.err != 0).float())
        torch.distributed.all_reduce(v, op=torch.distributed.ReduceOp.MAX, group=self.pg)     # (stream-ordered)
        self._vote_pin.copy_(v, non_blocking=True)
        self._vote.record()
        if res is None:
            return False
        if res[1] > 0:
            raise RuntimeError(getattr(self, "_cap_error", None) or
                               "static step: a pool capacity was exceeded on another rank; the step is invalid — "
                               "raise the pool margins or use static_graph=False")
        if res[2] > 0:
            raise RuntimeError("peer-memory bandit exchange: a rank's update did not arrive within the timeout")
        if res[3] > 0:
            raise RuntimeError("peer-memory gradient exchange: a rank's gradient did not arrive within the timeout")
        return bool(res[0] > 0)

    def _dp_exchange(self):
        """The only exchanges of a data-parallel step: one all-reduce of the flat gradient buffer and one
        all-gather of the packed sparse bandit updates."""
        self.grads.all_reduce_mean_(self.pg)
        if self._exchange is not None:
            torch.distributed.all_gather_into_tensor(self._exchange.recv, self._exchange.send, group=self.pg)

    def _capture_full(self):
        """Captures, per pool set p: S[p] (sampling of one batch into set p: the prologue of an un-prefetched step),
        G[p] / GN[p] (the step on set p with / without sampling the next batch into set 1-p) — or, data parallel,
        A1[p] (forward + exponents of the bandit update written into the exchange's send buffer), A2[p] (backward),
        B1[p] / B1N[p] (apply every rank's gathered updates [+ sample the next batch]) and B2 (Adam), with the two
        collectives launched between them:

            A1 ─ all_gather (NCCL stream) ─ B1: apply all ranks' updates ─ sample_blocks(t+1)   ┐ side stream
            └─── A2: backward ─ all_reduce of the gradients ─ B2: Adam                          ┘ main stream
        """
        from . import _native
        dm, g, smp = self.dm, self.dm.g, self.dm.sampler
        L = len(smp.nodes_per_layer)
        smp._bind(g)
        if getattr(self, "_step_dev", None) is None:
            self._step_dev = torch.zeros(1, dtype=torch.int64, device=g.device)      # Philox step of the NEXT sampling
            self._drop_dev = torch.zeros(1, dtype=torch.int64, device=g.device)      # dropout's Philox step (one per step)
            if getattr(self.model, "_drop_step_t", None) is not None:               # continue the eager steps' count
                self._drop_dev.copy_(self.model._drop_step_t)
        self._step_dev.fill_(smp.step)
        if hasattr(self.model, "_drop_step"):
            self.model._drop_step_t, self.model._external_drop_step = self._drop_dev, True
        bandit_mode = smp._mode == _native.MODE_BANDIT
        for pset in self._sets:
            for l, pb in enumerate(pset.padded):
                pool = pset.pools[l]
                pb._n_edges_dev = smp._wsp.counter_ptr(pset.ctr_base + l, "n_edges")
                if bandit_mode:
                    pb.edata["q_ij"] = pool.e32[4].view(torch.float32)
                    pb.srcdata[smp.node_prob] = pool.node_prob
                dict.pop(pb.srcdata, "embed_norm", None)
                dict.pop(pb.edata, "a_ij", None)
        self.last_pred = None

        dp = self.world > 1 or self._force_dp
        bandit = "bandit" in dm.sampler_name
        if getattr(self, "_side_t", None) is None:
            # the step's main line (forward, backward, Adam) is captured on a high-priority stream; the side branch
            # (bandit update, next batch's sampling, transposes) on normal-priority ones
            # (BLISS_SAMPLE_PRIO=1 gives the priority to the sampling branch instead)
            sp = -5 if os.environ.get("BLISS_SAMPLE_PRIO") == "1" else 0
            self._side_t, self._side_s = torch.cuda.Stream(priority=sp), torch.cuda.Stream(priority=sp)
            self._side_b = torch.cuda.Stream()
            self._main_hp = torch.cuda.Stream(priority=0 if sp else -5)
            self._side_apply = torch.cuda.Stream(priority=sp)
        def clear_events(pset):
            for pb in pset.padded:                # events recorded in one capture must not be waited for in another
                pb._ready = pb._t_ready = None

        from .model import GCN
        # (GCN's forward pass reads the out-degrees off the transpose; data parallel, the main line's tail — gradient
        # exchange, Adam over all ranks' slots — is longer than the sampling chain's, so the transpose stays there)
        defer0 = not isinstance(getattr(self.model, "module", self.model), GCN) and not dp \
            or os.environ.get("BLISS_DEFER_T0") == "1"

        plan_ahead = os.environ.get("BLISS_PLAN_AHEAD", "1") != "0"

        def plan_top(pset):
            """The top layer's plan of the next sampling into ``pset`` (depends on the batch only): at the head of the
            step, on the stream the sampling will follow on."""
            smp.plan_top_static(g, pset.seeds_in, pset.pools, self._step_dev, ctr_base=pset.ctr_base,
                                ctr_mirror=self._ctr_pin("G", pset))

        def sample_into(pset, kind, layer_pre=None, planned=False):
            """Sampling of one batch (``pset.seeds_in``) into ``pset``: every layer's front half on the current stream,
            back halves and transposes on ``_side_t``; complete (joined) on return.  The input layer's transpose —
            the last thing on the sampling chain, read by the backward pass only — is left to the step that trains
            on the set (``launch_deferred``).  Every layer's finish kernel writes the layer's counters to the pinned buffer
            ``(kind, set)`` (device-mapped host memory), so nothing has to be enqueued between two step graphs."""
            cur = torch.cuda.current_stream()
            self._side_t.wait_stream(cur)
            with torch.cuda.stream(self._side_t):
                pset.seeds.copy_(pset.seeds_in, non_blocking=True)
            pset.deferred = smp.enqueue_static(g, pset.seeds_in, pset.pools, self._step_dev, transpose_stream=self._side_t,
                                               defer_last_transpose=defer0, ctr_base=pset.ctr_base, layer_pre=layer_pre,
                                               ctr_mirror=self._ctr_pin(kind, pset), top_planned=planned)
            self._step_dev.add_(1)                # (read by the layers' select kernels only: all launched by now)
            self._gather_inputs(pset, True)       # beside the input layer's fill / transposes (needs its source list only)
            cur.wait_stream(self._side_t)
            clear_events(pset)

        if self._loss_pins is None:
            self._loss_pins = [torch.zeros(1).pin_memory() for _ in range(2)]
            self._side_c = torch.cuda.Stream()

        def emit_loss(loss, p):
            """The step's loss to pinned host memory (``host_loss``): a 4-byte copy node beside the backward pass."""
            self._side_c.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side_c):
                self._loss_pins[p].copy_(loss.detach().reshape(1), non_blocking=True)

        def mirror_wmax(after):
            """The layers' running weight maxima to pinned memory, behind the step's last update (a copy node off the
            sampling chain): the host's range guard reads them without a sync (sampler.tick_renorm)."""
            self._side_c.wait_stream(after)
            with torch.cuda.stream(self._side_c):
                smp.mirror_wmax_()

        def launch_deferred(pset):
            """The input layer's transpose of the set this step trains on, beside the forward pass."""
            if pset.deferred:
                self._side_t.wait_stream(torch.cuda.current_stream())
                for fn in pset.deferred:
                    fn()                          # (on _side_t; records the block's _t_ready for the backward pass)

        def body_sample(p):
            sample_into(self._sets[p], "S")

        def body_step(p, prefetch):               # single rank: the whole step in one graph
            """forward ─ backward ─ Adam on the main stream.  A layer's bandit update is launched (side_b) as soon as the
            model has stored what it reads from the forward pass (that layer's embed_norm; a_ij for GAT); the top
            layer's update and, right behind it, the sampling of the NEXT batch (top layer first) go to side_s — so
            sampling starts while the last layer's forward pass is still running."""
            pset, other = self._sets[p], self._sets[1 - p]
            clear_events(pset)
            main = torch.cuda.current_stream()
            launch_deferred(pset)
            planned = bool(prefetch and plan_ahead)
            if planned:
                self._side_s.wait_stream(main)
                with torch.cuda.stream(self._side_s):
                    plan_top(other)
            inputs = (pset.x, pset.x_norm, pset.y) if pset.x is not None else None
            fired = []

            def make(l, pb):
                def hook():
                    cur = torch.cuda.current_stream()
                    fired.append(l)
                    if l == 0 and late0:
                        return                    # (launched with layer 1's, see below)
                    if l < L - 1:
                        self._side_b.wait_stream(cur)
                        with torch.cuda.stream(self._side_b):
                            if l == 1 and late0:
                                smp.update_exp3_weights(0, pset.padded[0], g)
                            smp.update_exp3_weights(l, pb, g)
                    else:                         # the top layer: its update, then the look-ahead sampling
                        self._side_s.wait_stream(cur)
                        if L > 1:
                            self._side_s.wait_stream(self._side_b)
                        with torch.cuda.stream(self._side_s):
                            smp.update_exp3_weights(l, pb, g)
                            mirror_wmax(self._side_s)
                            if prefetch:
                                sample_into(other, "G", planned=planned)
                return hook

            key = "a_ij" if smp.model == "gat" else "embed_norm"
            # The input layer's update (the largest: a random read-modify-write per edge of the 4096-fan-out block) could
            # start with the step — its embed_norm comes with the gathered inputs — but it would share the memory system
            # with the input layer's aggregation, the longest kernel of the forward pass, which the sampling chain is
            # waiting behind.  It is launched with layer 1's update instead, beside the smaller layers' forward pass.
            late0 = L > 2 and key == "embed_norm" and os.environ.get("BLISS_EXP3_L0_EARLY") != "1"
            if bandit:
                for l, pb in enumerate(pset.padded):
                    (pb.edata if key == "a_ij" else pb.srcdata).on_set[key] = make(l, pb)

            def after_forward():
                emit_loss(self._fwd_loss, p)
                if bandit:
                    assert sorted(fired) == list(range(L)), f"bandit update hooks fired for layers {fired}"
                elif prefetch:
                    self._side_s.wait_stream(main)
                    with torch.cuda.stream(self._side_s):
                        sample_into(other, "G", planned=planned)

            try:
                loss, pred, y = self._padded_fwd_bwd(True, after_forward=after_forward, pset=pset, inputs=inputs)
            finally:
                for pb in pset.padded:
                    pb.srcdata.on_set.pop("embed_norm", None)
                    pb.edata.on_set.pop("a_ij", None)
            self._drop_dev.add_(1)                # (before the joins: nothing but the joins is left behind the sampling chain)
            if bandit and L > 1:                  # (a stream that was not forked inside this capture must not be joined)
                main.wait_stream(self._side_b)
            if bandit or prefetch:
                main.wait_stream(self._side_s)
            if pset.deferred:
                main.wait_stream(self._side_t)
            main.wait_stream(self._side_c)
            return loss, pred, y

        early_emit = bandit and smp.model != "gat"      # GAT's alpha needs a_ij: emitted after the forward pass

        def body_a1(p):
            pset = self._sets[p]
            clear_events(pset)
            main = torch.cuda.current_stream()
            launch_deferred(pset)
            if early_emit:      # a layer's exponents are emitted as soon as the model has stored its embed_norm
                def make(l, pb):
                    def hook():
                        self._side_b.wait_stream(torch.cuda.current_stream())
                        with torch.cuda.stream(self._side_b):
                            smp.exp3_emit_layer(l, pb, g, self._exchange)
                    return hook
                for l, pb in enumerate(pset.padded):
                    pb.srcdata.on_set["embed_norm"] = make(l, pb)
            try:
                loss, pred, y = self._padded_fwd(pset, (pset.x, pset.x_norm, pset.y) if pset.x is not None else None)
            finally:
                for pb in pset.padded:
                    pb.srcdata.on_set.pop("embed_norm", None)
            if bandit and not early_emit:
                self._side_b.wait_stream(main)
                with torch.cuda.stream(self._side_b):
                    smp.exp3_emit(pset.padded, g, self._exchange)
            if bandit:
                main.wait_stream(self._side_b)
            if pset.deferred:                     # (the backward pass is its own graph: joined here, event dropped)
                main.wait_stream(self._side_t)
                clear_events(pset)
            emit_loss(loss, p)
            main.wait_stream(self._side_c)
            return loss, pred.detach(), y

        def body_a2(loss):
            self._zero_grads()
            ops.LAST_SPMM_BWD = None          # (an event of an earlier pass / capture must not be waited for)
            loss.backward()

        p2p = bandit and self._exchange is not None and self._exchange.p2p

        def body_b1(p, prefetch):
            """All ranks' updates applied, then the next batch sampled.  Peer-memory exchange: every layer's apply kernel
            first waits for all ranks' flags of that layer; the layers are applied in the order the forward pass emits
            them (input layer first), so the big input-layer passes run beside the forward pass and only the small top
            layer's is left when the top layer's flags arrive and sampling (top layer first) can start."""
            if bandit:
                smp.exp3_apply(self._exchange, L)
                smp.mirror_wmax_()
            if prefetch:
                sample_into(self._sets[1 - p], "G")
            if not (bandit or prefetch):
                self._drop_dev.add_(0)            # (a captured graph must hold at least one node)

        gradx = self._gradx if self._gradx else None

        def body_b2(adds=True):
            if gradx is not None:                 # push the flat gradient into every rank's window; Adam adds the slots
                gradx.push(self._flat_grad)
                self.optimizer.step_p2p(gradx)
                self._grads_clean = True
            else:
                self._optimizer_step()
            if adds:
                self._drop_dev.add_(1)
                if p2p:
                    self._exchange.step_dev.add_(1)   # next step: other parity half of the windows, next flag value

        # Both exchanges through peer memory: no collective call is left between the parts of the step, so the whole
        # data-parallel step is ONE graph like the single-rank one (three graph boundaries fewer on the main line).
        one_graph = dp and gradx is not None and (p2p or not bandit)
        self._dp_one_graph = one_graph

        def body_step_p2p(p, prefetch):
            """main: forward - backward - gradient push - Adam over all ranks' slots.  A layer's updates are emitted into
            every rank's window (side_b) as soon as the model has stored its embed_norm; right behind the local emit
            (side_apply) that layer's wait-for-all-ranks + apply, and behind the top layer's the sampling of the next
            batch.  A wait kernel is therefore never scheduled ahead of the local kernel whose flag it waits for."""
            pset, other = self._sets[p], self._sets[1 - p]
            clear_events(pset)
            main = torch.cuda.current_stream()
            launch_deferred(pset)
            forked = []
            planned = bool(prefetch and plan_ahead)
            if planned:                           # the next batch's top-layer plan: needs the batch only
                self._side_apply.wait_stream(main)
                forked.append(True)
                with torch.cuda.stream(self._side_apply):
                    plan_top(other)

            def apply_from(l0, l1, after):
                """Layers l0..l1-1 applied behind ``after`` (the stream their local emit ran on); then, behind the top
                layer, the look-ahead sampling."""
                self._side_apply.wait_stream(after)
                forked.append(True)
                with torch.cuda.stream(self._side_apply):
                    for l in range(l0, l1):
                        smp.exp3_apply_layer(self._exchange, l)
                    if l1 == L:
                        mirror_wmax(self._side_apply)
                    if l1 == L and prefetch:
                        sample_into(other, "G", planned=planned)

            if early_emit:
                def make(l, pb):
                    def hook():
                        self._side_b.wait_stream(torch.cuda.current_stream())
                        with torch.cuda.stream(self._side_b):
                            smp.exp3_emit_layer(l, pb, g, self._exchange)
                        apply_from(l, l + 1, self._side_b)
                    return hook
                for l, pb in enumerate(pset.padded):
                    pb.srcdata.on_set["embed_norm"] = make(l, pb)
            try:
                loss, pred, y = self._padded_fwd(pset, (pset.x, pset.x_norm, pset.y) if pset.x is not None else None)
            finally:
                for pb in pset.padded:
                    pb.srcdata.on_set.pop("embed_norm", None)
            emit_loss(loss, p)
            if bandit and not early_emit:         # GAT: alpha needs a_ij, known after the forward pass
                self._side_b.wait_stream(main)
                with torch.cuda.stream(self._side_b):
                    smp.exp3_emit(pset.padded, g, self._exchange)
                apply_from(0, L, self._side_b)
            elif not bandit and prefetch:
                self._side_apply.wait_stream(main)
                forked.append(True)
                with torch.cuda.stream(self._side_apply):
                    sample_into(other, "G", planned=planned)
            self._zero_grads()
            ops.LAST_SPMM_BWD = None          # (an event of an earlier pass / capture must not be waited for)
            loss.backward()
            body_b2(adds=False)
            self._drop_dev.add_(1)
            if bandit:
                main.wait_stream(self._side_b)
            if forked:
                main.wait_stream(self._side_apply)
            if pset.deferred:
                main.wait_stream(self._side_t)
            main.wait_stream(self._side_c)
            if p2p:
                self._exchange.step_dev.add_(1)   # (after the apply kernels, which read the parity from it)
            return loss.detach(), pred.detach(), y

        def dp_step_eager(p, prefetch):
            if one_graph:
                return body_step_p2p(p, prefetch)
            loss_w, _, _ = body_a1(p)
            if self._exchange is not None and not p2p:
                torch.distributed.all_gather_into_tensor(self._exchange.recv, self._exchange.send, group=self.pg)
            body_a2(loss_w)
            body_b1(p, prefetch)
            if gradx is None:
                self.grads.all_reduce_mean_(self.pg)
            body_b2()

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        state = (smp.state_dict()["exp3_w_csc"].clone(), smp._l1.clone()) if smp._w_csc is not None else None
        with torch.cuda.stream(side):                       # warm-up runs off the default stream …
            import copy
            opt_state = copy.deepcopy(self.optimizer.state_dict())
            params = [p_.detach().clone() for p_ in self.grads.params]
            drop0 = self._drop_dev.clone()
            for pset in self._sets:
                pset.seeds_in.copy_(self.dm.train_nid[: dm.batch_size])
            body_sample(0)
            for p in (0, 1):
                if dp:
                    dp_step_eager(p, True)
                else:
                    body_step(p, True)
            # … and must not change the training state: restore parameters, Adam moments, bandit weights
            for p_, q in zip(self.grads.params, params):
                p_.data.copy_(q)
            self.optimizer.load_state_dict(opt_state)
            if state is not None:
                for l, w in enumerate(state[0]):
                    smp._w_csc[l].copy_(w)
                smp._l1.copy_(state[1])
            self._step_dev.fill_(smp.step)
            self._dev_step_mirror = smp.step
            self._drop_dev.copy_(drop0)
        torch.cuda.current_stream().wait_stream(side)

        graphs, counts, outs = {}, {}, {}

        def capture(key, fn, pool=None, stream=None):
            before = _native.STATS.launches
            gr = torch.cuda.CUDAGraph()
            kw = {}
            if pool is not None:
                kw["pool"] = pool
            with torch.cuda.graph(gr, stream=stream or self._main_hp, **kw):
                res = fn()
            counts[key] = _native.STATS.launches - before
            _native.STATS.launches = before
            graphs[key] = gr
            return res

        # (Capturing NCCL's collectives into ONE graph with both halves was measured at N=2: no faster than replays
        # with the collectives launched in between, and the process hung in NCCL teardown.)
        for p in (0, 1):
            capture(("S", p), lambda: body_sample(p))
            if not dp:
                outs[p] = capture(("G", p), lambda: body_step(p, True))
                outs[("N", p)] = capture(("GN", p), lambda: body_step(p, False))
            elif one_graph:
                outs[p] = capture(("G", p), lambda: body_step_p2p(p, True))
                outs[("N", p)] = capture(("GN", p), lambda: body_step_p2p(p, False))
            else:
                loss, pred, y = capture(("A1", p), lambda: body_a1(p))
                outs[p] = (loss.detach(), pred, y)
                # the backward pass of A1's autograd graph: same memory pool
                capture(("A2", p), lambda: body_a2(loss), pool=graphs[("A1", p)].pool())
                del loss
                capture(("B1", p), lambda: body_b1(p, True), stream=self._side_apply)
                capture(("B1N", p), lambda: body_b1(p, False), stream=self._side_apply)
        if dp and not one_graph:
            capture(("B2", 0), body_b2)
        self._graphs, self._graph_kernel_counts, self._static_out = graphs, counts, outs
        self.graph_kernels = counts.get(("G", 0), 0) or sum(counts.get((k, 0), 0) for k in ("A1", "A2", "B1", "B2"))
        # (ranks enter the first replay together: a rank still capturing would leave its peers' wait kernels spinning)
        if dp and torch.distributed.is_initialized():
            torch.cuda.synchronize()
            torch.distributed.barrier(group=self.pg)

    # ---- checkpoint / resume (the reference checkpoints the model only; the bandit state is part of training) ----
    def state_dict(self):
        """Everything a resumed run needs to continue the same trajectory: parameters, Adam moments and step,
        lr schedule, the sampler's EXP3 weights / L1 norms / Philox step, and the step counters."""
        self.flush()
        self._drop_prefetch()       # blocks sampled ahead are re-drawn after a restore (same Philox step, same weights)
        smp = self.dm.sampler
        return {"model": {k: v.detach().clone() for k, v in self.model.state_dict().items()},
                "optimizer": self.optimizer.state_dict(), "scheduler": self.scheduler.state_dict(),
                "sampler": smp.state_dict() if getattr(smp, "_w_csc", None) is not None else {"step": smp.step},
                "num_steps": self.num_steps, "cum_nodes": list(self.cum_sampled_nodes),
                "cum_edges": list(self.cum_sampled_edges), "epoch": self.dm._epoch,
                "drop_step": (int(self.model._drop_step_t.item())
                              if getattr(self.model, "_drop_step_t", None) is not None else 0)}

    def load_state_dict(self, sd):
        """In place: parameters, moments and bandit weights keep their addresses, so a captured step graph
        stays valid; the device-side Philox step follows at the next step."""
        self.flush()
        self._drop_prefetch()
        smp = self.dm.sampler
        self.model.load_state_dict(sd["model"])
        self.optimizer.load_state_dict(sd["optimizer"])
        self.scheduler.load_state_dict(sd["scheduler"])
        if "exp3_w_csc" in sd["sampler"]:
            smp.load_state_dict(sd["sampler"], self.dm.g)
        else:
            smp.step = int(sd["sampler"]["step"])
        self.num_steps = int(sd["num_steps"])
        self.cum_sampled_nodes, self.cum_sampled_edges = list(sd["cum_nodes"]), list(sd["cum_edges"])
        self.dm._epoch = int(sd["epoch"])
        if getattr(self.model, "_drop_step_t", None) is not None:      # dropout's Philox step (eager and captured path)
            self.model._drop_step_t.fill_(int(sd.get("drop_step", 0)))
        elif hasattr(self.model, "_drop_step") and next(self.model.parameters()).is_cuda:
            self.model._drop_step(next(self.model.parameters()).device).fill_(int(sd.get("drop_step", 0)))
        self._grads_clean = False
        self._dev_step_mirror = None

    @torch.no_grad()
    def validate(self) -> float:
        """Per-epoch validation with the same stochastic sampler (``train_lightning.py:179-203,410-422``)."""
        self.flush()
        self._drop_prefetch()
        smp = self.dm.sampler
        seed = smp.rng_seed
        smp.rng_seed = seed & 0xFFFFFFFF          # every rank validates with rank 0's draws: one val_acc, one stop decision
        self.model.eval()
        tp = fp_fn = 0.0
        try:
            for seeds in self.dm.val_batches():
                _, _, mfgs = smp.sample_blocks(self.dm.g, seeds)
                pred = self.model(mfgs, mfgs[0].srcdata["features"])
                y = mfgs[-1].dstdata["labels"]
                # micro-F1 accumulated over the epoch like torchmetrics (:68-72,179-203): counts, not a mean of batch scores
                if self.dm.multilabel:
                    p, t = torch.sigmoid(pred) > 0.5, y > 0.5
                    tp += float((p & t).sum())
                    fp_fn += float(p.sum() + t.sum())
                else:
                    tp += float((pred.argmax(1) == y).sum())
                    fp_fn += 2.0 * y.shape[0]
        finally:
            smp.rng_seed = seed
            self.model.train()
        return 2.0 * tp / max(fp_fn, 1.0)


class _PoolSet:
    """One of the two buffer sets of the pipelined static step: per-layer capacity pools, the capacity-padded
    blocks over them, the seed buffer and the first of its counter blocks in the sampler workspace."""

    def __init__(self, pools, padded, seeds, ctr_base):
        self.pools, self.padded, self.seeds, self.ctr_base = pools, padded, seeds, ctr_base
        # ``seeds_in`` is where a batch arrives and what the sampling graph reads; that graph also copies it to
        # ``seeds`` (the top block's destination ids, read by the step that trains on the set), so the next batch but
        # one can be staged into ``seeds_in`` while that step is still running
        self.seeds_in = torch.zeros_like(seeds)
        self.deferred = []                        # the input layer's transpose, launched by the step that trains on the set
        self.x = self.x_norm = self.y = None      # input features / their row norms / labels, gathered with the blocks


class _CounterBlocks(list):
    """What ``Trainer.last_blocks`` holds after a whole-step graph replay: the per-layer sizes (the
    blocks themselves live in the capacity pools)."""

    class _B:
        def __init__(self, c):
            self._c = c

        def num_src_nodes(self):
            return int(self._c.n_src)

        def num_dst_nodes(self):
            return int(self._c.n_seeds)

        def num_edges(self):
            return int(self._c.n_edges)

    def __init__(self, ctrs):
        super().__init__(self._B(c) for c in ctrs)


def build_argparser() -> argparse.ArgumentParser:
    """The reference's 30 flags, names and defaults verbatim (``train_lightning.py:489-552``), plus
    ``--seed`` and ``--normalize``; ``--dataset`` also accepts ``synthetic:<name>[:scale]``."""
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpu", type=int, default=0 if torch.cuda.is_available() else -1)
    ap.add_argument("--model", type=str, default="sage", choices=["sage", "gcn", "gat"])
    ap.add_argument("--dataset", type=str, default="cora")
    ap.add_argument("--num-epochs", type=int, default=-1)
    ap.add_argument("--num-steps", type=int, default=-1)
    ap.add_argument("--min-steps", type=int, default=0)
    ap.add_argument("--num-hidden", type=int, default=256)
    ap.add_argument("--num-layers", type=int, default=3)
    ap.add_argument("--num-in-heads", type=int, default=4)
    ap.add_argument("--num-out-heads", type=int, default=1)
    ap.add_argument("--attn-dropout", type=float, default=0.1)
    ap.add_argument("--negative-slope", type=float, default=0.2)
    ap.add_argument("--residual", action="store_true", default=False)
    ap.add_argument("--allow-zero-in-degree", action="store_true", default=False)
    ap.add_argument("--fan-out", type=str, default="16384,8192,4096")
    ap.add_argument("--eta", type=float, default=0.1)
    ap.add_argument("--batch-size", type=int, default=1024)
    ap.add_argument("--lr", type=float, default=0.002)
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--num-workers", type=int, default=0)
    ap.add_argument("--data-cpu", action="store_true")
    ap.add_argument("--sampler", type=str, default="poisson-bandit",
                    choices=["full", "neighbor", "bandit", "poisson-bandit", "ladies", "poisson-ladies"])
    ap.add_argument("--importance-sampling", type=int, default=1)
    ap.add_argument("--logdir", type=str, default="tb_logs")
    ap.add_argument("--vertex-limit", type=int, default=-1)
    ap.add_argument("--use-uva", action="store_true")
    ap.add_argument("--cache-size", type=int, default=0)
    ap.add_argument("--undirected", action="store_true")
    ap.add_argument("--val-acc-target", type=float, default=1)
    ap.add_argument("--early-stopping-patience", type=int, default=1000)
    ap.add_argument("--disable-checkpoint", action="store_true")
    ap.add_argument("--precision", type=str, default="medium")
    ap.add_argument("--k-runs", type=int, default=1)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--normalize", type=str, default="lazy", choices=["lazy", "literal"])
    return ap


class BatchSizeController:
    """``BatchSizeCallback`` (``train_lightning.py:425-486``) without Lightning: running mean / variance of the
    number of input nodes per batch (Welford, ``:437-451``); at the end of an epoch, when ``--vertex-limit`` is set and
    the mean is off the limit by more than ``factor`` standard errors, the batch size is rescaled by
    ``limit / mean`` (``:473-486``)."""

    def __init__(self, limit, factor=3):
        self.limit, self.factor = limit, factor
        self.clear()

    def clear(self):
        self.n, self.m, self.s = 0, 0.0, 0.0

    def push(self, x):
        self.n += 1
        m = self.m
        self.m += (x - m) / self.n
        self.s += (x - m) * (x - self.m)

    @property
    def var(self):
        return self.s / (self.n - 1)

    @property
    def std(self):
        return math.sqrt(self.var)

    def on_train_epoch_end(self, batch_size: int) -> int:
        """The batch size of the next epoch (unchanged unless the limit test fires)."""
        if self.limit > 0 and self.n >= 2 and abs(self.limit - self.m) * self.n >= self.std * self.factor:
            batch_size = max(1, int(batch_size * self.limit / self.m))
            self.clear()
        return batch_size


class EarlyStopping:
    """Lightning ``EarlyStopping(monitor='val_acc', mode='max', stopping_threshold, patience)`` as the reference
    configures it (``train_lightning.py:627-634``): stop once the metric EXCEEDS the threshold, or after ``patience``
    validation checks without a new best."""

    def __init__(self, stopping_threshold, patience):
        self.threshold, self.patience = stopping_threshold, patience
        self.best, self.wait = -float("inf"), 0

    def check(self, value: float) -> bool:
        if value > self.best:
            self.best, self.wait = value, 0
        else:
            self.wait += 1
        return (self.threshold is not None and value > self.threshold) or self.wait >= self.patience


_NO_EFFECT_FLAGS = {   # accepted for command-line compatibility with the reference, inert here — said once, loudly
    "data_cpu": "the graph and its features are resident in HBM (the hot path has no CPU path)",
    "use_uva": "the graph and its features are resident in HBM: nothing to address through UVA",
    "cache_size": "no host-side feature store, hence no GPU feature cache",
    "num_workers": "sampling runs on the GPU inside the step graph: there are no sampler worker processes",
    "allow_zero_in_degree": "parsed and unused by the reference as well (train_lightning.py:510-511)",
}


def main(argv=None):
    import json
    import sys
    args = build_argparser().parse_args(argv)
    if args.gpu < 0 or not torch.cuda.is_available():
        raise SystemExit("this build runs the BLISS hot path on a CUDA device only (--gpu >= 0); "
                         "the CPU restatement lives in oracle/ and is test infrastructure")
    if args.precision != "highest":
        torch.set_float32_matmul_precision(args.precision)                       # :554-555
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", args.gpu))
    defaults = build_argparser().parse_args([])
    for flag, why in _NO_EFFECT_FLAGS.items():
        if rank == 0 and getattr(args, flag) != getattr(defaults, flag):
            print(f"warning: --{flag.replace('_', '-')} has no effect in this build: {why}", file=sys.stderr)
    pg = None
    if world > 1:
        torch.distributed.init_process_group("nccl")
        pg = torch.distributed.group.WORLD
    device = torch.device(f"cuda:{local}")
    torch.cuda.set_device(device)
    subdir = "paper_{}_{}_{}_{}_steps_{}_bs_{}_layers_{}_lr_{}_eta_{}".format(       # :636-646
        args.model, args.dataset.replace(":", "-"), args.sampler, args.importance_sampling, args.num_steps, args.batch_size,
        args.num_layers, args.lr, args.eta)
    results = []
    for run in range(args.k_runs):                                               # :562
        if rank == 0:
            print("=" * 20 + f"run_{run + 1} for eta_{args.eta}" + "=" * 20)
        dm = DataModule(args.dataset, args.undirected, args.data_cpu, args.use_uva,
                        [int(_) for _ in args.fan_out.split(",")], args.eta, device, args.batch_size,
                        args.num_workers, args.sampler, args.importance_sampling, args.cache_size, args.num_steps,
                        args.model, seed=args.seed + run, rank=rank, world_size=world, normalize=args.normalize)
        torch.manual_seed(args.seed + 3 + run)
        model = build_model(args.model, dm.in_feats, args.num_hidden, dm.n_classes, args.num_layers, args.dropout,
                            args.num_in_heads, args.num_out_heads, args.attn_dropout, args.negative_slope,
                            args.residual).to(device)
        tr = Trainer(dm, model, args.lr, pg, static_graph=True)     # whole-step CUDA graphs (ragged batches run eagerly)
        logdir = None
        if rank == 0:                                               # TensorBoardLogger(logdir, name=subdir) -> version_k
            base = os.path.join(args.logdir, subdir)
            os.makedirs(base, exist_ok=True)
            vers = [int(d.split("_")[-1]) for d in os.listdir(base) if d.startswith("version_")]
            logdir = os.path.join(base, f"version_{max(vers) + 1 if vers else 0}")
            os.makedirs(os.path.join(logdir, "checkpoints"), exist_ok=True)
            json.dump(vars(args), open(os.path.join(logdir, "hparams.json"), "w"), indent=1)
        log = open(os.path.join(logdir, "metrics.jsonl"), "w") if logdir else None

        def emit(**kw):
            if log:
                log.write(json.dumps(kw) + "\n")
                log.flush()

        stopper = EarlyStopping(args.val_acc_target, args.early_stopping_patience)
        limiter = BatchSizeController(args.vertex_limit)
        best_ckpt, best_val = None, -float("inf")
        step, epoch, done = 0, 0, False
        t_prev = time.time()
        while not done:
            batches = list(dm.train_batches())
            if not batches:
                raise SystemExit(f"no full training batch: {dm.train_nid.numel()} training nodes, batch size {dm.batch_size}")
            for i, seeds in enumerate(batches):
                # the next batch of the epoch is announced: its blocks are sampled beside this step's backward pass
                loss = tr.training_step(seeds, batches[i + 1] if i + 1 < len(batches) else None)
                step += 1
                if tr.last_blocks is not None:
                    limiter.push(tr.last_blocks[0].num_src_nodes())              # :464 (one step late when pipelined)
                if rank == 0 and (step % 50 == 0 or step == 1):
                    acc = micro_f1(tr.last_pred.detach(), tr.last_labels, dm.multilabel)
                    t = time.time()
                    edges = sum(tr.num_sampled_edges(i_) for i_ in range(len(tr.cum_sampled_edges)))
                    print(f"step {step} loss {loss.item():.4f} train_acc {acc:.4f} iter_time {(t - t_prev):.4f} "
                          f"num_edges {edges:.0f}")
                    emit(step=step, train_loss=float(loss.item()), train_acc=acc, num_edges=edges)
                    t_prev = t
                if 0 < args.num_steps <= step:
                    done = True
                    break
            epoch += 1
            tr.scheduler.step()                                                  # StepLR per epoch (:205-216)
            new_bs = limiter.on_train_epoch_end(dm.batch_size)                   # --vertex-limit (:473-486)
            if new_bs != dm.batch_size:
                if rank == 0:
                    print(f"epoch {epoch}: vertex limit {args.vertex_limit}: batch size {dm.batch_size} -> {new_bs}")
                tr.set_batch_size(new_bs)
            if dm.val_nid.numel():
                val = tr.validate()
                stop = stopper.check(val)
                if world > 1:     # every rank validated the same batches with the same draws; agree on the decision anyway
                    flag = torch.tensor([1.0 if stop else 0.0, val], device=device)
                    torch.distributed.broadcast(flag, src=0, group=pg)
                    stop, val = bool(flag[0].item() > 0), float(flag[1].item())
                if rank == 0:
                    print(f"epoch {epoch} val_acc {val:.4f}")
                    emit(epoch=epoch, step=step, val_acc=val)
                if val > best_val:                       # ModelCheckpoint(monitor='val_acc', save_top_k=1, mode='max') :622-625
                    best_val = val
                    if not args.disable_checkpoint and rank == 0:
                        best_ckpt = os.path.join(logdir, "checkpoints", f"epoch={epoch - 1}-step={step}.ckpt")
                        for old in os.listdir(os.path.dirname(best_ckpt)):
                            os.remove(os.path.join(os.path.dirname(best_ckpt), old))
                        torch.save(tr.state_dict(), best_ckpt)
                if stop and step >= args.min_steps:                              # EarlyStopping :627-634, min_steps :654
                    done = True
            if 0 < args.num_epochs <= epoch:
                done = True
        tr.flush()
        if not args.disable_checkpoint and dm.val_nid.numel():                   # reload the best checkpoint :662-685
            if world > 1:
                box = [best_ckpt]
                torch.distributed.broadcast_object_list(box, src=0, group=pg)
                best_ckpt = box[0]
            if best_ckpt is not None:
                if rank == 0:
                    print("Evaluating model in", os.path.dirname(os.path.dirname(best_ckpt)))
                model.load_state_dict(torch.load(best_ckpt, map_location=device)["model"])
        with torch.no_grad():                                                    # :686-705
            pred = model.inference(dm.g, device, 128, args.use_uva, args.num_workers)
            out = {}
            for nid, split in zip([dm.train_nid, dm.val_nid, dm.test_nid], ["Train", "Validation", "Test"]):
                if nid.numel():
                    out[split] = micro_f1(pred[nid.long()], dm.g.ndata["labels"][nid.long()], dm.multilabel)
                    if rank == 0:
                        print(f"{split} accuracy: {out[split]}")
        emit(final=out, steps=step, epochs=epoch)
        if log:
            log.close()
        results.append(out)
    if args.k_runs > 1 and rank == 0:       # the reference reduces the k runs' TensorBoard logs to mean / std (:711-733)
        for split in results[0]:
            vals = torch.tensor([r[split] for r in results], dtype=torch.float64)
            print(f"{split} accuracy over {args.k_runs} runs: mean {vals.mean():.4f} std {vals.std(unbiased=False):.4f}")
    if world > 1:
        torch.distributed.destroy_process_group()
    return results


if __name__ == "__main__":
    main()
