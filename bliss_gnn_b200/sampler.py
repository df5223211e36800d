"""Drop-in samplers: ``BanditLadiesSampler`` / ``PoissonBanditLadiesSampler``
(reference ``bandit_sampler.py:29-424``) and ``LadiesSampler`` / ``PoissonLadiesSampler``
(``ladies_sampler.py:24-183``) over the sm_100a kernels of ``csrc/sampler.cu`` and
``csrc/bandit.cu``.

Same class names, constructor signatures, public attributes and method names as the reference;
``sample_blocks(g, seed_nodes, exclude_eids=None) -> (input_nodes, output_nodes, blocks)`` is the
``dgl.dataloading.BlockSampler`` protocol and ``exp3(mfgs, g)`` the post-step bandit update
(``train_lightning.py:463-471``).  What differs underneath:

* the stage methods (``exp3_probabilities`` → ``compute_prob`` → ``select_neighbors`` →
  ``generate_block``) pass a :class:`Frontier` handle (device workspace) instead of DGL sub-graphs;
* EXP3 weights live CSC-ordered and un-normalised with a running L1 norm (``normalize='lazy'``);
  ``exp3_weights`` exposes the reference's ``[L, |E|]`` edge-id-ordered normalised view.
  ``normalize='literal'`` re-normalises densely after every update like ``bandit_sampler.py:249``;
* randomness is one Philox4x32-10 draw per (seed, step, layer, node id) — ``rng_seed`` — or an
  injected dense ``[|V|]`` array of uniforms (``inject_uniforms``), never a torch generator.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

import math

import torch

from . import _native as N
from .graph import EID, NID, Block, Graph


class _Workspace:
    """Device buffers of one (sampler, graph) pair; every array is |V|-sized so no capacity can be
    exceeded, and nothing |V|-sized is cleared per step (the finish kernel restores the
    invariant for the touched nodes)."""

    def __init__(self, g: Graph):
        dev = g.device
        if dev.type != "cuda":
            raise RuntimeError("the BLISS samplers run on a CUDA graph only (no CPU fallback): use g.to('cuda')")
        V = g.num_nodes()
        i32 = dict(dtype=torch.int32, device=dev)
        self.acc = torch.empty(V, dtype=torch.int64, device=dev)
        self.first_pos = torch.empty(V, dtype=torch.int64, device=dev)
        self.node_info = torch.empty(2 * V, **i32)
        self.sel_bits = torch.empty((V + 31) // 32, **i32)
        self.cand_bits = torch.empty((V + 31) // 32, **i32)
        n_chunk_cap = g.num_edges() // 256 + V + 8          # every column cut into 256-edge chunks
        self.keep_bits = torch.zeros(8 * n_chunk_cap, **i32)
        self.row_a = torch.empty(V, dtype=torch.int64, device=dev)
        self.row_d = torch.empty(V, **i32)
        self.chunk_first = torch.empty(V + 1, **i32)
        self.chunk_rec = torch.empty(n_chunk_cap * 8, **i32)      # 32-byte record per chunk
        self.part_w = torch.empty(n_chunk_cap, dtype=torch.float64, device=dev)
        self.part_q = torch.empty(n_chunk_cap, dtype=torch.float64, device=dev)
        self.cand = torch.empty(V, **i32)
        self.p_cand = torch.empty(V, dtype=torch.float32, device=dev)
        self.sel = torch.empty(V, **i32)
        self.row_w = torch.empty(V, dtype=torch.float32, device=dev)
        self.row_q = torch.empty(V, dtype=torch.float32, device=dev)
        self.row_cnt = torch.empty(V, **i32)
        self.part_cnt = torch.empty(n_chunk_cap, **i32)
        self.part_t = torch.empty(n_chunk_cap, dtype=torch.float64, device=dev)
        self.chunk_pre = torch.empty(n_chunk_cap, **i32)
        self.row_t = torch.empty(V, dtype=torch.float32, device=dev)
        self.src_nid = torch.empty(V, **i32)
        self.node_prob = torch.empty(V, dtype=torch.float32, device=dev)
        self.key_scratch = None
        self.MAX_LAYERS = 16
        csz = C.sizeof(N.Counters)
        self.ctr_all = torch.zeros(self.MAX_LAYERS, csz, dtype=torch.uint8, device=dev)   # one block per layer
        self.ctr = self.ctr_all[0]
        self.ctr_host = torch.zeros(csz, dtype=torch.uint8).pin_memory()
        self.ctr_all_host = torch.zeros(self.MAX_LAYERS, csz, dtype=torch.uint8).pin_memory()
        self.ws = N.Workspace(
            acc=N.ptr(self.acc), first_pos=N.ptr(self.first_pos), node_info=N.ptr(self.node_info),
            sel_bits=N.ptr(self.sel_bits), cand_bits=N.ptr(self.cand_bits), keep_bits=N.ptr(self.keep_bits), cand=N.ptr(self.cand), p_cand=N.ptr(self.p_cand), sel=N.ptr(self.sel),
            row_a=N.ptr(self.row_a), row_d=N.ptr(self.row_d),
            chunk_first=N.ptr(self.chunk_first), chunk_rec=N.ptr(self.chunk_rec), part_w=N.ptr(self.part_w),
            part_q=N.ptr(self.part_q), row_w=N.ptr(self.row_w), row_q=N.ptr(self.row_q),
            row_cnt=N.ptr(self.row_cnt), part_cnt=N.ptr(self.part_cnt), part_t=N.ptr(self.part_t), chunk_pre=N.ptr(self.chunk_pre),
            row_t=N.ptr(self.row_t), cap_seeds=V, cap_sel=V, ctr=N.ptr(self.ctr))
        self.gview = N.Graph(num_nodes=V, num_edges=g.num_edges(), indptr=N.ptr(g.indptr),
                             indices=N.ptr(g.indices), eid=N.ptr(g.eid))
        self._keep = (g.indptr, g.indices, g.eid)
        N.call("bliss_workspace_init", C.byref(self.ws), V, N.stream())

    def ws_layer(self, layer: int, n_seeds_dev=None, step_dev=None, counters=None, ctr_mirror=None) -> "N.Workspace":
        """The workspace descriptor with layer ``layer``'s own counters block and, for sync-free
        chaining, the device addresses the kernels read the seed count / Philox step from.
        ``counters``: the workspace whose per-layer counters blocks to use (the step's second workspace
        shares the first one's)."""
        ws = N.Workspace.from_buffer_copy(self.ws)
        ws.ctr = (counters or self).ctr_all[layer].data_ptr()
        ws.n_seeds_dev = n_seeds_dev
        ws.step_dev = step_dev
        ws.ctr_mirror = ctr_mirror                # (pinned host address or None: see include/bliss_b200.h)
        return ws

    def counter_ptr(self, layer: int, field: str) -> int:
        return self.ctr_all[layer].data_ptr() + getattr(N.Counters, field).offset

    def read_all_counters(self, n_layers: int):
        """One D2H copy + one stream sync for all layers' counters (end of a sync-free step)."""
        self.ctr_all_host.copy_(self.ctr_all, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        raw = self.ctr_all_host.numpy()
        return [N.Counters.from_buffer_copy(raw[l].tobytes()) for l in range(n_layers)]

    def read_counters(self) -> N.Counters:
        self.ctr_host.copy_(self.ctr, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return N.Counters.from_buffer_copy(self.ctr_host.numpy().tobytes())


class LayerPool:
    """Persistent capacity-sized buffers of one layer's block (static addresses, so the model's
    forward / backward over the *padded* block can be captured once in a CUDA graph and replayed).
    Rows beyond the sampled counts are valid padding: empty destination rows (indptr tail = E_b,
    mean divisor 1), source ids that stay valid node ids, edge slots beyond E_b never referenced."""

    def __init__(self, device, cap_dst: int, cap_src: int, cap_edges: int, bandit: bool = True):
        self.cap_dst, self.cap_src, self.cap_edges, self.bandit = int(cap_dst), int(cap_src), int(cap_edges), bandit
        i32 = dict(dtype=torch.int32, device=device)
        self.meta = torch.zeros(3 * (self.cap_dst + 1), **i32)
        n = self.cap_dst + 1
        self.indptr, self.seg_ptr = self.meta[:n], self.meta[n:2 * n]
        self.inv_deg = self.meta[2 * n:].view(torch.float32)[:self.cap_dst]
        self.inv_deg.fill_(1.0)
        self.e32 = torch.zeros((5, self.cap_edges), **i32)
        self.csc_pos = torch.zeros(self.cap_edges, dtype=torch.int64, device=device)
        self.src = torch.zeros(2 * self.cap_src, **i32)
        self.src_nid = self.src[:self.cap_src]
        self.node_prob = self.src[self.cap_src:].view(torch.float32)
        self.t_indptr = torch.zeros(self.cap_src + 1, **i32)
        self.t_cursor = torch.zeros(self.cap_src, **i32)
        # source x destination bitmap of the block + word prefixes: the transpose is read off the bitmap
        self.t_words = (self.cap_dst + 31) // 32
        self.t_bits = torch.zeros(self.cap_src * self.t_words, **i32)
        self.t_pre = torch.zeros(self.cap_src * self.t_words, **i32)
        self.t_dst = torch.zeros(self.cap_edges, **i32)
        self.t_perm = torch.zeros(self.cap_edges, **i32)
        self.t_seg_ptr = torch.zeros(self.cap_src + 1, **i32)
        self.t_w = torch.zeros(self.cap_edges, dtype=torch.float32, device=device)   # block weights in transpose order
        self.padded = None      # the capacity-sized Block the captured graph runs on

    def fits(self, n_dst, n_src, n_edges) -> bool:
        return n_dst <= self.cap_dst and n_src <= self.cap_src and n_edges <= self.cap_edges


class Frontier:
    """What ``exp3_probabilities`` / ``compute_prob`` hand on (the reference's ``insg`` + ``edge_prob``)."""

    def __init__(self, g, wsp, seeds, n_seeds, layer, mode, weights):
        self.g, self.wsp, self.seeds, self.n_seeds = g, wsp, seeds, n_seeds
        self.layer, self.mode, self.weights = layer, mode, weights
        self.counters: Optional[N.Counters] = None


class BanditLadiesSampler:
    """``bandit_sampler.py:29-367``: EXP3-bandit layer-importance sampler, multinomial selection."""

    _poisson = False
    _mode = N.MODE_BANDIT
    attach_weights = True     # blocks carry ``edge_weights`` (importance weights); False for the uniform samplers
    DENSE_COLLECT_MAX = 1 << 22   # up to 4 M nodes the 8 B/node accumulator scan is cheaper than marking bits

    def __init__(self, nodes_per_layer, importance_sampling=True, weight="w", out_weight="edge_weights",
                 node_embedding="nfeat", node_prob="node_prob", replace=False, eta=0.4, num_steps=5000,
                 model="sage", rng_seed: int = 0, normalize: str = "lazy", renorm_every: int = 64):
        self.nodes_per_layer = nodes_per_layer
        self.importance_sampling = importance_sampling
        self.edge_weight = weight
        self.output_weight = out_weight
        self.node_prob = node_prob
        self.node_embedding = node_embedding
        self.replace = replace
        self.eta = eta
        self.T = num_steps
        self.model = model
        self.eps = 0.9999
        if replace:
            raise NotImplementedError("replace=True (multinomial with replacement) is not used by the reference CLI")
        if normalize not in ("lazy", "literal"):
            raise ValueError("normalize must be 'lazy' or 'literal'")
        self.rng_seed = int(rng_seed)
        self.step = 0
        self.normalize = normalize
        self.renorm_every = int(renorm_every)
        self.inject_uniforms = None   # dense [|V|] float32 (or {layer: tensor}), test hook
        self.collect = "auto"         # candidate collection: 'dense' scan, 'bitmap', or by |V| ('auto')
        self.process_group = None                              # set for data-parallel bandit exchange
        self._w_csc: Optional[torch.Tensor] = None             # [L, |E|] un-normalised, CSC order
        self._l1: Optional[torch.Tensor] = None                # [L] float64 running L1 norms
        self._updated = None                                   # [L] bool: weights ever updated
        self._updates_since_renorm = 0
        self._wsp: Optional[_Workspace] = None
        self._g: Optional[Graph] = None
        self.last_counters: List[Optional[N.Counters]] = [None] * len(nodes_per_layer)

    # ---- state --------------------------------------------------------------------------
    def _bind(self, g: Graph):
        if self._g is not g:
            self._wsp = _Workspace(g)
            self._wsp2 = None          # the whole-step graph's second workspace belongs to the previous graph
            self._g = g
            self._w_csc = None
        if self._w_csc is None and self._mode == N.MODE_BANDIT:
            L, E = len(self.nodes_per_layer), g.num_edges()
            e_pad = (E + 63) // 64 * 64      # every layer's row starts 256-byte aligned
            self._w_buf = torch.ones(L, e_pad, dtype=torch.float32, device=g.device)  # bandit_sampler.py:343
            self._w_csc = [self._w_buf[l, :E] for l in range(L)]
            self._l1 = torch.full((L,), float(E), dtype=torch.float64, device=g.device)
            # running max of every layer's weights (kept by the update kernels) + a pinned copy a step graph refreshes:
            # the lazy normalisation re-scales when a weight nears the top of the fp32 range (tick_renorm)
            self._wmax = torch.ones(L, dtype=torch.float32, device=g.device)
            self._wmax_host = torch.ones(L, dtype=torch.float32).pin_memory() if g.device.type == "cuda" \
                else torch.ones(L, dtype=torch.float32)
            self._updated = [False] * L
            self._norm_partial = torch.empty(1024, dtype=torch.float64, device=g.device)
        return self._wsp

    @property
    def exp3_weights(self):
        """The reference's ``[L, |E|]`` edge-id-ordered weights (``bandit_sampler.py:43,343,249``):
        all ones before the first update, L1-normalised afterwards."""
        if self._w_csc is None:
            return None
        g = self._g
        out = torch.empty(len(self._w_csc), g.num_edges(), dtype=torch.float32, device=g.device)
        for l in range(len(self._w_csc)):
            w = self._w_csc[l]
            if self._updated[l]:
                w = (w.double() / self._l1[l].clamp_min(1e-12)).float()
            out[l, g.eid.long()] = w
        return out

    @exp3_weights.setter
    def exp3_weights(self, value):
        if value is None:
            self._w_csc = None
            return
        g = self._g
        if g is None:
            raise RuntimeError("bind a graph first: call sample_blocks once or use set_exp3_weights(g, w)")
        self.set_exp3_weights(g, value)

    def set_exp3_weights(self, g: Graph, value: torch.Tensor):
        """Load ``[L, |E|]`` weights given in edge-id order (checkpoint restore / tests)."""
        self._bind(g)
        v = value.to(device=g.device, dtype=torch.float32)
        for l in range(v.shape[0]):
            self._w_csc[l].copy_(v[l, g.eid.long()])
        self._l1.copy_(torch.stack([w.double().abs().sum() for w in self._w_csc]))   # in place (captured graphs)
        self._refresh_wmax()
        self._updated = [True] * v.shape[0]

    def _refresh_wmax(self):
        self._wmax.copy_(torch.stack([w.max() for w in self._w_csc]).clamp_min(0).float())
        self._wmax_host.copy_(self._wmax)

    def state_dict(self):
        return {"exp3_w_csc": torch.stack(list(self._w_csc)), "l1": self._l1.clone(), "updated": list(self._updated),
                "step": self.step, "rng_seed": self.rng_seed, "updates_since_renorm": self._updates_since_renorm}

    def load_state_dict(self, sd, g: Graph):
        self._bind(g)
        for l, w in enumerate(sd["exp3_w_csc"]):
            self._w_csc[l].copy_(w)
        self._l1.copy_(sd["l1"])          # in place: a captured step graph holds this buffer's address
        self._refresh_wmax()
        self._updates_since_renorm = int(sd.get("updates_since_renorm", 0))
        self._updated = list(sd["updated"])
        self.step = int(sd["step"])
        self.rng_seed = int(sd["rng_seed"])

    # ---- stage 1: edge probabilities ------------------------------------------------------
    def exp3_probabilities(self, idx, g, seed_nodes):
        """``bandit_sampler.py:101-138``: registers the frontier (replaces in_subgraph +
        compact_graphs); q_ij itself is produced on the fly by the later passes.  Returns
        ``(edge_prob, insg)`` as one :class:`Frontier` handle twice, to keep the call shape."""
        wsp = self._bind(g)
        n = int(seed_nodes.numel())
        fr = Frontier(g, wsp, seed_nodes, n, idx, self._mode, self._w_csc[idx])
        N.call("bliss_frontier_plan", C.byref(wsp.gview), N.ptr(seed_nodes), n, C.byref(wsp.ws), N.stream())
        return fr, fr

    # ---- stage 2: node probabilities ---------------------------------------------------------
    def _frontier_prob(self, fr: Frontier):
        mode = fr.mode | (0 if self.importance_sampling else N.MODE_UNIFORM)
        if self.collect == 'bitmap' or (self.collect == 'auto' and fr.g.num_nodes() > self.DENSE_COLLECT_MAX):
            mode |= N.COLLECT_BITMAP
        N.call("bliss_frontier_prob", C.byref(fr.wsp.gview), N.ptr(fr.seeds), fr.n_seeds, N.ptr(fr.weights),
                                            float(self.eta), mode, C.byref(fr.wsp.ws), N.stream())

    def compute_prob(self, insg: Frontier, seed_nodes, edge_prob, num):
        """``bandit_sampler.py:47-82`` (+ the Poisson scale search ``:381-406`` in the subclass)."""
        self._frontier_prob(insg)
        if not self._poisson:      # candidate probabilities for the top-k selection
            N.call("bliss_poisson_scale", insg.n_seeds, int(num), float(self.eps), 0, C.byref(insg.wsp.ws),
                   N.stream())
        insg.fanout = int(num)
        return insg

    # ---- stage 3: selection -------------------------------------------------------------------
    def _u_ptr(self, g, layer=None):
        u = self.inject_uniforms
        if isinstance(u, dict):
            u = u.get(layer)
        if u is None:
            return None
        if u.dtype != torch.float32 or u.numel() != g.num_nodes() or u.device != g.device:
            raise ValueError("inject_uniforms must be a float32 [num_nodes] tensor on the graph's device")
        return N.ptr(u)

    def select_neighbors(self, prob: Frontier, num):
        """``bandit_sampler.py:84-99``: multinomial without replacement = top-k of p / Exp(1)."""
        wsp = prob.wsp
        if wsp.key_scratch is None:
            wsp.key_scratch = torch.empty(prob.g.num_nodes() + 4, dtype=torch.float32, device=prob.g.device)
        N.call("bliss_select_topk", prob.n_seeds, int(num), self.rng_seed, self.step, prob.layer,
                                          self._u_ptr(prob.g, prob.layer), N.ptr(wsp.key_scratch), C.byref(wsp.ws),
                                          N.stream())
        return prob

    # ---- stage 4: block construction ------------------------------------------------------------
    def generate_block(self, insg: Frontier, neighbor_nodes_idx, seed_nodes, P_sg=None, W_sg=None):
        """``bandit_sampler.py:269-339`` (``ladies_sampler.py:71-107``)."""
        fr, wsp = insg, insg.wsp
        out, bufs = self._block_out(fr)
        N.call("bliss_block_count", C.byref(wsp.gview), N.ptr(fr.seeds), fr.n_seeds, N.ptr(fr.weights),
               float(self.eta), fr.mode, C.byref(wsp.ws), N.stream())
        N.call("bliss_block_index", N.ptr(fr.seeds), fr.n_seeds, C.byref(wsp.ws), C.byref(out), N.stream())
        return self._finish_block(fr, out, bufs)

    def _block_out(self, fr: Frontier, pool: Optional[LayerPool] = None):
        """Output descriptor of one layer: indptr / SpMM segment prefix / mean divisor are sized by the seeds,
        source arrays go to |V|-sized scratch until the counts are known.  With a :class:`LayerPool`
        the arrays are the pool's persistent capacity buffers and indptr is padded to the capacity."""
        wsp, dev, n_s = fr.wsp, fr.g.device, fr.n_seeds
        if pool is not None and n_s <= pool.cap_dst:
            # out_deg -> the pool's transpose cursor: the fill kernel counts every source's block edges,
            # so the transpose starts from its scan (no separate count pass)
            out = N.BlockOut(indptr=N.ptr(pool.indptr), src_nid=N.ptr(wsp.src_nid), node_prob=N.ptr(wsp.node_prob),
                             seg_ptr=N.ptr(pool.seg_ptr), inv_deg=N.ptr(pool.inv_deg), out_deg=N.ptr(pool.t_cursor),
                             t_bits=N.ptr(pool.t_bits), t_words=pool.t_words, cap_edges=0, cap_src=fr.g.num_nodes(), pad_src=pool.cap_src, pad_rows=pool.cap_dst)
            return out, (pool.indptr[:n_s + 1], pool.seg_ptr[:n_s + 1], pool.inv_deg[:n_s])
        meta = torch.empty(3 * (n_s + 1), dtype=torch.int32, device=dev)   # indptr | seg_ptr | inv_deg (as f32)
        indptr, seg_ptr, inv_deg = meta[:n_s + 1], meta[n_s + 1:2 * (n_s + 1)], meta[2 * (n_s + 1):].view(torch.float32)
        out = N.BlockOut(indptr=N.ptr(indptr), src_nid=N.ptr(wsp.src_nid), node_prob=N.ptr(wsp.node_prob),
                         seg_ptr=N.ptr(seg_ptr), inv_deg=N.ptr(inv_deg), cap_edges=0, cap_src=fr.g.num_nodes())
        return out, (indptr, seg_ptr, inv_deg[:n_s])

    def _finish_block(self, fr: Frontier, out, bufs, pool: Optional[LayerPool] = None):
        """Read the layer's counters (the one host sync), size the edge arrays, fill + finish."""
        wsp, g, dev, n_s = fr.wsp, fr.g, fr.g.device, fr.n_seeds
        indptr, seg_ptr, inv_deg = bufs
        ctr = wsp.read_counters()
        if ctr.error:
            raise RuntimeError(f"BLISS sampler capacity error {ctr.error} in layer {fr.layer}")
        fr.counters = ctr
        self.last_counters[fr.layer] = ctr
        n_src, E = int(ctr.n_src), int(ctr.n_edges)
        bandit = fr.mode == N.MODE_BANDIT
        pooled = pool is not None and out.pad_rows > 0 and pool.fits(n_s, n_src, E)
        if pool is not None and not pooled:
            self.pool_overflow = True          # the caller grows the pool and re-captures
            out.out_deg, out.t_bits = None, None
            if out.pad_rows > 0:               # indptr lives in the pool but the block does not fit: detach it
                indptr, seg_ptr, inv_deg = indptr.clone(), seg_ptr[:n_s + 1].clone(), inv_deg.clone()
        # one allocation for the 4-byte edge arrays, one for the 8-byte CSC positions
        if pooled:
            e32, csc_pos = pool.e32, pool.csc_pos[:E]
        else:
            e32 = torch.empty((5 if bandit else 4, max(E, 1)), dtype=torch.int32, device=dev)
            csc_pos = torch.empty(E, dtype=torch.int64, device=dev)
        edge_src, edge_dst, eid = e32[0, :E], e32[1, :E], e32[2, :E]
        edge_w = e32[3, :E].view(torch.float32)
        q_ij = e32[4, :E].view(torch.float32) if bandit else None
        out.edge_src, out.edge_dst, out.csc_pos = N.ptr(edge_src), N.ptr(edge_dst), N.ptr(csc_pos)
        out.eid, out.edge_w, out.q_ij = N.ptr(eid), N.ptr(edge_w), N.ptr(q_ij)
        out.cap_edges = E
        N.call("bliss_sample_layer_back", C.byref(wsp.gview), N.ptr(fr.seeds), n_s, N.ptr(fr.weights),
               float(self.eta), fr.mode, C.byref(wsp.ws), C.byref(out), N.stream())
        if pooled:
            src_nid, node_prob = pool.src_nid[:n_src], pool.node_prob[:n_src]
        else:
            src = torch.empty(2 * n_src, dtype=torch.int32, device=dev)
            src_nid = src[:n_src]
            node_prob = src[n_src:].view(torch.float32)
        src_nid.copy_(wsp.src_nid[:n_src])
        node_prob.copy_(wsp.node_prob[:n_src])
        block = Block(indptr, edge_src, edge_dst, src_nid, fr.seeds, graph=g, csc_pos=csc_pos)
        block.seg_ptr = seg_ptr
        block._mean_scale = inv_deg
        block.edata[EID] = eid                                                   # :337
        block.edata[self.output_weight] = edge_w                                 # :324
        self._attach(block, q_ij, node_prob)
        return block

    def _attach(self, block, q_ij, node_prob):
        block.edata["q_ij"] = q_ij                                               # :326
        block.srcdata[self.node_prob] = node_prob                                # :328

    # ---- driver ---------------------------------------------------------------------------------
    def _prep_seeds(self, g, seed_nodes):
        s = torch.as_tensor(seed_nodes)
        if s.device != g.device:
            s = (s.pin_memory() if s.device.type == "cpu" and not s.is_pinned() else s).to(g.device, non_blocking=True)
        return s.to(torch.int32).contiguous()

    def sample_blocks(self, g, seed_nodes, exclude_eids=None, pools=None):
        """``bandit_sampler.py:341-367``.  ``pools`` (one :class:`LayerPool` per layer, optional) makes
        the blocks views of persistent capacity buffers (static-shape CUDA-graph replay, train.py)."""
        self._bind(g)
        seed_nodes = self._prep_seeds(g, seed_nodes)
        output_nodes = seed_nodes
        blocks = []
        self.pool_overflow = False
        fused = self._stages_not_overridden()
        self.pool_used = bool(fused and pools)
        for block_id in reversed(range(len(self.nodes_per_layer))):              # :350
            num = self.nodes_per_layer[block_id]
            if fused:   # same kernels, two FFI calls per layer instead of eight
                block = self._sample_layer_fused(g, seed_nodes, block_id, num, self._w_csc[block_id],
                                                 pools[block_id] if pools else None)
                seed_nodes = block.srcdata[NID]
                blocks.insert(0, block)
                continue
            edge_prob, insg = self.exp3_probabilities(block_id, g, seed_nodes)   # :354
            node_prob = self.compute_prob(insg, seed_nodes, edge_prob, num)      # :356
            chosen = self.select_neighbors(node_prob, num)                       # :360
            block = self.generate_block(insg, chosen, seed_nodes, node_prob, edge_prob)   # :362
            seed_nodes = block.srcdata[NID]                                      # :364
            blocks.insert(0, block)                                              # :366
        self.step += 1
        return seed_nodes, output_nodes, blocks

    _STAGES = ("exp3_probabilities", "compute_prob", "select_neighbors", "generate_block")

    def _stages_not_overridden(self) -> bool:
        """The stage methods are the reference's extension points; when a subclass overrides one, the
        per-stage path runs so the override is honoured."""
        if getattr(self, "force_stage_path", False):      # bench.py: time every kernel on its own
            return False
        return all(getattr(type(self), m).__module__ == __name__ for m in self._STAGES)

    def _sample_layer_fused(self, g, seed_nodes, block_id, num, weights, pool: Optional[LayerPool] = None):
        wsp = self._wsp
        n = int(seed_nodes.numel())
        fr = Frontier(g, wsp, seed_nodes, n, block_id, self._mode, weights)
        mode = fr.mode | (0 if self.importance_sampling else N.MODE_UNIFORM)
        if self.collect == "bitmap" or (self.collect == "auto" and g.num_nodes() > self.DENSE_COLLECT_MAX):
            mode |= N.COLLECT_BITMAP
        if not self._poisson and wsp.key_scratch is None:
            wsp.key_scratch = torch.empty(g.num_nodes() + 4, dtype=torch.float32, device=g.device)
        out, bufs = self._block_out(fr, pool)
        N.call("bliss_sample_layer_front", C.byref(wsp.gview), N.ptr(seed_nodes), n, N.ptr(weights), float(self.eta),
               mode, int(num), float(self.eps), int(self._poisson), self.rng_seed, self.step, block_id,
               self._u_ptr(g, block_id), N.ptr(wsp.key_scratch), C.byref(wsp.ws), C.byref(out), N.stream())
        return self._finish_block(fr, out, bufs, pool)

    # ---- sync-free path (CUDA-graph capture of the whole step) -----------------------------------
    def plan_top_static(self, g, seeds_static, pools, step_dev, ctr_base: int = 0, ctr_mirror=None):
        """The top layer's frontier plan (row extents, chunk lists, counters) of :meth:`enqueue_static`, enqueued on
        its own: it depends on the batch only, so the whole-step graph runs it at the head of the step and the
        sampling chain behind the bandit update starts with the probability passes."""
        wsp = self._bind(g)
        L = len(self.nodes_per_layer)
        ws = wsp.ws_layer(ctr_base + L - 1, None, N.ptr(step_dev), counters=wsp,
                          ctr_mirror=None if ctr_mirror is None else ctr_mirror[L - 1].data_ptr())
        N.call("bliss_frontier_plan", C.byref(wsp.gview), N.ptr(seeds_static), pools[L - 1].cap_dst, C.byref(ws), N.stream())

    def enqueue_static(self, g, seeds_static, pools, step_dev, transpose_stream=None, defer_last_transpose=False,
                       ctr_base: int = 0, layer_pre=None, ctr_mirror=None, top_planned: bool = False):
        """Enqueue the sampling of every layer into the capacity pools with NO host synchronisation:
        each layer reads its true seed count from the previous layer's device counters, the Philox
        step from ``step_dev``, and the transpose its edge count from the counters.  Capturable in a
        CUDA graph; the caller reads all counters once per step (``_Workspace.read_all_counters``).

        ``transpose_stream``: side stream for every layer's back half (fill, workspace restore, transpose); the
        padded blocks get ``_ready`` / ``_t_ready`` events their readers wait for, and the caller must join the
        stream before the step ends.  ``defer_last_transpose``: do not launch the input layer's transpose here but
        return it (a list of callables) for the caller to launch after the forward pass.  ``ctr_base``: first
        counters block to use (layer l writes block ``ctr_base + l``): the pipelined step samples the NEXT step's
        blocks into a second pool set while this step's backward pass still reads the first set's counts.
        ``layer_pre``: ``{layer: callable}`` run on the current stream right before that layer is sampled (the
        data-parallel step applies all ranks' bandit updates of a layer just before the layer's weights are read).
        ``top_planned``: the top layer's plan was already enqueued (:meth:`plan_top_static`).
        ``ctr_mirror``: pinned host tensor ``[>= L, sizeof(counters)]``; layer l's finish kernel writes its counters
        to row l (device-mapped host memory), so the caller needs no device-to-host copy after the step."""
        wsp = self._bind(g)
        L = len(self.nodes_per_layer)
        bandit = self._mode == N.MODE_BANDIT
        weights_static = None if bandit else g.csc_edata(self.edge_weight)
        side = transpose_stream
        if side is not None and getattr(self, "_wsp2", None) is None:
            # Consecutive layers alternate between two workspaces, so a layer's back half (fill, workspace restore,
            # transpose) can run on the side stream while the next layer's front half already samples: the next
            # layer only needs this layer's source list and counters, which the front half produced.
            self._wsp2 = _Workspace(g)
        done = {}                                   # layer -> event: its fill and workspace restore are finished
        deferred = []                               # transposes the caller launches later (``defer_last_transpose``)
        for block_id in reversed(range(L)):
            pool, top = pools[block_id], block_id == L - 1
            w = wsp if (side is None or (L - 1 - block_id) % 2 == 0) else self._wsp2
            seeds = seeds_static if top else pools[block_id + 1].src_nid
            n_cap = pool.cap_dst
            ws = w.ws_layer(ctr_base + block_id, None if top else wsp.counter_ptr(ctr_base + block_id + 1, "n_src"),
                            N.ptr(step_dev), counters=wsp,
                            ctr_mirror=None if ctr_mirror is None else ctr_mirror[block_id].data_ptr())
            weights = self._w_csc[block_id] if bandit else weights_static
            mode = self._mode | (0 if self.importance_sampling else N.MODE_UNIFORM)
            if self.collect == "bitmap" or (self.collect == "auto" and g.num_nodes() > self.DENSE_COLLECT_MAX):
                mode |= N.COLLECT_BITMAP
            if top and top_planned:
                mode |= N.MODE_PLANNED
            if not self._poisson and w.key_scratch is None:
                w.key_scratch = torch.empty(g.num_nodes() + 4, dtype=torch.float32, device=g.device)
            e32 = pool.e32
            out = N.BlockOut(indptr=N.ptr(pool.indptr), edge_src=N.ptr(e32[0]), edge_dst=N.ptr(e32[1]),
                             csc_pos=N.ptr(pool.csc_pos), eid=N.ptr(e32[2]), q_ij=N.ptr(e32[4]) if bandit else None,
                             edge_w=N.ptr(e32[3]), src_nid=N.ptr(pool.src_nid), node_prob=N.ptr(pool.node_prob),
                             out_deg=N.ptr(pool.t_cursor), t_bits=N.ptr(pool.t_bits), t_words=pool.t_words,
                             seg_ptr=N.ptr(pool.seg_ptr), inv_deg=N.ptr(pool.inv_deg),
                             cap_edges=pool.cap_edges, cap_src=pool.cap_src, pad_src=pool.cap_src,
                             pad_rows=pool.cap_dst)
            main = torch.cuda.current_stream()
            if layer_pre and block_id in layer_pre:
                layer_pre[block_id]()
            if block_id + 2 in done:                # this workspace was last used two layers ago: its restore must be done
                main.wait_event(done[block_id + 2])
            N.call("bliss_sample_layer_front", C.byref(w.gview), N.ptr(seeds), n_cap, N.ptr(weights), float(self.eta),
                   mode, int(self.nodes_per_layer[block_id]), float(self.eps), int(self._poisson), self.rng_seed,
                   0, block_id, self._u_ptr(g, block_id), N.ptr(w.key_scratch), C.byref(ws), C.byref(out), N.stream())
            back = side if side is not None else main
            if side is not None:
                side.wait_stream(main)
            with torch.cuda.stream(back):
                N.call("bliss_block_fill", C.byref(w.gview), N.ptr(seeds), n_cap, N.ptr(weights), float(self.eta),
                       self._mode, C.byref(ws), C.byref(out), N.stream())
                if side is not None:                # the forward aggregation over this block waits for this event only
                    pool.ready = torch.cuda.Event()
                    pool.ready.record(back)
                    if pool.padded is not None:
                        pool.padded._ready = pool.ready
                N.call("bliss_block_finish", n_cap, self._mode, C.byref(ws), C.byref(out), N.stream())

            def transpose(pool=pool, e32=e32, block_id=block_id, back=back):
                # read by the backward pass only (the caller joins ``transpose_stream`` before it)
                with torch.cuda.stream(back):
                    N.call("bliss_block_transpose", N.ptr(e32[0]), N.ptr(e32[1]), pool.cap_edges, pool.cap_src,
                           pool.cap_dst, N.ptr(pool.t_indptr), N.ptr(pool.t_cursor), N.ptr(pool.t_bits), N.ptr(pool.t_pre),
                           pool.t_words, N.ptr(pool.t_dst), N.ptr(pool.t_perm), N.ptr(pool.t_seg_ptr), 1,
                           wsp.counter_ptr(ctr_base + block_id, "n_edges"), N.ptr(e32[3]), N.ptr(pool.t_w), N.stream())
                    if pool.padded is not None:     # (weights tensor it was built from, transposed copy)
                        pool.padded._t_w = (e32[3].data_ptr(), pool.t_w)
                    if side is not None:            # readers of the transpose (backward pass, GCN out-degrees) wait for this
                        ev = torch.cuda.Event()
                        ev.record(back)
                        if pool.padded is not None:
                            pool.padded._t_ready = ev

            if side is not None:
                with torch.cuda.stream(back):
                    done[block_id] = torch.cuda.Event()
                    done[block_id].record(back)
            if side is not None and block_id == 0 and defer_last_transpose:
                # the input layer's transpose would run beside the first (bandwidth-bound) aggregation of the forward
                # pass: the caller launches it after the forward pass instead
                deferred.append(transpose)
            else:
                transpose()
        return deferred

    # ---- bandit update ------------------------------------------------------------------------
    def calculate_alpha(self, mfg):
        """``bandit_sampler.py:140-158``.  SAGE/GCN: the static edge weight; GAT: from a_ij, q_ij."""
        if self.model == "gat":
            n_dst = mfg.num_dst_nodes()
            asum = torch.empty(n_dst, dtype=torch.float32, device=mfg.device)
            qsum = torch.empty(n_dst, dtype=torch.float32, device=mfg.device)
            a = mfg.edata["a_ij"].detach().contiguous()
            N.call("bliss_gat_alpha_sums", N.ptr(mfg.indptr), N.ptr(a), N.ptr(mfg.edata["q_ij"]), n_dst,
                                                 N.ptr(asum), N.ptr(qsum), N.stream())
            return ("gat", a, asum, qsum)
        return ("static", None, None, None)

    def _reward_call(self, idx, mfg, g, alpha, weights, rewards=None, x_out=None, l1=None, n_edges_dev=None,
                     count_out=None, pos_out=None, p2p=None):
        kind, a, asum, qsum = alpha
        wsp = self._bind(g)
        w_static = g.csc_edata(self.edge_weight) if kind == "static" else None
        emb = mfg.srcdata["embed_norm"]
        emb = emb.detach()
        if emb.dtype != torch.float32:
            emb = emb.float()
        if n_edges_dev is None:
            n_edges_dev = getattr(mfg, "_n_edges_dev", None)     # capacity-padded block: true count on the device
        N.call("bliss_reward_update",
            C.byref(wsp.gview), N.ptr(mfg.indptr), N.ptr(mfg.edge_src), N.ptr(mfg.edge_dst), N.ptr(mfg.csc_pos),
            N.ptr(mfg.dstdata[NID]), N.ptr(mfg.edata["q_ij"]), N.ptr(mfg.srcdata[self.node_prob]),
            N.ptr(emb.contiguous()), N.ptr(w_static), N.ptr(a), N.ptr(asum), N.ptr(qsum),
            1 if kind == "gat" else 0, 0.01, mfg.num_dst_nodes(), mfg.num_edges(), N.ptr(weights),
            N.ptr(rewards), N.ptr(x_out), N.ptr(l1), n_edges_dev, count_out, N.ptr(pos_out),
            C.byref(p2p) if p2p is not None else None,
            N.ptr(self._wmax[idx:idx + 1]) if weights is not None else None, N.stream())

    def calculate_rewards(self, idx, mfg, g, alpha):
        """``bandit_sampler.py:160-193``: stores ``mfg.edata['rewards']`` (emit-only kernel call)."""
        r = torch.empty(mfg.num_edges(), dtype=torch.float32, device=mfg.device)
        self._reward_call(idx, mfg, g, alpha, None, rewards=r)
        mfg.edata["rewards"] = r

    def update_exp3_weights(self, idx, mfg, g, alpha=None):
        """``bandit_sampler.py:195-249``: w[e] *= exp(min(1, r/P · δ/n)) then L1-normalise."""
        if alpha is None:
            alpha = self.calculate_alpha(mfg)
        pg = self.process_group
        if pg is not None and torch.distributed.get_world_size(pg) > 1:
            self._update_distributed(idx, mfg, g, alpha, pg)
        else:
            self._reward_call(idx, mfg, g, alpha, self._w_csc[idx], l1=self._l1[idx:idx + 1])
        self._updated[idx] = True
        if self.normalize == "literal":
            self._renormalize(idx)

    def _update_distributed(self, idx, mfg, g, alpha, pg):
        """Every rank sampled from the same frozen weights; all ranks apply all ranks' updates
        (``w *= exp(x)`` commutes).  One all-gather of the sparse (CSC position, exponent) pairs."""
        from .parallel import gather_updates
        x = torch.empty(mfg.num_edges(), dtype=torch.float32, device=mfg.device)
        self._reward_call(idx, mfg, g, alpha, None, x_out=x)
        for pos_r, x_r in gather_updates(mfg.csc_pos, x, pg):
            if pos_r.numel():
                N.call("bliss_apply_updates", N.ptr(pos_r), N.ptr(x_r), pos_r.numel(), N.ptr(self._w_csc[idx]),
                       N.ptr(self._l1[idx:idx + 1]), N.ptr(self._wmax[idx:idx + 1]), N.stream())

    def _renormalize(self, idx):
        w = self._w_csc[idx]
        L = N.lib()
        N.call("bliss_l1_norm", N.ptr(w), w.numel(), N.ptr(self._norm_partial), N.ptr(self._l1[idx:idx + 1]),
                                N.stream())
        N.call("bliss_scale_by_inv", N.ptr(w), w.numel(), N.ptr(self._l1[idx:idx + 1]), 1e-12, N.stream())
        self._l1[idx:idx + 1].fill_(1.0)      # (a fill launch: capturable, unlike assigning a Python scalar)
        self._wmax[idx:idx + 1].fill_(1.0)    # (an upper bound: the re-scaled weights sum to 1)

    def exp3_emit(self, mfgs, g, exchange):
        """Data-parallel, sync-free: compute every layer's clamped exponents into the exchange's send
        buffer (positions as int32 + exponents) and the edge counts into its header."""
        if g.num_edges() >= 2 ** 31:
            raise NotImplementedError("the packed bandit exchange carries int32 CSC positions: |E| must be < 2^31")
        for idx, mfg in enumerate(mfgs):
            self.exp3_emit_layer(idx, mfg, g, exchange)

    def exp3_emit_layer(self, idx, mfg, g, exchange):
        """One layer of :meth:`exp3_emit` (needs the layer's ``embed_norm`` — and ``a_ij`` for GAT — only)."""
        assert mfg.num_edges() <= exchange.caps[idx]
        if exchange.p2p:      # straight into every rank's window over NVLink (flags raised by the kernel's last CTA)
            self._reward_call(idx, mfg, g, self.calculate_alpha(mfg), None, p2p=exchange.p2p_struct(idx))
            return
        self._reward_call(idx, mfg, g, self.calculate_alpha(mfg), None, x_out=exchange.x[idx],
                          count_out=exchange.header.data_ptr() + 8 * idx, pos_out=exchange.pos[idx])

    def exp3_apply(self, exchange, n_layers: int):
        """Apply all ranks' gathered updates (one kernel per layer, counts read from the headers)."""
        for idx in range(n_layers):
            self.exp3_apply_layer(exchange, idx)

    def exp3_apply_layer(self, exchange, idx: int):
        if exchange.p2p:      # wait for every rank's flag of this layer, then apply the slots of the own window
            N.call("bliss_apply_updates_p2p", C.byref(exchange.p2p_struct(idx)), exchange.caps[idx],
                   N.ptr(self._w_csc[idx]), N.ptr(self._l1[idx:idx + 1]), N.ptr(exchange.err),
                   N.ptr(self._wmax[idx:idx + 1]), N.stream())
        else:
            N.call("bliss_apply_updates_packed", N.ptr(exchange.recv), exchange.stride, exchange.world,
                   8 * idx, exchange.pos_off[idx], exchange.x_off[idx], exchange.caps[idx],
                   N.ptr(self._w_csc[idx]), N.ptr(self._l1[idx:idx + 1]), N.ptr(self._wmax[idx:idx + 1]), N.stream())
        self._updated[idx] = True
        if self.normalize == "literal":
            self._renormalize(idx)

    #: ln(FLT_MAX) with a little room
    _LOG_RANGE = 88.0

    def mirror_wmax_(self):
        """Refresh the pinned copy of the running weight maxima (a copy node of the step graph; stream-ordered)."""
        self._wmax_host.copy_(self._wmax, non_blocking=True)

    def tick_renorm(self, n_layers: int, mirrored: bool = False):
        """Lazy mode's range safety (a weight grows by at most e per update, ``bandit_sampler.py:244-246``; every
        rank's update lands on this copy of the weights, so W ranks can grow an edge by e^W per step).  The update
        kernels keep the running maximum of every layer's weights; the weights are physically re-normalised — a pass
        over 3 x |E| floats — only when that maximum leaves room for fewer updates than can happen before the next
        look, not on a fixed schedule.

        ``mirrored``: a step graph refreshes the pinned copy of the maxima every step (``mirror_wmax_``): it is read
        here without a sync, on every call, and may be up to 6 steps old.  Otherwise the device value is read (one
        sync) every ``renorm_every`` updates."""
        if self.normalize != "lazy" or self._w_csc is None:     # (the LADIES / uniform samplers have no bandit state)
            return
        pg = self.process_group
        W = torch.distributed.get_world_size(pg) if pg is not None else 1
        if mirrored and self._LOG_RANGE - 7.0 * W >= 1.0:
            if float(self._wmax_host[:n_layers].max()) <= math.exp(self._LOG_RANGE - 7.0 * W):
                return
        else:
            self._updates_since_renorm += W
            if self._updates_since_renorm < self.renorm_every:
                return
            self._updates_since_renorm = 0
            room = self._LOG_RANGE - self.renorm_every - W
            if room >= 1.0 and float(self._wmax[:n_layers].max()) <= math.exp(room):
                return
        for idx in range(n_layers):
            self._renormalize(idx)
        self._wmax_host.fill_(1.0)
        self._updates_since_renorm = 0

    def exp3(self, mfgs, g, exchange=None, count_renorm=True):
        """``bandit_sampler.py:251-267``: reward + weight update of every layer, one fused kernel each.
        ``exchange`` (a ``parallel.BanditExchange``) selects the data-parallel fast path: exponents are
        written into the exchange's send buffer, ONE all-gather moves all layers, one kernel per layer
        applies every rank's update."""
        self._bind(g)
        if (exchange is not None and g.num_edges() < 2 ** 31
                and all(m.num_edges() <= exchange.caps[i] for i, m in enumerate(mfgs))):
            for idx, mfg in enumerate(mfgs):      # (eager steps always go through the NCCL all-gather)
                self._reward_call(idx, mfg, g, self.calculate_alpha(mfg), None, x_out=exchange.x[idx],
                                  pos_out=exchange.pos[idx])
            exchange.exchange([m.num_edges() for m in mfgs])
            for idx in range(len(mfgs)):
                N.call("bliss_apply_updates_packed", N.ptr(exchange.recv), exchange.stride, exchange.world,
                       8 * idx, exchange.pos_off[idx], exchange.x_off[idx], exchange.caps[idx],
                       N.ptr(self._w_csc[idx]), N.ptr(self._l1[idx:idx + 1]), N.ptr(self._wmax[idx:idx + 1]), N.stream())
                self._updated[idx] = True
                if self.normalize == "literal":
                    self._renormalize(idx)
        else:
            for idx, mfg in enumerate(mfgs):
                alpha = self.calculate_alpha(mfg)
                self.update_exp3_weights(idx, mfg, g, alpha)
        if count_renorm:
            self.tick_renorm(len(mfgs))


class PoissonBanditLadiesSampler(BanditLadiesSampler):
    """``bandit_sampler.py:369-424``: Poisson (independent inclusion) variant — the CLI default."""

    _poisson = True

    def select_neighbors(self, prob: Frontier, num):
        """``bandit_sampler.py:408-425``: ``bernoulli(P) == 1``  ⇔  ``u < P``."""
        # scale search (:391-406) + selection in one thread-block-cluster launch
        N.call("bliss_poisson_select", prob.n_seeds, int(num), float(self.eps), self.rng_seed, self.step,
               prob.layer, self._u_ptr(prob.g, prob.layer), C.byref(prob.wsp.ws), N.stream())
        return prob


class LadiesSampler(BanditLadiesSampler):
    """``ladies_sampler.py:24-123``: static weights ``g.edata[weight]``, no bandit state."""

    _mode = N.MODE_LADIES

    def __init__(self, nodes_per_layer, importance_sampling=True, weight="w", out_weight="edge_weights",
                 replace=False, allow_zero_in_degree=False, rng_seed: int = 0):
        super().__init__(nodes_per_layer, importance_sampling, weight, out_weight, replace=replace,
                         rng_seed=rng_seed)
        self.allow_zero_in_degree = allow_zero_in_degree

    def compute_prob(self, g, seed_nodes, weight, num):
        """``ladies_sampler.py:34-52``: returns ``(prob, insg)`` (one Frontier handle twice)."""
        wsp = self._bind(g)
        n = int(seed_nodes.numel())
        fr = Frontier(g, wsp, seed_nodes, n, self._layer, N.MODE_LADIES, weight)
        N.call("bliss_frontier_plan", C.byref(wsp.gview), N.ptr(seed_nodes), n, C.byref(wsp.ws), N.stream())
        self._frontier_prob(fr)
        if not self._poisson:
            N.call("bliss_poisson_scale", n, int(num), float(self.eps), 0, C.byref(wsp.ws), N.stream())
        return fr, fr

    def _attach(self, block, q_ij, node_prob):
        pass                                                                     # ladies_sampler.py:99-106

    def sample_blocks(self, g, seed_nodes, exclude_eids=None):
        """``ladies_sampler.py:109-123``."""
        self._bind(g)
        seed_nodes = self._prep_seeds(g, seed_nodes)
        output_nodes = seed_nodes
        blocks = []
        W = g.csc_edata(self.edge_weight)                                        # :114 (CSC order)
        if W.dtype != torch.float32:
            W = W.float()
        if not hasattr(self, "_w_checked"):
            if float(W.max()) > 1.0:
                raise ValueError("LADIES edge weights must be <= 1 (fixed-point column sums; DESIGN.md §4)")
            self._w_checked = True
        fused = self._stages_not_overridden()
        for block_id in reversed(range(len(self.nodes_per_layer))):
            self._layer = block_id
            num = self.nodes_per_layer[block_id]
            if fused:
                block = self._sample_layer_fused(g, seed_nodes, block_id, num, W)
                seed_nodes = block.srcdata[NID]
                blocks.insert(0, block)
                continue
            prob, insg = self.compute_prob(g, seed_nodes, W, num)                # :115
            chosen = self.select_neighbors(prob, num)                            # :117
            block = self.generate_block(insg, chosen, seed_nodes, prob, W)       # :118-120
            seed_nodes = block.srcdata[NID]
            blocks.insert(0, block)
        self.step += 1
        return seed_nodes, output_nodes, blocks

    def exp3(self, mfgs, g):
        raise AttributeError("LadiesSampler has no bandit state (train_lightning.py:469 only calls exp3 for bandit samplers)")


class PoissonLadiesSampler(LadiesSampler):
    """``ladies_sampler.py:125-183``."""

    _poisson = True

    def __init__(self, nodes_per_layer, importance_sampling=True, weight="w", out_weight="edge_weights",
                 allow_zero_in_degree=False, rng_seed: int = 0):
        # the reference forwards allow_zero_in_degree into the ``replace`` slot (:134-136); the value
        # is False on every CLI path, so it is simply kept as an attribute here
        super().__init__(nodes_per_layer, importance_sampling, weight, out_weight, replace=False,
                         allow_zero_in_degree=allow_zero_in_degree, rng_seed=rng_seed)

    select_neighbors = PoissonBanditLadiesSampler.select_neighbors


class NeighborSampler(LadiesSampler):
    """``dgl.dataloading.NeighborSampler(fanouts)`` as the reference builds it for ``--sampler neighbor``
    (``train_lightning.py:351-357``): layer by layer from the output side, every seed keeps ``min(fanout, in-degree)``
    of its in-edges chosen uniformly without replacement (fanout ``-1``: all), the block's sources are the seeds
    followed by the new sources in first-occurrence order (``dgl.to_block``).  No importance weights: the blocks carry
    no ``edge_weights``, so the models aggregate with the plain mean (``model.py:321-329``).

    Same workspace, plan, block-build and transpose kernels as the LADIES path; the per-edge choice is
    ``bliss_neighbor_select`` (``csrc/sampler.cu``): key = Philox(seed; CSC position, layer, step), the row keeps its
    ``fanout`` smallest keys.  DGL draws from its own generator, so the reference's sets are not reproducible bit for
    bit by construction; the contract (uniform k-subsets, block layout) is checked against ``oracle.samplers``."""

    _mode = N.MODE_LADIES | N.MODE_NEIGHBOR
    _poisson = True          # (no top-k scratch; the flag is not consulted on the neighbour path)
    attach_weights = False   # capacity-padded blocks of the step graph carry no edge_weights either

    def __init__(self, fanouts, edge_dir="in", prob=None, mask=None, replace=False, prefetch_node_feats=None,
                 prefetch_labels=None, prefetch_edge_feats=None, output_device=None, rng_seed: int = 0):
        if edge_dir != "in" or prob is not None or mask is not None or replace:
            raise NotImplementedError("NeighborSampler: only edge_dir='in', uniform, without replacement "
                                      "(what train_lightning.py:351-357 constructs)")
        super().__init__([int(f) for f in fanouts], rng_seed=rng_seed)
        self.fanouts = self.nodes_per_layer

    def _finish_block(self, fr, out, bufs, pool=None):
        block = super()._finish_block(fr, out, bufs, pool)
        dict.pop(block.edata, self.output_weight, None)      # uniform sampling: unweighted aggregation
        return block

    def sample_blocks(self, g, seed_nodes, exclude_eids=None, pools=None):
        self._bind(g)
        seed_nodes = self._prep_seeds(g, seed_nodes)
        output_nodes = seed_nodes
        blocks = []
        W = g.csc_edata(self.edge_weight)
        self.pool_overflow, self.pool_used = False, bool(pools)
        for block_id in reversed(range(len(self.nodes_per_layer))):
            block = self._sample_layer_fused(g, seed_nodes, block_id, self.nodes_per_layer[block_id], W,
                                             pools[block_id] if pools else None)
            seed_nodes = block.srcdata[NID]
            blocks.insert(0, block)
        self.step += 1
        return seed_nodes, output_nodes, blocks

    def _stages_not_overridden(self) -> bool:
        return not getattr(self, "force_stage_path", False)


class MultiLayerFullNeighborSampler(NeighborSampler):
    """``dgl.dataloading.MultiLayerFullNeighborSampler(n_layers)`` (``--sampler full``, ``train_lightning.py:349-350``):
    every layer takes the whole in-neighbourhood of its seeds."""

    def __init__(self, num_layers, **kw):
        super().__init__([-1] * int(num_layers), **kw)
