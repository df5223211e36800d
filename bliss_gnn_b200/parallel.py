"""Data-parallel host logic (one process per GPU, replicated graph) — backend agnostic, so the same
code runs over NCCL on the B200s and over gloo in the CPU tests.

The reference is single-device (``train_lightning.py:648-651``); this is new work required by the
north-star: every rank samples and aggregates its own seed batches, and only (1) the gradients and
(2) the sparse bandit updates are exchanged.
"""
from __future__ import annotations

import os
import warnings
from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_batches(n_batches: int, rank: int, world: int) -> range:
    """Rank r takes batches r, r+R, … of the shared epoch permutation; the tail that does not fill a
    whole round is dropped so every rank runs the same number of steps."""
    usable = n_batches - (n_batches % world)
    return range(rank, usable, world)


def flat_layout(params):
    """Offsets of the parameters in the flat buffers (gradients here, parameters and moments in :class:`FlatAdam`):
    every tensor starts on a 16-byte boundary, and a 2-D parameter marked ``_bliss_pad_cols = k`` (the input layer's
    weights when the feature rows are padded to a 16-byte multiple, ``train.Trainer``) gets ``k`` extra zero columns per
    row IN its storage — the model then hands cuBLAS the padded matrix as it lies instead of building a padded copy
    (a fill and a copy kernel per weight at the head of every step).  Returns ``[(offset, rows, cols, pad)]``, total."""
    out, off = [], 0
    for p in params:
        k = int(getattr(p, "_bliss_pad_cols", 0)) if p.dim() == 2 else 0
        n = p.shape[0] * (p.shape[1] + k) if k else p.numel()
        out.append((off, k))
        off += (n + 3) // 4 * 4
    return out, off


def flat_view(flat, off, p, k):
    """(view shaped like ``p``, the padded ``[rows, cols + k]`` matrix it lives in or None)."""
    if k:
        full = flat[off:off + p.shape[0] * (p.shape[1] + k)].view(p.shape[0], p.shape[1] + k)
        return full[:, :p.shape[1]], full
    return flat[off:off + p.numel()].view_as(p), None


class FlatGrads:
    """All parameter gradients as views of one flat buffer: a single all-reduce per step."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.layout, n = flat_layout(self.params)
        self.flat = torch.zeros(n, dtype=self.params[0].dtype, device=self.params[0].device)
        for p, (off, k) in zip(self.params, self.layout):
            p.grad, p._bliss_padded_grad = flat_view(self.flat, off, p, k)

    def zero_(self):
        self.flat.zero_()

    def all_reduce_mean_(self, group=None):
        world = dist.get_world_size(group) if group is not None else 1
        if world > 1:
            dist.all_reduce(self.flat, group=group)
            self.flat.div_(world)


def gather_updates(pos: torch.Tensor, x: torch.Tensor, group) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """All-gather every rank's sparse bandit update ``(csc position, clamped exponent)`` of one layer.
    Lengths differ per rank: sizes are gathered first, payloads are padded to the maximum."""
    world = dist.get_world_size(group)
    n_loc = torch.tensor([pos.numel()], dtype=torch.int64, device=pos.device)
    sizes = [torch.zeros_like(n_loc) for _ in range(world)]
    dist.all_gather(sizes, n_loc, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    pos_pad = torch.zeros(cap, dtype=torch.int64, device=pos.device)
    x_pad = torch.zeros(cap, dtype=torch.float32, device=pos.device)
    pos_pad[:pos.numel()] = pos
    x_pad[:x.numel()] = x
    pos_all = [torch.empty_like(pos_pad) for _ in range(world)]
    x_all = [torch.empty_like(x_pad) for _ in range(world)]
    dist.all_gather(pos_all, pos_pad, group=group)
    dist.all_gather(x_all, x_pad, group=group)
    return [(pos_all[r][:sizes[r]], x_all[r][:sizes[r]]) for r in range(world)]


class BanditExchange:
    """One all-gather per step for the sparse bandit updates of all layers.

    Send buffer of a rank: ``[int64 count per layer (64 B header) | layer 0: int32 pos[cap0], fp32 x[cap0] |
    layer 1 … ]`` — 8 bytes per sampled edge on the wire.  The reward kernel writes ``pos[l]`` (the edge's CSC
    position; the graph must have fewer than 2^31 edges) and ``x[l]`` in place, so nothing is packed or copied
    before the collective, and the apply kernel reads the counts from the gathered headers — no host-side sizes,
    no host sync."""

    HEADER = 64

    def __init__(self, caps, world, device, group):
        self.caps, self.world, self.group = [int(c) for c in caps], int(world), group
        self.device, self.p2p = device, False
        assert len(caps) * 8 <= self.HEADER
        off, self.pos_off, self.x_off = self.HEADER, [], []
        for c in self.caps:
            self.pos_off.append(off)
            off += 4 * c
            self.x_off.append(off)
            off += 4 * c
            off = (off + 15) // 16 * 16
        self.stride = off
        self.send = torch.zeros(off, dtype=torch.uint8, device=device)
        self.recv = torch.zeros(self.world * off, dtype=torch.uint8, device=device)
        self.header = self.send[:self.HEADER].view(torch.int64)
        self.pos = [self.send[o:o + 4 * c].view(torch.int32) for o, c in zip(self.pos_off, self.caps)]
        self.x = [self.send[o:o + 4 * c].view(torch.float32) for o, c in zip(self.x_off, self.caps)]
        self._hdr_host = torch.zeros(self.HEADER // 8, dtype=torch.int64)
        if device.type == "cuda":
            self._hdr_host = self._hdr_host.pin_memory()
        if group is not None and device.type == "cuda" and os.environ.get("BLISS_P2P", "1") != "0":
            try:
                self._init_p2p(group)
            except Exception as e:      # no symmetric memory on this system: the NCCL all-gather path stays
                warnings.warn(f"peer-memory bandit exchange unavailable ({type(e).__name__}: {e}); using NCCL all-gather")
                self.p2p = False

    def _init_p2p(self, group):
        """Symmetric-memory window per rank (``torch.distributed._symmetric_memory``: CUDA VMM allocations mapped into
        every rank of the node): [2 parities][world slots of ``stride`` bytes] ++ flags[2][L][world] uint64.  The reward
        kernel of every rank stores its updates into its slot of EVERY window over NVLink and raises a flag per layer;
        the apply kernel polls the local flags (``csrc/bandit.cu``) — the all-gather is fused into the producer.
        (``BLISS_P2P_PULL=1``: the producer stores locally only and the consumer reads the peers' windows — measured at
        N = 8: 1.31 ms per step against 0.74 ms for push, the scattered CAS loop then waits on NVLink read latency.)"""
        import torch.distributed._symmetric_memory as symm
        from . import _native as N
        L, W = len(self.caps), self.world
        self.flags_off = 2 * W * self.stride
        nbytes = self.flags_off + 2 * L * W * 8
        window = symm.empty(nbytes, dtype=torch.uint8, device=self.device)
        try:
            hdl = symm.rendezvous(window, group)
        except Exception:
            symm.enable_symm_mem_for_group(group.group_name)
            hdl = symm.rendezvous(window, group)
        window.zero_()
        self.window, self._hdl, self.rank = window, hdl, int(hdl.rank)
        self.peer_base = torch.tensor([int(p) for p in hdl.buffer_ptrs], dtype=torch.int64, device=self.device)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.device)     # exchange step: parity + flag value
        self.done_ctr = torch.zeros(L, dtype=torch.int32, device=self.device)
        self.err = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.mc_base = _multicast_base(hdl)
        self._structs = [N.P2P(peer_base=self.peer_base.data_ptr(), world=W, rank=self.rank, mc_base=self.mc_base,
                               parity_stride=W * self.stride, rank_stride=self.stride, count_off=8 * l,
                               pos_off=self.pos_off[l], x_off=self.x_off[l], flags_off=self.flags_off, layer=l,
                               n_layers=L, step_dev=self.step_dev.data_ptr(),
                               done_ctr=self.done_ctr[l:].data_ptr(),
                               pull=1 if os.environ.get("BLISS_P2P_PULL", "0") == "1" else 0) for l in range(L)]
        torch.cuda.synchronize(self.device)
        dist.barrier(group)               # every window is zeroed before anybody stores into it
        self.p2p = True

    def p2p_struct(self, layer: int):
        return self._structs[layer]

    def exchange(self, counts):
        for l, c in enumerate(counts):
            self._hdr_host[l] = int(c)
        self.header.copy_(self._hdr_host, non_blocking=True)
        dist.all_gather_into_tensor(self.recv, self.send, group=self.group)


def _multicast_base(hdl) -> int:
    """NVSwitch multicast address of a symmetric-memory window (0 when the fabric has none, or with BLISS_P2P_MC=0):
    a store to it is replicated into every rank's window by the switch, so a producer issues one store per value
    instead of one per rank."""
    if os.environ.get("BLISS_P2P_MC", "1") == "0":
        return 0
    try:
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
    except Exception:
        return 0
    return mc


class GradExchange:
    """Symmetric-memory windows for the gradient all-reduce fused into the Adam launch (``csrc/optim.cu``): every rank
    pushes its flat gradient into its slot of every rank's window, Adam adds the slots in rank order."""

    def __init__(self, n_params: int, world: int, device, group):
        import torch.distributed._symmetric_memory as symm
        from . import _native as N
        self.world, self.n = int(world), int(n_params)
        self.slot_bytes = (4 * self.n + 15) // 16 * 16
        self.flags_off = 2 * self.world * self.slot_bytes
        window = symm.empty(self.flags_off + 2 * self.world * 8, dtype=torch.uint8, device=device)
        try:
            hdl = symm.rendezvous(window, group)
        except Exception:
            symm.enable_symm_mem_for_group(group.group_name)
            hdl = symm.rendezvous(window, group)
        window.zero_()
        self.window, self._hdl, self.rank = window, hdl, int(hdl.rank)
        self.peer_base = torch.tensor([int(p) for p in hdl.buffer_ptrs], dtype=torch.int64, device=device)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=device)
        self.done_ctr = torch.zeros(1, dtype=torch.int32, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.mc_base = _multicast_base(hdl)
        self.struct = N.GradP2P(peer_base=self.peer_base.data_ptr(), world=self.world, rank=self.rank, mc_base=self.mc_base,
                                parity_stride=self.world * self.slot_bytes, slot_bytes=self.slot_bytes,
                                flags_off=self.flags_off, step_dev=self.step_dev.data_ptr(),
                                done_ctr=self.done_ctr.data_ptr())
        torch.cuda.synchronize(device)
        dist.barrier(group)

    def push(self, flat_grad: torch.Tensor):
        import ctypes as C
        from . import _native as N
        N.call("bliss_grad_push", N.ptr(flat_grad), flat_grad.numel(), C.byref(self.struct), N.stream())


class FlatAdam(torch.optim.Optimizer):
    """``torch.optim.Adam(params, lr)`` (``train_lightning.py:206``) over flat buffers: the parameters are
    re-homed as views of one flat fp32 tensor (like the gradients of :class:`FlatGrads`), the moments are
    flat too, and ``step()`` is ONE launch of ``bliss_adam_step`` that also clears the gradients.  ``lr``
    and the step count are device scalars, so a captured step keeps following ``param_groups[0]['lr']``
    (``StepLR``): call :meth:`sync_lr` before replaying."""

    def __init__(self, flat_grads: FlatGrads, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        params = flat_grads.params
        if not params or not params[0].is_cuda:
            raise RuntimeError("FlatAdam runs on CUDA parameters only (torch.optim.Adam covers the CPU tests)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        dev = params[0].device
        self.flat_g = flat_grads.flat
        self.flat_p = torch.zeros_like(self.flat_g)         # (padding columns / alignment gaps stay zero: their
        with torch.no_grad():                               # gradients and moments are zero, so Adam leaves them)
            for p, (off, k) in zip(params, flat_grads.layout):
                view, full = flat_view(self.flat_p, off, p, k)
                view.copy_(p.detach())
                p.data = view
                p._bliss_padded = full
        self.exp_avg = torch.zeros_like(self.flat_g)
        self.exp_avg_sq = torch.zeros_like(self.flat_g)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self._lr_host = float(lr)

    def sync_lr(self):
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_host:
            self.lr_dev.fill_(lr)
            self._lr_host = lr

    @torch.no_grad()
    def step(self, closure=None):
        from . import _native as N
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()
        g = self.param_groups[0]
        N.call("bliss_adam_step", N.ptr(self.flat_p), N.ptr(self.flat_g), N.ptr(self.exp_avg), N.ptr(self.exp_avg_sq),
               self.flat_p.numel(), N.ptr(self.lr_dev), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
               N.ptr(self.step_dev), 1, N.stream())

    @torch.no_grad()
    def step_p2p(self, gx: "GradExchange"):
        """Adam over the MEAN of all ranks' gradients read from the peer-memory window (after ``gx.push``): replaces
        ``all_reduce`` + ``div_`` + ``step``; advances the exchange's step counter."""
        import ctypes as C
        from . import _native as N
        g = self.param_groups[0]
        N.call("bliss_adam_step_p2p", N.ptr(self.flat_p), N.ptr(self.flat_g), N.ptr(self.exp_avg), N.ptr(self.exp_avg_sq),
               self.flat_p.numel(), N.ptr(self.lr_dev), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
               N.ptr(self.step_dev), C.byref(gx.struct), N.ptr(gx.err), N.stream())
        gx.step_dev.add_(1)

    def state_dict(self):
        return {"exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(), "step": self.step_dev.clone(),
                "lr": float(self.param_groups[0]["lr"])}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_dev.copy_(sd["step"])
        self.param_groups[0]["lr"] = sd["lr"]
        self._lr_host = None
        self.sync_lr()
