"""World-size-2 gloo tests (CPU) of the data-parallel host logic: batch sharding, the flat gradient
all-reduce, and the sparse bandit-update exchange.  R>1 semantics (SURVEY.md §7 hard part 7): every rank
samples from the same frozen EXP3 weights, then all ranks apply all R updates — checked against the oracle
doing exactly that in one process."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bliss_gnn_b200.parallel import BanditExchange, FlatGrads, gather_updates, shard_batches
from oracle import samplers as osamp
from tests.util import philox_uniform_fn, random_graph

WORLD = 2
FAN = [48, 24]


def _oracle_update(g, seeds, rank, weights=None):
    """One rank's sparse update (edge id, clamped exponent) per layer, from frozen weights."""
    s = osamp.PoissonBanditLadiesSampler(FAN, eta=0.1, uniform_fn=philox_uniform_fn(100 + rank, 0))
    if weights is not None:
        s.exp3_weights = weights.clone()
    _, _, blocks = s.sample_blocks(g, seeds)
    frozen = s.exp3_weights.clone()
    gen = torch.Generator().manual_seed(7 + rank)
    out = []
    for l, b in enumerate(blocks):
        b.srcdata["embed_norm"] = torch.rand(b.num_src_nodes(), generator=gen) + 0.5
        s.calculate_rewards(l, b, g, s.calculate_alpha(b))
        s.update_exp3_weights(l, b, g)
        out.append((b.edata["_ID"].long().clone(), s.trace["delta_reward"][l].clone()))
    return frozen, out


def _worker(rank, init_file, result_file):
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=WORLD)
    torch.set_num_threads(1)
    g = random_graph(500, 3000, seed=3)
    # (1) sharding: disjoint, same count on every rank
    mine = list(shard_batches(11, rank, WORLD))
    assert mine == list(range(rank, 10, WORLD))
    # (2) flat gradient all-reduce == mean of the per-rank gradients
    torch.manual_seed(0)
    lin = torch.nn.Linear(5, 3)
    fg = FlatGrads(lin.parameters())
    fg.zero_()
    x = torch.full((4, 5), float(rank + 1))
    lin(x).sum().backward()
    local = fg.flat.clone()
    fg.all_reduce_mean_(dist.group.WORLD)
    both = [torch.empty_like(local) for _ in range(WORLD)]
    dist.all_gather(both, local)
    torch.testing.assert_close(fg.flat, sum(both) / WORLD)
    assert lin.weight.grad.data_ptr() == fg.flat.data_ptr()          # grads are views of the flat buffer
    # (3) bandit exchange: ragged all-gather of (position, exponent), applied multiplicatively by all
    seeds = torch.arange(rank * 40, rank * 40 + 32)
    frozen, upd = _oracle_update(g, seeds, rank)
    w = frozen.clone()
    for l, (eid, xe) in enumerate(upd):
        for eid_r, x_r in gather_updates(eid, xe.float(), dist.group.WORLD):
            w[l, eid_r] = w[l, eid_r] * torch.exp(x_r)               # edge ids are unique within a rank's block
        w[l] = w[l] / w[l].double().sum().float()
    # (4) the packed one-collective exchange carries exactly the same (position, exponent) lists
    caps = [max(int(e.numel()) for e, _ in upd) + 7 + rank * 0] * len(upd)
    t = torch.tensor(caps)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ex = BanditExchange(t.tolist(), WORLD, torch.device("cpu"), dist.group.WORLD)
    for l, (eid, xe) in enumerate(upd):
        ex.pos[l][:eid.numel()] = eid.to(torch.int32)
        ex.x[l][:eid.numel()] = xe.float()
    ex.exchange([e.numel() for e, _ in upd])
    for l, (eid, xe) in enumerate(upd):
        ref = gather_updates(eid, xe.float(), dist.group.WORLD)
        for r in range(WORLD):
            base = ex.recv[r * ex.stride:(r + 1) * ex.stride]
            n = int(base[:64].view(torch.int64)[l])
            pos_r = base[ex.pos_off[l]:ex.pos_off[l] + 4 * n].view(torch.int32).long()
            x_r = base[ex.x_off[l]:ex.x_off[l] + 4 * n].view(torch.float32)
            assert n == ref[r][0].numel() and torch.equal(pos_r, ref[r][0]) and torch.equal(x_r, ref[r][1])
    if rank == 0:
        torch.save(w, result_file)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_gloo_sharding_grads_and_bandit_exchange():
    with tempfile.TemporaryDirectory() as d:
        init_file, result_file = os.path.join(d, "init"), os.path.join(d, "w.pt")
        mp.spawn(_worker, args=(init_file, result_file), nprocs=WORLD, join=True)
        w_dp = torch.load(result_file)
    # single-process statement of the same semantics: R batches from frozen weights, then R updates
    g = random_graph(500, 3000, seed=3)
    ref = None
    updates = []
    for rank in range(WORLD):
        frozen, upd = _oracle_update(g, torch.arange(rank * 40, rank * 40 + 32), rank)
        ref = frozen.clone() if ref is None else ref
        updates.append(upd)
    for l in range(len(FAN)):
        for upd in updates:
            eid, xe = upd[l]
            ref[l, eid] = ref[l, eid] * torch.exp(xe.float())
        ref[l] = ref[l] / ref[l].double().sum().float()
    torch.testing.assert_close(w_dp, ref, rtol=1e-6, atol=0)
    assert abs(float(w_dp[0].double().sum()) - 1.0) < 1e-5
