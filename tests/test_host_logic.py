"""CPU tests of the host-side logic around the C ABI (no CUDA needed): scratch sizing of the balanced SpMM,
the packed bandit-exchange layout, the block-frame hooks of the data-parallel step, flag parsing."""
import torch

from bliss_gnn_b200 import ops
from bliss_gnn_b200.graph import Block, toy_graph
from bliss_gnn_b200.parallel import BanditExchange, shard_batches
from bliss_gnn_b200.train import Trainer, build_argparser


def test_spmm_tiling_covers_every_vector_width():
    """``ops._spmm_tiling`` must size the partial-sum scratch for whichever (VEC, NCH) ``launch_spmm`` picks
    (csrc/aggregate.cu: VEC in {4,2,1} by divisibility/alignment, NCH = next power of two <= 8)."""
    for d in (1, 7, 41, 64, 256, 300, 602, 604, 1030, 4096):
        width, tiles = ops._spmm_tiling(d)
        for vec in (4, 2, 1):
            if d % vec:
                continue
            per_lane = -(-d // (32 * vec))
            nch = 1
            while nch < per_lane and nch < 8:
                nch *= 2
            tile = 32 * vec * nch
            n = -(-d // tile)
            assert width >= n * tile and tiles >= n and width >= d


def test_exchange_layout_is_8_bytes_per_edge_and_aligned():
    caps = [1000, 333, 7]
    ex = BanditExchange(caps, world=3, device=torch.device("cpu"), group=None)
    assert ex.recv.numel() == 3 * ex.stride and ex.stride % 16 == 0
    end = BanditExchange.HEADER
    for l, c in enumerate(caps):
        assert ex.pos_off[l] >= end and ex.pos_off[l] % 4 == 0 and ex.x_off[l] == ex.pos_off[l] + 4 * c
        assert ex.pos[l].dtype == torch.int32 and ex.x[l].dtype == torch.float32
        assert ex.pos[l].numel() == c and ex.x[l].numel() == c
        end = ex.x_off[l] + 4 * c
    assert ex.stride - BanditExchange.HEADER <= 8 * sum(caps) + 16 * len(caps)
    ex.pos[1][:3] = torch.tensor([5, 6, 7], dtype=torch.int32)            # views alias the send buffer
    assert ex.send[ex.pos_off[1]:ex.pos_off[1] + 12].view(torch.int32).tolist() == [5, 6, 7]


def test_block_frame_hook_fires_on_assignment():
    idx = torch.zeros(1, dtype=torch.int32)
    b = Block(torch.tensor([0, 0], dtype=torch.int32), idx[:0], idx[:0], idx, idx)
    seen = []
    b.srcdata.on_set["embed_norm"] = lambda: seen.append(b.srcdata["embed_norm"].item())
    b.srcdata["embed_norm"] = torch.tensor(3.0)
    b.srcdata["other"] = torch.tensor(1.0)
    assert seen == [3.0]
    b.srcdata.on_set.pop("embed_norm")
    b.srcdata["embed_norm"] = torch.tensor(4.0)
    assert seen == [3.0]


def test_label_gather_plain_path_and_sharding():
    labels = torch.arange(10) * 3
    nid = torch.tensor([4, 0, 9], dtype=torch.int32)
    assert Trainer._gather_labels(labels.float(), nid).tolist() == [12.0, 0.0, 27.0]   # non-int64: plain indexing
    assert list(shard_batches(11, 1, 4)) == [1, 5] and list(shard_batches(3, 0, 4)) == []


def test_cli_flags_match_the_reference_defaults():
    """``train_lightning.py:489-552``: flag names and defaults."""
    a = build_argparser().parse_args([])
    assert (a.model, a.sampler, a.fan_out, a.batch_size, a.num_hidden, a.num_layers) == \
        ("sage", "poisson-bandit", "16384,8192,4096", 1024, 256, 3)
    assert (a.eta, a.lr, a.dropout, a.importance_sampling, a.precision) == (0.1, 0.002, 0.1, 1, "medium")
    assert toy_graph().num_nodes() == 5


def test_early_stopping_and_vertex_limit_controllers():
    """``EarlyStopping(monitor='val_acc', mode='max', stopping_threshold, patience)`` (train_lightning.py:627-634) and
    ``BatchSizeCallback`` (:425-486) restated without Lightning."""
    from bliss_gnn_b200.train import BatchSizeController, EarlyStopping
    es = EarlyStopping(stopping_threshold=0.9, patience=2)
    assert [es.check(v) for v in (0.5, 0.6, 0.6, 0.55)] == [False, False, False, True]      # two checks without a new best
    assert EarlyStopping(0.9, 100).check(0.95) and not EarlyStopping(1, 100).check(1.0)      # strictly above the threshold
    c = BatchSizeController(limit=1000)
    for x in (2000, 2100, 1900, 2050):
        c.push(x)
    assert abs(c.m - 2012.5) < 1e-9 and c.n == 4
    assert c.on_train_epoch_end(64) == int(64 * 1000 / 2012.5) and c.n == 0                  # rescaled, statistics cleared
    off = BatchSizeController(limit=-1)
    off.push(5.0)
    off.push(7.0)
    assert off.on_train_epoch_end(64) == 64                                                  # --vertex-limit -1: never


def test_sampler_factory_follows_the_flag():
    """``--sampler`` choices of the reference CLI (train_lightning.py:349-370, :538-543)."""
    from bliss_gnn_b200 import sampler as S
    from bliss_gnn_b200.train import make_sampler
    want = {"full": S.MultiLayerFullNeighborSampler, "neighbor": S.NeighborSampler, "bandit": S.BanditLadiesSampler,
            "poisson-bandit": S.PoissonBanditLadiesSampler, "ladies": S.LadiesSampler, "poisson-ladies": S.PoissonLadiesSampler}
    for name, cls in want.items():
        s = make_sampler(name, [16, 8, 4], eta=0.1)
        assert type(s) is cls, name
        assert len(s.nodes_per_layer) == 3
    assert make_sampler("full", [16, 8]).nodes_per_layer == [-1, -1]
    assert not make_sampler("neighbor", [5, 5]).attach_weights and make_sampler("ladies", [5, 5]).attach_weights
    args = build_argparser().parse_args(["--sampler", "neighbor", "--vertex-limit", "5000", "--early-stopping-patience", "3"])
    assert args.sampler == "neighbor" and args.vertex_limit == 5000 and args.early_stopping_patience == 3


def test_flat_layout_pads_marked_weights_in_place():
    """parallel.flat_layout: a 2-D parameter marked ``_bliss_pad_cols`` gets its zero columns inside the flat
    storage (the gradient view and the padded matrix alias the same memory), every tensor starts 16-byte aligned."""
    import torch
    from bliss_gnn_b200.parallel import FlatGrads, flat_layout, flat_view
    w0 = torch.nn.Parameter(torch.arange(6 * 5, dtype=torch.float32).view(6, 5))
    b0 = torch.nn.Parameter(torch.ones(6))
    w1 = torch.nn.Parameter(torch.ones(3, 6))
    w0._bliss_pad_cols = 3
    layout, total = flat_layout([w0, b0, w1])
    assert [off % 4 for off, _ in layout] == [0, 0, 0] and layout[0][1] == 3 and layout[1][1] == 0
    assert total >= 6 * 8 + 6 + 18
    fg = FlatGrads([w0, b0, w1])
    assert w0.grad.shape == (6, 5) and w0.grad.stride() == (8, 1) and w0._bliss_padded_grad.shape == (6, 8)
    w0.grad.fill_(2.0)
    assert float(w0._bliss_padded_grad[:, :5].min()) == 2.0 and float(w0._bliss_padded_grad[:, 5:].abs().max()) == 0.0
    assert float(fg.flat.sum()) == 2.0 * 30
    assert b0._bliss_padded_grad is None and b0.grad.is_contiguous()
    flat = torch.zeros(total)
    view, full = flat_view(flat, layout[0][0], w0, 3)
    view.copy_(w0.detach())
    assert torch.equal(full[:, :5], w0.detach()) and float(full[:, 5:].abs().max()) == 0.0


def test_lazy_renorm_guard_decides_from_the_weight_maximum():
    """sampler.tick_renorm (host logic, no device): with the pinned copy of the running maxima (step graphs) the
    re-scale fires only above e^(88 - 7W); on the schedule path it looks at the device value every ``renorm_every``
    updates and re-scales only when the room left is smaller than the next horizon."""
    import math
    import torch
    from bliss_gnn_b200.sampler import PoissonBanditLadiesSampler
    smp = PoissonBanditLadiesSampler([8, 4, 2], eta=0.1)
    L = 3
    smp._w_csc = [torch.ones(4) for _ in range(L)]
    smp._wmax, smp._wmax_host = torch.ones(L), torch.ones(L)
    calls = []
    smp._renormalize = lambda idx: calls.append(idx)
    for _ in range(500):                       # step graphs, small weights: never re-scaled, nothing counted
        smp.tick_renorm(L, mirrored=True)
    assert calls == [] and smp._updates_since_renorm == 0
    smp._wmax_host[1] = math.exp(82.0)         # above e^(88 - 7): re-scale every layer, the copy is reset
    smp.tick_renorm(L, mirrored=True)
    assert calls == [0, 1, 2] and float(smp._wmax_host.max()) == 1.0
    calls.clear()
    for _ in range(smp.renorm_every - 1):      # schedule path (eager steps): counts updates …
        smp.tick_renorm(L)
    assert calls == [] and smp._updates_since_renorm == smp.renorm_every - 1
    smp.tick_renorm(L)                         # … and at the horizon finds room for another one: counter reset only
    assert calls == [] and smp._updates_since_renorm == 0
    smp._wmax[0] = math.exp(30.0)              # room for fewer than renorm_every more updates: re-scale at the horizon
    for _ in range(smp.renorm_every):
        smp.tick_renorm(L)
    assert calls == [0, 1, 2]
    smp.normalize = "literal"                  # (the literal mode re-normalises after every update itself)
    calls.clear()
    smp.tick_renorm(L)
    assert calls == []
