import sys, os, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bliss_gnn_b200 import _native as N
from bliss_gnn_b200.sampler import NeighborSampler, Frontier
from tests.util import random_graph
N.build()
g = random_graph(3000, 20000, seed=5, hubs=6, hub_degree=1500).to("cuda:0")
seeds = torch.arange(0, 48, dtype=torch.int32, device="cuda:0")
smp = NeighborSampler([10], rng_seed=13)
wsp = smp._bind(g)
W = g.csc_edata("w")
n = 48
N.call("bliss_frontier_plan", C.byref(wsp.gview), N.ptr(seeds), n, C.byref(wsp.ws), N.stream())
torch.cuda.synchronize(); c = wsp.read_counters(); print("plan: n_seeds", c.n_seeds, "n_cand", c.n_cand, "n_sel", c.n_sel, "chunks", c.n_chunks, "e_in", c.e_in, "err", c.error)
N.call("bliss_neighbor_select", C.byref(wsp.gview), n, 10, 13, 0, 0, C.byref(wsp.ws), N.stream())
torch.cuda.synchronize(); c = wsp.read_counters(); print("select: n_cand", c.n_cand, "n_sel", c.n_sel, "err", c.error)
kb = wsp.keep_bits[: 8 * c.n_chunks].cpu()
print("kept bits total", sum(bin(int(x) & 0xffffffff).count("1") for x in kb.tolist()), "expected", int(g.in_degrees(seeds.long()).clamp(max=10).sum()))
print("sel bits set", sum(bin(int(x) & 0xffffffff).count("1") for x in wsp.sel_bits.cpu().tolist()))
N.call("bliss_block_count", C.byref(wsp.gview), N.ptr(seeds), n, N.ptr(W), 0.1, 5, C.byref(wsp.ws), N.stream())
torch.cuda.synchronize(); c = wsp.read_counters()
sel = wsp.sel[: c.n_sel].long()
fp = wsp.first_pos[sel]
print("count: n_sel", c.n_sel, "unset first_pos", int((fp == -1).sum()), "row_cnt sum", int(wsp.row_cnt[:n].sum()), "part_cnt sum", int(wsp.part_cnt[:c.n_chunks].sum()))
kb2 = wsp.keep_bits[: 8 * c.n_chunks].cpu()
print("keep bits after count", sum(bin(int(x) & 0xffffffff).count("1") for x in kb2.tolist()), "same as before", bool((kb == kb2).all()))
ni = wsp.node_info.view(-1, 2)[sel]
print("node_info of selected:", ni[:5].tolist())
