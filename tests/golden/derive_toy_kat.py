"""Derives tests/golden/toy_kat.json: one poisson-bandit layer + EXP3 update on the reference's
ToyDataset (load_graph.py:91-119) in plain-Python float64 scalar arithmetic, straight from the
formulas of bandit_sampler.py (line numbers in comments).  It imports neither the oracle nor the
product, so it pins both.  NOT DGL output (DGL is not installable here) — see SURVEY.md §8(c).

    python tests/golden/derive_toy_kat.py
"""
import json
import math
import os

# toy graph after remove_self_loop + add_self_loop (train_lightning.py:334-335)
edges = [(2, 0), (3, 0), (3, 1), (4, 1)] + [(v, v) for v in range(5)]          # eid -> (src, dst)
V = 5
in_edges = {v: [e for e, (s, d) in enumerate(edges) if d == v] for v in range(V)}   # CSC column order
in_deg = [len(in_edges[v]) for v in range(V)]
w_static = [1.0 / in_deg[d] for (s, d) in edges]                                # normalized_edata :20-27
eta, fanout, seeds = 0.4, 2, [0, 1]
exp3 = [1.0] * len(edges)                                                       # :343
u_inject = {0: 0.9, 1: 0.9, 2: 0.5, 3: 0.5, 4: 0.2}                             # per global node id

# in_subgraph + compact_graphs (:123-125): edge list in seed order / CSC order, first-occurrence ids
ins = [(e, edges[e][0], i) for i, s in enumerate(seeds) for e in in_edges[s]]   # (eid, src, dst_local)
nodes = list(seeds)
for _, s, _ in ins:
    if s not in nodes:
        nodes.append(s)
# q_ij (:129-137)
q = []
for e, s, i in ins:
    wsum = sum(exp3[e2] for e2 in in_edges[seeds[i]])
    q.append(eta / in_deg[seeds[i]] + (1 - eta) * exp3[e] / wsum)
# p_j (:67-75)
qsum = [sum(q[k] for k, (_, _, i2) in enumerate(ins) if i2 == i) for i in range(len(seeds))]
prob = [math.sqrt(sum((q[k] / qsum[i]) ** 2 for k, (_, s, i) in enumerate(ins) if s == n)) for n in nodes]
# Poisson scale search (:391-401)
c, iters = 1.0, 0
for it in range(50):
    iters = it + 1
    S = sum(min(p * c, 1.0) for p in prob)
    if min(S, fanout) / max(S, fanout) >= 0.9999:
        break
    c *= fanout / S
P = [1.0 if n in seeds else min(p * c, 1.0) for n, p in zip(nodes, prob)]       # :403-406
selected = [k for k, n in enumerate(nodes) if u_inject[n] < P[k]]               # :422-424  (u < P)
# generate_block (:285-337): keep edges whose source is selected, relabel, normalise
src_nodes = [nodes[k] for k in selected]                                         # seeds first, first-occurrence order
kept = [(k, e, s, i) for k, (e, s, i) in enumerate(ins) if nodes.index(s) in selected]
wt = [q[k] / P[nodes.index(s)] for k, e, s, i in kept]                           # :314
for i in range(len(seeds)):
    rows = [j for j, (_, _, _, i2) in enumerate(kept) if i2 == i]
    tot = sum(wt[j] for j in rows)
    for j in rows:
        wt[j] *= len(rows) / tot                                                 # :316-320
block = {
    "src_nid": src_nodes, "dst_nid": seeds,
    "edge_src_local": [src_nodes.index(s) for _, _, s, _ in kept],
    "edge_dst_local": [i for _, _, _, i in kept],
    "eid": [e for _, e, _, _ in kept],
    "q_ij": [q[k] for k, _, _, _ in kept],
    "edge_weights": wt,
    "node_prob": [P[k] for k in selected],
}
# EXP3 update with embed_norm = 1, alpha = static w (:157-249)
kdeg = [sum(1 for x in kept if x[3] == i) for i in range(len(seeds))]
rewards, xs = [], []
for j, (k, e, s, i) in enumerate(kept):
    r = (w_static[e] ** 2 / kdeg[i]) * (1.0 / q[k] ** 2)                         # :186-191
    x = min(1.0, (r / P[nodes.index(s)]) * (0.01 / in_deg[seeds[i]]))            # :240-244
    rewards.append(r)
    xs.append(x)
    exp3[e] *= math.exp(x)                                                       # :246-248
tot = sum(exp3)
exp3 = [v / tot for v in exp3]                                                   # :249

out = {
    "provenance": "hand formulas of bandit_sampler.py in float64 (tests/golden/derive_toy_kat.py); NOT DGL output",
    "eta": eta, "fanout": fanout, "seeds": seeds, "u_inject": [u_inject[v] for v in range(V)],
    "insg_eid": [e for e, _, _ in ins], "insg_nodes": nodes, "q_insg": q, "prob_unnormalised": prob,
    "c": c, "iters": iters, "P": P, "selected_local": selected, "block": block,
    "rewards": rewards, "x": xs, "exp3_after": exp3, "w_static": w_static,
}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "toy_kat.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
