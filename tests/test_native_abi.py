"""CPU checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports every
symbol ``include/bliss_b200.h`` declares (no compute calls without a GPU); the product path fails
loudly without CUDA instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "bliss_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"^\s*int\s+(bliss_\w+)\s*\(", txt, flags=re.M)))


def test_header_declares_expected_entry_points():
    syms = _header_symbols()
    assert len(syms) >= 24
    for must in ("bliss_frontier_prob", "bliss_poisson_scale", "bliss_select_poisson", "bliss_block_fill",
                 "bliss_spmm", "bliss_gatv2_fwd", "bliss_reward_update"):
        assert must in syms


def test_library_exports_every_declared_symbol(native_lib):
    from bliss_gnn_b200 import _native
    for name in _header_symbols():
        assert hasattr(native_lib, name), f"{name} declared in bliss_b200.h but not exported"
        assert name in _native.PROTOTYPES, f"{name} has no ctypes prototype"
    assert set(_native.PROTOTYPES) == set(_header_symbols())
    assert native_lib.bliss_version() == 100


def test_struct_layouts_match_header(tmp_path):
    """ctypes mirrors vs the real header: compile a C probe with gcc and compare sizeof/offsetof."""
    import subprocess
    from bliss_gnn_b200 import _native as N
    probe = tmp_path / "probe.c"
    probe.write_text(r"""
#include <stdio.h>
#include <stddef.h>
#include "bliss_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(bliss_graph), sizeof(bliss_counters),
         sizeof(bliss_workspace), sizeof(bliss_block_out), offsetof(bliss_counters, c), offsetof(bliss_counters, error),
         offsetof(bliss_workspace, ctr), offsetof(bliss_block_out, cap_edges), sizeof(bliss_p2p),
         offsetof(bliss_p2p, flags_off), offsetof(bliss_p2p, done_ctr), sizeof(bliss_grad_p2p),
         offsetof(bliss_grad_p2p, step_dev), offsetof(bliss_workspace, ctr_mirror), offsetof(bliss_p2p, mc_base),
         offsetof(bliss_grad_p2p, mc_base));
  return 0;
}
""")
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(probe), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(N.Graph), ctypes.sizeof(N.Counters), ctypes.sizeof(N.Workspace), ctypes.sizeof(N.BlockOut),
            N.Counters.c.offset, N.Counters.error.offset, N.Workspace.ctr.offset, N.BlockOut.cap_edges.offset,
            ctypes.sizeof(N.P2P), N.P2P.flags_off.offset, N.P2P.done_ctr.offset, ctypes.sizeof(N.GradP2P),
            N.GradP2P.step_dev.offset, N.Workspace.ctr_mirror.offset, N.P2P.mc_base.offset, N.GradP2P.mc_base.offset]
    assert got == want


def test_bad_arguments_are_rejected_without_a_gpu(native_lib):
    assert native_lib.bliss_philox_fill(0, 0, 0, None, -1, None, None) < 0
    assert native_lib.bliss_spmm(None, None, None, None, None, None, 0, None, -1, 8, None, None, 0, None, None) < 0
    assert native_lib.bliss_adam_step(None, None, None, None, 4, None, 0.9, 0.999, 1e-8, None, 1, None) < 0
    assert native_lib.bliss_gather_rows(None, None, 4, 8, None, None, None) < 0


def test_product_path_refuses_cpu():
    """No CPU fallback: sampling on a CPU graph or aggregating CPU tensors raises."""
    from bliss_gnn_b200 import ops
    from bliss_gnn_b200.graph import normalized_edata, toy_graph
    from bliss_gnn_b200.sampler import PoissonBanditLadiesSampler
    g = toy_graph()
    g.edata["w"] = normalized_edata(g)
    with pytest.raises(RuntimeError, match="CUDA"):
        PoissonBanditLadiesSampler([2]).sample_blocks(g, torch.tensor([0, 1]))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.row_norm(torch.ones(4, 4))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "bliss_gnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
