"""ctypes binding of ``libbliss_b200.so`` (the C ABI in ``include/bliss_b200.h``).

Only raw device pointers (``tensor.data_ptr()``), sizes, scalars and the current CUDA stream
cross this boundary — no torch types.  The library is built in-tree by :func:`build`
(``nvcc -gencode arch=compute_100a,code=sm_100a``); there is NO fallback: if it is missing or a
call fails, the op raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libbliss_b200.so")
SOURCES = ["sampler.cu", "aggregate.cu", "bandit.cu", "gat.cu", "optim.cu", "epilogue.cu", "profile.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]

MODE_BANDIT, MODE_LADIES, MODE_UNIFORM, MODE_NEIGHBOR, MODE_PLANNED, COLLECT_BITMAP = 0, 1, 2, 4, 8, 16
AGG_SUM, AGG_MEAN = 0, 1


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)] + [os.path.join(_ROOT, "include", "bliss_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into ``libbliss_b200.so`` (in-tree)."""
    if not force and not _stale():
        return LIB_PATH
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(_CSRC, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(_CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out.decode(), file=sys.stderr)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out.decode()}")
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs, "-lcudart"]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if out.returncode != 0:
        raise RuntimeError(f"link failed:\n{out.stdout.decode()}")
    return LIB_PATH


# ---- C structs (mirror include/bliss_b200.h) ---------------------------------------------
class Graph(C.Structure):
    _fields_ = [("num_nodes", C.c_int64), ("num_edges", C.c_int64), ("indptr", C.c_void_p),
                ("indices", C.c_void_p), ("eid", C.c_void_p)]


class Counters(C.Structure):
    _fields_ = [("n_seeds", C.c_int32), ("n_cand", C.c_int32), ("n_sel", C.c_int32), ("n_src", C.c_int32),
                ("n_heavy", C.c_int32), ("n_light", C.c_int32), ("take_all", C.c_int32), ("iters", C.c_int32),
                ("e_in", C.c_int64), ("n_edges", C.c_int64), ("c", C.c_double), ("s_last", C.c_double),
                ("queue", C.c_int32 * 8), ("error", C.c_int32), ("n_chunks", C.c_int32)]


class Workspace(C.Structure):
    _fields_ = [("acc", C.c_void_p), ("first_pos", C.c_void_p), ("node_info", C.c_void_p),
                ("sel_bits", C.c_void_p), ("cand_bits", C.c_void_p), ("keep_bits", C.c_void_p), ("cand", C.c_void_p), ("p_cand", C.c_void_p), ("sel", C.c_void_p),
                ("row_a", C.c_void_p), ("row_d", C.c_void_p),
                ("chunk_first", C.c_void_p), ("chunk_rec", C.c_void_p), ("part_w", C.c_void_p), ("part_q", C.c_void_p),
                ("row_w", C.c_void_p), ("row_q", C.c_void_p),
                ("row_cnt", C.c_void_p), ("part_cnt", C.c_void_p), ("part_t", C.c_void_p), ("chunk_pre", C.c_void_p), ("row_t", C.c_void_p), ("cap_seeds", C.c_int64),
                ("cap_sel", C.c_int64), ("ctr", C.c_void_p), ("n_seeds_dev", C.c_void_p), ("step_dev", C.c_void_p),
                ("ctr_mirror", C.c_void_p)]


class BlockOut(C.Structure):
    _fields_ = [("indptr", C.c_void_p), ("edge_src", C.c_void_p), ("edge_dst", C.c_void_p),
                ("csc_pos", C.c_void_p), ("eid", C.c_void_p), ("q_ij", C.c_void_p), ("edge_w", C.c_void_p),
                ("src_nid", C.c_void_p), ("node_prob", C.c_void_p), ("out_deg", C.c_void_p),
                ("t_bits", C.c_void_p), ("t_words", C.c_int64),
                ("seg_ptr", C.c_void_p), ("inv_deg", C.c_void_p), ("cap_edges", C.c_int64), ("cap_src", C.c_int64), ("pad_src", C.c_int64), ("pad_rows", C.c_int64)]


class P2P(C.Structure):
    _fields_ = [("peer_base", C.c_void_p), ("world", C.c_int32), ("rank", C.c_int32), ("parity_stride", C.c_int64),
                ("rank_stride", C.c_int64), ("count_off", C.c_int64), ("pos_off", C.c_int64), ("x_off", C.c_int64),
                ("flags_off", C.c_int64), ("layer", C.c_int32), ("n_layers", C.c_int32), ("step_dev", C.c_void_p),
                ("done_ctr", C.c_void_p), ("pull", C.c_int32), ("pad_", C.c_int32), ("mc_base", C.c_uint64)]


class GradP2P(C.Structure):
    _fields_ = [("peer_base", C.c_void_p), ("world", C.c_int32), ("rank", C.c_int32), ("parity_stride", C.c_int64),
                ("slot_bytes", C.c_int64), ("flags_off", C.c_int64), ("step_dev", C.c_void_p), ("done_ctr", C.c_void_p),
                ("mc_base", C.c_uint64)]


_P, _I32, _I64, _U32, _U64, _F, _D = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_float, C.c_double
_GP, _WP, _BP = C.POINTER(Graph), C.POINTER(Workspace), C.POINTER(BlockOut)

#: every symbol ``include/bliss_b200.h`` declares, with its argument types
PROTOTYPES = {
    "bliss_version": [],
    "bliss_workspace_init": [_WP, _I64, _P],
    "bliss_frontier_plan": [_GP, _P, _I32, _WP, _P],
    "bliss_frontier_prob": [_GP, _P, _I32, _P, _F, _I32, _WP, _P],
    "bliss_poisson_scale": [_I32, _I32, _D, _I32, _WP, _P],
    "bliss_select_poisson": [_I32, _U64, _U64, _U32, _P, _WP, _P],
    "bliss_poisson_select": [_I32, _I32, _D, _U64, _U64, _U32, _P, _WP, _P],
    "bliss_select_topk": [_I32, _I32, _U64, _U64, _U32, _P, _P, _WP, _P],
    "bliss_philox_fill": [_U64, _U64, _U32, _P, _I64, _P, _P],
    "bliss_neighbor_select": [_GP, _I32, _I32, _U64, _U64, _U32, _WP, _P],
    "bliss_block_count": [_GP, _P, _I32, _P, _F, _I32, _WP, _P],
    "bliss_block_index": [_P, _I32, _WP, _BP, _P],
    "bliss_block_fill": [_GP, _P, _I32, _P, _F, _I32, _WP, _BP, _P],
    "bliss_block_finish": [_I32, _I32, _WP, _BP, _P],
    "bliss_sample_layer_front": [_GP, _P, _I32, _P, _F, _I32, _I32, _D, _I32, _U64, _U64, _U32, _P, _P, _WP, _BP, _P],
    "bliss_sample_layer_back": [_GP, _P, _I32, _P, _F, _I32, _WP, _BP, _P],
    "bliss_block_transpose": [_P, _P, _I64, _I32, _I32, _P, _P, _P, _P, _I64, _P, _P, _P, _I32, _P, _P, _P, _P],
    "bliss_gather_rows": [_P, _P, _I64, _I32, _P, _P, _P],
    "bliss_row_norm": [_P, _I64, _I32, _P, _P],
    "bliss_spmm": [_P, _P, _P, _P, _P, _P, _I32, _P, _I32, _I32, _P, _P, _I64, _P, _P],
    "bliss_gatv2_fwd": [_P, _P, _P, _P, _P, _F, _I32, _I32, _I32, _P, _P, _P, _P, _P],
    "bliss_gatv2_bwd_dst": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _I32, _I32, _I32, _P, _P, _P, _P],
    "bliss_gatv2_bwd_src": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _I32, _I32, _I32, _I32, _P, _P],
    "bliss_gat_alpha_sums": [_P, _P, _P, _I32, _P, _P, _P],
    "bliss_reward_update": [_GP, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _F, _I32, _I64,
                            _P, _P, _P, _P, _P, _P, _P, C.POINTER(P2P), _P, _P],
    "bliss_apply_updates_p2p": [C.POINTER(P2P), _I64, _P, _P, _P, _P, _P],
    "bliss_apply_updates": [_P, _P, _I64, _P, _P, _P, _P],
    "bliss_apply_updates_packed": [_P, _I64, _I32, _I64, _I64, _I64, _I64, _P, _P, _P, _P],
    "bliss_l1_norm": [_P, _I64, _P, _P, _P],
    "bliss_scale_by_inv": [_P, _I64, _P, _D, _P],
    "bliss_adam_step": [_P, _P, _P, _P, _I64, _P, _F, _F, _F, _P, _I32, _P],
    "bliss_splitk_accumulate": [_P, _I32, _I32, _I32, _I32, _P, _P],
    "bliss_grad_push": [_P, _I64, C.POINTER(GradP2P), _P],
    "bliss_adam_step_p2p": [_P, _P, _P, _P, _I64, _P, _F, _F, _F, _P, C.POINTER(GradP2P), _P, _P],
    "bliss_sage_epilogue_parts": [],
    "bliss_xent_mean": [_P, _P, _I32, _I32, _P, _P, _P, _P],
    "bliss_sage_epilogue_fwd": [_P, _P, _P, _I32, _I32, _I32, _F, _U64, _P, _U32, _P, _P, _P],
    "bliss_sage_epilogue_bwd": [_P, _P, _I32, _I32, _I32, _F, _P, _P, _P, _P],
    "bliss_profile_enable": [_I32],
    "bliss_profile_read": [C.c_char_p, _I32, _P, _P, _I32],
    "bliss_l2_gather_probe": [_P, _I32, _I32, _I32, _P, _I32, _P],
}

_lib = None


def lib():
    """The loaded library; raises (never falls back) when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for the BLISS hot path)")
        _lib = C.CDLL(LIB_PATH)
        for name, argtypes in PROTOTYPES.items():
            fn = getattr(_lib, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
    return _lib


class BlissNativeError(RuntimeError):
    pass


#: kernels each entry point launches (for the ``gpu_launches`` count of bench.py)
LAUNCHES = {"bliss_frontier_prob": 4, "bliss_sample_layer_front": 11, "bliss_poisson_select": 2, "bliss_frontier_plan": 3, "bliss_sample_layer_back": 2,
            "bliss_select_topk": 3, "bliss_block_transpose": 3, "bliss_l1_norm": 2, "bliss_version": 0,
            "bliss_adam_step": 2, "bliss_adam_step_p2p": 3, "bliss_spmm": 2, "bliss_apply_updates_p2p": 2, "bliss_sage_epilogue_bwd": 2,
            "bliss_sage_epilogue_parts": 0, "bliss_xent_mean": 2}


class _Stats:
    """Launch counter and optional per-entry-point CUDA-event timing (bench.py roofline)."""

    def __init__(self):
        self.launches = 0
        self.timing = False
        self.events = {}

    def reset(self, timing=False):
        self.launches = 0
        self.timing = timing
        self.events = {}

    def elapsed_ms(self):
        """{entry point: (calls, total ms)} — call after a synchronize."""
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in self.events.items()}


STATS = _Stats()


def profile_enable(on: bool):
    """Per-kernel CUDA-event timing inside the library (csrc/profile.cu); clears earlier records."""
    lib().bliss_profile_enable(1 if on else 0)


def profile_read():
    """{kernel name: (launches, total ms)} of everything recorded since :func:`profile_enable` (synchronises)."""
    cap = 256
    names = C.create_string_buffer(1 << 14)
    ms = (C.c_float * cap)()
    calls = (C.c_int32 * cap)()
    n = lib().bliss_profile_read(names, len(names), C.cast(ms, C.c_void_p), C.cast(calls, C.c_void_p), cap)
    if n < 0:
        raise BlissNativeError(f"bliss_profile_read failed: {n}")
    keys = names.value.decode().split("\n") if n else []
    return {k: (int(calls[i]), float(ms[i])) for i, k in enumerate(keys)}


def call(name: str, *args):
    """Invoke one C-ABI entry point on the current stream; raises on any non-zero return."""
    fn = getattr(lib(), name)
    if STATS.timing:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        STATS.events.setdefault(name, []).append((e0, e1))
    else:
        rc = fn(*args)
    STATS.launches += LAUNCHES.get(name, 1)
    check(rc, name)


def check(rc: int, what: str):
    if rc != 0:
        kind = "bad argument" if rc < 0 else "cudaError"
        raise BlissNativeError(f"{what} failed: {kind} {rc}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    """Raw handle of torch's current CUDA stream (the C call, not the slow Python Stream object)."""
    import torch
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())
