// sampler.cu — per-layer sampling of the BLISS hot path on sm_100a:
//   (1) layer-importance probabilities over the frontier's CSC neighbourhood,
//   (2) Poisson scale search + Philox inclusion sampling (or top-k for the multinomial samplers),
//   (3) block construction: filter, ordered compaction, relabelling, importance-weight
//       normalisation.
// Replaces bandit_sampler.py:47-138,269-425 and ladies_sampler.py:34-183 of the reference
// (which reach DGL in_subgraph/compact_graphs/g-SpMM/g-SDDMM/to_block and torch sort/unique/
// bernoulli).  Work distribution: the frontier's rows (in-edge lists of the seeds) are cut into
// 256-edge chunks with a 32-byte record each; every edge pass gives one warp one chunk, dealt
// round-robin to the resident warps.  All sizes live in device counters, so no size ever travels to
// the host inside a layer (or a step: the whole training step replays as one CUDA graph).
#include <cooperative_groups.h>
#include "common.cuh"
#include "profile.cuh"

namespace cg = cooperative_groups;

namespace bliss {

struct GraphView {
  const int64_t* __restrict__ indptr;
  const int32_t* __restrict__ indices;
  const int32_t* __restrict__ eid;
  int64_t num_nodes;
};

// ------------------------------------------------------------------------------------------
// plan, two launches: (a) k_plan_rows, one thread per seed over as many CTAs as needed — a single
// SM retires about one scattered sector per cycle, so registering thousands of seeds from one CTA
// costs more than the three probability passes of a small layer — registers the seed (candidate
// slot = local id = seed rank, P = 1, selected bit), stores the row's CSC start / degree and its
// number of 256-edge warp-chunks; (b) k_plan_scan, one CTA, turns the chunk counts into the prefix
// array chunk_first and resets the layer's counters.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_plan_rows(GraphView g, const int32_t* __restrict__ seeds, int n_seeds,
                                                  bliss_workspace ws) {
  pdl_wait();
  pdl_trigger();
  // sync-free chaining of layers: the true seed count may live on the device (the previous
  // layer's n_src); the host value is then only the capacity
  if (ws.n_seeds_dev) n_seeds = min(n_seeds, *ws.n_seeds_dev);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_seeds) return;
  const int s = seeds[i];
  const long long a = g.indptr[s];
  const int d = (int)min((long long)(g.indptr[s + 1] - a), (long long)INT_MAX);
  ws.cand[i] = s;
  *reinterpret_cast<int2*>(&ws.node_info[2 * s]) = make_int2(i, __float_as_int(1.0f));
  atomicOr(&ws.sel_bits[s >> 5], 1u << (s & 31));
  ws.row_a[i] = a;
  ws.row_d[i] = d;
  ws.row_cnt[i] = 0;
  ws.chunk_first[i] = max(1, (d + BLISS_CHUNK - 1) / BLISS_CHUNK);
}

// 4 consecutive ints per thread; 128-bit accesses when the array allows it (warp-uniform test)
__device__ __forceinline__ void load4(const int32_t* __restrict__ p, int i0, int n, int (&v)[4]) {
  if (i0 + 3 < n && (((uintptr_t)p) & 15) == 0) {
    const int4 t = *reinterpret_cast<const int4*>(p + i0);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = (i0 + u < n) ? p[i0 + u] : 0;
  }
}
__device__ __forceinline__ void store4(int32_t* __restrict__ p, int i0, int n, const int (&v)[4]) {
  if (i0 + 3 < n && (((uintptr_t)p) & 15) == 0) {
    *reinterpret_cast<int4*>(p + i0) = make_int4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i0 + u < n) p[i0 + u] = v[u];
  }
}

__global__ void __launch_bounds__(1024) k_plan_scan(int n_seeds, bliss_workspace ws) {
  pdl_wait();
  pdl_trigger();
  __shared__ int s_scan[40];
  __shared__ unsigned long long s_e[32];
  bliss_counters* ctr = ws.ctr;
  if (ws.n_seeds_dev) n_seeds = min(n_seeds, *ws.n_seeds_dev);
  unsigned long long e_in = 0;
  int chunk_base = 0;
  for (int b = 0; b < n_seeds; b += blockDim.x * 4) {
    const int i0 = b + threadIdx.x * 4;
    int nch[4], d[4], pre[4];
    load4(ws.chunk_first, i0, n_seeds, nch);
    load4(ws.row_d, i0, n_seeds, d);
    const int sum = nch[0] + nch[1] + nch[2] + nch[3];
    e_in += (unsigned long long)d[0] + d[1] + d[2] + d[3];
    int tc;
    int pc = chunk_base + block_excl_scan(sum, s_scan, &tc);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      pre[u] = pc;
      pc += nch[u];
    }
    store4(ws.chunk_first, i0, n_seeds, pre);
    chunk_base += tc;
  }
  e_in = block_sum(e_in, s_e);
  if (threadIdx.x == 0) {
    ctr->n_seeds = n_seeds;
    ctr->n_cand = n_seeds;
    ctr->n_sel = 0;
    ctr->n_src = n_seeds;
    ctr->n_heavy = 0;
    ctr->n_light = 0;
    ctr->take_all = 0;
    ctr->iters = 0;
    ctr->e_in = (int64_t)e_in;
    ctr->n_edges = 0;
    ctr->c = 1.0;
    ctr->s_last = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) ctr->queue[q] = 0;
    ctr->error = 0;
    ctr->n_chunks = chunk_base;
    ws.chunk_first[n_seeds] = chunk_base;
  }
}

// ------------------------------------------------------------------------------------------
// (1) frontier probabilities.
//   BANDIT: W_i = Σ_j w_ij ; q_ij = η/n_i + (1-η) w_ij/W_i ; Q_i = Σ_j q_ij ;
//           acc[src] += fx((q_ij/Q_i)^2)                          bandit_sampler.py:129-137,67-73
//   LADIES: acc[src] += fx(w_ij^2)                                ladies_sampler.py:46-47
//   UNIFORM flag: acc[src] |= fx(1)                               bandit_sampler.py:79-81
// The scatter is a fire-and-forget 64-bit integer reduction (RED: order independent, nothing to
// wait for).  The candidate list is collected afterwards (k_collect_candidates): a dense scan of
// the accumulators when |V| is moderate (BLISS_COLLECT_BITMAP clear), else from a candidate
// bitmap the scatter marks with a second fire-and-forget RED.OR.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void scatter_term(bool uniform, bool bitmap, float t, int src, double fx_scale,
                                             const bliss_workspace& ws) {
  if (uniform)
    atomicOr((unsigned long long*)&ws.acc[src], (unsigned long long)fx_scale);
  else
    atomicAdd((unsigned long long*)&ws.acc[src], fx_term(t, fx_scale));
  if (bitmap) {  // sparse-frontier mode: mark the candidate bit unless L2 already shows it set
    const unsigned bit = 1u << (src & 31);
    if (!(__ldcg(&ws.cand_bits[src >> 5]) & bit)) atomicOr(&ws.cand_bits[src >> 5], bit);
  }
}

// Every edge pass of a layer (three probability passes, kept-edge count, block fill) works on
// fixed 256-edge warp-chunks of the frontier's rows: one warp per chunk, 8 values per lane in
// registers, no CTA barrier inside a chunk — perfectly balanced whatever the degree distribution,
// no per-row chain of dependent round trips, no giant-row tail.  A row quantity (Σw, Σq, kept
// count, ΣW~) is the fixed-order sum of its chunks' partials, so results are deterministic.
//   pass 1  partW[c] = Σ w                       (weights streamed once from HBM)
//   pass 2  W_i = Σ_c partW ; q = η/n + (1-η) w / W_i ; partQ[c] = Σ q      (weights from L2)
//   pass 3  Q_i = Σ_c partQ ; RED acc[src] += fx((q / Q_i)^2)              (indices from HBM)
// CTAs pull groups of 8 consecutive chunks from a device-side queue (one atomic per group, issued
// at the top of an iteration and consumed at its end, so its latency hides behind the chunk).
struct ChunkRef {   // == bliss_chunk_rec (32 bytes, written by k_plan_chunks): everything a pass needs about a chunk
  int64_t a;        // CSC position of the chunk's first edge
  int len;          // edges in the chunk (<= 256)
  int row;          // seed rank
  int d;            // in-degree of the row
  int k0;           // offset of the chunk inside its row
  int c_first, c_last;  // chunk range of the row [c_first, c_last)
};
static_assert(sizeof(ChunkRef) == 32, "chunk record layout");
__device__ __forceinline__ ChunkRef chunk_ref(const bliss_workspace& ws, int c) {
  // one 32-byte record per chunk (two broadcast 128-bit loads) instead of a search plus four dependent loads
  const int4* __restrict__ p = reinterpret_cast<const int4*>(ws.chunk_rec) + 2 * (int64_t)c;
  const int4 lo = __ldg(p), hi = __ldg(p + 1);
  ChunkRef r;
  r.a = ((int64_t)(unsigned)lo.x) | ((int64_t)lo.y << 32);
  r.len = lo.z;
  r.row = lo.w;
  r.d = hi.x;
  r.k0 = hi.y;
  r.c_first = hi.z;
  r.c_last = hi.w;
  return r;
}
// one warp per row: the row's chunk records
__global__ void __launch_bounds__(256) k_plan_chunks(bliss_workspace ws) {
  pdl_wait();
  pdl_trigger();
  const int n_seeds = ws.ctr->n_seeds;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < n_seeds; i += nwarps) {
    const int64_t a = ws.row_a[i];
    const int d = ws.row_d[i];
    const int c_first = ws.chunk_first[i], c_last = ws.chunk_first[i + 1];
    for (int c = c_first + lane_id(); c < c_last; c += 32) {
      const int k0 = (c - c_first) * BLISS_CHUNK;
      const int64_t pos = a + k0;
      int4* p = reinterpret_cast<int4*>(ws.chunk_rec) + 2 * (int64_t)c;
      p[0] = make_int4((int)(pos & 0xffffffffll), (int)(pos >> 32), min(BLISS_CHUNK, d - k0), i);
      p[1] = make_int4(d, k0, c_first, c_last);
    }
  }
}
// Σ of a row's chunk partials: lanes take the chunks round-robin (the first three per lane are loaded
// together: 96 chunks = 24 K edges cover all but the largest hubs in one round trip), then a
// butterfly — a fixed order (deterministic), every lane ends with the same bits.
__device__ __forceinline__ double row_partials(const double* __restrict__ part, int c_first, int c_last) {
  const int c = c_first + lane_id();
  const double v0 = (c < c_last) ? __ldg(part + c) : 0.0;
  const double v1 = (c + 32 < c_last) ? __ldg(part + c + 32) : 0.0;
  const double v2 = (c + 64 < c_last) ? __ldg(part + c + 64) : 0.0;
  double t = (v0 + v1) + v2;
  for (int cc = c + 96; cc < c_last; cc += 32) t += __ldg(part + cc);
  return warp_sum(t);
}
__device__ __forceinline__ float row_total(const double* __restrict__ part, int c_first, int c_last) {
  return __double2float_rn(row_partials(part, c_first, c_last));
}

// for (ChunkLoop q(cursor, n_chunks); q.more(); q.next()) { const int c = q.chunk(); ... }
// Chunks are dealt round-robin to the resident warps (chunks cost about the same, and thousands of
// same-address queue atomics per pass serialise in L2 for longer than the imbalance they remove).
struct ChunkLoop {
  int n_chunks, item;
  __device__ __forceinline__ ChunkLoop(int*, int n_chunks_)
      : n_chunks(n_chunks_), item((blockIdx.x * blockDim.x + threadIdx.x) >> 5) {}
  __device__ __forceinline__ bool more() const { return item < n_chunks; }   // warp-uniform
  __device__ __forceinline__ int chunk() const { return item; }
  __device__ __forceinline__ void next() { item += (gridDim.x * blockDim.x) >> 5; }
};

__global__ void __launch_bounds__(BLISS_CTA, 6) k_prob_pass1(const float* __restrict__ W, bliss_workspace ws) {
  pdl_wait();
  pdl_trigger();
  const int lane = lane_id();
  const int n_chunks = ws.ctr->n_chunks;
  for (ChunkLoop q(&ws.ctr->queue[0], n_chunks); q.more(); q.next()) {
    const int c = q.chunk();
    const ChunkRef r = chunk_ref(ws, c);
    const float* __restrict__ wr = W + r.a;
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < BLISS_CHUNK / 32; ++j) {
      const int k = lane + 32 * j;
      acc += (k < r.len) ? (double)__ldg(wr + k) : 0.0;
    }
    acc = warp_sum(acc);
    if (lane == 0) ws.part_w[c] = acc;
  }
}

__global__ void __launch_bounds__(BLISS_CTA, 6) k_prob_pass2(const float* __restrict__ W, float eta, float one_minus_eta,
                                                         bliss_workspace ws) {
  pdl_wait();
  pdl_trigger();
  const int lane = lane_id();
  const int n_chunks = ws.ctr->n_chunks;
  for (ChunkLoop q(&ws.ctr->queue[1], n_chunks); q.more(); q.next()) {
    const int c = q.chunk();
    const ChunkRef r = chunk_ref(ws, c);
    const float* __restrict__ wr = W + r.a;
    float v[BLISS_CHUNK / 32];
#pragma unroll
    for (int j = 0; j < BLISS_CHUNK / 32; ++j) {
      const int k = lane + 32 * j;
      v[j] = (k < r.len) ? __ldg(wr + k) : 0.0f;
    }
    const float row_w = row_total(ws.part_w, r.c_first, r.c_last);
    const float eta_n = __fdiv_rn(eta, (float)r.d);
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < BLISS_CHUNK / 32; ++j) {
      const int k = lane + 32 * j;
      acc += (k < r.len) ? (double)edge_q(v[j], row_w, eta_n, one_minus_eta) : 0.0;
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      ws.part_q[c] = acc;
      if (c == r.c_first) ws.row_w[r.row] = row_w;
    }
  }
}

__global__ void __launch_bounds__(BLISS_CTA, 5) k_prob_pass3(GraphView g, const float* __restrict__ W, float eta,
                                                         float one_minus_eta, int mode_flags, bliss_workspace ws) {
  pdl_wait();
  pdl_trigger();
  const int mode = mode_flags & 1;
  const bool uniform = (mode_flags & BLISS_MODE_UNIFORM);
  const bool bitmap = (mode_flags & BLISS_COLLECT_BITMAP);
  const int lane = lane_id();
  const int n_chunks = ws.ctr->n_chunks;
  const double fx_scale = (double)(1ull << fx_bits_for(ws.ctr->n_seeds));
  for (ChunkLoop q(&ws.ctr->queue[2], n_chunks); q.more(); q.next()) {
    const int c = q.chunk();
    const ChunkRef r = chunk_ref(ws, c);
    const float* __restrict__ wr = W + r.a;
    const int32_t* __restrict__ idx = g.indices + r.a;
    float v[BLISS_CHUNK / 32];
    int src[BLISS_CHUNK / 32];
#pragma unroll
    for (int j = 0; j < BLISS_CHUNK / 32; ++j) {
      const int k = lane + 32 * j;
      src[j] = (k < r.len) ? __ldg(idx + k) : 0;
      v[j] = (k < r.len && !uniform) ? __ldg(wr + k) : 0.0f;
    }
    float row_w = 1.0f, row_q = 1.0f, eta_n = 0.0f;
    if (mode == BLISS_MODE_BANDIT && !uniform) {
      row_w = row_total(ws.part_w, r.c_first, r.c_last);
      row_q = row_total(ws.part_q, r.c_first, r.c_last);
      eta_n = __fdiv_rn(eta, (float)r.d);
      if (lane == 0 && c == r.c_first) ws.row_q[r.row] = row_q;
    }
#pragma unroll
    for (int j = 0; j < BLISS_CHUNK / 32; ++j) {
      const int k = lane + 32 * j;
      if (k < r.len) {
        float t = 0.0f;
        if (uniform) {
        } else if (mode == BLISS_MODE_BANDIT) {
          const float qn = __fdiv_rn(edge_q(v[j], row_w, eta_n, one_minus_eta), row_q);
          t = __fmul_rn(qn, qn);
        } else {
          t = __fmul_rn(v[j], v[j]);
        }
        scatter_term(uniform, bitmap, t, src[j], fx_scale, ws);
      }
    }
  }
}

// Candidate list = seeds (already listed by the plan) ++ every non-seed node that received a
// scatter.  Dense mode: a node is a candidate iff its accumulator is non-zero (one coalesced scan
// of |V| x 8 B).  Bitmap mode: iterate the set bits of cand_bits (|V| / 8 B) and clear them.
// A CTA handles 4 words (128 nodes) per warp and appends its candidates with ONE atomic on the
// shared counter (thousands of same-address atomics would serialise in L2).
#define BLISS_COLLECT_WORDS 4
__global__ void __launch_bounds__(256) k_collect_candidates(int64_t num_nodes, int bitmap, bliss_workspace ws) {
  pdl_wait();
  pdl_trigger();
  __shared__ int s_scan[40];
  __shared__ int s_base;
  const int lane = lane_id();
  const int64_t n_words = (num_nodes + 31) >> 5;
  const int64_t words_per_cta = (int64_t)BLISS_COLLECT_WORDS * (blockDim.x >> 5);
  const int n_seeds = ws.ctr->n_seeds;
  const double fx_inv_scale = 1.0 / (double)(1ull << fx_bits_for(n_seeds));
  // the raw probability p_j = sqrt(acc_j) goes straight into the candidate's slot (the scan holds acc_j
  // anyway), so the scale search starts from a coalesced array instead of |N_c| dependent gathers
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_seeds; i += (int64_t)gridDim.x * blockDim.x)
    ws.p_cand[i] = fx_to_prob(ws.acc[ws.cand[i]], fx_inv_scale);
  for (int64_t w0 = blockIdx.x * words_per_cta; w0 < n_words; w0 += gridDim.x * words_per_cta) {
    unsigned m[BLISS_COLLECT_WORDS];
    unsigned long long av[BLISS_COLLECT_WORDS];
    int cnt = 0;
#pragma unroll
    for (int u = 0; u < BLISS_COLLECT_WORDS; ++u) {
      const int64_t w = w0 + (int64_t)warp_id() * BLISS_COLLECT_WORDS + u;
      const int64_t v = (w << 5) + lane;
      bool hit = false;
      av[u] = 0ull;
      if (w < n_words) {
        if (bitmap) {
          const unsigned word = ws.cand_bits[w];  // warp-uniform
          __syncwarp();
          if (word != 0u && lane == 0) ws.cand_bits[w] = 0u;
          hit = (word >> lane) & 1u;
          if (hit) av[u] = ws.acc[v];
        } else if (v < num_nodes) {
          av[u] = ws.acc[v];
          hit = av[u] != 0ull;
        }
      }
      const bool is_new = hit && ws.node_info[2 * v] < 0;  // seeds are listed already
      m[u] = __ballot_sync(0xffffffffu, is_new);
      cnt += __popc(m[u]);
    }
    int tot;
    int off = block_excl_scan(lane == 0 ? cnt : 0, s_scan, &tot);   // lane 0 of each warp carries the warp's count
    off = __shfl_sync(0xffffffffu, off, 0);
    if (tot == 0) continue;   // CTA-uniform
    if (threadIdx.x == 0) s_base = atomicAdd(&ws.ctr->n_cand, tot);
    __syncthreads();
    int pos = s_base + off;
#pragma unroll
    for (int u = 0; u < BLISS_COLLECT_WORDS; ++u) {
      const int64_t v = ((w0 + (int64_t)warp_id() * BLISS_COLLECT_WORDS + u) << 5) + lane;
      if ((m[u] >> lane) & 1u) {
        const int slot = pos + __popc(m[u] & ((1u << lane) - 1u));
        ws.cand[slot] = (int)v;
        ws.p_cand[slot] = fx_to_prob(av[u], fx_inv_scale);
      }
      pos += __popc(m[u]);
    }
    __syncthreads();   // s_base is rewritten by the next round
  }
}

// ------------------------------------------------------------------------------------------
// (2a) candidate probabilities + Poisson scale search, on the device, no host round trips.
//      One CTA; the candidate array is L2 resident.   bandit_sampler.py:391-401
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_poisson_scale(int n_seeds, int fanout, double eps, int poisson,
                                                       bliss_workspace ws, double fx_inv_scale) {
  n_seeds = ws.ctr->n_seeds;   // the plan's (possibly device-side) count, not the host capacity
  fx_inv_scale = 1.0 / (double)(1ull << fx_bits_for(n_seeds));
  __shared__ unsigned long long s_red[32];
  bliss_counters* ctr = ws.ctr;
  const int n_cand = ctr->n_cand;   // p_cand was filled by k_collect_candidates
  (void)fx_inv_scale;
  if (!poisson) return;
  if (n_cand <= fanout) {  // "prob.shape[0] <= num: return one"
    if (threadIdx.x == 0) {
      ctr->take_all = 1;
      ctr->c = 1.0;
    }
    return;
  }
  __syncthreads();
  double c = 1.0, S = 0.0;
  int it = 0;
  const double s_scale = (double)(1ull << BLISS_S_FIX_BITS);
  const double s_inv = 1.0 / s_scale;
  const double num = (double)fanout;
  for (int i = 0; i < 50; ++i) {
    it = i + 1;
    const float cf = (float)c;
    unsigned long long part = 0;
    for (int j = threadIdx.x; j < n_cand; j += blockDim.x) {
      float v = fminf(__fmul_rn(ws.p_cand[j], cf), 1.0f);
      part += __double2ull_rn((double)v * s_scale);
    }
    unsigned long long tot = block_sum(part, s_red);
    S = __ull2double_rn(tot) * s_inv;
    if (fmin(S, num) / fmax(S, num) >= eps) break;
    c *= num / S;
  }
  if (threadIdx.x == 0) {
    ctr->c = c;
    ctr->iters = it;
    ctr->s_last = S;
  }
}

// ------------------------------------------------------------------------------------------
// (2b) Poisson selection: u < P.   bandit_sampler.py:403-406,422-424
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void push_selected(bool sel, int nid, const bliss_workspace& ws) {
  unsigned m = __ballot_sync(0xffffffffu, sel);
  if (m) {
    int leader = __ffs(m) - 1;
    int base = 0;
    if (lane_id() == leader) base = atomicAdd(&ws.ctr->n_sel, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (sel) {
      int slot = base + __popc(m & ((1u << lane_id()) - 1u));
      if (slot < ws.cap_sel) {
        ws.sel[slot] = nid;
        atomicOr(&ws.sel_bits[nid >> 5], 1u << (nid & 31));
      } else {
        ws.ctr->error = BLISS_ERR_SEL_CAPACITY;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// (2a) Poisson scale search in ONE launch of a single thread-block cluster (16 CTAs x 1024 threads,
// 8 if 16 is not schedulable); the selection itself is the grid-wide k_select_poisson.  Every
// thread keeps up to NREG candidate probabilities in registers, so one search iteration is
// NREG multiplies per thread + a CTA reduction + a DSMEM exchange of one 64-bit partial per CTA
// and one cluster barrier — no global memory traffic and no host round trip inside the search
// (the reference does up to 50 .item() syncs per layer, bandit_sampler.py:396-401).
// Candidates beyond NREG per thread are re-read from p_cand (L2) each iteration.
// ------------------------------------------------------------------------------------------
#define BLISS_SCALE_NREG 16
#define BLISS_SCALE_MAX_CLUSTER 16
__global__ void __launch_bounds__(1024, 1) k_scale_search(int fanout, double eps, bliss_workspace ws) {
  pdl_wait();
  pdl_trigger();
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const unsigned nblk = cluster.num_blocks();
  __shared__ unsigned long long s_red[32];
  __shared__ unsigned long long s_parts[2][BLISS_SCALE_MAX_CLUSTER];
  bliss_counters* ctr = ws.ctr;
  const int n_cand = ctr->n_cand;
  const int gtid = rank * blockDim.x + threadIdx.x;
  const int nthreads = nblk * blockDim.x;

  float preg[BLISS_SCALE_NREG];   // p_cand was filled by k_collect_candidates: coalesced loads
#pragma unroll
  for (int r = 0; r < BLISS_SCALE_NREG; ++r) {
    const int j = gtid + r * nthreads;
    preg[r] = (j < n_cand) ? ws.p_cand[j] : 0.0f;
  }

  const bool take_all = n_cand <= fanout;  // "prob.shape[0] <= num: return one"  (:392-393)
  double c = 1.0, S = 0.0;
  int it = 0;
  if (!take_all) {
    const float s_scale = 1099511627776.0f;  // 2^40: v * 2^40 is exact in fp32, so is its integer rounding
    const double s_inv = 1.0 / 1099511627776.0;
    const double num = (double)fanout;
    for (int i = 0; i < 50; ++i) {
      it = i + 1;
      const float cf = (float)c;
      unsigned long long part = 0;
#pragma unroll
      for (int r = 0; r < BLISS_SCALE_NREG; ++r)
        part += __float2ull_rn(__fmul_rn(fminf(__fmul_rn(preg[r], cf), 1.0f), s_scale));
      for (int j = gtid + BLISS_SCALE_NREG * nthreads; j < n_cand; j += nthreads)
        part += __float2ull_rn(__fmul_rn(fminf(__fmul_rn(ws.p_cand[j], cf), 1.0f), s_scale));
      // warp totals -> warp 0 adds them and publishes this CTA's partial into every CTA of the cluster
      part = warp_sum(part);
      if (lane_id() == 0) s_red[warp_id()] = part;
      __syncthreads();
      if (warp_id() == 0) {
        const unsigned long long tot = warp_sum(s_red[lane_id()]);
        if (lane_id() < nblk) *cluster.map_shared_rank(&s_parts[i & 1][rank], lane_id()) = tot;
      }
      cluster.sync();
      unsigned long long all = 0;
      for (unsigned b = 0; b < nblk; ++b) all += s_parts[i & 1][b];
      S = __ull2double_rn(all) * s_inv;
      if (fmin(S, num) / fmax(S, num) >= eps) break;   // :398
      c *= num / S;                                    // :401
    }
  }
  if (gtid == 0) {
    ctr->c = c;
    ctr->iters = it;
    ctr->s_last = S;
    ctr->take_all = take_all ? 1 : 0;
  }
  cluster.sync();  // no CTA may exit while its shared memory can still be written remotely
}

// Grid-wide: 4 candidates per thread and round; the selected ones of a CTA round are appended with ONE
// atomic on the shared counter (flags -> CTA scan -> base).  Only selected nodes get their
// (selected mark, P) pair written: nothing reads it for the others.
__global__ void __launch_bounds__(256) k_select_poisson(int n_seeds, unsigned long long seed,
                                                       unsigned long long step, unsigned layer,
                                                       const float* __restrict__ u_inject,
                                                       bliss_workspace ws) {
  pdl_wait();
  pdl_trigger();
  __shared__ int s_scan[40];
  __shared__ int s_base;
  n_seeds = ws.ctr->n_seeds;   // the plan's (possibly device-side) count, not the host capacity
  if (ws.step_dev) step = *ws.step_dev;   // Philox step counter kept on the device (CUDA-graph replay)
  bliss_counters* ctr = ws.ctr;
  const int n_cand = ctr->n_cand;
  const float cf = (float)ctr->c;
  const int take_all = ctr->take_all;
  constexpr int PER = 4;
  const int per_round = blockDim.x * PER;
  for (int j0 = n_seeds + blockIdx.x * per_round; j0 < n_cand; j0 += gridDim.x * per_round) {
    int nid[PER];
    float P[PER];
    unsigned mask = 0;
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int j = j0 + u * blockDim.x + threadIdx.x;
      nid[u] = 0;
      P[u] = 0.0f;
      if (j < n_cand) {
        nid[u] = ws.cand[j];
        P[u] = take_all ? 1.0f : fminf(__fmul_rn(ws.p_cand[j], cf), 1.0f);
      }
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int j = j0 + u * blockDim.x + threadIdx.x;
      if (j < n_cand) {
        const float uu = u_inject ? u_inject[nid[u]] : philox_uniform(seed, step, layer, (unsigned)nid[u]);
        if (uu < P[u]) mask |= 1u << u;   // u < P  (:422-424)
      }
    }
    int tot;
    int off = block_excl_scan(__popc(mask), s_scan, &tot);
    if (tot == 0) continue;   // CTA-uniform
    if (threadIdx.x == 0) s_base = atomicAdd(&ctr->n_sel, tot);
    __syncthreads();
    int slot = s_base + off;
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      if ((mask >> u) & 1u) {
        if (slot < ws.cap_sel) {
          ws.sel[slot] = nid[u];
          *reinterpret_cast<int2*>(&ws.node_info[2 * nid[u]]) = make_int2(-2, __float_as_int(P[u]));
          atomicOr(&ws.sel_bits[nid[u] >> 5], 1u << (nid[u] & 31));
        } else {
          ctr->error = BLISS_ERR_SEL_CAPACITY;
        }
        ++slot;
      }
    }
    __syncthreads();   // s_base is rewritten by the next round
  }
}

// ------------------------------------------------------------------------------------------
// (2c) multinomial-without-replacement selection = the k largest p / Exp(1).
//      bandit_sampler.py:84-99, ladies_sampler.py:54-69.  Radix select over the key bits.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_topk_keys(int n_seeds, unsigned long long seed, unsigned long long step,
                                                  unsigned layer, const float* __restrict__ u_inject,
                                                  float* __restrict__ keys, bliss_workspace ws) {
  n_seeds = ws.ctr->n_seeds;   // the plan's (possibly device-side) count, not the host capacity
  if (ws.step_dev) step = *ws.step_dev;
  const int n_cand = ws.ctr->n_cand;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_cand; j += gridDim.x * blockDim.x) {
    int nid = ws.cand[j];
    float u = u_inject ? u_inject[nid] : philox_uniform(seed, step, layer, (unsigned)nid);
    float e = -log1pf(-u);
    keys[j] = __fdiv_rn(ws.p_cand[j], e);
    // every candidate starts unselected; seeds keep their local id (they stay block nodes)
    ws.node_info[2 * nid + 1] = __float_as_int(ws.p_cand[j]);
    if (j >= n_seeds) ws.node_info[2 * nid] = -1;
  }
}

// One CTA: find the bit pattern T of the k-th largest key (keys are >= 0, so their float bits
// order like unsigned ints) and how many keys equal to T are still needed.
__global__ void __launch_bounds__(1024) k_topk_threshold(int fanout, const float* __restrict__ keys,
                                                        unsigned* __restrict__ thr_out, bliss_workspace ws) {
  __shared__ int hist[256];
  __shared__ unsigned s_prefix;
  __shared__ int s_need;
  const int n_cand = ws.ctr->n_cand;
  int k = min(fanout, n_cand);
  if (threadIdx.x == 0) {
    s_prefix = 0;
    s_need = k;
  }
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    const unsigned mask_hi = (pass == 0) ? 0u : (0xffffffffu << (shift + 8));
    for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    const unsigned prefix = s_prefix;
    for (int j = threadIdx.x; j < n_cand; j += blockDim.x) {
      unsigned bits = __float_as_uint(keys[j]);
      if ((bits & mask_hi) == prefix) atomicAdd(&hist[(bits >> shift) & 255u], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int need = s_need, b = 255;
      for (; b > 0; --b) {
        if (hist[b] >= need) break;
        need -= hist[b];
      }
      s_prefix = prefix | ((unsigned)b << shift);
      s_need = need;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    thr_out[0] = s_prefix;        // threshold bits
    thr_out[1] = (unsigned)s_need;  // how many keys == threshold to take (lowest slots first)
    thr_out[2] = 0;               // tie cursor
  }
}

// keys > T are selected; among keys == T the `need` lowest candidate slots (deterministic).
__global__ void __launch_bounds__(1024) k_topk_mark(int n_seeds, const float* __restrict__ keys,
                                                   unsigned* __restrict__ thr, bliss_workspace ws) {
  n_seeds = ws.ctr->n_seeds;   // the plan's (possibly device-side) count, not the host capacity
  // single CTA so ties are taken in slot order
  __shared__ int s_scan[40];
  const int n_cand = ws.ctr->n_cand;
  const unsigned T = thr[0];
  const int need = (int)thr[1];
  int tie_base = 0;
  const int n_pad = (n_cand + blockDim.x - 1) / blockDim.x * blockDim.x;
  for (int j = threadIdx.x; j < n_pad; j += blockDim.x) {
    bool valid = j < n_cand;
    unsigned bits = valid ? __float_as_uint(keys[j]) : 0u;
    int is_tie = valid && bits == T;
    int tot;
    int rank = block_excl_scan(is_tie, s_scan, &tot);
    bool sel = valid && (bits > T || (is_tie && tie_base + rank < need));
    tie_base += tot;
    int nid = valid ? ws.cand[j] : 0;
    if (valid && j < n_seeds) {
      // seeds stay block nodes; an unselected seed loses its out-edges (bandit_sampler.py:295-298)
      if (!sel) atomicAnd(&ws.sel_bits[nid >> 5], ~(1u << (nid & 31)));
      sel = false;
    } else if (sel) {
      ws.node_info[2 * nid] = -2;
    }
    push_selected(sel, nid, ws);
  }
}

__global__ void k_philox_fill(unsigned long long seed, unsigned long long step, unsigned layer,
                              const int32_t* __restrict__ nids, int64_t n, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = philox_uniform(seed, step, layer, (unsigned)nids[i]);
}

// ------------------------------------------------------------------------------------------
// (2d) uniform neighbour sampling (`--sampler neighbor | full`, train_lightning.py:349-357: DGL NeighborSampler /
//      MultiLayerFullNeighborSampler): every seed keeps min(fanout, degree) of its in-edges, chosen uniformly
//      without replacement; fanout <= 0 keeps all.  One warp per row.  The draw: every in-edge gets the key
//      Philox4x32-10(seed; CSC position, layer | 0x8000, step) (32 bits); the row keeps its `fanout` smallest keys,
//      ties in CSC order — the k smallest of d i.i.d. keys are a uniform k-subset.  The threshold is found by a
//      4-pass radix select over the row's keys (recomputed per pass: nothing edge-sized is stored), then one pass
//      writes the row's keep bits (the layout k_block_count produces) and registers every kept edge's source as a
//      block node (selected bit, list of selected nodes, candidate list for the workspace restore).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned edge_key(unsigned long long seed, unsigned long long step, unsigned layer, int64_t pos) {
  const uint4 r = philox4x32_10(make_uint4((unsigned)pos, layer | 0x8000u | ((unsigned)(pos >> 32) << 16), (unsigned)step,
                                           (unsigned)(step >> 32)),
                                make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
  return r.x;
}
__global__ void __launch_bounds__(256) k_neighbor_select(GraphView g, int fanout, unsigned long long seed,
                                                        unsigned long long step, unsigned layer, bliss_workspace ws) {
  pdl_wait();
  pdl_trigger();
  if (ws.step_dev) step = *ws.step_dev;
  bliss_counters* ctr = ws.ctr;
  const int n_seeds = ctr->n_seeds;
  const int lane = lane_id();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < n_seeds; i += nwarps) {
    const int64_t a = ws.row_a[i];
    const int d = ws.row_d[i];
    const int c_first = ws.chunk_first[i];
    unsigned T = 0xffffffffu;     // keep keys < T, and `need` of the keys == T (first in CSC order)
    int need = 0x7fffffff;
    if (fanout > 0 && d > fanout) {
      unsigned prefix = 0;
      int k = fanout;             // rank of the threshold among the keys that match the prefix so far
      for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        const unsigned mask_hi = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        // 256-bin histogram of the next byte over the keys matching the prefix: 8 bins per lane
        int hist[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) hist[b] = 0;
        for (int j0 = 0; j0 < d; j0 += 32) {
          const int j = j0 + lane;
          unsigned key = 0;
          bool in = false;
          if (j < d) {
            key = edge_key(seed, step, layer, a + j);
            in = (key & mask_hi) == prefix;
          }
          const unsigned byte = (key >> shift) & 255u;
          // every lane owns bins [8 lane, 8 lane + 8): count the matching keys of this round by ballots over the byte's owner
          for (unsigned m = __ballot_sync(0xffffffffu, in); m;) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const unsigned bsrc = __shfl_sync(0xffffffffu, byte, src);
            if ((int)(bsrc >> 3) == lane) ++hist[bsrc & 7];
          }
        }
        // smallest byte value whose cumulative count reaches k
        int mine = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) mine += hist[b];
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += up;
        }
        const unsigned reach = __ballot_sync(0xffffffffu, incl >= k);
        const int owner = __ffs(reach) - 1;            // lane whose bins contain the k-th key
        int below = __shfl_sync(0xffffffffu, incl - mine, owner), bin = 0;
        if (lane == owner) {
          int run = below;
          for (int b = 0; b < 8; ++b) {
            if (run + hist[b] >= k) { bin = b; below = run; break; }
            run += hist[b];
          }
        }
        bin = __shfl_sync(0xffffffffu, bin, owner);
        below = __shfl_sync(0xffffffffu, below, owner);
        prefix |= (unsigned)(owner * 8 + bin) << shift;
        k -= below;
      }
      T = prefix;
      need = k;                    // keys equal to T still to take
    }
    int ties = 0;
    for (int j0 = 0; j0 < d; j0 += 32) {
      const int j = j0 + lane;
      bool keep = false, tie = false;
      int src = 0;
      if (j < d) {
        src = __ldg(g.indices + a + j);
        if (fanout <= 0 || d <= fanout) {
          keep = true;
        } else {
          const unsigned key = edge_key(seed, step, layer, a + j);
          keep = key < T;
          tie = key == T;
        }
      }
      const unsigned tm = __ballot_sync(0xffffffffu, tie);
      if (tie && ties + __popc(tm & ((1u << lane) - 1u)) < need) keep = true;
      ties += __popc(tm);
      const unsigned km = __ballot_sync(0xffffffffu, keep);
      // keep bits: chunk c_first + j0 / 256, word (j0 % 256) / 32
      if (lane == 0) ws.keep_bits[(int64_t)(c_first + (j0 >> 8)) * (BLISS_CHUNK / 32) + ((j0 & 255) >> 5)] = km;
      bool fresh = false;
      if (keep) {
        const unsigned bit = 1u << (src & 31);
        const unsigned old = atomicOr(&ws.sel_bits[src >> 5], bit);
        fresh = !(old & bit);      // seeds were marked by the plan, so a fresh node is a non-seed source
      }
      const unsigned fm = __ballot_sync(0xffffffffu, fresh);
      if (fm) {
        int base = 0;
        if (lane == __ffs(fm) - 1) {
          base = atomicAdd(&ctr->n_sel, __popc(fm));
          atomicAdd(&ctr->n_cand, __popc(fm));
        }
        base = __shfl_sync(0xffffffffu, base, __ffs(fm) - 1);
        if (fresh) {
          const int slot = base + __popc(fm & ((1u << lane) - 1u));
          if (slot < ws.cap_sel) {
            ws.sel[slot] = src;
            ws.cand[n_seeds + slot] = src;
            *reinterpret_cast<int2*>(&ws.node_info[2 * src]) = make_int2(-2, __float_as_int(1.0f));
          } else {
            ctr->error = BLISS_ERR_SEL_CAPACITY;
          }
        }
      }
    }
    // the tail words of the row's last chunk stay zero from the previous use?  No: clear them explicitly
    const int last_word = (d + 31) >> 5;
    const int n_words_row = (ws.chunk_first[i + 1] - c_first) * (BLISS_CHUNK / 32);
    for (int wd = last_word + lane; wd < n_words_row; wd += 32)
      ws.keep_bits[(int64_t)c_first * (BLISS_CHUNK / 32) + wd] = 0u;
  }
}

// ------------------------------------------------------------------------------------------
// (3a) kept in-edges (source selected) of every 256-edge chunk of the frontier: keep bits, kept
//      count and — bandit mode — the chunk's partial of ΣW~ (W~ = q_ij / P_src); first occurrence
//      of every selected non-seed source: key = (row+1, position in row).  bandit_sampler.py:289-314
// ------------------------------------------------------------------------------------------
struct FillCtx {
  GraphView g;
  const float* __restrict__ W;
  float eta, one_minus_eta;
  int mode;
};

// Kept edges are sparse in a chunk (~8 % of the bits at Reddit fan-outs), so the dependent random
// accesses (node_info -> first_pos -> atomicMin, weight -> q -> W~) are not done inside the word
// loop, where only 2-3 lanes per warp would be active: the kept positions go to a per-warp
// shared-memory list and are processed densely afterwards, one kept edge per lane.
// SMEM_BITS: the selected-node bitmap (|V| / 8 bytes) is copied into shared memory once per CTA — the
// per-edge bit test is a 32-way gather, which an SM's L1 serves at one sector per cycle (it alone
// would cost more than streaming the indices) but shared memory serves at bank speed.
template <bool SMEM_BITS, bool EDGE_BITS = false>   // EDGE_BITS: the keep bits were chosen per edge (k_neighbor_select)
__global__ void __launch_bounds__(BLISS_CTA, 6) k_block_count(FillCtx c, bliss_workspace ws, int bit_words) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ uint32_t s_bits[];
  __shared__ unsigned char s_k[BLISS_WARPS][BLISS_CHUNK];   // chunk-relative positions of the kept edges
  bliss_counters* ctr = ws.ctr;
  const int n_chunks = ctr->n_chunks;
  const int lane = lane_id();
  const bool bandit = (c.mode == BLISS_MODE_BANDIT);
  unsigned char* lk = s_k[warp_id()];
  if (blockIdx.x * BLISS_WARPS >= n_chunks) return;   // CTA-uniform: no first chunk for any warp
  if (SMEM_BITS) {
    const int4* __restrict__ src4 = reinterpret_cast<const int4*>(ws.sel_bits);
    for (int i = threadIdx.x; i < (bit_words >> 2); i += BLISS_CTA) reinterpret_cast<int4*>(s_bits)[i] = __ldg(src4 + i);
    for (int i = (bit_words & ~3) + threadIdx.x; i < bit_words; i += BLISS_CTA) s_bits[i] = ws.sel_bits[i];
    __syncthreads();
  }
  for (ChunkLoop q(&ctr->queue[3], n_chunks); q.more(); q.next()) {
    const int ch = q.chunk();
    const ChunkRef r = chunk_ref(ws, ch);
    const int32_t* __restrict__ idx = c.g.indices + r.a;
    int src[BLISS_CHUNK / 32];
#pragma unroll
    for (int j = 0; j < BLISS_CHUNK / 32; ++j) {
      const int k = lane + 32 * j;
      src[j] = (k < r.len) ? __ldg(idx + k) : -1;
    }
    int cnt = 0;
    unsigned myword = 0;
    const unsigned pre = (EDGE_BITS && lane < BLISS_CHUNK / 32)
                             ? ws.keep_bits[(int64_t)ch * (BLISS_CHUNK / 32) + lane] : 0u;
#pragma unroll
    for (int j = 0; j < BLISS_CHUNK / 32; ++j) {
      bool keep = false;
      const unsigned pre_j = EDGE_BITS ? __shfl_sync(0xffffffffu, pre, j) : 0u;   // (every lane takes part in the shuffle)
      if (EDGE_BITS)
        keep = src[j] >= 0 && ((pre_j >> lane) & 1u);
      else if (src[j] >= 0)
        keep = SMEM_BITS ? ((s_bits[src[j] >> 5] >> (src[j] & 31)) & 1u) : test_bit(ws.sel_bits, src[j]);
      const unsigned bits = __ballot_sync(0xffffffffu, keep);
      if (keep) lk[cnt + __popc(bits & ((1u << lane) - 1u))] = (unsigned char)(lane + 32 * j);
      if (lane == j) myword = bits;
      cnt += __popc(bits);
    }
    if (lane < BLISS_CHUNK / 32) ws.keep_bits[(int64_t)ch * (BLISS_CHUNK / 32) + lane] = myword;
    __syncwarp();
    double t = 0.0;
    if (cnt) {
      const unsigned long long key_hi = (unsigned long long)(r.row + 1) << 32;
      const float row_w = (bandit && r.d > 0) ? __ldg(ws.row_w + r.row) : 1.0f;
      const float eta_n = __fdiv_rn(c.eta, (float)r.d);
      for (int j = lane; j < cnt; j += 32) {
        const int k = lk[j];
        const int sj = __ldg(idx + k);                       // L1 hit: this warp streamed it a moment ago
        // independent loads, one round trip: (selected mark, P), current first-occurrence key, weight
        const int2 info = __ldg(reinterpret_cast<const int2*>(ws.node_info) + sj);
        const unsigned long long cur = __ldcg((const unsigned long long*)&ws.first_pos[sj]);
        const float wv = bandit ? __ldg(c.W + r.a + k) : 0.0f;
        const unsigned long long key = key_hi | (unsigned)(r.k0 + k);
        if (info.x < 0 && key < cur) atomicMin((unsigned long long*)&ws.first_pos[sj], key);   // selected non-seed
        if (bandit)   // also with importance_sampling=0: q_ij is still built from W (:354-358)
          t += (double)__fdiv_rn(edge_q(wv, row_w, eta_n, c.one_minus_eta), __int_as_float(info.y));
      }
    }
    if (bandit) t = warp_sum(t);
    if (lane == 0) {
      ws.part_cnt[ch] = cnt;
      if (bandit) ws.part_t[ch] = t;
      if (cnt) atomicAdd(&ws.row_cnt[r.row], cnt);
    }
    __syncwarp();   // the list is rewritten by the warp's next chunk
  }
}

// ------------------------------------------------------------------------------------------
// (3b) block indptr + SpMM segment prefix (CTA 0, one scan round per 4096 rows) while all CTAs
//      copy the seeds' ids / probabilities, write the capacity padding and rank the selected
//      sources by first occurrence:
//      src order = [seeds in seed order] ++ [selected non-seeds by first occurrence]
//      (compact_graphs / to_block ordering).  Rank by counting: one warp per key,
//      the keys of all selected sources pass through shared memory in tiles.
// ------------------------------------------------------------------------------------------
#define BLISS_RANK_TILE 2048
__global__ void __launch_bounds__(1024) k_block_index(const int32_t* __restrict__ seeds, int n_seeds,
                                                     bliss_workspace ws, bliss_block_out out) {
  pdl_wait();
  pdl_trigger();
  n_seeds = ws.ctr->n_seeds;   // the plan's (possibly device-side) count, not the host capacity
  __shared__ unsigned long long s_keys[BLISS_RANK_TILE];
  __shared__ int s_scan[40];
  bliss_counters* ctr = ws.ctr;
  const int n_sel = min(ctr->n_sel, (int)ws.cap_sel);
  const int lane = lane_id();
  if (blockIdx.x == 0) {
    // An overflowing block (more kept edges than the pool holds) must stay memory-safe for the kernels of a replayed
    // step that run before the host sees the error flag: row extents are clamped to the capacity (the fill kernel
    // skips the block, so the rows then point at older, in-range edges) and the counters report clamped sizes.
    const int ecap = (out.cap_edges > 0) ? (int)min((int64_t)0x7fffffff, out.cap_edges) : 0x7fffffff;
    int base = 0, sbase = 0;
    for (int b = 0; b < n_seeds; b += blockDim.x * 4) {
      const int i0 = b + threadIdx.x * 4;
      int v[4], len[4], sg[4], pre[4], spre[4];
      load4(ws.row_cnt, i0, n_seeds, v);
      int vsum = 0, ssum = 0;
#pragma unroll
      for (int u = 0; u < 4; ++u) vsum += v[u];
      int tot, st = 0;
      int p = base + block_excl_scan(vsum, s_scan, &tot);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        pre[u] = min(p, ecap);
        len[u] = min(p + v[u], ecap) - pre[u];   // == v[u] unless the block overflows
        sg[u] = (i0 + u < n_seeds) ? max(1, (len[u] + BLISS_SPMM_SEG - 1) / BLISS_SPMM_SEG) : 0;
        ssum += sg[u];
        p += v[u];
      }
      int sp = sbase + (out.seg_ptr ? block_excl_scan(ssum, s_scan, &st) : 0);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        spre[u] = sp;
        sp += sg[u];
      }
      store4(out.indptr, i0, n_seeds, pre);
      if (out.seg_ptr) store4(out.seg_ptr, i0, n_seeds, spre);
      if (out.inv_deg) {
#pragma unroll
        for (int u = 0; u < 4; ++u) pre[u] = __float_as_int(1.0f / (float)max(len[u], 1));   // fn.mean divisor
        store4(reinterpret_cast<int32_t*>(out.inv_deg), i0, n_seeds, pre);
      }
      base += tot;
      sbase += st;
    }
    const int base_c = min(base, ecap);
    // capacity padding (static-shape replay): rows beyond n_seeds are empty (one empty segment each)
    for (int64_t i = n_seeds + threadIdx.x; i <= max((int64_t)n_seeds, out.pad_rows); i += blockDim.x) {
      out.indptr[i] = base_c;
      if (out.seg_ptr) out.seg_ptr[i] = sbase + (int)(i - n_seeds);
    }
    if (threadIdx.x == 0) {
      ctr->n_edges = base_c;
      ctr->n_src = (int)min((int64_t)n_seeds + n_sel, out.cap_src > 0 ? out.cap_src : (int64_t)0x7fffffff);
      if (n_seeds + n_sel > out.cap_src) ctr->error |= BLISS_ERR_SEL_CAPACITY;
      if (base > out.cap_edges && out.cap_edges > 0) ctr->error |= BLISS_ERR_EDGE_CAPACITY;
    }
  }
  // CTA 0 only scans (it is the longest single piece); the other CTAs share everything else
  if (blockIdx.x == 0 && gridDim.x > 1) return;
  const int wb = (gridDim.x > 1) ? blockIdx.x - 1 : 0, wn = (gridDim.x > 1) ? gridDim.x - 1 : 1;   // worker CTA index / count
  {  // the seeds' rows of the source arrays and the padding that does not depend on the scan
    const int64_t gt = wb * (int64_t)blockDim.x + threadIdx.x, gn = (int64_t)wn * blockDim.x;
    for (int64_t i = gt; i < n_seeds; i += gn) {
      const int s = seeds[i];
      out.src_nid[i] = s;
      out.node_prob[i] = __int_as_float(ws.node_info[2 * s + 1]);
      if (out.out_deg) out.out_deg[i] = 0;
    }
    if (out.inv_deg)   // the mean divisor of a padded row is 1
      for (int64_t i = n_seeds + gt; i < out.pad_rows; i += gn) out.inv_deg[i] = 1.0f;
    if (out.out_deg)   // padded sources have no edges (their counts feed the transpose scan)
      for (int64_t i = n_seeds + n_sel + gt; i < out.pad_src; i += gn) out.out_deg[i] = 0;
    // per row (one warp each): the kept-edge prefix of its chunks and its ΣW~, so the fill starts every
    // chunk from two loads instead of re-adding the row's partials chunk by chunk
    for (int i = (int)(gt >> 5); i < n_seeds; i += (int)(gn >> 5)) {
      const int c_first = ws.chunk_first[i], c_last = ws.chunk_first[i + 1];
      int run = 0;
      double t = 0.0;
      for (int c0 = c_first; c0 < c_last; c0 += 32) {
        const int cix = c0 + lane;
        const int n = (cix < c_last) ? __ldg(ws.part_cnt + cix) : 0;
        if (cix < c_last) t += __ldg(ws.part_t + cix);   // not written by the LADIES count: unused there
        int incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += up;
        }
        if (cix < c_last) ws.chunk_pre[cix] = run + incl - n;
        run += __shfl_sync(0xffffffffu, incl, 31);
      }
      t = warp_sum(t);
      if (lane == 0) ws.row_t[i] = __double2float_rn(t);
    }
    if (out.t_bits) {  // source x destination bitmap of the block (transpose): clear this block's source rows
      // (whole rows: an earlier, larger block may have left bits beyond this block's destinations)
      const int64_t rows = min((int64_t)n_seeds + n_sel, out.pad_src > 0 ? out.pad_src : out.cap_src);
      for (int64_t i = gt; i < rows * out.t_words; i += gn) out.t_bits[i] = 0u;
    }
  }
  // ranking: one warp per key, 32 keys per CTA round; every lane counts 1/32 of each tile
  constexpr int PER_T = BLISS_RANK_TILE / 1024;
  const int n_groups = (n_sel + 31) / 32;
  for (int grp = wb; grp < n_groups; grp += wn) {
    const int j = grp * 32 + warp_id();
    const bool valid = j < n_sel;
    const int nid = valid ? ws.sel[j] : 0;
    const unsigned long long key = valid ? ws.first_pos[nid] : 0ull;
    int rank = 0;
    for (int t0 = 0; t0 < n_sel; t0 += BLISS_RANK_TILE) {
      const int tn = min(BLISS_RANK_TILE, n_sel - t0);
      __syncthreads();
      // the tile's keys are two dependent gathers (sel -> first_pos): both per thread in flight together
      int tn_id[PER_T];
      unsigned long long tk[PER_T];
#pragma unroll
      for (int u = 0; u < PER_T; ++u) {
        const int t = threadIdx.x + u * 1024;
        tn_id[u] = (t < tn) ? ws.sel[t0 + t] : -1;
      }
#pragma unroll
      for (int u = 0; u < PER_T; ++u) tk[u] = (tn_id[u] >= 0) ? ws.first_pos[tn_id[u]] : ~0ull;
#pragma unroll
      for (int u = 0; u < PER_T; ++u) s_keys[threadIdx.x + u * 1024] = tk[u];
      __syncthreads();
      if (valid) {
#pragma unroll 8
        for (int t = lane; t < tn; t += 32) rank += (s_keys[t] < key);
      }
    }
    rank = warp_sum(rank);
    if (valid && lane == 0) {
      const int local = n_seeds + rank;
      ws.node_info[2 * nid] = local;
      if (local < out.cap_src) {
        out.src_nid[local] = nid;
        out.node_prob[local] = __int_as_float(ws.node_info[2 * nid + 1]);
      }
      if (out.out_deg && (out.pad_src == 0 || local < out.pad_src)) out.out_deg[local] = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------
// (3c) fill: ordered compaction of the kept in-edges into the block CSR, chunk by chunk — a
//      chunk's first slot is the row's start plus the kept counts of the row's earlier chunks —
//      with relabelled sources, q_ij and the final block weight
//      W~ = (q_ij / P_src) * d / ΣW~  (bandit_sampler.py:306-320)  or  (w / P_src) * d  (ladies_sampler.py:97).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLISS_CTA, 6) k_block_fill(FillCtx c, bliss_workspace ws, bliss_block_out out) {
  __shared__ unsigned char s_k[BLISS_WARPS][BLISS_CHUNK];
  bliss_counters* ctr = ws.ctr;
  const int n_chunks = ctr->n_chunks;
  const int lane = lane_id();
  const bool bandit = (c.mode == BLISS_MODE_BANDIT);
  if (ctr->n_edges > out.cap_edges || ctr->error) return;  // capacity error already flagged by the index kernel
  unsigned char* lk = s_k[warp_id()];
  for (ChunkLoop q(&ctr->queue[4], n_chunks); q.more(); q.next()) {
    const int ch = q.chunk();
    // round trip 1: everything that depends on the chunk id only
    const int cnt = __ldg(ws.part_cnt + ch);
    const ChunkRef r = chunk_ref(ws, ch);
    const unsigned myword = (lane < BLISS_CHUNK / 32) ? __ldg(ws.keep_bits + (int64_t)ch * (BLISS_CHUNK / 32) + lane) : 0u;
    if (cnt == 0) continue;   // warp-uniform
    const int pre = __ldg(ws.chunk_pre + ch);   // kept edges of the row before this chunk (k_block_index)
    // round trip 2: the row's block offset, kept count and ΣW~, and this lane's first kept edge
    // (index, weight, edge id) — all independent, issued together
    const int row_base = __ldg(out.indptr + r.row);
    const int tot = __ldg(ws.row_cnt + r.row);
    const float row_t = bandit ? __ldg(ws.row_t + r.row) : 1.0f;
    const float row_w = (bandit && r.d > 0) ? __ldg(ws.row_w + r.row) : 1.0f;
    int n = 0;
#pragma unroll
    for (int j = 0; j < BLISS_CHUNK / 32; ++j) {
      const unsigned bits = __shfl_sync(0xffffffffu, myword, j);
      if ((bits >> lane) & 1u) lk[n + __popc(bits & ((1u << lane) - 1u))] = (unsigned char)(lane + 32 * j);
      n += __popc(bits);
    }
    __syncwarp();
    int64_t p = r.a + ((lane < cnt) ? lk[lane] : 0);
    int src = __ldg(c.g.indices + p);
    float wv = __ldg(c.W + p);
    int eidv = (out.eid && c.g.eid) ? __ldg(c.g.eid + p) : (int32_t)p;
    // round trip 3: (local id, P) of the source
    int2 info = __ldg(reinterpret_cast<const int2*>(ws.node_info) + src);
    float f = (float)tot;                                   // ladies: W~ *= d
    if (bandit) f = __fdiv_rn(f, row_t);                    // bandit: W~ *= d / ΣW~
    const float eta_n = __fdiv_rn(c.eta, (float)r.d);
    const int base = row_base + pre;
    for (int j = lane; j < cnt; j += 32) {
      if (j >= 32) {   // chunks that keep more than 32 edges: the remaining ones, one round trip chain each
        p = r.a + lk[j];
        src = __ldg(c.g.indices + p);
        wv = __ldg(c.W + p);
        eidv = (out.eid && c.g.eid) ? __ldg(c.g.eid + p) : (int32_t)p;
        info = __ldg(reinterpret_cast<const int2*>(ws.node_info) + src);
      }
      const float qv = bandit ? edge_q(wv, row_w, eta_n, c.one_minus_eta) : wv;
      const float wt = __fmul_rn(__fdiv_rn(qv, __int_as_float(info.y)), f);
      const int slot = base + j;
      out.edge_src[slot] = info.x;
      out.edge_dst[slot] = r.row;
      out.csc_pos[slot] = p;
      if (out.eid) out.eid[slot] = eidv;
      if (out.q_ij) out.q_ij[slot] = qv;
      out.edge_w[slot] = wt;
      if (out.out_deg) atomicAdd(&out.out_deg[info.x], 1);
      if (out.t_bits) atomicOr(&out.t_bits[(int64_t)info.x * out.t_words + (r.row >> 5)], 1u << (r.row & 31));
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// (3d) finish: restore the workspace invariant for every node this layer touched (the block
//      weights were already normalised row by row in k_block_fill).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_block_finish(int mode, bliss_workspace ws, bliss_block_out out) {
  bliss_counters* ctr = ws.ctr;
  const int n_cand = ctr->n_cand;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  (void)mode;  // the rows are normalised by k_block_fill itself; this kernel only restores the workspace
  if (ws.ctr_mirror && t0 < (int64_t)(sizeof(bliss_counters) / sizeof(int32_t)))   // counters are final: host copy
    reinterpret_cast<volatile int32_t*>(ws.ctr_mirror)[t0] = reinterpret_cast<const int32_t*>(ctr)[t0];
  for (int64_t j = t0; j < n_cand; j += stride) {
    int nid = ws.cand[j];
    ws.acc[nid] = 0ull;
    ws.first_pos[nid] = ~0ull;
    ws.node_info[2 * nid] = -1;
    ws.node_info[2 * nid + 1] = 0;
    ws.sel_bits[nid >> 5] = 0u;
  }
}

__global__ void k_ws_init(bliss_workspace ws, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += stride) {
    ws.acc[v] = 0ull;
    ws.first_pos[v] = ~0ull;
    ws.node_info[2 * v] = -1;
    ws.node_info[2 * v + 1] = 0;
    if ((v & 31) == 0) {
      ws.sel_bits[v >> 5] = 0u;
      ws.cand_bits[v >> 5] = 0u;
    }
  }
}

// ------------------------------------------------------------------------------------------
// source-major transpose of a block for the backward aggregation
// ------------------------------------------------------------------------------------------
__global__ void k_t_count(const int32_t* __restrict__ edge_src, int64_t n_edges, int32_t* __restrict__ cnt) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_edges; e += stride)
    atomicAdd(&cnt[edge_src[e]], 1);
}
// cnt_cursor holds the per-source counts on entry and the fill cursors (= row starts) on exit;
// seg (may be NULL) receives the 32-edge segment prefix of the source rows (balanced SpMM).
__global__ void __launch_bounds__(1024) k_t_scan(int32_t* cnt_cursor, int n, int32_t* __restrict__ indptr,
                                                int32_t* __restrict__ seg) {
  __shared__ int s_scan[40];
  int base = 0, sbase = 0;
  for (int b = 0; b < n; b += blockDim.x * 4) {
    const int i0 = b + threadIdx.x * 4;
    int v[4], pre[4], spre[4];
    load4(cnt_cursor, i0, n, v);
    int vsum = 0, ssum = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      vsum += v[u];
      ssum += (i0 + u < n) ? max(1, (v[u] + BLISS_SPMM_SEG - 1) / BLISS_SPMM_SEG) : 0;
    }
    int tot, st = 0;
    int p = base + block_excl_scan(vsum, s_scan, &tot);
    int sp = sbase + (seg ? block_excl_scan(ssum, s_scan, &st) : 0);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      pre[u] = p;
      spre[u] = sp;
      p += v[u];
      sp += max(1, (v[u] + BLISS_SPMM_SEG - 1) / BLISS_SPMM_SEG);
    }
    store4(indptr, i0, n, pre);
    store4(cnt_cursor, i0, n, pre);
    if (seg) store4(seg, i0, n, spre);
    base += tot;
    sbase += st;
  }
  if (threadIdx.x == 0) {
    indptr[n] = base;
    if (seg) seg[n] = sbase;
  }
}
// The transpose is driven by a source x destination bitmap of the block (bits[s][d]: edge s->d is in
// the block; a block is a simple graph, so a bit is an edge).  A source's row of the transpose is
// its set bits in ascending destination = ascending edge id (block edges are destination-major), so
// the backward sums are deterministic without a sort and without returning atomics:
//   k_t_mark   bits[src_e][dst_e] = 1                      (eager path; the pooled path marks in k_block_fill)
//   k_t_rows   per source: word prefix popcounts -> pre[s][w], and t_dst[t_indptr[s] + k] = k-th set bit
//   k_t_place  per edge:   t_perm[t_indptr[s] + pre[s][d/32] + popc(bits[s][d/32] below d)] = e
__global__ void k_t_mark(const int32_t* __restrict__ edge_src, const int32_t* __restrict__ edge_dst, int64_t n_edges,
                         uint32_t* __restrict__ bits, int64_t t_words) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_edges; e += stride) {
    const int d = edge_dst[e];
    atomicOr(&bits[(int64_t)edge_src[e] * t_words + (d >> 5)], 1u << (d & 31));
  }
}
__global__ void __launch_bounds__(256) k_t_rows(const int32_t* __restrict__ t_indptr, int n_src, int n_words,
                                               const uint32_t* __restrict__ bits, int64_t t_words,
                                               int32_t* __restrict__ pre, int32_t* __restrict__ t_dst) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int lane = lane_id();
  for (int s = warp; s < n_src; s += nwarps) {
    const int a = t_indptr[s], b = t_indptr[s + 1];
    if (a == b) continue;   // no edges: nothing reads this row's prefix
    int base = a;
    for (int w0 = 0; w0 < n_words; w0 += 32) {
      const int w = w0 + lane;
      unsigned word = (w < n_words) ? bits[(int64_t)s * t_words + w] : 0u;
      const int c = __popc(word);
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      int k = base + incl - c;
      if (w < n_words) pre[(int64_t)s * t_words + w] = k - a;
      while (word) {   // this lane's destinations, ascending
        const int bit = __ffs(word) - 1;
        word &= word - 1;
        t_dst[k++] = (w << 5) + bit;
      }
      base += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
}
__global__ void k_t_place(const int32_t* __restrict__ edge_src, const int32_t* __restrict__ edge_dst, int64_t n_edges,
                          const int64_t* __restrict__ n_edges_dev, const int32_t* __restrict__ t_indptr,
                          const uint32_t* __restrict__ bits, const int32_t* __restrict__ pre, int64_t t_words,
                          int32_t* __restrict__ t_perm, const float* __restrict__ edge_w, float* __restrict__ t_w) {
  if (n_edges_dev) n_edges = min(n_edges, *n_edges_dev);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_edges; e += stride) {
    const int s = edge_src[e], d = edge_dst[e];
    const int64_t wi = (int64_t)s * t_words + (d >> 5);
    const int slot = t_indptr[s] + pre[wi] + __popc(bits[wi] & ((1u << (d & 31)) - 1u));
    t_perm[slot] = (int32_t)e;
    if (t_w) t_w[slot] = edge_w[e];   // block weights in transpose order: the backward SpMM streams them
  }
}

}  // namespace bliss

// ==========================================================================================
// C ABI
// ==========================================================================================
using namespace bliss;

// Launch with a programmatic edge to the previous kernel of the stream (see pdl_wait in common.cuh).
// BLISS_PDL=0 in the environment falls back to ordinary stream order.
static bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("BLISS_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define BLISS_LAUNCH_PDL(kernel, grid, block, smem, st, ...)                         \
  do {                                                                               \
    BLISS_KSCOPE(#kernel, st);                                                       \
    cudaError_t e__ = launch_pdl(kernel, grid, block, smem, st, __VA_ARGS__);        \
    if (e__ != cudaSuccess) return (int)e__;                                         \
  } while (0)

static inline GraphView view_of(const bliss_graph* g) {
  GraphView v;
  v.indptr = g->indptr;
  v.indices = g->indices;
  v.eid = g->eid;
  v.num_nodes = g->num_nodes;
  return v;
}
// Grid of the chunk passes: the chunk count lives on the device, so the grid is sized from the
// host's bound on the rows (a CTA takes 8 chunks per round; CTAs that find the queue empty exit at
// once) and capped at what is resident at once (6 CTAs of 256 threads per SM at <= 40 registers).
static inline int chunk_grid(int64_t n_seeds, int per_sm = 6) {
  int64_t b = 4 * n_seeds + 8;
  if (b < 1) b = 1;
  if (b > BLISS_SM_COUNT * per_sm) b = BLISS_SM_COUNT * per_sm;
  return (int)b;
}
static inline int grid_for(int64_t n, int threads, int max_blocks) {
  int64_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

extern "C" {

int bliss_version(void) { return BLISS_B200_VERSION; }

int bliss_workspace_init(const bliss_workspace* ws, int64_t num_nodes, void* stream) {
  if (!ws || num_nodes <= 0) return -1;
  k_ws_init<<<grid_for(num_nodes, 256, BLISS_SM_COUNT * 8), 256, 0, (cudaStream_t)stream>>>(*ws, num_nodes);
  return 0;
}

int bliss_frontier_plan(const bliss_graph* g, const int32_t* seeds, int32_t n_seeds,
                        const bliss_workspace* ws, void* stream) {
  if (!g || !seeds || !ws || n_seeds < 0 || n_seeds > ws->cap_seeds) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_seeds > 0) {
    BLISS_LAUNCH_PDL(k_plan_rows, dim3((n_seeds + 255) / 256), dim3(256), 0, st, view_of(g), seeds, n_seeds, *ws);
  }
  BLISS_LAUNCH_PDL(k_plan_scan, dim3(1), dim3(1024), 0, st, n_seeds, *ws);
  if (n_seeds > 0) {
    BLISS_LAUNCH_PDL(k_plan_chunks, dim3(grid_for((int64_t)n_seeds * 32, 256, BLISS_SM_COUNT * 8)), dim3(256), 0, st, *ws);
  }
  return 0;
}

int bliss_frontier_prob(const bliss_graph* g, const int32_t* seeds, int32_t n_seeds,
                        const float* edge_weight_csc, float eta, int32_t mode,
                        const bliss_workspace* ws, void* stream) {
  if (!g || !seeds || !ws || n_seeds < 0) return -1;
  if (!edge_weight_csc) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  const float one_minus_eta = (float)(1.0 - (double)eta);
  // persistent grids over the plan's chunks (count read on the device): every resident CTA pulls
  // groups of 8 chunks from the layer's queue
  const int blocks = chunk_grid(n_seeds);
  const bool bandit = (mode & 1) == BLISS_MODE_BANDIT;
  if (bandit) {   // the row sums are also needed by the block fill when importance sampling is off
    BLISS_LAUNCH_PDL(k_prob_pass1, dim3(blocks), dim3(256), 0, st, edge_weight_csc, *ws);
    BLISS_LAUNCH_PDL(k_prob_pass2, dim3(blocks), dim3(256), 0, st, edge_weight_csc, eta, one_minus_eta, *ws);
  }
  BLISS_LAUNCH_PDL(k_prob_pass3, dim3(chunk_grid(n_seeds, 5)), dim3(256), 0, st, view_of(g), edge_weight_csc, eta,
                   one_minus_eta, (int)mode, *ws);
  BLISS_LAUNCH_PDL(k_collect_candidates, dim3(grid_for(g->num_nodes, 32 * BLISS_COLLECT_WORDS * 8, BLISS_SM_COUNT * 8)),
                   dim3(256), 0, st, (int64_t)g->num_nodes, (mode & BLISS_COLLECT_BITMAP) ? 1 : 0, *ws);
  return 0;
}

int bliss_poisson_scale(int32_t n_seeds, int32_t fanout, double eps, int32_t poisson,
                        const bliss_workspace* ws, void* stream) {
  if (!ws || fanout < 0) return -1;
  const double fx_inv = 1.0 / (double)(1ull << fx_bits_for(n_seeds));
  BLISS_LAUNCH(k_poisson_scale, 1, 1024, 0, (cudaStream_t)stream, n_seeds, fanout, eps, poisson, *ws, fx_inv);
  return 0;
}

int bliss_select_poisson(int32_t n_seeds, uint64_t seed, uint64_t step, uint32_t layer,
                         const float* u_inject, const bliss_workspace* ws, void* stream) {
  if (!ws) return -1;
  BLISS_LAUNCH_PDL(k_select_poisson, dim3(BLISS_SM_COUNT * 2), dim3(256), 0, (cudaStream_t)stream, (int)n_seeds,
                   (unsigned long long)seed, (unsigned long long)step, (unsigned)layer, u_inject, *ws);
  return 0;
}

static int g_cluster_size = 0;  // 16, 8, or -1 (clusters unavailable)

int bliss_poisson_select(int32_t n_seeds, int32_t fanout, double eps, uint64_t seed, uint64_t step,
                         uint32_t layer, const float* u_inject, const bliss_workspace* ws, void* stream) {
  if (!ws || fanout < 0) return -1;
  if (g_cluster_size == 0) {
    g_cluster_size = -1;
    if (cudaFuncSetAttribute(k_scale_search, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
      for (int cs = BLISS_SCALE_MAX_CLUSTER; cs >= 8 && g_cluster_size < 0; cs >>= 1) {
        cudaLaunchConfig_t q = {};
        q.gridDim = dim3(cs);
        q.blockDim = dim3(1024);
        cudaLaunchAttribute a[1];
        a[0].id = cudaLaunchAttributeClusterDimension;
        a[0].val.clusterDim.x = cs;
        a[0].val.clusterDim.y = a[0].val.clusterDim.z = 1;
        q.attrs = a;
        q.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, k_scale_search, &q) == cudaSuccess && n > 0) g_cluster_size = cs;
      }
    }
    (void)cudaGetLastError();
  }
  if (g_cluster_size < 0) {  // no cluster support: the single-CTA search
    int rc = bliss_poisson_scale(n_seeds, fanout, eps, 1, ws, stream);
    if (rc) return rc;
    return bliss_select_poisson(n_seeds, seed, step, layer, u_inject, ws, stream);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(g_cluster_size);
  cfg.blockDim = dim3(1024);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = g_cluster_size;
  attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  {
    BLISS_KSCOPE("k_scale_search", stream);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_scale_search, (int)fanout, eps, *ws);
    if (e != cudaSuccess) return (int)e;
  }
  return bliss_select_poisson(n_seeds, seed, step, layer, u_inject, ws, stream);
}

int bliss_select_topk(int32_t n_seeds, int32_t fanout, uint64_t seed, uint64_t step, uint32_t layer,
                      const float* u_inject, float* key_scratch, const bliss_workspace* ws, void* stream) {
  if (!ws || !key_scratch) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  // the threshold triple lives in the tail of the caller's key scratch ([V + 4] floats)
  unsigned* thr = reinterpret_cast<unsigned*>(key_scratch);
  float* keys = key_scratch + 4;
  BLISS_LAUNCH(k_topk_keys, BLISS_SM_COUNT * 4, 256, 0, st, n_seeds, seed, step, layer, u_inject, keys, *ws);
  BLISS_LAUNCH(k_topk_threshold, 1, 1024, 0, st, fanout, keys, thr, *ws);
  BLISS_LAUNCH(k_topk_mark, 1, 1024, 0, st, n_seeds, keys, thr, *ws);
  return 0;
}

int bliss_philox_fill(uint64_t seed, uint64_t step, uint32_t layer, const int32_t* nids, int64_t n,
                      float* out, void* stream) {
  if (n < 0 || (n > 0 && (!nids || !out))) return -1;
  if (n == 0) return 0;
  BLISS_LAUNCH(k_philox_fill, grid_for(n, 256, BLISS_SM_COUNT * 8), 256, 0, (cudaStream_t)stream, seed, step, layer, nids, n, out);
  return 0;
}

int bliss_neighbor_select(const bliss_graph* g, int32_t n_seeds, int32_t fanout, uint64_t seed, uint64_t step,
                          uint32_t layer, const bliss_workspace* ws, void* stream) {
  if (!g || !ws || n_seeds < 0) return -1;
  BLISS_LAUNCH_PDL(k_neighbor_select, dim3(grid_for((int64_t)n_seeds * 32, 256, BLISS_SM_COUNT * 8)), dim3(256), 0,
                   (cudaStream_t)stream, view_of(g), (int)fanout, (unsigned long long)seed, (unsigned long long)step,
                   (unsigned)layer, *ws);
  return 0;
}

int bliss_block_count(const bliss_graph* g, const int32_t* seeds, int32_t n_seeds,
                      const float* edge_weight_csc, float eta, int32_t mode,
                      const bliss_workspace* ws, void* stream) {
  if (!g || !seeds || !ws || !edge_weight_csc) return -1;
  FillCtx c;
  c.g = view_of(g);
  c.W = edge_weight_csc;
  c.eta = eta;
  c.one_minus_eta = (float)(1.0 - (double)eta);
  c.mode = mode & 1;
  if (mode & BLISS_MODE_NEIGHBOR) {   // the keep bits were chosen per edge by bliss_neighbor_select
    BLISS_LAUNCH_PDL((k_block_count<false, true>), dim3(chunk_grid(n_seeds)), dim3(BLISS_CTA), 0, (cudaStream_t)stream, c,
                     *ws, 0);
    return 0;
  }
  // selected-node bitmap in shared memory when six CTAs of it fit an SM (|V| <= ~280 K), else tested in L1/L2
  const int bit_words = (int)((g->num_nodes + 31) / 32);
  const size_t smem = (size_t)bit_words * sizeof(uint32_t);
  if (smem <= 35 * 1024) {
    BLISS_LAUNCH_PDL(k_block_count<true>, dim3(chunk_grid(n_seeds)), dim3(BLISS_CTA), smem, (cudaStream_t)stream, c, *ws,
                     bit_words);
  } else {
    BLISS_LAUNCH_PDL(k_block_count<false>, dim3(chunk_grid(n_seeds)), dim3(BLISS_CTA), 0, (cudaStream_t)stream, c, *ws,
                     bit_words);
  }
  return 0;
}

int bliss_block_index(const int32_t* seeds, int32_t n_seeds, const bliss_workspace* ws,
                      const bliss_block_out* out, void* stream) {
  if (!seeds || !ws || !out || !out->indptr || !out->src_nid || !out->node_prob) return -1;
  BLISS_LAUNCH_PDL(k_block_index, dim3(BLISS_SM_COUNT), dim3(1024), 0, (cudaStream_t)stream, seeds, (int)n_seeds, *ws, *out);
  return 0;
}

int bliss_block_fill(const bliss_graph* g, const int32_t* seeds, int32_t n_seeds,
                     const float* edge_weight_csc, float eta, int32_t mode,
                     const bliss_workspace* ws, const bliss_block_out* out, void* stream) {
  if (!g || !seeds || !ws || !out || !edge_weight_csc) return -1;
  if (!out->edge_src || !out->edge_dst || !out->csc_pos || !out->edge_w) return -1;
  FillCtx c;
  c.g = view_of(g);
  c.W = edge_weight_csc;
  c.eta = eta;
  c.one_minus_eta = (float)(1.0 - (double)eta);
  c.mode = mode & 1;
  BLISS_LAUNCH(k_block_fill, chunk_grid(n_seeds), BLISS_CTA, 0, (cudaStream_t)stream, c, *ws, *out);
  return 0;
}

int bliss_block_finish(int32_t n_seeds, int32_t mode, const bliss_workspace* ws,
                       const bliss_block_out* out, void* stream) {
  if (!ws || !out) return -1;
  (void)n_seeds;
  BLISS_LAUNCH(k_block_finish, BLISS_SM_COUNT * 4, 256, 0, (cudaStream_t)stream, mode & 1, *ws, *out);
  return 0;
}

int bliss_block_transpose(const int32_t* edge_src, const int32_t* edge_dst, int64_t n_edges,
                          int32_t n_src, int32_t n_dst, int32_t* t_indptr, int32_t* t_cursor,
                          uint32_t* t_bits, int32_t* t_pre, int64_t t_words, int32_t* t_dst, int32_t* t_perm,
                          int32_t* t_seg_ptr, int32_t have_counts, const int64_t* n_edges_dev, const float* edge_w,
                          float* t_w, void* stream) {
  if (n_edges < 0 || n_src < 0 || !t_indptr || !t_cursor) return -1;
  if ((edge_w == nullptr) != (t_w == nullptr)) return -1;
  if (n_edges > 0 && (!edge_src || !edge_dst || !t_bits || !t_pre || !t_dst || !t_perm)) return -1;
  if (n_edges > 0 && t_words < (n_dst + 31) / 32) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
  if (!have_counts) {   // otherwise the block fill already counted (bliss_block_out.out_deg) and marked (t_bits)
    e = cudaMemsetAsync(t_cursor, 0, sizeof(int32_t) * (size_t)(n_src > 0 ? n_src : 1), st);
    if (e != cudaSuccess) return (int)e;
    if (n_edges > 0) {
      e = cudaMemsetAsync(t_bits, 0, sizeof(uint32_t) * (size_t)n_src * (size_t)t_words, st);
      if (e != cudaSuccess) return (int)e;
      BLISS_LAUNCH(k_t_count, grid_for(n_edges, 256, BLISS_SM_COUNT * 8), 256, 0, st, edge_src, n_edges, t_cursor);
      BLISS_LAUNCH(k_t_mark, grid_for(n_edges, 256, BLISS_SM_COUNT * 8), 256, 0, st, edge_src, edge_dst, n_edges, t_bits, t_words);
    }
  }
  BLISS_LAUNCH(k_t_scan, 1, 1024, 0, st, t_cursor, n_src, t_indptr, t_seg_ptr);
  if (n_edges == 0) return 0;
  const int n_words = (n_dst + 31) / 32;
  BLISS_LAUNCH(k_t_rows, grid_for((int64_t)n_src * 32, 256, BLISS_SM_COUNT * 8), 256, 0, st, t_indptr, n_src, n_words, t_bits, t_words,
                                                                              t_pre, t_dst);
  BLISS_LAUNCH(k_t_place, grid_for(n_edges, 256, BLISS_SM_COUNT * 8), 256, 0, st, edge_src, edge_dst, n_edges, n_edges_dev, t_indptr,
                                                                   t_bits, t_pre, t_words, t_perm, edge_w, t_w);
  return 0;
}

// One call per layer phase for the host fast path (fewer FFI crossings, same kernels):
// front = plan -> probabilities (+ candidate collection) -> selection -> kept-edge count -> index.
int bliss_sample_layer_front(const bliss_graph* g, const int32_t* seeds, int32_t n_seeds,
                             const float* edge_weight_csc, float eta, int32_t mode, int32_t fanout, double eps,
                             int32_t poisson, uint64_t seed, uint64_t step, uint32_t layer, const float* u_inject,
                             float* key_scratch, const bliss_workspace* ws, const bliss_block_out* out,
                             void* stream) {
  int rc = (mode & BLISS_MODE_PLANNED) ? 0 : bliss_frontier_plan(g, seeds, n_seeds, ws, stream);
  if (rc) return rc;
  mode &= ~BLISS_MODE_PLANNED;
  if (mode & BLISS_MODE_NEIGHBOR) {   // uniform fan-out per seed / full neighbourhood: no probabilities, no node selection
    rc = bliss_neighbor_select(g, n_seeds, fanout, seed, step, layer, ws, stream);
    if (rc) return rc;
    rc = bliss_block_count(g, seeds, n_seeds, edge_weight_csc, eta, mode, ws, stream);
    if (rc) return rc;
    return bliss_block_index(seeds, n_seeds, ws, out, stream);
  }
  rc = bliss_frontier_prob(g, seeds, n_seeds, edge_weight_csc, eta, mode, ws, stream);
  if (rc) return rc;
  if (poisson) {
    rc = bliss_poisson_select(n_seeds, fanout, eps, seed, step, layer, u_inject, ws, stream);
  } else {
    rc = bliss_poisson_scale(n_seeds, fanout, eps, 0, ws, stream);
    if (rc) return rc;
    rc = bliss_select_topk(n_seeds, fanout, seed, step, layer, u_inject, key_scratch, ws, stream);
  }
  if (rc) return rc;
  rc = bliss_block_count(g, seeds, n_seeds, edge_weight_csc, eta, mode, ws, stream);
  if (rc) return rc;
  return bliss_block_index(seeds, n_seeds, ws, out, stream);
}

// back = fill -> normalise + workspace restore.
int bliss_sample_layer_back(const bliss_graph* g, const int32_t* seeds, int32_t n_seeds,
                            const float* edge_weight_csc, float eta, int32_t mode, const bliss_workspace* ws,
                            const bliss_block_out* out, void* stream) {
  if (out && out->cap_edges > 0) {
    int rc = bliss_block_fill(g, seeds, n_seeds, edge_weight_csc, eta, mode, ws, out, stream);
    if (rc) return rc;
  }
  return bliss_block_finish(n_seeds, mode, ws, out, stream);
}

}  // extern "C"
