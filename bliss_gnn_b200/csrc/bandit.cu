// bandit.cu — EXP3 reward / weight update of the BLISS hot path on sm_100a.
// Replaces calculate_alpha / calculate_rewards / update_exp3_weights (bandit_sampler.py:140-249):
// five g-SDDMM launches, two scalar g-SpMMs, an index_put and a dense F.normalize over |E| per
// layer per step in the reference become one pass over the block's edges that reads every
// operand once and read-modify-writes the CSC-ordered weight in place.
#include <float.h>
#include <string.h>
#include "common.cuh"
#include "profile.cuh"

namespace bliss {

__device__ __forceinline__ float nan_to_num_default(float x) {  // torch.nan_to_num(x)
  if (isnan(x)) return 0.0f;
  if (isinf(x)) return x > 0 ? FLT_MAX : -FLT_MAX;
  return x;
}
__device__ __forceinline__ float nan_to_num_posinf0(float x) {  // torch.nan_to_num(x, posinf=0)
  if (isnan(x)) return 0.0f;
  if (isinf(x)) return x > 0 ? 0.0f : -FLT_MAX;
  return x;
}

// GAT alpha (bandit_sampler.py:148-154) needs Σ a_ij and Σ q_ij per destination: warp per row,
// fp64 accumulation rounded once (numeric contract).
__global__ void __launch_bounds__(256) k_row_sums2(const int32_t* __restrict__ indptr, const float* __restrict__ a,
                                                  const float* __restrict__ q, int n_rows,
                                                  float* __restrict__ asum, float* __restrict__ qsum) {
  const int lane = lane_id();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < n_rows; r += nwarps) {
    double sa = 0.0, sq = 0.0;
    for (int e = indptr[r] + lane; e < indptr[r + 1]; e += 32) {
      sa += (double)a[e];
      sq += (double)q[e];
    }
    sa = warp_sum(sa);
    sq = warp_sum(sq);
    if (lane == 0) {
      asum[r] = __double2float_rn(sa);
      qsum[r] = __double2float_rn(sq);
    }
  }
}

struct RewardArgs {
  const int64_t* __restrict__ g_indptr;
  const int32_t* __restrict__ blk_indptr;
  const int32_t* __restrict__ edge_src;
  const int32_t* __restrict__ edge_dst;
  const int64_t* __restrict__ csc_pos;
  const int32_t* __restrict__ dst_nid;
  const float* __restrict__ q_ij;
  const float* __restrict__ node_prob;
  const float* __restrict__ embed_norm;
  const float* __restrict__ w_static;
  const float* __restrict__ a_ij;
  const float* __restrict__ asum;
  const float* __restrict__ qsum;
  int alpha_mode;
  float delta;
  int64_t n_edges;
  const int64_t* n_edges_dev;   // true edge count on the device (n_edges is then the capacity)
  int64_t* count_out;           // data-parallel exchange header slot (or NULL)
  int32_t* pos_out;             // data-parallel exchange: CSC position of every edge as int32 (or NULL)
  float* exp3_w;
  float* rewards;
  float* x_out;
  double* l1_delta;
  float* wmax;                  // running max of the updated weights (or NULL)
  bliss_p2p p2p;                // peer-memory exchange (world == 0: off)
};

// ---- peer-memory exchange of the sparse bandit updates (data parallel) -------------------------------------
// Every rank owns a window in symmetric memory that all ranks can address:
//   window = [parity 0 | parity 1] x [slot of rank 0 | ... | slot of rank W-1] ++ flags[2][L][W] (uint64)
// and a slot has the layout of the packed exchange buffer (int64 counts, then per layer int32 pos[cap], fp32 x[cap]).
// The reward kernel stores every edge's (position, exponent) straight into ITS slot of EVERY rank's window over
// NVLink (fire-and-forget stores: the "all-gather" is fused into the kernel that produces the data); the last CTA to
// finish publishes flag = step + 1 in every window.  A consumer polls its own flags (one CTA), then applies the W
// slots of its own window.  Parity = step & 1: ranks are never two steps apart (the gradient all-reduce of every
// step is a barrier), so a slot is not overwritten while a slower rank still reads it.
// Running upper bound of a layer's EXP3 weights (range guard of the lazy normalisation: the weights are re-scaled
// when it nears the top of the fp32 range, not on a schedule).  Weights are positive, so their bit patterns order
// like integers.  One atomic per CTA.
__device__ __forceinline__ void publish_wmax(float m, float* wmax, float* s_max /* [32] */) {
  if (!wmax) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) s_max[warp] = m;
  __syncthreads();
  if (warp == 0) {
    m = (lane < (int)((blockDim.x + 31) >> 5)) ? s_max[lane] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0 && m > 0.0f) atomicMax(reinterpret_cast<int*>(wmax), __float_as_int(m));
  }
}

__device__ __forceinline__ unsigned char* p2p_slot(const bliss_p2p& q, int peer, int parity, int src_rank) {
  return reinterpret_cast<unsigned char*>(q.peer_base[peer]) + (int64_t)parity * q.parity_stride +
         (int64_t)src_rank * q.rank_stride;
}
__device__ __forceinline__ unsigned long long* p2p_flag(const bliss_p2p& q, int peer, int parity, int layer, int src_rank) {
  return reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(q.peer_base[peer]) + q.flags_off) +
         ((int64_t)parity * q.n_layers + layer) * q.world + src_rank;
}

__global__ void __launch_bounds__(256) k_reward_update(RewardArgs p) {
  __shared__ double s_red[32];
  __shared__ float s_max[32];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double dsum = 0.0;
  float wtop = 0.0f;
  const int64_t n_edges = p.n_edges_dev ? min(p.n_edges, *p.n_edges_dev) : p.n_edges;
  if (p.count_out && blockIdx.x == 0 && threadIdx.x == 0) *p.count_out = n_edges;
  const int W = p.p2p.world;
  const long long xstep = W ? *p.p2p.step_dev : 0;
  const int parity = (int)(xstep & 1);
  const int r_lo = (W && p.p2p.pull) ? p.p2p.rank : 0, r_hi = (W && p.p2p.pull) ? p.p2p.rank + 1 : W;   // windows written
  if (W && blockIdx.x == 0 && (int)threadIdx.x >= r_lo && (int)threadIdx.x < r_hi)   // this layer's edge count
    *reinterpret_cast<int64_t*>(p2p_slot(p.p2p, threadIdx.x, parity, p.p2p.rank) + p.p2p.count_off) = n_edges;
  unsigned char* mc_slot = (W && p.p2p.mc_base && !p.p2p.pull)     // my slot in the multicast view of the windows
                               ? reinterpret_cast<unsigned char*>(p.p2p.mc_base) + (int64_t)parity * p.p2p.parity_stride +
                                     (int64_t)p.p2p.rank * p.p2p.rank_stride
                               : nullptr;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_edges; e += stride) {
    const int i = p.edge_dst[e];
    const int u = p.edge_src[e];
    const int64_t pos = p.csc_pos[e];
    const float q = p.q_ij[e];
    float alpha;
    if (p.alpha_mode == 0) {
      alpha = __ldg(p.w_static + pos);                                           // :157
    } else {
      float ad = nan_to_num_default(__fdiv_rn(p.a_ij[e], p.asum[i]));            // :152-153
      alpha = __fmul_rn(ad, p.qsum[i]);                                          // :154
    }
    const float k_i = (float)(p.blk_indptr[i + 1] - p.blk_indptr[i]);            // :180
    const float a_k = nan_to_num_posinf0(__fdiv_rn(__fmul_rn(alpha, alpha), k_i));   // :186-187
    const float h = p.embed_norm[u];
    const float hq = __fdiv_rn(__fmul_rn(h, h), __fmul_rn(q, q));                // :189
    const float r = __fmul_rn(a_k, hq);                                          // :191
    if (p.rewards) p.rewards[e] = r;
    const int dn = p.dst_nid[i];
    const float n_i = (float)(p.g_indptr[dn + 1] - p.g_indptr[dn]);              // :223
    const float r_hat = __fdiv_rn(r, p.node_prob[u]);                            // :240
    float x = __fmul_rn(r_hat, __fdiv_rn(p.delta, n_i));                         // :242
    if (x > 1.0f) x = 1.0f;                                                      // :244
    if (p.x_out) p.x_out[e] = x;
    if (p.pos_out) p.pos_out[e] = (int32_t)pos;
    if (mc_slot) {                                  // one store each, replicated into every window by the switch
      multimem_st_b32(reinterpret_cast<int32_t*>(mc_slot + p.p2p.pos_off) + e, (uint32_t)(int32_t)pos);
      multimem_st_f32(reinterpret_cast<float*>(mc_slot + p.p2p.x_off) + e, x);
    } else {
      for (int r = r_lo; r < r_hi; ++r) {           // my slot in rank r's window (pull mode: only my own window)
        unsigned char* slot = p2p_slot(p.p2p, r, parity, p.p2p.rank);
        reinterpret_cast<int32_t*>(slot + p.p2p.pos_off)[e] = (int32_t)pos;
        reinterpret_cast<float*>(slot + p.p2p.x_off)[e] = x;
      }
    }
    if (p.exp3_w) {
      const float w_old = p.exp3_w[pos];
      const float w_new = __fmul_rn(w_old, expf(x));                             // :246-248
      p.exp3_w[pos] = w_new;
      dsum += (double)w_new - (double)w_old;
      wtop = fmaxf(wtop, w_new);
    }
  }
  if (p.l1_delta) {
    dsum = block_sum(dsum, s_red);
    if (threadIdx.x == 0 && dsum != 0.0) atomicAdd(p.l1_delta, dsum);
  }
  if (p.exp3_w) publish_wmax(wtop, p.wmax, s_max);
  if (W) {   // publish: every CTA's stores are ordered before its ticket; the last CTA raises the flags
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned t = atomicAdd(p.p2p.done_ctr, 1u);
      if (t == gridDim.x - 1) {
        *p.p2p.done_ctr = 0u;
        __threadfence_system();
        for (int r = 0; r < W; ++r)
          *reinterpret_cast<volatile unsigned long long*>(p2p_flag(p.p2p, r, parity, p.p2p.layer, p.p2p.rank)) =
              (unsigned long long)(xstep + 1);
        __threadfence_system();
      }
    }
  }
}

// One CTA polls this rank's flags of one layer until every rank has published this step (bounded: a peer that
// never arrives raises *error instead of hanging the device), so that the apply kernel behind it never spins with
// a full grid.
__global__ void __launch_bounds__(32) k_p2p_wait(bliss_p2p q, int32_t* error) {
  const long long want = *q.step_dev + 1;
  const int parity = (int)((want - 1) & 1);
  if ((int)threadIdx.x < q.world) {
    volatile unsigned long long* f = p2p_flag(q, q.rank, parity, q.layer, threadIdx.x);
    const long long t0 = clock64();
    while ((long long)*f != want) {
      if (clock64() - t0 > 8000000000ll) {      // ~4 s at 2 GHz
        if (error) atomicOr(error, 1 << q.layer);
        break;
      }
      __nanosleep(100);
    }
  }
  __threadfence_system();
}

// apply all ranks' updates of one layer from this rank's own window (after k_p2p_wait)
__global__ void __launch_bounds__(256) k_apply_updates_p2p(bliss_p2p q, int64_t cap, float* exp3_w, double* l1_delta,
                                                          float* wmax) {
  __shared__ double s_red[32];
  __shared__ float s_max[32];
  float wtop = 0.0f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t total = (int64_t)q.world * cap;
  const int parity = (int)(*q.step_dev & 1);
  double dsum = 0.0;
  // Four updates per thread and round: their window reads, their reads of the old weights (random 4-byte reads of a
  // |E|-sized array: DRAM latency) and their first compare-and-swap attempts are all in flight together.
  constexpr int U = 4;
  for (int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t0 < total; t0 += U * stride) {
    float* addr[U];
    float f[U];
    unsigned old[U], got[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t t = t0 + u * stride;
      addr[u] = nullptr;
      if (t < total) {
        const int r = (int)(t / cap);
        const int64_t k = t - (int64_t)r * cap;
        // push: every rank stored its slot into my window; pull: rank r's slot is read from rank r's window over NVLink
        const unsigned char* base = p2p_slot(q, q.pull ? r : q.rank, parity, r);
        const int64_t n = __ldcg(reinterpret_cast<const long long*>(base + q.count_off));   // (L1 may hold the slot's old lines)
        if (k < n) {
          addr[u] = exp3_w + __ldcg(reinterpret_cast<const int32_t*>(base + q.pos_off) + k);
          f[u] = expf(__ldcg(reinterpret_cast<const float*>(base + q.x_off) + k));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (addr[u]) old[u] = __float_as_uint(*addr[u]);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (addr[u])
        got[u] = atomicCAS(reinterpret_cast<unsigned*>(addr[u]), old[u],
                           __float_as_uint(__fmul_rn(__uint_as_float(old[u]), f[u])));
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!addr[u]) continue;
      unsigned seen = got[u], assumed = old[u];
      while (seen != assumed) {   // two ranks sampled the same edge (or the same position twice in this round): retry
        assumed = seen;
        seen = atomicCAS(reinterpret_cast<unsigned*>(addr[u]), assumed,
                         __float_as_uint(__fmul_rn(__uint_as_float(assumed), f[u])));
      }
      const float w_old = __uint_as_float(seen);
      dsum += (double)__fmul_rn(w_old, f[u]) - (double)w_old;
      wtop = fmaxf(wtop, __fmul_rn(w_old, f[u]));
    }
  }
  if (l1_delta) {
    dsum = block_sum(dsum, s_red);
    if (threadIdx.x == 0 && dsum != 0.0) atomicAdd(l1_delta, dsum);
  }
  publish_wmax(wtop, wmax, s_max);
}

__global__ void __launch_bounds__(256) k_apply_updates(const int64_t* __restrict__ pos, const float* __restrict__ x,
                                                      int64_t n, float* exp3_w, double* l1_delta, float* wmax) {
  __shared__ double s_red[32];
  __shared__ float s_max[32];
  float wtop = 0.0f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double dsum = 0.0;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += stride) {
    // the same CSC position may be updated by several ranks: multiplicative -> atomic CAS loop
    float* addr = exp3_w + pos[k];
    const float f = expf(x[k]);
    unsigned old = __float_as_uint(*addr), assumed;
    do {
      assumed = old;
      old = atomicCAS(reinterpret_cast<unsigned*>(addr), assumed,
                      __float_as_uint(__fmul_rn(__uint_as_float(assumed), f)));
    } while (old != assumed);
    const float w_old = __uint_as_float(old);
    dsum += (double)__fmul_rn(w_old, f) - (double)w_old;
    wtop = fmaxf(wtop, __fmul_rn(w_old, f));
  }
  if (l1_delta) {
    dsum = block_sum(dsum, s_red);
    if (threadIdx.x == 0 && dsum != 0.0) atomicAdd(l1_delta, dsum);
  }
  publish_wmax(wtop, wmax, s_max);
}

// Apply every rank's update of one layer straight from the all-gathered exchange buffer:
// per rank a header of int64 counts, then per layer int32 positions and fp32 exponents (8 bytes per
// sampled edge on the wire).
__global__ void __launch_bounds__(256) k_apply_updates_packed(const unsigned char* __restrict__ recv,
                                                             int64_t rank_stride, int world, int64_t count_off,
                                                             int64_t pos_off, int64_t x_off, int64_t cap,
                                                             float* exp3_w, double* l1_delta, float* wmax) {
  __shared__ double s_red[32];
  __shared__ float s_max[32];
  float wtop = 0.0f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t total = (int64_t)world * cap;
  double dsum = 0.0;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += stride) {
    const int r = (int)(t / cap);
    const int64_t k = t - (int64_t)r * cap;
    const unsigned char* base = recv + (int64_t)r * rank_stride;
    const int64_t n = *reinterpret_cast<const int64_t*>(base + count_off);
    if (k >= n) continue;
    const int64_t pos = reinterpret_cast<const int32_t*>(base + pos_off)[k];
    const float f = expf(reinterpret_cast<const float*>(base + x_off)[k]);
    float* addr = exp3_w + pos;
    unsigned old = __float_as_uint(*addr), assumed;
    do {  // two ranks may have sampled the same edge: multiplicative update through a CAS loop
      assumed = old;
      old = atomicCAS(reinterpret_cast<unsigned*>(addr), assumed,
                      __float_as_uint(__fmul_rn(__uint_as_float(assumed), f)));
    } while (old != assumed);
    const float w_old = __uint_as_float(old);
    dsum += (double)__fmul_rn(w_old, f) - (double)w_old;
    wtop = fmaxf(wtop, __fmul_rn(w_old, f));
  }
  if (l1_delta) {
    dsum = block_sum(dsum, s_red);
    if (threadIdx.x == 0 && dsum != 0.0) atomicAdd(l1_delta, dsum);
  }
  publish_wmax(wtop, wmax, s_max);
}

// literal F.normalize(w, p=1): fixed two-level tree -> deterministic
#define BLISS_NORM_BLOCKS 1024
__global__ void __launch_bounds__(256) k_l1_partial(const float* __restrict__ w, int64_t n, double* __restrict__ partial) {
  __shared__ double s_red[32];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  double acc = 0.0;
  const int64_t n4 = n >> 2;
  const float4* __restrict__ w4 = reinterpret_cast<const float4*>(w);
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n4; v += stride / 4) {
    float4 t = __ldg(w4 + v);
    acc += ((double)fabsf(t.x) + (double)fabsf(t.y)) + ((double)fabsf(t.z) + (double)fabsf(t.w));
  }
  if (blockIdx.x == 0) {
    for (int64_t k = (n4 << 2) + threadIdx.x; k < n; k += blockDim.x) acc += (double)fabsf(w[k]);
  }
  acc = block_sum(acc, s_red);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
__global__ void __launch_bounds__(1024) k_l1_final(const double* __restrict__ partial, int n, double* __restrict__ out) {
  __shared__ double s_red[32];
  double acc = (threadIdx.x < n) ? partial[threadIdx.x] : 0.0;
  acc = block_sum(acc, s_red);
  if (threadIdx.x == 0) out[0] = acc;
}
__global__ void __launch_bounds__(256) k_scale_by_inv(float* __restrict__ w, int64_t n, const double* __restrict__ norm,
                                                     double eps) {
  // F.normalize: w / max(||w||_1, eps) with the denominator in the weights' dtype
  const float den = (float)fmax(norm[0], eps);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n >> 2;
  float4* __restrict__ w4 = reinterpret_cast<float4*>(w);
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n4; v += stride) {
    float4 t = w4[v];
    t.x = __fdiv_rn(t.x, den); t.y = __fdiv_rn(t.y, den); t.z = __fdiv_rn(t.z, den); t.w = __fdiv_rn(t.w, den);
    w4[v] = t;
  }
  if (blockIdx.x == 0)
    for (int64_t k = (n4 << 2) + threadIdx.x; k < n; k += blockDim.x) w[k] = __fdiv_rn(w[k], den);
}

}  // namespace bliss

using namespace bliss;

static inline int grid_for(int64_t n, int threads, int max_blocks) {
  int64_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

extern "C" {

int bliss_gat_alpha_sums(const int32_t* blk_indptr, const float* a_ij, const float* q_ij, int32_t n_dst,
                         float* asum, float* qsum, void* stream) {
  if (!blk_indptr || !a_ij || !q_ij || !asum || !qsum || n_dst < 0) return -1;
  if (n_dst == 0) return 0;
  k_row_sums2<<<grid_for((int64_t)n_dst * 32, 256, BLISS_SM_COUNT * 8), 256, 0, (cudaStream_t)stream>>>(
      blk_indptr, a_ij, q_ij, n_dst, asum, qsum);
  BLISS_CHECK_LAUNCH();
  return 0;
}

int bliss_reward_update(const bliss_graph* g, const int32_t* blk_indptr, const int32_t* edge_src,
                        const int32_t* edge_dst, const int64_t* csc_pos, const int32_t* dst_nid,
                        const float* q_ij, const float* node_prob, const float* embed_norm,
                        const float* w_static_csc, const float* a_ij, const float* asum, const float* qsum,
                        int32_t alpha_mode, float delta, int32_t n_dst, int64_t n_edges, float* exp3_w_csc,
                        float* rewards, float* x_out, double* l1_delta, const int64_t* n_edges_dev,
                        int64_t* count_out, int32_t* pos_out, const bliss_p2p* p2p, float* wmax, void* stream) {
  if (!g || n_edges < 0 || n_dst < 0) return -1;
  if (p2p && (p2p->world <= 0 || p2p->world > 32 || !p2p->peer_base || !p2p->step_dev || !p2p->done_ctr ||
              p2p->rank < 0 || p2p->rank >= p2p->world || p2p->layer < 0 || p2p->layer >= p2p->n_layers))
    return -1;
  if (n_edges == 0) return 0;
  if (!blk_indptr || !edge_src || !edge_dst || !csc_pos || !dst_nid || !q_ij || !node_prob || !embed_norm) return -1;
  if (alpha_mode == 0 && !w_static_csc) return -1;
  if (alpha_mode == 1 && (!a_ij || !asum || !qsum)) return -1;
  RewardArgs p;
  p.g_indptr = g->indptr;
  p.blk_indptr = blk_indptr;
  p.edge_src = edge_src;
  p.edge_dst = edge_dst;
  p.csc_pos = csc_pos;
  p.dst_nid = dst_nid;
  p.q_ij = q_ij;
  p.node_prob = node_prob;
  p.embed_norm = embed_norm;
  p.w_static = w_static_csc;
  p.a_ij = a_ij;
  p.asum = asum;
  p.qsum = qsum;
  p.alpha_mode = alpha_mode;
  p.delta = delta;
  p.n_edges = n_edges;
  p.n_edges_dev = n_edges_dev;
  p.count_out = count_out;
  p.pos_out = pos_out;
  p.exp3_w = exp3_w_csc;
  p.rewards = rewards;
  p.x_out = x_out;
  p.l1_delta = l1_delta;
  p.wmax = wmax;
  if (p2p) {
    p.p2p = *p2p;
  } else {
    memset(&p.p2p, 0, sizeof(p.p2p));
  }
  BLISS_KSCOPE("k_reward_update", stream);
  k_reward_update<<<grid_for(n_edges, 256, BLISS_SM_COUNT * 8), 256, 0, (cudaStream_t)stream>>>(p);
  BLISS_CHECK_LAUNCH();
  return 0;
}

int bliss_apply_updates(const int64_t* pos, const float* x, int64_t n, float* exp3_w_csc, double* l1_delta,
                        float* wmax, void* stream) {
  if (n < 0 || !exp3_w_csc) return -1;
  if (n == 0) return 0;
  if (!pos || !x) return -1;
  k_apply_updates<<<grid_for(n, 256, BLISS_SM_COUNT * 8), 256, 0, (cudaStream_t)stream>>>(pos, x, n, exp3_w_csc, l1_delta, wmax);
  BLISS_CHECK_LAUNCH();
  return 0;
}

int bliss_apply_updates_packed(const void* recv, int64_t rank_stride_bytes, int32_t world, int64_t count_off,
                               int64_t pos_off, int64_t x_off, int64_t cap, float* exp3_w_csc, double* l1_delta,
                               float* wmax, void* stream) {
  if (!recv || !exp3_w_csc || world <= 0 || cap < 0 || rank_stride_bytes <= 0) return -1;
  if ((count_off & 7) || (pos_off & 3) || (x_off & 3)) return -1;
  if (cap == 0) return 0;
  BLISS_KSCOPE("k_apply_updates_packed", stream);
  k_apply_updates_packed<<<grid_for((int64_t)world * cap, 256, BLISS_SM_COUNT * 8), 256, 0, (cudaStream_t)stream>>>(
      (const unsigned char*)recv, rank_stride_bytes, world, count_off, pos_off, x_off, cap, exp3_w_csc, l1_delta, wmax);
  BLISS_CHECK_LAUNCH();
  return 0;
}

int bliss_apply_updates_p2p(const bliss_p2p* p2p, int64_t cap, float* exp3_w_csc, double* l1_delta, int32_t* error,
                            float* wmax, void* stream) {
  if (!p2p || !exp3_w_csc || cap < 0 || p2p->world <= 0 || p2p->world > 32 || !p2p->peer_base || !p2p->step_dev) return -1;
  if (cap == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  {
    BLISS_KSCOPE("k_p2p_wait", st);
    k_p2p_wait<<<1, 32, 0, st>>>(*p2p, error);
    BLISS_CHECK_LAUNCH();
  }
  BLISS_KSCOPE("k_apply_updates_p2p", st);
  k_apply_updates_p2p<<<grid_for((int64_t)p2p->world * cap, 256, BLISS_SM_COUNT * 4), 256, 0, st>>>(*p2p, cap, exp3_w_csc,
                                                                                                 l1_delta, wmax);
  BLISS_CHECK_LAUNCH();
  return 0;
}

int bliss_l1_norm(const float* w, int64_t n, double* partial, double* out, void* stream) {
  if (!w || n < 0 || !partial || !out) return -1;
  if ((uintptr_t)w % 16 != 0) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  k_l1_partial<<<BLISS_NORM_BLOCKS, 256, 0, st>>>(w, n, partial);
  BLISS_CHECK_LAUNCH();
  k_l1_final<<<1, 1024, 0, st>>>(partial, BLISS_NORM_BLOCKS, out);
  BLISS_CHECK_LAUNCH();
  return 0;
}

int bliss_scale_by_inv(float* w, int64_t n, const double* norm, double eps, void* stream) {
  if (!w || n < 0 || !norm) return -1;
  if ((uintptr_t)w % 16 != 0) return -1;
  if (n == 0) return 0;
  k_scale_by_inv<<<grid_for(n / 4 + 1, 256, BLISS_SM_COUNT * 16), 256, 0, (cudaStream_t)stream>>>(w, n, norm, eps);
  BLISS_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
