// aggregate.cu — sparse aggregation of the BLISS hot path on sm_100a: input-feature gather with
// fused row norm, embed_norm, and the SpMM used by SAGE / GCN forward and (on the transposed
// block) backward.  Replaces DGL g-SpMM u_mul_e/sum + fn.mean (model.py:321-329,428-436 through
// dglnn.SAGEConv / GraphConv), th.norm (model.py:318,425) and the lazy DGL frame gather
// (train_lightning.py:138).  All of it is HBM/L2-bound gather work: warp per row (or per 32-edge
// row segment), 128-bit loads along the feature dimension, edge metadata loaded coalesced and
// broadcast by shuffle.
#include "common.cuh"
#include "profile.cuh"

namespace bliss {

template <int VEC>
struct VecT;
template <>
struct VecT<1> { using T = float; };
template <>
struct VecT<2> { using T = float2; };
template <>
struct VecT<4> { using T = float4; };

template <int VEC>
__device__ __forceinline__ void vload(float* r, const float* p) {
  using T = typename VecT<VEC>::T;
  T v = __ldg(reinterpret_cast<const T*>(p));
  const float* f = reinterpret_cast<const float*>(&v);
#pragma unroll
  for (int i = 0; i < VEC; ++i) r[i] = f[i];
}
template <int VEC>
__device__ __forceinline__ void vstore(float* p, const float* r) {
  using T = typename VecT<VEC>::T;
  T v;
  float* f = reinterpret_cast<float*>(&v);
#pragma unroll
  for (int i = 0; i < VEC; ++i) f[i] = r[i];
  *reinterpret_cast<T*>(p) = v;
}

// out[i,:] = table[nid[i],:]; optionally row_norm[i] = ||out[i,:]||_2
template <int VEC>
__global__ void __launch_bounds__(256) k_gather_rows(const float* __restrict__ table, const int32_t* __restrict__ nid,
                                                    int64_t n_rows, int dim, float* __restrict__ out,
                                                    float* __restrict__ row_norm) {
  const int lane = lane_id();
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rows; r += nwarps) {
    const float* __restrict__ src = table + (int64_t)(nid ? nid[r] : r) * dim;
    float* __restrict__ dst = out ? out + r * dim : nullptr;
    float ss = 0.0f;
#pragma unroll 4
    for (int c = lane * VEC; c < dim; c += 32 * VEC) {
      float v[VEC];
      vload<VEC>(v, src + c);
      if (dst) vstore<VEC>(dst + c, v);
#pragma unroll
      for (int i = 0; i < VEC; ++i) ss += v[i] * v[i];
    }
    if (row_norm) {
      ss = warp_sum(ss);
      if (lane == 0) row_norm[r] = sqrtf(ss);
    }
  }
}

// y[r, c0:c0+W] = dscale_r * Σ_{e in row r} w_e * sscale[col_e] * x[col_e, c0:c0+W]
// One warp per (row or row segment, column tile), accumulators NCH*VEC floats per lane; the edge
// metadata of 32 edges is loaded coalesced and broadcast by shuffle.
// Up to 32 edges [e0, min(b, e0+32)): metadata loaded coalesced, then the source rows are gathered
// MLP edges at a time with every load issued before the first use (a warp's gathers are a chain of
// L2 round trips otherwise: ~1 us per 4 edges).
template <int VEC, int NCH, int MLP_OVERRIDE = 0>
__device__ __forceinline__ void spmm_chunk(float (&acc)[NCH][VEC], int e0, int b, int lane, int c0, int dim,
                                           const int32_t* __restrict__ col, const int32_t* __restrict__ perm,
                                           const float* __restrict__ w, const float* __restrict__ sscale,
                                           const float* __restrict__ x) {
  constexpr int MLP = MLP_OVERRIDE ? MLP_OVERRIDE : ((NCH * VEC <= 8) ? 4 : (NCH * VEC <= 16 ? 2 : 1));
  const int e = e0 + lane;
  int my_c = 0;
  float my_w = 0.0f;
  if (e < b) {
    my_c = __ldg(col + e);
    my_w = w ? __ldg(w + (perm ? __ldg(perm + e) : e)) : 1.0f;
    if (sscale) my_w *= __ldg(sscale + my_c);
  }
  const int n = min(32, b - e0);
  for (int j0 = 0; j0 < n; j0 += MLP) {
    float v[MLP][NCH][VEC];
    float ww[MLP];
#pragma unroll
    for (int u = 0; u < MLP; ++u) {
      // lanes beyond the chunk hold column 0 / weight 0: a harmless (cached) load that adds nothing
      const int c = __shfl_sync(0xffffffffu, my_c, (j0 + u) & 31);
      ww[u] = (j0 + u < n) ? __shfl_sync(0xffffffffu, my_w, (j0 + u) & 31) : 0.0f;
      const float* __restrict__ xr = x + (int64_t)c * dim + c0;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int cc = (ch * 32 + lane) * VEC;
        if (c0 + cc < dim) {
          vload<VEC>(v[u][ch], xr + cc);
        } else {
#pragma unroll
          for (int i = 0; i < VEC; ++i) v[u][ch][i] = 0.0f;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < MLP; ++u)
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[ch][i] = fmaf(ww[u], v[u][ch][i], acc[ch][i]);
  }
}

template <int VEC, int NCH>
__global__ void __launch_bounds__(256) k_spmm(const int32_t* __restrict__ indptr, const int32_t* __restrict__ col,
                                             const int32_t* __restrict__ perm, const float* __restrict__ w,
                                             const float* __restrict__ sscale, const float* __restrict__ dscale,
                                             int agg, const float* __restrict__ x, int n_rows, int dim,
                                             float* __restrict__ y) {
  constexpr int TILE = 32 * VEC * NCH;
  const int lane = lane_id();
  const int c0 = blockIdx.y * TILE;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < n_rows; r += nwarps) {
    const int a = indptr[r], b = indptr[r + 1];
    float acc[NCH][VEC];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[ch][i] = 0.0f;
    for (int e0 = a; e0 < b; e0 += 32) spmm_chunk<VEC, NCH>(acc, e0, b, lane, c0, dim, col, perm, w, sscale, x);
    float s = dscale ? dscale[r] : 1.0f;
    if (agg == BLISS_AGG_MEAN) s = s / (float)max(b - a, 1);
    float* __restrict__ yr = y + (int64_t)r * dim + c0;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int cc = (ch * 32 + lane) * VEC;
      if (c0 + cc < dim) {
        float v[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = acc[ch][i] * s;
        vstore<VEC>(yr + cc, v);
      }
    }
  }
}

// Balanced variant for sampled blocks: every row is cut into segments of BLISS_SPMM_SEG (32) edges
// (seg_ptr = prefix of max(1, ceil(len / 32)) over the rows, built with the block).  A CTA takes 8
// consecutive segments, one per warp, whatever rows they belong to; the partial sums meet in
// shared memory and each run of segments of one row is added up in segment order by the run's
// first warp.  A row that lies inside one group is finished there.  A hub row (thousands of edges:
// a popular destination is linked to most sampled sources) spans many groups, so it no longer
// streams megabytes through one SM: every group stores ONE partial per run (slot = the run's
// first segment), and k_spmm_combine adds a row's partials in group order — a fixed order, so the
// result is deterministic without atomics.
template <int VEC, int NCH>
__global__ void __launch_bounds__(256, 5) k_spmm_seg(const int32_t* __restrict__ indptr, const int32_t* __restrict__ seg_ptr,
                                                 const int32_t* __restrict__ col, const int32_t* __restrict__ perm,
                                                 const float* __restrict__ w, const float* __restrict__ sscale,
                                                 const float* __restrict__ dscale, int agg, const float* __restrict__ x,
                                                 int n_rows, int dim, float* __restrict__ partial, int64_t item_cap,
                                                 float* __restrict__ y) {
  constexpr int TILE = 32 * VEC * NCH;
  using T = typename VecT<VEC>::T;
  extern __shared__ float s_part[];   // [8 warps][TILE]
  __shared__ int s_row[8];
  const int lane = lane_id(), wid = warp_id();
  const int c0 = blockIdx.y * TILE;
  const int n_items = seg_ptr[n_rows];
  const int n_groups = (n_items + 7) >> 3;
  float* __restrict__ tile_partial = partial + (int64_t)blockIdx.y * item_cap * TILE;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int item = grp * 8 + wid;
    const bool valid = item < n_items;
    int r = -1, a = 0, b = 0, s0 = 0, nseg = 1;
    float acc[NCH][VEC];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[ch][i] = 0.0f;
    if (valid) {
      int lo = 0, hi = n_rows - 1;   // row of the item: last r with seg_ptr[r] <= item (warp-uniform search)
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(seg_ptr + mid) <= item) lo = mid; else hi = mid - 1;
      }
      r = lo;
      a = indptr[r];
      b = indptr[r + 1];
      s0 = seg_ptr[r];
      nseg = seg_ptr[r + 1] - s0;
      const int e_lo = a + (item - s0) * BLISS_SPMM_SEG;
      spmm_chunk<VEC, NCH>(acc, e_lo, min(b, e_lo + BLISS_SPMM_SEG), lane, c0, dim, col, perm, w, sscale, x);
    }
    if (lane == 0) s_row[wid] = r;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) vstore<VEC>(s_part + wid * TILE + (ch * 32 + lane) * VEC, acc[ch]);
    __syncthreads();
    if (valid && (wid == 0 || s_row[wid - 1] != r)) {   // first segment of this row's run in the group
      int len = 1;
      while (wid + len < 8 && s_row[wid + len] == r) ++len;
      for (int k = 1; k < len; ++k) {                    // run total, segment order
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const T v = *reinterpret_cast<const T*>(s_part + (wid + k) * TILE + (ch * 32 + lane) * VEC);   // shared memory
          const float* f = reinterpret_cast<const float*>(&v);
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[ch][i] += f[i];
        }
      }
      if (item == s0 && len == nseg) {                   // the whole row lies in this group
        float s = dscale ? dscale[r] : 1.0f;
        if (agg == BLISS_AGG_MEAN) s = s / (float)max(b - a, 1);
        float* __restrict__ yr = y + (int64_t)r * dim + c0;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const int cc = (ch * 32 + lane) * VEC;
          if (c0 + cc < dim) {
            float v[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) v[i] = acc[ch][i] * s;
            vstore<VEC>(yr + cc, v);
          }
        }
      } else {
        float* __restrict__ pr = tile_partial + (int64_t)item * TILE;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) vstore<VEC>(pr + (ch * 32 + lane) * VEC, acc[ch]);
      }
    }
    __syncthreads();
  }
}

// ---- independent-warp variant of the balanced SpMM ---------------------------------------------------------
// Every warp walks the 32-edge segments round-robin on its own: no shared memory, no CTA barrier (in k_spmm_seg the
// eight warps of a group wait for the slowest segment of every round), fewer registers, so more warps are resident
// and each of them keeps its gathers in flight independently.  A row with a single segment is finished by its warp;
// every segment of a longer row stores its partial sum (slot = segment index) and k_spmm_combine_items adds a row's
// partials in segment order — a fixed order, deterministic, no atomics.
template <int VEC, int NCH, int LB>
__global__ void __launch_bounds__(256, LB) k_spmm_item(const int32_t* __restrict__ indptr, const int32_t* __restrict__ seg_ptr,
                                                  const int32_t* __restrict__ col, const int32_t* __restrict__ perm,
                                                  const float* __restrict__ w, const float* __restrict__ sscale,
                                                  const float* __restrict__ dscale, int agg, const float* __restrict__ x,
                                                  int n_rows, int dim, float* __restrict__ partial, int64_t item_cap,
                                                  float* __restrict__ y) {
  constexpr int TILE = 32 * VEC * NCH;
  const int lane = lane_id();
  const int c0 = blockIdx.y * TILE;
  const int n_items = seg_ptr[n_rows];
  float* __restrict__ tile_partial = partial + (int64_t)blockIdx.y * item_cap * TILE;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int item = gwarp; item < n_items; item += nwarps) {
    int lo = 0, hi = n_rows - 1;   // row of the item: last r with seg_ptr[r] <= item (warp-uniform search, L1 resident)
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (__ldg(seg_ptr + mid) <= item) lo = mid; else hi = mid - 1;
    }
    const int r = lo;
    const int a = __ldg(indptr + r), b = __ldg(indptr + r + 1);
    const int s0 = __ldg(seg_ptr + r), nseg = __ldg(seg_ptr + r + 1) - s0;
    float acc[NCH][VEC];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[ch][i] = 0.0f;
    const int e_lo = a + (item - s0) * BLISS_SPMM_SEG;
    spmm_chunk<VEC, NCH, (LB == 3 && NCH * VEC <= 8) ? 8 : 0>(acc, e_lo, min(b, e_lo + BLISS_SPMM_SEG), lane, c0, dim, col, perm,
                                                               w, sscale, x);
    if (nseg == 1) {
      float sc = dscale ? __ldg(dscale + r) : 1.0f;
      if (agg == BLISS_AGG_MEAN) sc = sc / (float)max(b - a, 1);
      float* __restrict__ yr = y + (int64_t)r * dim + c0;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int cc = (ch * 32 + lane) * VEC;
        if (c0 + cc < dim) {
          float v[VEC];
#pragma unroll
          for (int i = 0; i < VEC; ++i) v[i] = acc[ch][i] * sc;
          vstore<VEC>(yr + cc, v);
        }
      }
    } else {
      float* __restrict__ pr = tile_partial + (int64_t)item * TILE;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) vstore<VEC>(pr + (ch * 32 + lane) * VEC, acc[ch]);
    }
  }
}

// rows with more than one segment: add the segments' partials in segment order.  One warp per row.
template <int VEC, int NCH>
__global__ void __launch_bounds__(256) k_spmm_combine_items(const int32_t* __restrict__ indptr,
                                                           const int32_t* __restrict__ seg_ptr,
                                                           const float* __restrict__ dscale, int agg, int n_rows, int dim,
                                                           const float* __restrict__ partial, int64_t item_cap,
                                                           float* __restrict__ y) {
  constexpr int TILE = 32 * VEC * NCH;
  using T = typename VecT<VEC>::T;
  const int lane = lane_id();
  const int c0 = blockIdx.y * TILE;
  const float* __restrict__ tile_partial = partial + (int64_t)blockIdx.y * item_cap * TILE;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < n_rows; r += nwarps) {
    const int s0 = seg_ptr[r], s1 = seg_ptr[r + 1];
    if (s1 - s0 <= 1) continue;   // finished by k_spmm_item
    float acc[NCH][VEC];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[ch][i] = 0.0f;
    constexpr int CU = (NCH * VEC <= 8) ? 8 : (NCH * VEC <= 16 ? 4 : 2);   // partials in flight
    for (int k0 = s0; k0 < s1; k0 += CU) {
      T pv[CU][NCH];
#pragma unroll
      for (int u = 0; u < CU; ++u)
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
          if (k0 + u < s1)
            pv[u][ch] = __ldg(reinterpret_cast<const T*>(tile_partial + (int64_t)(k0 + u) * TILE + (ch * 32 + lane) * VEC));
#pragma unroll
      for (int u = 0; u < CU; ++u) {
        if (k0 + u < s1) {
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) {
            const float* f = reinterpret_cast<const float*>(&pv[u][ch]);
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[ch][i] += f[i];
          }
        }
      }
    }
    const int a = indptr[r], b = indptr[r + 1];
    float sc = dscale ? dscale[r] : 1.0f;
    if (agg == BLISS_AGG_MEAN) sc = sc / (float)max(b - a, 1);
    float* __restrict__ yr = y + (int64_t)r * dim + c0;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int cc = (ch * 32 + lane) * VEC;
      if (c0 + cc < dim) {
        float v[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = acc[ch][i] * sc;
        vstore<VEC>(yr + cc, v);
      }
    }
  }
}

// ---- bulk-copy (TMA) variant of the balanced SpMM -------------------------------------------------------
// Same decomposition as k_spmm_seg (32-edge segments, 8 consecutive segments per CTA round, runs of one row added in
// shared memory, partials + k_spmm_combine for rows that span several rounds), but the source rows are not gathered
// through registers: every warp owns a ring of R row slots in shared memory and fetches each source row with ONE
// cp.async.bulk (1 KB at D = 256) that completes on the slot's mbarrier.  The bytes in flight per SM are then bounded
// by shared memory (3 CTAs x 8 warps x 8 slots = 192 rows), not by the registers a warp can spend on loads, the load
// path issues one instruction per row instead of 2 x 32 LDG.128, and the accumulators are all the registers the
// kernel needs.  Rows must be 16-byte multiples and 16-byte aligned (dim % 4 == 0); dim <= 512.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// arm the slot's barrier with the row's byte count and start the bulk copy global -> shared that completes on it
__device__ __forceinline__ void bulk_row_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  const uint32_t b = smem_u32(bar);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the slot's earlier generic-proxy reads precede the async write
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(b)
               : "memory");
}

template <int NCH4>   // float4 column chunks per lane: dim <= 128 * NCH4
__global__ void __launch_bounds__(256, 3) k_spmm_tma(const int32_t* __restrict__ indptr, const int32_t* __restrict__ seg_ptr,
                                                   const int32_t* __restrict__ col, const int32_t* __restrict__ perm,
                                                   const float* __restrict__ w, const float* __restrict__ sscale,
                                                   const float* __restrict__ dscale, int agg, const float* __restrict__ x,
                                                   int n_rows, int dim, float* __restrict__ partial, int64_t item_cap,
                                                   float* __restrict__ y, int R) {
  constexpr int TILE = 128 * NCH4;                      // row stride of the partials (== k_spmm_combine<4, NCH4>)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* ring = reinterpret_cast<float*>(smem_raw);     // [8 warps][R slots][dim]
  float* s_part = ring + (size_t)8 * R * dim;           // [8 warps][TILE]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_part + 8 * TILE);   // [8 warps][R]
  __shared__ int s_row[8];
  const int lane = lane_id(), wid = warp_id();
  float* my_ring = ring + (size_t)wid * R * dim;
  uint64_t* my_bars = bars + wid * R;
  const uint32_t rowbytes = (uint32_t)dim * 4u;
  if (lane < R) mbar_init(&my_bars[lane], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  uint32_t phase_bits = 0;                              // parity of every slot's next completion
  const int n_items = seg_ptr[n_rows];
  const int n_groups = (n_items + 7) >> 3;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int item = grp * 8 + wid;
    const bool valid = item < n_items;
    int r = -1, a = 0, b = 0, s0 = 0, nseg = 1;
    float4 acc[NCH4];
#pragma unroll
    for (int ch = 0; ch < NCH4; ++ch) acc[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
      int lo = 0, hi = n_rows - 1;   // row of the item: last r with seg_ptr[r] <= item (warp-uniform search)
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(seg_ptr + mid) <= item) lo = mid; else hi = mid - 1;
      }
      r = lo;
      a = indptr[r];
      b = indptr[r + 1];
      s0 = seg_ptr[r];
      nseg = seg_ptr[r + 1] - s0;
      const int e_lo = a + (item - s0) * BLISS_SPMM_SEG;
      const int n = min(BLISS_SPMM_SEG, b - e_lo);      // edges of this segment (may be 0 for an empty row)
      int my_c = 0;
      float my_w = 0.0f;
      if (lane < n) {
        const int e = e_lo + lane;
        my_c = __ldg(col + e);
        my_w = w ? __ldg(w + (perm ? __ldg(perm + e) : e)) : 1.0f;
        if (sscale) my_w *= __ldg(sscale + my_c);
        if (lane < R) bulk_row_g2s(my_ring + (size_t)lane * dim, x + (int64_t)my_c * dim, rowbytes, &my_bars[lane]);
      }
      for (int j = 0; j < n; ++j) {
        const int slot = j & (R - 1);
        mbar_wait(&my_bars[slot], (phase_bits >> slot) & 1u);
        phase_bits ^= 1u << slot;
        const float wj = __shfl_sync(0xffffffffu, my_w, j);
        const float* __restrict__ row = my_ring + (size_t)slot * dim;
#pragma unroll
        for (int ch = 0; ch < NCH4; ++ch) {
          const int c = ch * 128 + lane * 4;
          if (c < dim) {
            const float4 v = *reinterpret_cast<const float4*>(row + c);
            acc[ch].x = fmaf(wj, v.x, acc[ch].x);
            acc[ch].y = fmaf(wj, v.y, acc[ch].y);
            acc[ch].z = fmaf(wj, v.z, acc[ch].z);
            acc[ch].w = fmaf(wj, v.w, acc[ch].w);
          }
        }
        __syncwarp();                                   // every lane has read the slot: it may be refilled
        if (lane == j + R && lane < n)
          bulk_row_g2s(my_ring + (size_t)slot * dim, x + (int64_t)my_c * dim, rowbytes, &my_bars[slot]);
      }
    }
    if (lane == 0) s_row[wid] = r;
#pragma unroll
    for (int ch = 0; ch < NCH4; ++ch) *reinterpret_cast<float4*>(s_part + wid * TILE + ch * 128 + lane * 4) = acc[ch];
    __syncthreads();
    if (valid && (wid == 0 || s_row[wid - 1] != r)) {   // first segment of this row's run in the group
      int len = 1;
      while (wid + len < 8 && s_row[wid + len] == r) ++len;
      for (int k = 1; k < len; ++k) {                    // run total, segment order
#pragma unroll
        for (int ch = 0; ch < NCH4; ++ch) {
          const float4 v = *reinterpret_cast<const float4*>(s_part + (wid + k) * TILE + ch * 128 + lane * 4);
          acc[ch].x += v.x;
          acc[ch].y += v.y;
          acc[ch].z += v.z;
          acc[ch].w += v.w;
        }
      }
      if (item == s0 && len == nseg) {                   // the whole row lies in this group
        float sc = dscale ? dscale[r] : 1.0f;
        if (agg == BLISS_AGG_MEAN) sc = sc / (float)max(b - a, 1);
        float* __restrict__ yr = y + (int64_t)r * dim;
#pragma unroll
        for (int ch = 0; ch < NCH4; ++ch) {
          const int c = ch * 128 + lane * 4;
          if (c < dim)
            *reinterpret_cast<float4*>(yr + c) = make_float4(acc[ch].x * sc, acc[ch].y * sc, acc[ch].z * sc, acc[ch].w * sc);
        }
      } else {
        float* __restrict__ pr = partial + (int64_t)item * TILE;
#pragma unroll
        for (int ch = 0; ch < NCH4; ++ch) *reinterpret_cast<float4*>(pr + ch * 128 + lane * 4) = acc[ch];
      }
    }
    __syncthreads();
  }
}

// Rows that span several groups of k_spmm_seg: add the runs' partials in group order.  Runs start
// at the row's first segment s0 and at every multiple of 8 after it.  One warp per row.
template <int VEC, int NCH>
__global__ void __launch_bounds__(256) k_spmm_combine(const int32_t* __restrict__ indptr, const int32_t* __restrict__ seg_ptr,
                                                     const float* __restrict__ dscale, int agg, int n_rows, int dim,
                                                     const float* __restrict__ partial, int64_t item_cap,
                                                     float* __restrict__ y) {
  constexpr int TILE = 32 * VEC * NCH;
  using T = typename VecT<VEC>::T;
  const int lane = lane_id();
  const int c0 = blockIdx.y * TILE;
  const float* __restrict__ tile_partial = partial + (int64_t)blockIdx.y * item_cap * TILE;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < n_rows; r += nwarps) {
    const int s0 = seg_ptr[r], s1 = seg_ptr[r + 1];
    if ((s0 >> 3) == ((s1 - 1) >> 3)) continue;   // the row lies in one group: finished by k_spmm_seg
    const int first_full = ((s0 >> 3) + 1) << 3;
    const int n_runs = 1 + ((s1 - 1) >> 3) - (s0 >> 3);
    float acc[NCH][VEC];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[ch][i] = 0.0f;
    constexpr int CU = (NCH * VEC <= 8) ? 8 : (NCH * VEC <= 16 ? 4 : 2);   // partials in flight
    for (int k0 = 0; k0 < n_runs; k0 += CU) {
      T pv[CU][NCH];
#pragma unroll
      for (int u = 0; u < CU; ++u) {
        const int k = k0 + u;
        const int64_t slot = (k == 0) ? s0 : first_full + 8 * (k - 1);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
          if (k < n_runs) pv[u][ch] = __ldg(reinterpret_cast<const T*>(tile_partial + slot * TILE + (ch * 32 + lane) * VEC));
      }
#pragma unroll
      for (int u = 0; u < CU; ++u) {
        if (k0 + u < n_runs) {
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) {
            const float* f = reinterpret_cast<const float*>(&pv[u][ch]);
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[ch][i] += f[i];
          }
        }
      }
    }
    const int a = indptr[r], b = indptr[r + 1];
    float s = dscale ? dscale[r] : 1.0f;
    if (agg == BLISS_AGG_MEAN) s = s / (float)max(b - a, 1);
    float* __restrict__ yr = y + (int64_t)r * dim + c0;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int cc = (ch * 32 + lane) * VEC;
      if (c0 + cc < dim) {
        float v[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = acc[ch][i] * s;
        vstore<VEC>(yr + cc, v);
      }
    }
  }
}

// L2 -> SM gather probe (bench.py): what a warp-per-row gather of 1 KB rows can pull out of L2 on this chip.
__global__ void __launch_bounds__(256) k_l2_gather_probe(const float* __restrict__ table, int n_rows, int dim,
                                                        int rows_per_warp, float* __restrict__ out, int n_warps) {
  const int lane = lane_id();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= n_warps) return;
  unsigned state = 0x9E3779B9u * (unsigned)(warp + 1);
  for (int c0 = 0; c0 < dim; c0 += 256) {     // 2 x float4 per lane per row and 256-column slice
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    for (int r0 = 0; r0 < rows_per_warp; r0 += 4) {
      float4 v[4][2];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        state = state * 1664525u + 1013904223u;
        const int row = (int)((state >> 8) % (unsigned)n_rows);
        const float4* __restrict__ p = reinterpret_cast<const float4*>(table + (int64_t)row * dim + c0);
        v[u][0] = __ldg(p + lane);
        v[u][1] = (c0 + 128 < dim) ? __ldg(p + 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a0.x += v[u][0].x; a0.y += v[u][0].y; a0.z += v[u][0].z; a0.w += v[u][0].w;
        a1.x += v[u][1].x; a1.y += v[u][1].y; a1.z += v[u][1].z; a1.w += v[u][1].w;
      }
    }
    float4* __restrict__ o = reinterpret_cast<float4*>(out + (int64_t)warp * dim + c0);
    o[lane] = a0;
    if (c0 + 128 < dim) o[32 + lane] = a1;
  }
}

}  // namespace bliss

using namespace bliss;

static inline int blocks_for_rows(int64_t n_rows, int max_blocks) {
  int64_t b = (n_rows + 7) / 8;  // 8 warps per 256-thread CTA
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

template <int VEC>
static int launch_spmm(const int32_t* indptr, const int32_t* col, const int32_t* perm, const float* w,
                       const float* sscale, const float* dscale, int agg, const float* x, int n_rows,
                       int dim, const int32_t* seg_ptr, float* partial, int64_t item_cap, float* y,
                       cudaStream_t st) {
  const int per_lane = (dim + 32 * VEC - 1) / (32 * VEC);
  int nch = 1;
  while (nch < per_lane && nch < 8) nch <<= 1;
  const int tile = 32 * VEC * nch;
  const int64_t units = seg_ptr ? item_cap : n_rows;   // one warp per unit
  dim3 grid(blocks_for_rows(units, BLISS_SM_COUNT * 8), (dim + tile - 1) / tile);
  dim3 grid_c(blocks_for_rows(n_rows, BLISS_SM_COUNT * 8), (dim + tile - 1) / tile);
  const size_t smem = seg_ptr ? (size_t)8 * tile * sizeof(float) : 0;
  static int item_mode = -1, item_lb = 4;   // BLISS_SPMM_MODE=group: the 8-segment CTA groups of k_spmm_seg; default: independent warps
  if (item_mode < 0) {
    const char* e = getenv("BLISS_SPMM_MODE");
    item_mode = (e && e[0] == 'g') ? 0 : 1;
    const char* l = getenv("BLISS_SPMM_LB");      // resident CTAs per SM the register budget is set for (4: no spills)
    if (l && l[0] == '5') item_lb = 5;
    if (l && l[0] == '3') item_lb = 3;      // 85 registers: 8 rows in flight per warp
  }
  const int64_t cta_cap = (int64_t)BLISS_SM_COUNT * item_lb;
  dim3 grid_i((unsigned)(cta_cap < (units + 7) / 8 ? cta_cap : ((units + 7) / 8 > 0 ? (units + 7) / 8 : 1)),
              (dim + tile - 1) / tile);
#define BLISS_SPMM_CASE(N)                                                                                    \
  case N:                                                                                                     \
    if (seg_ptr && item_mode) {                                                                               \
      {                                                                                                       \
        BLISS_KSCOPE("k_spmm_item", st);                                                                      \
        if (item_lb == 5)                                                                                     \
          k_spmm_item<VEC, N, 5><<<grid_i, 256, 0, st>>>(indptr, seg_ptr, col, perm, w, sscale, dscale, agg,  \
                                                         x, n_rows, dim, partial, item_cap, y);               \
        else if (item_lb == 3)                                                                                \
          k_spmm_item<VEC, N, 3><<<grid_i, 256, 0, st>>>(indptr, seg_ptr, col, perm, w, sscale, dscale, agg,  \
                                                         x, n_rows, dim, partial, item_cap, y);               \
        else                                                                                                  \
          k_spmm_item<VEC, N, 4><<<grid_i, 256, 0, st>>>(indptr, seg_ptr, col, perm, w, sscale, dscale, agg,  \
                                                         x, n_rows, dim, partial, item_cap, y);               \
      }                                                                                                       \
      BLISS_KSCOPE("k_spmm_combine", st);                                                                     \
      k_spmm_combine_items<VEC, N><<<grid_c, 256, 0, st>>>(indptr, seg_ptr, dscale, agg, n_rows, dim,        \
                                                           partial, item_cap, y);                             \
    } else if (seg_ptr) {                                                                                     \
      {                                                                                                       \
        BLISS_KSCOPE("k_spmm_seg", st);                                                                       \
        k_spmm_seg<VEC, N><<<grid, 256, smem, st>>>(indptr, seg_ptr, col, perm, w, sscale, dscale, agg, x,   \
                                                    n_rows, dim, partial, item_cap, y);                      \
      }                                                                                                       \
      BLISS_KSCOPE("k_spmm_combine", st);                                                                     \
      k_spmm_combine<VEC, N><<<grid_c, 256, 0, st>>>(indptr, seg_ptr, dscale, agg, n_rows, dim, partial,     \
                                                     item_cap, y);                                            \
    } else {                                                                                                  \
      BLISS_KSCOPE("k_spmm", st);                                                                             \
      k_spmm<VEC, N><<<grid, 256, 0, st>>>(indptr, col, perm, w, sscale, dscale, agg, x, n_rows, dim, y);    \
    }                                                                                                         \
    break;
  switch (nch) {
    BLISS_SPMM_CASE(1)
    BLISS_SPMM_CASE(2)
    BLISS_SPMM_CASE(4)
    BLISS_SPMM_CASE(8)
  }
#undef BLISS_SPMM_CASE
  BLISS_CHECK_LAUNCH();
  return 0;
}

static bool spmm_tma_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("BLISS_SPMM_TMA");     // opt-in: measured 2.5x slower than the LDG gather (profiles/r2_spmm_variants.md)
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on == 1;
}

// bulk-copy SpMM over 32-edge segments (dim % 4 == 0, dim <= 512, 16-byte aligned operands); combine pass as before
template <int NCH4>
static int launch_spmm_tma(const int32_t* indptr, const int32_t* col, const int32_t* perm, const float* w,
                           const float* sscale, const float* dscale, int agg, const float* x, int n_rows, int dim,
                           const int32_t* seg_ptr, float* partial, int64_t item_cap, float* y, cudaStream_t st) {
  constexpr int TILE = 128 * NCH4;
  const int R = (dim <= 256) ? 8 : 4;
  const size_t smem = (size_t)8 * R * dim * 4 + (size_t)8 * TILE * 4 + (size_t)8 * R * 8;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_spmm_tma<NCH4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int64_t groups = (item_cap + 7) / 8;
  int blocks = (int)(groups < BLISS_SM_COUNT * 3 ? (groups > 0 ? groups : 1) : BLISS_SM_COUNT * 3);
  {
    BLISS_KSCOPE("k_spmm_tma", st);
    k_spmm_tma<NCH4><<<blocks, 256, smem, st>>>(indptr, seg_ptr, col, perm, w, sscale, dscale, agg, x, n_rows, dim, partial,
                                               item_cap, y, R);
    BLISS_CHECK_LAUNCH();
  }
  dim3 grid_c(blocks_for_rows(n_rows, BLISS_SM_COUNT * 8), 1);
  BLISS_KSCOPE("k_spmm_combine", st);
  k_spmm_combine<4, NCH4><<<grid_c, 256, 0, st>>>(indptr, seg_ptr, dscale, agg, n_rows, dim, partial, item_cap, y);
  BLISS_CHECK_LAUNCH();
  return 0;
}

extern "C" {

int bliss_gather_rows(const float* table, const int32_t* nid, int64_t n_rows, int32_t dim, float* out,
                      float* row_norm, void* stream) {
  if (n_rows < 0 || dim <= 0 || !table) return -1;
  if (n_rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = blocks_for_rows(n_rows, BLISS_SM_COUNT * 16);
  const bool al16 = ((uintptr_t)table % 16 == 0) && (!out || (uintptr_t)out % 16 == 0);
  BLISS_KSCOPE("k_gather_rows", st);
  if (dim % 4 == 0 && al16)
    k_gather_rows<4><<<blocks, 256, 0, st>>>(table, nid, n_rows, dim, out, row_norm);
  else if (dim % 2 == 0)
    k_gather_rows<2><<<blocks, 256, 0, st>>>(table, nid, n_rows, dim, out, row_norm);
  else
    k_gather_rows<1><<<blocks, 256, 0, st>>>(table, nid, n_rows, dim, out, row_norm);
  BLISS_CHECK_LAUNCH();
  return 0;
}

int bliss_l2_gather_probe(const float* table, int32_t n_rows, int32_t dim, int32_t rows_per_warp, float* out,
                          int32_t n_warps, void* stream) {
  if (!table || !out || n_rows <= 0 || dim <= 0 || dim % 128 || rows_per_warp <= 0 || n_warps <= 0) return -1;
  if (((uintptr_t)table | (uintptr_t)out) % 16) return -1;
  BLISS_LAUNCH(k_l2_gather_probe, (n_warps + 7) / 8, 256, 0, (cudaStream_t)stream, table, n_rows, dim, rows_per_warp, out,
               n_warps);
  return 0;
}

int bliss_row_norm(const float* x, int64_t n_rows, int32_t dim, float* out, void* stream) {
  if (!out) return -1;
  return bliss_gather_rows(x, nullptr, n_rows, dim, nullptr, out, stream);
}

int bliss_spmm(const int32_t* indptr, const int32_t* col, const int32_t* perm, const float* w,
               const float* sscale, const float* dscale, int32_t agg, const float* x, int32_t n_rows,
               int32_t dim, const int32_t* seg_ptr, float* partial, int64_t item_cap, float* y, void* stream) {
  if (n_rows < 0 || dim <= 0 || !indptr || !y) return -1;
  if (n_rows == 0) return 0;
  if (!col || !x) return -1;
  if (seg_ptr && (!partial || item_cap <= 0)) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  const bool al16 = ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && (!seg_ptr || (uintptr_t)partial % 16 == 0);
  if (seg_ptr && dim % 4 == 0 && al16 && dim <= 512 && spmm_tma_enabled()) {
    if (dim <= 128)
      return launch_spmm_tma<1>(indptr, col, perm, w, sscale, dscale, agg, x, n_rows, dim, seg_ptr, partial, item_cap, y, st);
    if (dim <= 256)
      return launch_spmm_tma<2>(indptr, col, perm, w, sscale, dscale, agg, x, n_rows, dim, seg_ptr, partial, item_cap, y, st);
    return launch_spmm_tma<4>(indptr, col, perm, w, sscale, dscale, agg, x, n_rows, dim, seg_ptr, partial, item_cap, y, st);
  }
  if (dim % 4 == 0 && al16)
    return launch_spmm<4>(indptr, col, perm, w, sscale, dscale, agg, x, n_rows, dim, seg_ptr, partial, item_cap, y, st);
  if (dim % 2 == 0)
    return launch_spmm<2>(indptr, col, perm, w, sscale, dscale, agg, x, n_rows, dim, seg_ptr, partial, item_cap, y, st);
  return launch_spmm<1>(indptr, col, perm, w, sscale, dscale, agg, x, n_rows, dim, seg_ptr, partial, item_cap, y, st);
}

}  // extern "C"
