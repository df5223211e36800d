// aggregate.cu — sparse aggregation of the BLISS hot path on sm_100a: input-feature gather with
// fused row norm, embed_norm, and the SpMM used by SAGE / GCN forward and (on the transposed
// block) backward.  Replaces DGL g-SpMM u_mul_e/sum + fn.mean (model.py:321-329,428-436 through
// dglnn.SAGEConv / GraphConv), th.norm (model.py:318,425) and the lazy DGL frame gather
// (train_lightning.py:138).  All of it is HBM/L2-bound gather work: warp per row, 128-bit
// loads along the feature dimension, edge metadata loaded coalesced and broadcast by shuffle.
#include "common.cuh"

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
// Light rows: one warp per (row, column tile), accumulators NCH*VEC floats per lane.
// Heavy rows (> BLISS_SPMM_HEAVY edges, listed in `heavy`): one CTA per row, the 8 warps take
// interleaved 32-edge chunks and their partial sums are combined through shared memory in warp
// order — a hub row no longer serialises on one warp and the result stays deterministic.
template <int VEC, int NCH>
__device__ __forceinline__ void spmm_chunk(float (&acc)[NCH][VEC], int e0, int b, int lane, int c0, int dim,
                                           const int32_t* __restrict__ col, const int32_t* __restrict__ perm,
                                           const float* __restrict__ w, const float* __restrict__ sscale,
                                           const float* __restrict__ x) {
  const int e = e0 + lane;
  int my_c = 0;
  float my_w = 0.0f;
  if (e < b) {
    my_c = __ldg(col + e);
    my_w = w ? __ldg(w + (perm ? __ldg(perm + e) : e)) : 1.0f;
    if (sscale) my_w *= __ldg(sscale + my_c);
  }
  const int n = min(32, b - e0);
#pragma unroll 4
  for (int j = 0; j < n; ++j) {
    const int c = __shfl_sync(0xffffffffu, my_c, j);
    const float ww = __shfl_sync(0xffffffffu, my_w, j);
    const float* __restrict__ xr = x + (int64_t)c * dim + c0;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int cc = (ch * 32 + lane) * VEC;
      if (c0 + cc < dim) {
        float v[VEC];
        vload<VEC>(v, xr + cc);
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[ch][i] = fmaf(ww, v[i], acc[ch][i]);
      }
    }
  }
}

template <int VEC, int NCH>
__global__ void __launch_bounds__(256) k_spmm(const int32_t* __restrict__ indptr, const int32_t* __restrict__ col,
                                             const int32_t* __restrict__ perm, const float* __restrict__ w,
                                             const float* __restrict__ sscale, const float* __restrict__ dscale,
                                             int agg, const float* __restrict__ x, int n_rows, int dim,
                                             const int32_t* __restrict__ heavy, float* __restrict__ y) {
  extern __shared__ float s_part[];  // [8 warps][TILE] partial sums of a heavy row
  constexpr int TILE = 32 * VEC * NCH;
  const int lane = lane_id();
  const int c0 = blockIdx.y * TILE;
  const int n_heavy = heavy ? heavy[0] : 0;

  for (int h = blockIdx.x; h < n_heavy; h += gridDim.x) {
    const int r = heavy[1 + h];
    const int a = indptr[r], b = indptr[r + 1];
    float acc[NCH][VEC];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[ch][i] = 0.0f;
    for (int e0 = a + 32 * warp_id(); e0 < b; e0 += 32 * 8)
      spmm_chunk<VEC, NCH>(acc, e0, b, lane, c0, dim, col, perm, w, sscale, x);
    __syncthreads();
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int i = 0; i < VEC; ++i) s_part[warp_id() * TILE + (ch * 32 + lane) * VEC + i] = acc[ch][i];
    __syncthreads();
    float s = dscale ? dscale[r] : 1.0f;
    if (agg == BLISS_AGG_MEAN) s = s / (float)max(b - a, 1);
    for (int cc = threadIdx.x; cc < TILE && c0 + cc < dim; cc += blockDim.x) {
      float t = 0.0f;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) t += s_part[wv * TILE + cc];
      y[(int64_t)r * dim + c0 + cc] = t * s;
    }
  }

  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < n_rows; r += nwarps) {
    const int a = indptr[r], b = indptr[r + 1];
    if (heavy && b - a > BLISS_SPMM_HEAVY) continue;
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
                       int dim, const int32_t* heavy, float* y, cudaStream_t st) {
  const int per_lane = (dim + 32 * VEC - 1) / (32 * VEC);
  int nch = 1;
  while (nch < per_lane && nch < 8) nch <<= 1;
  const int tile = 32 * VEC * nch;
  dim3 grid(blocks_for_rows(n_rows, BLISS_SM_COUNT * 16), (dim + tile - 1) / tile);
  const size_t smem = heavy ? (size_t)8 * tile * sizeof(float) : 0;
#define BLISS_SPMM_CASE(N)                                                                               \
  case N:                                                                                                \
    k_spmm<VEC, N><<<grid, 256, smem, st>>>(indptr, col, perm, w, sscale, dscale, agg, x, n_rows, dim,  \
                                            heavy, y);                                                   \
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

extern "C" {

int bliss_gather_rows(const float* table, const int32_t* nid, int64_t n_rows, int32_t dim, float* out,
                      float* row_norm, void* stream) {
  if (n_rows < 0 || dim <= 0 || !table) return -1;
  if (n_rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = blocks_for_rows(n_rows, BLISS_SM_COUNT * 16);
  const bool al16 = ((uintptr_t)table % 16 == 0) && (!out || (uintptr_t)out % 16 == 0);
  if (dim % 4 == 0 && al16)
    k_gather_rows<4><<<blocks, 256, 0, st>>>(table, nid, n_rows, dim, out, row_norm);
  else if (dim % 2 == 0)
    k_gather_rows<2><<<blocks, 256, 0, st>>>(table, nid, n_rows, dim, out, row_norm);
  else
    k_gather_rows<1><<<blocks, 256, 0, st>>>(table, nid, n_rows, dim, out, row_norm);
  BLISS_CHECK_LAUNCH();
  return 0;
}

int bliss_row_norm(const float* x, int64_t n_rows, int32_t dim, float* out, void* stream) {
  if (!out) return -1;
  return bliss_gather_rows(x, nullptr, n_rows, dim, nullptr, out, stream);
}

int bliss_spmm(const int32_t* indptr, const int32_t* col, const int32_t* perm, const float* w,
               const float* sscale, const float* dscale, int32_t agg, const float* x, int32_t n_rows,
               int32_t dim, const int32_t* heavy, float* y, void* stream) {
  if (n_rows < 0 || dim <= 0 || !indptr || !y) return -1;
  if (n_rows == 0) return 0;
  if (!col || !x) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  const bool al16 = ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0);
  if (dim % 4 == 0 && al16) return launch_spmm<4>(indptr, col, perm, w, sscale, dscale, agg, x, n_rows, dim, heavy, y, st);
  if (dim % 2 == 0) return launch_spmm<2>(indptr, col, perm, w, sscale, dscale, agg, x, n_rows, dim, heavy, y, st);
  return launch_spmm<1>(indptr, col, perm, w, sscale, dscale, agg, x, n_rows, dim, heavy, y, st);
}

}  // extern "C"
