// gat.cu — fused GATv2 attention of the BLISS hot path on sm_100a.
// Replaces, in custom_GATv2Conv.forward (model.py:80-99): g-SDDMM u_add_v materialising
// E x H x D, leaky_relu, (e * attn).sum(-1), edge_softmax (five launches in DGL) and g-SpMM
// u_mul_e/sum.  One warp owns one (destination, head): it streams the in-edges once, computes
// the logit on the fly, keeps an online softmax (running max / sum) and the weighted feature
// sum in registers — E x H x D is never written.  The pre-softmax logits ARE written ([E,H])
// because the reference returns them as "attention" (model.py:108-110) and the bandit reads
// their head mean as a_ij (model.py:224-227).  Backward: a destination-major pass (grad of the
// logits, of the destination term and of attn) and a source-major pass over the transposed
// block (grad of the source features) — no atomics on feature rows.
#include "common.cuh"

namespace bliss {

__device__ __forceinline__ float lrelu(float z, float slope) { return z > 0.0f ? z : z * slope; }

template <int NCH>
__global__ void __launch_bounds__(256) k_gatv2_fwd(const int32_t* __restrict__ indptr, const int32_t* __restrict__ col,
                                                  const float* __restrict__ feat, const float* __restrict__ attn,
                                                  const float* __restrict__ drop_mask, float slope, int n_dst,
                                                  int H, int D, float* __restrict__ out, float* __restrict__ logits,
                                                  float* __restrict__ row_max, float* __restrict__ row_sum) {
  const int lane = lane_id();
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_items = (int64_t)n_dst * H;
  for (int64_t it = warp; it < n_items; it += nwarps) {
    const int i = (int)(it / H), h = (int)(it % H);
    float er[NCH], at[NCH], acc[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int d = c * 32 + lane;
      er[c] = d < D ? __ldg(feat + ((int64_t)i * H + h) * D + d) : 0.0f;
      at[c] = d < D ? __ldg(attn + h * D + d) : 0.0f;
      acc[c] = 0.0f;
    }
    float m = -INFINITY, l = 0.0f;
    const int a = indptr[i], b = indptr[i + 1];
    for (int e = a; e < b; ++e) {
      const int u = __ldg(col + e);
      const float* __restrict__ fu = feat + ((int64_t)u * H + h) * D;
      float el[NCH];
      float s = 0.0f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int d = c * 32 + lane;
        el[c] = d < D ? __ldg(fu + d) : 0.0f;
        s = fmaf(at[c], lrelu(el[c] + er[c], slope), s);
      }
      s = warp_sum(s);
      if (lane == 0) logits[(int64_t)e * H + h] = s;
      const float m_new = fmaxf(m, s);
      const float corr = expf(m - m_new);
      const float p = expf(s - m_new);
      l = l * corr + p;
      const float pm = drop_mask ? p * __ldg(drop_mask + (int64_t)e * H + h) : p;
#pragma unroll
      for (int c = 0; c < NCH; ++c) acc[c] = fmaf(pm, el[c], acc[c] * corr);
      m = m_new;
    }
    const float inv = (b > a) ? 1.0f / l : 0.0f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int d = c * 32 + lane;
      if (d < D) out[((int64_t)i * H + h) * D + d] = acc[c] * inv;
    }
    if (lane == 0) {
      row_max[it] = m;
      row_sum[it] = l;
    }
  }
}

template <int NCH>
__global__ void __launch_bounds__(256) k_gatv2_bwd_dst(const int32_t* __restrict__ indptr, const int32_t* __restrict__ col,
                                                      const float* __restrict__ feat, const float* __restrict__ attn,
                                                      const float* __restrict__ drop_mask, const float* __restrict__ logits,
                                                      const float* __restrict__ row_max, const float* __restrict__ row_sum,
                                                      const float* __restrict__ out, const float* __restrict__ grad_out,
                                                      float slope, int n_dst, int H, int D,
                                                      float* __restrict__ grad_logit, float* __restrict__ grad_feat,
                                                      float* __restrict__ grad_attn) {
  const int lane = lane_id();
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_items = (int64_t)n_dst * H;
  for (int64_t it = warp; it < n_items; it += nwarps) {
    const int i = (int)(it / H), h = (int)(it % H);
    float er[NCH], at[NCH], go[NCH], ger[NCH], gat[NCH];
    float dsum = 0.0f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int d = c * 32 + lane;
      const int64_t o = ((int64_t)i * H + h) * D + d;
      er[c] = d < D ? __ldg(feat + o) : 0.0f;
      at[c] = d < D ? __ldg(attn + h * D + d) : 0.0f;
      go[c] = d < D ? __ldg(grad_out + o) : 0.0f;
      dsum = fmaf(go[c], d < D ? __ldg(out + o) : 0.0f, dsum);
      ger[c] = 0.0f;
      gat[c] = 0.0f;
    }
    dsum = warp_sum(dsum);
    const float m = row_max[it];
    const float inv_l = 1.0f / row_sum[it];
    for (int e = indptr[i]; e < indptr[i + 1]; ++e) {
      const int u = __ldg(col + e);
      const float* __restrict__ fu = feat + ((int64_t)u * H + h) * D;
      float el[NCH];
      float dot = 0.0f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int d = c * 32 + lane;
        el[c] = d < D ? __ldg(fu + d) : 0.0f;
        dot = fmaf(go[c], el[c], dot);
      }
      dot = warp_sum(dot);
      const float a = expf(logits[(int64_t)e * H + h] - m) * inv_l;
      const float mk = drop_mask ? __ldg(drop_mask + (int64_t)e * H + h) : 1.0f;
      const float ds = a * (mk * dot - dsum);
      if (lane == 0) grad_logit[(int64_t)e * H + h] = ds;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const float z = el[c] + er[c];
        ger[c] = fmaf(ds * at[c], z > 0.0f ? 1.0f : slope, ger[c]);
        gat[c] = fmaf(ds, lrelu(z, slope), gat[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int d = c * 32 + lane;
      if (d < D) {
        grad_feat[((int64_t)i * H + h) * D + d] = ger[c];
        atomicAdd(grad_attn + h * D + d, gat[c]);
      }
    }
  }
}

template <int NCH>
__global__ void __launch_bounds__(256) k_gatv2_bwd_src(const int32_t* __restrict__ t_indptr, const int32_t* __restrict__ t_dst,
                                                      const int32_t* __restrict__ t_perm, const float* __restrict__ feat,
                                                      const float* __restrict__ attn, const float* __restrict__ drop_mask,
                                                      const float* __restrict__ logits, const float* __restrict__ row_max,
                                                      const float* __restrict__ row_sum, const float* __restrict__ grad_out,
                                                      const float* __restrict__ grad_logit, float slope, int n_src,
                                                      int n_dst, int H, int D, float* __restrict__ grad_feat) {
  const int lane = lane_id();
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_items = (int64_t)n_src * H;
  for (int64_t it = warp; it < n_items; it += nwarps) {
    const int u = (int)(it / H), h = (int)(it % H);
    float el[NCH], at[NCH], g[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int d = c * 32 + lane;
      const int64_t o = ((int64_t)u * H + h) * D + d;
      el[c] = d < D ? __ldg(feat + o) : 0.0f;
      at[c] = d < D ? __ldg(attn + h * D + d) : 0.0f;
      g[c] = (d < D && u < n_dst) ? grad_feat[o] : 0.0f;  // destination-term gradient from bwd_dst
    }
    for (int k = t_indptr[u]; k < t_indptr[u + 1]; ++k) {
      const int e = __ldg(t_perm + k);
      const int i = __ldg(t_dst + k);
      const int64_t ih = (int64_t)i * H + h;
      const float a = expf(logits[(int64_t)e * H + h] - row_max[ih]) / row_sum[ih];
      const float mk = drop_mask ? __ldg(drop_mask + (int64_t)e * H + h) : 1.0f;
      const float am = a * mk;
      const float ds = grad_logit[(int64_t)e * H + h];
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int d = c * 32 + lane;
        if (d < D) {
          const float go = __ldg(grad_out + ih * D + d);
          const float z = el[c] + __ldg(feat + ih * D + d);
          g[c] = fmaf(am, go, g[c]);
          g[c] = fmaf(ds * at[c], z > 0.0f ? 1.0f : slope, g[c]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int d = c * 32 + lane;
      if (d < D) grad_feat[((int64_t)u * H + h) * D + d] = g[c];
    }
  }
}

}  // namespace bliss

using namespace bliss;

static inline int warp_grid(int64_t items, int max_blocks) {
  int64_t b = (items + 7) / 8;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

#define BLISS_GAT_DISPATCH(KERNEL, ...)                         \
  do {                                                          \
    const int per_lane = (dim + 31) / 32;                       \
    if (per_lane <= 1) KERNEL<1><<<grid, 256, 0, st>>>(__VA_ARGS__);      \
    else if (per_lane <= 2) KERNEL<2><<<grid, 256, 0, st>>>(__VA_ARGS__); \
    else if (per_lane <= 4) KERNEL<4><<<grid, 256, 0, st>>>(__VA_ARGS__); \
    else if (per_lane <= 8) KERNEL<8><<<grid, 256, 0, st>>>(__VA_ARGS__); \
    else return -2; /* head width > 256 not supported */       \
  } while (0)

extern "C" {

int bliss_gatv2_fwd(const int32_t* indptr, const int32_t* col, const float* feat, const float* attn,
                    const float* drop_mask, float negative_slope, int32_t n_dst, int32_t heads, int32_t dim,
                    float* out, float* logits, float* row_max, float* row_sum, void* stream) {
  if (!indptr || !feat || !attn || !out || !row_max || !row_sum || n_dst < 0 || heads <= 0 || dim <= 0) return -1;
  if (n_dst == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = warp_grid((int64_t)n_dst * heads, BLISS_SM_COUNT * 16);
  BLISS_GAT_DISPATCH(k_gatv2_fwd, indptr, col, feat, attn, drop_mask, negative_slope, n_dst, heads, dim, out,
                     logits, row_max, row_sum);
  BLISS_CHECK_LAUNCH();
  return 0;
}

int bliss_gatv2_bwd_dst(const int32_t* indptr, const int32_t* col, const float* feat, const float* attn,
                        const float* drop_mask, const float* logits, const float* row_max, const float* row_sum,
                        const float* out, const float* grad_out, float negative_slope, int32_t n_dst,
                        int32_t heads, int32_t dim, float* grad_logit, float* grad_feat, float* grad_attn,
                        void* stream) {
  if (!indptr || !feat || !attn || !logits || !row_max || !row_sum || !out || !grad_out || !grad_feat || !grad_attn)
    return -1;
  if (n_dst <= 0) return n_dst < 0 ? -1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = warp_grid((int64_t)n_dst * heads, BLISS_SM_COUNT * 16);
  BLISS_GAT_DISPATCH(k_gatv2_bwd_dst, indptr, col, feat, attn, drop_mask, logits, row_max, row_sum, out, grad_out,
                     negative_slope, n_dst, heads, dim, grad_logit, grad_feat, grad_attn);
  BLISS_CHECK_LAUNCH();
  return 0;
}

int bliss_gatv2_bwd_src(const int32_t* t_indptr, const int32_t* t_dst, const int32_t* t_perm, const float* feat,
                        const float* attn, const float* drop_mask, const float* logits, const float* row_max,
                        const float* row_sum, const float* grad_out, const float* grad_logit,
                        float negative_slope, int32_t n_src, int32_t n_dst, int32_t heads, int32_t dim,
                        float* grad_feat, void* stream) {
  if (!t_indptr || !feat || !attn || !logits || !row_max || !row_sum || !grad_out || !grad_logit || !grad_feat)
    return -1;
  if (n_src <= 0) return n_src < 0 ? -1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = warp_grid((int64_t)n_src * heads, BLISS_SM_COUNT * 16);
  BLISS_GAT_DISPATCH(k_gatv2_bwd_src, t_indptr, t_dst, t_perm, feat, attn, drop_mask, logits, row_max, row_sum,
                     grad_out, grad_logit, negative_slope, n_src, n_dst, heads, dim, grad_feat);
  BLISS_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
