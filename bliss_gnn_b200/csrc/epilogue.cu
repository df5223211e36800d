// epilogue.cu — the element-wise tail of a hidden SAGE layer on sm_100a, fused:
//   forward   y = dropout(relu(a + b + bias))   and   row_norm[r] = ||y_r||_2
//             (a = fc_self(h_dst) without bias, b = mean-aggregated neighbours; model.py:321-332, and the
//              next layer's embed_norm, model.py:318 — three torch kernels and a norm pass in one)
//   backward  gz = gy * [y > 0] / (1 - p)       and   bias gradient = column sums of gz
//             (dropout + relu + add backward and the bias reduction in one pass; y > 0 exactly when the
//              element was kept by the dropout and passed the relu, so no mask is stored)
// Dropout draws come from Philox4x32-10 keyed by (seed; element group, layer, step), the step being a
// device scalar, so the launch is CUDA-graph replayable.  One warp per row, 128-bit accesses.
#include "common.cuh"

namespace bliss {

#define BLISS_EPI_MAXIT 8   // dim <= 4 * 32 * 8 = 1024

__global__ void __launch_bounds__(256) k_sage_epilogue_fwd(const float* __restrict__ a, const float* __restrict__ b,
                                                          const float* __restrict__ bias, int n_rows, int dim,
                                                          int relu, float p_drop, unsigned long long seed,
                                                          const int64_t* __restrict__ step_dev, unsigned layer,
                                                          float* __restrict__ y, float* __restrict__ row_norm) {
  const int lane = lane_id();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const unsigned long long step = step_dev ? (unsigned long long)*step_dev : 0ull;
  const float keep = 1.0f - p_drop, scale = (p_drop > 0.0f) ? 1.0f / keep : 1.0f;
  const int groups_per_row = dim >> 2;
  for (int r = warp; r < n_rows; r += nwarps) {
    const int64_t off = (int64_t)r * dim;
    float ss = 0.0f;
    for (int c = lane * 4; c < dim; c += 128) {
      const float4 va = *reinterpret_cast<const float4*>(a + off + c);
      const float4 vb = *reinterpret_cast<const float4*>(b + off + c);
      const float4 bs = bias ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      float v[4] = {va.x + vb.x + bs.x, va.y + vb.y + bs.y, va.z + vb.z + bs.z, va.w + vb.w + bs.w};
      if (relu) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], 0.0f);
      }
      if (p_drop > 0.0f) {
        const unsigned g = (unsigned)(r * groups_per_row + (c >> 2));
        const uint4 rnd = philox4x32_10(make_uint4(g, layer, (unsigned)step, (unsigned)(step >> 32)),
                                        make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
        const unsigned w[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float u = (float)(w[i] >> 8) * 5.9604644775390625e-08f;
          v[i] = (u < keep) ? v[i] * scale : 0.0f;
        }
      }
      *reinterpret_cast<float4*>(y + off + c) = make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
      for (int i = 0; i < 4; ++i) ss += v[i] * v[i];
    }
    if (row_norm) {
      ss = warp_sum(ss);
      if (lane == 0) row_norm[r] = sqrtf(ss);
    }
  }
}

// gz = gy * [y > 0] * scale (or gy * scale without relu/dropout gate when gate == 0); every CTA writes its
// column sums of gz to bias_partial[blockIdx.x][dim] (rows are dealt to warps round-robin, the 8 warps
// of a CTA are added in warp order, k_colsum_final adds the CTAs in order: deterministic).
__global__ void __launch_bounds__(256) k_sage_epilogue_bwd(const float* __restrict__ gy, const float* __restrict__ y,
                                                          int n_rows, int dim, int gate, float scale,
                                                          float* __restrict__ gz, float* __restrict__ bias_partial) {
  extern __shared__ float s_col[];   // [8 warps][dim]
  const int lane = lane_id(), wid = warp_id();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  float4 acc[BLISS_EPI_MAXIT];
#pragma unroll
  for (int it = 0; it < BLISS_EPI_MAXIT; ++it) acc[it] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = warp; r < n_rows; r += nwarps) {
    const int64_t off = (int64_t)r * dim;
#pragma unroll
    for (int it = 0; it < BLISS_EPI_MAXIT; ++it) {
      const int c = lane * 4 + it * 128;
      if (c < dim) {
        float4 g = *reinterpret_cast<const float4*>(gy + off + c);
        if (gate) {
          const float4 yy = *reinterpret_cast<const float4*>(y + off + c);
          g.x = yy.x > 0.0f ? g.x * scale : 0.0f;
          g.y = yy.y > 0.0f ? g.y * scale : 0.0f;
          g.z = yy.z > 0.0f ? g.z * scale : 0.0f;
          g.w = yy.w > 0.0f ? g.w * scale : 0.0f;
        }
        *reinterpret_cast<float4*>(gz + off + c) = g;
        acc[it].x += g.x;
        acc[it].y += g.y;
        acc[it].z += g.z;
        acc[it].w += g.w;
      }
    }
  }
  if (!bias_partial) return;
#pragma unroll
  for (int it = 0; it < BLISS_EPI_MAXIT; ++it) {
    const int c = lane * 4 + it * 128;
    if (c < dim) *reinterpret_cast<float4*>(s_col + wid * dim + c) = acc[it];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_col[w * dim + c];
    bias_partial[(int64_t)blockIdx.x * dim + c] = t;
  }
}

// out[c] = Σ_k partial[k][c], k ascending within each of 8 interleaved groups, groups added in order
// (a fixed order).  A CTA takes 32 columns x 8 groups, so the ~150 partials of a column are read as
// 19 independent loads per thread instead of one serial chain.
__global__ void __launch_bounds__(256) k_colsum_final(const float* __restrict__ partial, int n_parts, int dim,
                                                     float* __restrict__ out) {
  __shared__ float s_g[8][33];
  const int cl = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float t = 0.0f;
  if (c < dim)
    for (int k = g; k < n_parts; k += 8) t += __ldg(partial + (int64_t)k * dim + c);
  s_g[g][cl] = t;
  __syncthreads();
  if (g == 0 && c < dim) {
    float r = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) r += s_g[w][cl];
    out[c] = r;
  }
}

// Cross-entropy with mean reduction over the seed batch (train_lightning.py:77-79,142: nn.CrossEntropyLoss())
// and its gradient in one pass: one warp per row, loss = mean_r (logsumexp(x_r) - x_r[y_r]),
// grad = (softmax(x_r) - onehot(y_r)) / n.  The row losses are added in row order by one warp
// (deterministic); the gradient is scaled by the upstream scalar in the backward wrapper.
__global__ void __launch_bounds__(256) k_xent(const float* __restrict__ logits, const int64_t* __restrict__ labels,
                                             int n_rows, int n_cls, float* __restrict__ row_loss,
                                             float* __restrict__ grad) {
  const int lane = lane_id();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const float inv_n = 1.0f / (float)n_rows;
  for (int r = warp; r < n_rows; r += nwarps) {
    const float* __restrict__ x = logits + (int64_t)r * n_cls;
    float m = -INFINITY;
    for (int c = lane; c < n_cls; c += 32) m = fmaxf(m, x[c]);
    m = warp_max(m);
    float s = 0.0f;
    for (int c = lane; c < n_cls; c += 32) s += expf(x[c] - m);
    s = warp_sum(s);
    const int y = (int)labels[r];
    if (lane == 0) row_loss[r] = (logf(s) + m) - x[y];
    const float inv_s = 1.0f / s;
    for (int c = lane; c < n_cls; c += 32)
      grad[(int64_t)r * n_cls + c] = (expf(x[c] - m) * inv_s - (c == y ? 1.0f : 0.0f)) * inv_n;
  }
}
__global__ void __launch_bounds__(32) k_xent_mean(const float* __restrict__ row_loss, int n_rows, float* __restrict__ loss) {
  float t = 0.0f;
  for (int r0 = 0; r0 < n_rows; r0 += 32) {   // row order within a lane, lanes by butterfly: fixed
    const int r = r0 + lane_id();
    t += (r < n_rows) ? row_loss[r] : 0.0f;
  }
  t = warp_sum(t);
  if (lane_id() == 0) *loss = t / (float)n_rows;
}

}  // namespace bliss

using namespace bliss;

extern "C" {

int bliss_sage_epilogue_parts(void) { return BLISS_SM_COUNT; }   // rows of the bias_partial scratch

int bliss_sage_epilogue_fwd(const float* a, const float* b, const float* bias, int32_t n_rows, int32_t dim,
                            int32_t relu, float p_drop, uint64_t seed, const int64_t* step_dev, uint32_t layer,
                            float* y, float* row_norm, void* stream) {
  if (n_rows < 0 || dim <= 0 || dim % 4 || dim > 128 * BLISS_EPI_MAXIT || !a || !b || !y) return -1;
  if (p_drop < 0.0f || p_drop >= 1.0f) return -1;
  if ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)y | (uintptr_t)bias) & 15)) return -2;
  if (n_rows == 0) return 0;
  int blocks = (n_rows + 7) / 8;
  if (blocks > BLISS_SM_COUNT * 8) blocks = BLISS_SM_COUNT * 8;
  k_sage_epilogue_fwd<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, b, bias, n_rows, dim, relu, p_drop, seed, step_dev,
                                                               layer, y, row_norm);
  BLISS_CHECK_LAUNCH();
  return 0;
}

int bliss_sage_epilogue_bwd(const float* grad_y, const float* y, int32_t n_rows, int32_t dim, int32_t gate,
                            float p_drop, float* grad_z, float* bias_partial /* [148, dim] or NULL */,
                            float* grad_bias /* [dim] or NULL */, void* stream) {
  if (n_rows < 0 || dim <= 0 || dim % 4 || dim > 128 * BLISS_EPI_MAXIT || !grad_y || !grad_z) return -1;
  if (gate && !y) return -1;
  if ((bias_partial == nullptr) != (grad_bias == nullptr)) return -1;
  if ((((uintptr_t)grad_y | (uintptr_t)y | (uintptr_t)grad_z) & 15)) return -2;
  cudaStream_t st = (cudaStream_t)stream;
  const float scale = (gate && p_drop > 0.0f) ? 1.0f / (1.0f - p_drop) : 1.0f;
  const size_t smem = (size_t)8 * dim * sizeof(float);
  k_sage_epilogue_bwd<<<BLISS_SM_COUNT, 256, smem, st>>>(grad_y, y, n_rows, dim, gate, scale, grad_z, bias_partial);
  BLISS_CHECK_LAUNCH();
  if (grad_bias) {
    k_colsum_final<<<(dim + 31) / 32, 256, 0, st>>>(bias_partial, BLISS_SM_COUNT, dim, grad_bias);
    BLISS_CHECK_LAUNCH();
  }
  return 0;
}


int bliss_xent_mean(const float* logits, const int64_t* labels, int32_t n_rows, int32_t n_cls, float* row_loss,
                    float* loss, float* grad, void* stream) {
  if (n_rows <= 0 || n_cls <= 0 || !logits || !labels || !row_loss || !loss || !grad) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (n_rows + 7) / 8;
  if (blocks > BLISS_SM_COUNT * 8) blocks = BLISS_SM_COUNT * 8;
  k_xent<<<blocks, 256, 0, st>>>(logits, labels, n_rows, n_cls, row_loss, grad);
  BLISS_CHECK_LAUNCH();
  k_xent_mean<<<1, 32, 0, st>>>(row_loss, n_rows, loss);
  BLISS_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
