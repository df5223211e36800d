// optim.cu — the optimizer step of the BLISS training step on sm_100a: Adam over ONE flat parameter
// buffer (train_lightning.py:206, torch.optim.Adam(lr) with the default betas / eps, no weight
// decay).  Every parameter, gradient and moment tensor of the model is a view into four flat fp32
// buffers, so the whole update is one coalesced 128-bit streaming kernel (28 B per parameter)
// instead of a multi-tensor launch per parameter list; the kernel also clears the gradient buffer
// for the next backward pass.  lr and the step count live on the device, so the launch is
// CUDA-graph replayable while a scheduler changes lr between replays.
#include "common.cuh"

namespace bliss {

__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                             float* __restrict__ v, int64_t n, const float* __restrict__ lr_dev,
                                             float beta1, float beta2, float eps, const int64_t* __restrict__ step_dev,
                                             int zero_grad) {
  __shared__ float s_c[2];
  if (threadIdx.x == 0) {
    const double t = (double)(*step_dev + 1);
    const double bc1 = 1.0 - pow((double)beta1, t), bc2 = 1.0 - pow((double)beta2, t);
    s_c[0] = (float)((double)*lr_dev / bc1);   // step size
    s_c[1] = (float)sqrt(bc2);                 // sqrt of the second-moment bias correction
  }
  __syncthreads();
  const float step_size = s_c[0], bc2_sqrt = s_c[1];
  const float w1 = 1.0f - beta1, w2 = 1.0f - beta2;
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i], gg = reinterpret_cast<float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* pf = reinterpret_cast<float*>(&pp);
    float* gf = reinterpret_cast<float*>(&gg);
    float* mf = reinterpret_cast<float*>(&mm);
    float* vf = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      mf[k] = mf[k] + w1 * (gf[k] - mf[k]);                 // exp_avg.lerp_(grad, 1 - beta1)
      vf[k] = beta2 * vf[k] + w2 * gf[k] * gf[k];           // exp_avg_sq = beta2 v + (1 - beta2) g^2
      pf[k] -= step_size * mf[k] / (sqrtf(vf[k]) / bc2_sqrt + eps);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i];
    const float mi = m[i] + w1 * (gi - m[i]);
    const float vi = beta2 * v[i] + w2 * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) / bc2_sqrt + eps);
    if (zero_grad) g[i] = 0.f;
  }
}

__global__ void k_adam_tick(int64_t* step_dev) { *step_dev += 1; }

// ---- gradient all-reduce through peer memory, fused into the optimizer step (data parallel) -----------------
// Every rank owns a window in symmetric memory: [2 parities][world slots of the flat gradient] ++ flags[2][world].
// k_grad_push stores this rank's flat gradient into ITS slot of EVERY window over NVLink and the last CTA raises
// flag = step + 1 everywhere; k_adam then reads the W slots of the rank's own window, adds them in rank order (a
// fixed order: every rank computes bit-identical means, parameters stay replicated) and applies Adam — no
// all-reduce call, no extra pass over the gradients.  parity = step & 1 (a rank cannot run two steps ahead: its
// Adam waits for every peer's push of the same step).
__device__ __forceinline__ float* gslot(const bliss_grad_p2p& q, int peer, int parity, int src) {
  return reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(q.peer_base[peer]) + (int64_t)parity * q.parity_stride +
                                  (int64_t)src * q.slot_bytes);
}
__device__ __forceinline__ unsigned long long* gflag(const bliss_grad_p2p& q, int peer, int parity, int src) {
  return reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(q.peer_base[peer]) + q.flags_off) +
         (int64_t)parity * q.world + src;
}
__global__ void __launch_bounds__(256) k_grad_push(const float* __restrict__ g, int64_t n, bliss_grad_p2p q) {
  const long long step = *q.step_dev;
  const int parity = (int)(step & 1);
  const int64_t n4 = n >> 2, stride = (int64_t)gridDim.x * blockDim.x;
  if (q.mc_base) {   // one store per value into the multicast view of the windows: the switch replicates it
    float* mc = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(q.mc_base) + (int64_t)parity * q.parity_stride +
                                         (int64_t)q.rank * q.slot_bytes);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride)
      multimem_st_f32x4(reinterpret_cast<float4*>(mc) + i, reinterpret_cast<const float4*>(g)[i]);
    for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride)
      multimem_st_f32(mc + i, g[i]);
  } else {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
      const float4 v = reinterpret_cast<const float4*>(g)[i];
      for (int r = 0; r < q.world; ++r) reinterpret_cast<float4*>(gslot(q, r, parity, q.rank))[i] = v;
    }
    for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride)
      for (int r = 0; r < q.world; ++r) gslot(q, r, parity, q.rank)[i] = g[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(q.done_ctr, 1u);
    if (t == gridDim.x - 1) {
      *q.done_ctr = 0u;
      __threadfence_system();
      for (int r = 0; r < q.world; ++r)
        *reinterpret_cast<volatile unsigned long long*>(gflag(q, r, parity, q.rank)) = (unsigned long long)(step + 1);
      __threadfence_system();
    }
  }
}
__global__ void __launch_bounds__(32) k_grad_wait(bliss_grad_p2p q, int32_t* error) {
  const long long want = *q.step_dev + 1;
  const int parity = (int)((want - 1) & 1);
  if ((int)threadIdx.x < q.world) {
    volatile unsigned long long* f = gflag(q, q.rank, parity, threadIdx.x);
    const long long t0 = clock64();
    while ((long long)*f != want) {
      if (clock64() - t0 > 8000000000ll) {
        if (error) atomicOr(error, 1 << 30);
        break;
      }
      __nanosleep(100);
    }
  }
  __threadfence_system();
}
__global__ void __launch_bounds__(256) k_adam_p2p(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                 float* __restrict__ v, int64_t n, const float* __restrict__ lr_dev,
                                                 float beta1, float beta2, float eps, const int64_t* __restrict__ step_dev,
                                                 bliss_grad_p2p q) {
  __shared__ float s_c[2];
  if (threadIdx.x == 0) {
    const double t = (double)(*step_dev + 1);
    const double bc1 = 1.0 - pow((double)beta1, t), bc2 = 1.0 - pow((double)beta2, t);
    s_c[0] = (float)((double)*lr_dev / bc1);
    s_c[1] = (float)sqrt(bc2);
  }
  __syncthreads();
  const float step_size = s_c[0], bc2_sqrt = s_c[1];
  const float w1 = 1.0f - beta1, w2 = 1.0f - beta2, world = (float)q.world;
  const int parity = (int)(*q.step_dev & 1);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float gi = __ldcg(gslot(q, q.rank, parity, 0) + i);
    for (int r = 1; r < q.world; ++r) gi += __ldcg(gslot(q, q.rank, parity, r) + i);   // rank order: identical on every rank
    gi = gi / world;                                                                    // all_reduce(sum) then div_(world)
    const float mi = m[i] + w1 * (gi - m[i]);
    const float vi = beta2 * v[i] + w2 * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) / bc2_sqrt + eps);
    g[i] = 0.f;
  }
}

// grad[r, c] += Σ_s part[s][r][c]  for c < cols (part rows are cols_pad wide): the chunk partials of a split-K
// weight gradient are added in chunk order (deterministic) straight into the flat gradient buffer — one
// launch instead of a reduction, a slice copy and an accumulate.
__global__ void __launch_bounds__(256) k_splitk_accumulate(const float* __restrict__ part, int n_parts, int rows,
                                                          int cols_pad, int cols, float* __restrict__ grad) {
  const int64_t plane = (int64_t)rows * cols_pad;
  const int64_t n = (int64_t)rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    const float* __restrict__ p = part + (int64_t)r * cols_pad + c;
    float t = 0.0f;
#pragma unroll 8
    for (int s = 0; s < n_parts; ++s) t += __ldg(p + s * plane);
    grad[i] += t;
  }
}

}  // namespace bliss

extern "C" int bliss_splitk_accumulate(const float* part, int32_t n_parts, int32_t rows, int32_t cols_pad, int32_t cols,
                                       float* grad, void* stream) {
  if (n_parts <= 0 || rows <= 0 || cols <= 0 || cols_pad < cols || !part || !grad) return -1;
  const int64_t n = (int64_t)rows * cols;
  int64_t blocks = (n + 255) / 256;
  if (blocks > BLISS_SM_COUNT * 8) blocks = BLISS_SM_COUNT * 8;
  bliss::k_splitk_accumulate<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(part, n_parts, rows, cols_pad, cols, grad);
  BLISS_CHECK_LAUNCH();
  return 0;
}

extern "C" int bliss_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                               const float* lr_dev, float beta1, float beta2, float eps, int64_t* step_dev,
                               int32_t zero_grad, void* stream) {
  if (n < 0 || !lr_dev || !step_dev) return -1;
  if (n > 0 && (!params || !grads || !exp_avg || !exp_avg_sq)) return -1;
  if (((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) return -2;
  cudaStream_t st = (cudaStream_t)stream;
  if (n > 0) {
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > BLISS_SM_COUNT * 8) blocks = BLISS_SM_COUNT * 8;
    bliss::k_adam<<<(int)blocks, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, lr_dev, beta1, beta2, eps,
                                               step_dev, zero_grad);
    BLISS_CHECK_LAUNCH();
  }
  bliss::k_adam_tick<<<1, 1, 0, st>>>(step_dev);
  BLISS_CHECK_LAUNCH();
  return 0;
}

extern "C" int bliss_grad_push(const float* grads, int64_t n, const bliss_grad_p2p* q, void* stream) {
  if (!grads || n <= 0 || !q || q->world <= 0 || q->world > 32 || !q->peer_base || !q->step_dev || !q->done_ctr) return -1;
  if (((uintptr_t)grads & 15) || (q->slot_bytes & 15) || q->slot_bytes < 4 * n) return -2;
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > BLISS_SM_COUNT * 2) blocks = BLISS_SM_COUNT * 2;
  bliss::k_grad_push<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(grads, n, *q);
  BLISS_CHECK_LAUNCH();
  return 0;
}

extern "C" int bliss_adam_step_p2p(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                   const float* lr_dev, float beta1, float beta2, float eps, int64_t* step_dev,
                                   const bliss_grad_p2p* q, int32_t* error, void* stream) {
  if (n <= 0 || !lr_dev || !step_dev || !params || !grads || !exp_avg || !exp_avg_sq) return -1;
  if (!q || q->world <= 0 || q->world > 32 || !q->peer_base || !q->step_dev || q->slot_bytes < 4 * n) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  bliss::k_grad_wait<<<1, 32, 0, st>>>(*q, error);
  BLISS_CHECK_LAUNCH();
  int64_t blocks = (n + 255) / 256;
  if (blocks > BLISS_SM_COUNT * 8) blocks = BLISS_SM_COUNT * 8;
  bliss::k_adam_p2p<<<(int)blocks, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, lr_dev, beta1, beta2, eps, step_dev,
                                                 *q);
  BLISS_CHECK_LAUNCH();
  bliss::k_adam_tick<<<1, 1, 0, st>>>(step_dev);
  BLISS_CHECK_LAUNCH();
  return 0;
}
