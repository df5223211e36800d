// common.cuh — device helpers shared by the BLISS sm_100a kernels.
//
// Numeric contract: every element-wise fp32 operation is a single IEEE
// round-to-nearest op (explicit __f*_rn intrinsics so nvcc never contracts into FMA), row sums
// are accumulated in fp64 and rounded once, column sums and the scale-search sum are 64-bit
// fixed-point integers (order independent -> bit-reproducible with atomics).
#pragma once
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>
#include "../../include/bliss_b200.h"

#define BLISS_SM_COUNT 148
#define BLISS_REG_BIT 0x8000000000000000ull   // acc[v] bit 63: reserved flag bit (masked off when acc is read)
#define BLISS_S_FIX_BITS 40                    // fixed point of the scale-search sum
#define BLISS_CTA 256                          // threads per CTA of the row kernels
#define BLISS_WARPS (BLISS_CTA / 32)
#define BLISS_CHUNK 256                        // edges per warp-chunk of the probability passes
#define BLISS_SPMM_SEG 32                       // edges per warp-segment of the balanced SpMM (rows are cut into segments)

#define BLISS_CHECK_LAUNCH()                      \
  do {                                            \
    cudaError_t e__ = cudaGetLastError();         \
    if (e__ != cudaSuccess) return (int)e__;      \
  } while (0)

namespace bliss {

// Programmatic dependent launch: a kernel launched with launch_pdl() may become resident while its
// predecessor in the stream is still draining; pdl_wait() (first statement, before any read of the
// predecessor's output) blocks until the predecessor grid has completed and its writes are visible;
// pdl_trigger() lets the successor be scheduled as soon as every CTA of this grid has started.  Every
// kernel of the sampling chain is at most one wave, so the waiting CTAs never hold a slot this grid
// still needs.  Without a programmatic edge both are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Sum over the whole CTA, result broadcast to every thread.  `scratch` needs >= 32 slots and is
// reusable right after the call returns.  Fixed tree -> deterministic.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  v = warp_sum(v);
  const int nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane_id() == 0) scratch[warp_id()] = v;
  __syncthreads();
  T r = (threadIdx.x < nw) ? scratch[threadIdx.x] : T(0);
  if (warp_id() == 0) {
    r = warp_sum(r);
    if (lane_id() == 0) scratch[0] = r;
  }
  __syncthreads();
  r = scratch[0];
  return r;
}

// Exclusive prefix sum of one int per thread over the CTA; returns the prefix and writes the
// CTA total to *total.  `scratch` needs >= 33 ints.
__device__ __forceinline__ int block_excl_scan(int v, int* scratch, int* total) {
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane_id() >= o) incl += t;
  }
  const int nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane_id() == 31) scratch[warp_id()] = incl;
  __syncthreads();
  if (warp_id() == 0) {
    int w = (lane_id() < nw) ? scratch[lane_id()] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane_id() >= o) wi += t;
    }
    if (lane_id() < nw) scratch[lane_id()] = wi - w;
    if (lane_id() == 31) scratch[32] = wi;
  }
  __syncthreads();
  int prefix = scratch[warp_id()] + incl - v;
  *total = scratch[32];
  return prefix;
}

// ---- fixed-point column accumulator ----------------------------------------------------
__host__ __device__ __forceinline__ int fx_bits_for(int n_seeds) {
  int bl = 0;
  for (unsigned x = (unsigned)n_seeds; x; x >>= 1) ++bl;
  if (bl < 1) bl = 1;
  return 62 - bl;
}
__device__ __forceinline__ unsigned long long fx_term(float t, double scale) {
  unsigned long long q = __double2ull_rn((double)t * scale);
  return q ? q : 1ull;
}
__device__ __forceinline__ float fx_to_prob(unsigned long long acc, double inv_scale) {
  double s = __ull2double_rn(acc & ~BLISS_REG_BIT) * inv_scale;
  return __fsqrt_rn(__double2float_rn(s));
}

// q_ij = eta/n_i + (1-eta) * (w_ij / sum_j w_ij)      bandit_sampler.py:131-137
__device__ __forceinline__ float edge_q(float w, float row_w, float eta_over_n, float one_minus_eta) {
  return __fadd_rn(eta_over_n, __fmul_rn(one_minus_eta, __fdiv_rn(w, row_w)));
}

// ---- Philox4x32-10 (Salmon et al., SC'11) -------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    unsigned hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    unsigned hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}
// One uniform in [0,1) per (seed, step, layer, node id): word 0, top 24 bits.
__device__ __forceinline__ float philox_uniform(unsigned long long seed, unsigned long long step,
                                                unsigned layer, unsigned nid) {
  uint4 r = philox4x32_10(make_uint4(nid, layer, (unsigned)step, (unsigned)(step >> 32)),
                          make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
  return (float)(r.x >> 8) * 5.9604644775390625e-08f;  // 2^-24, exact
}

__device__ __forceinline__ bool test_bit(const uint32_t* __restrict__ bits, int v) {
  return (__ldg(bits + (v >> 5)) >> (v & 31)) & 1u;
}

// Stores to an NVSwitch multicast address (symmetric-memory windows, data-parallel exchanges): the switch replicates
// one store into every rank's window.  Weak stores, ordered like any other store of the thread by the
// __threadfence_system() in front of the flag that publishes them.
__device__ __forceinline__ void multimem_st_b32(void* a, uint32_t v) {
  asm volatile("multimem.st.weak.global.b32 [%0], %1;" ::"l"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void multimem_st_f32(void* a, float v) {
  asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void multimem_st_f32x4(void* a, float4 v) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

}  // namespace bliss
