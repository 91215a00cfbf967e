// Mixture-of-experts predictive moments: the per-sample part of MixtureofLinearTransforms.predict
// (transforms/MixtureofLinearTransforms.py:100-106)
//
//     mu_s    = sum_k p_sk m_sk                                   m_sk = E[y | x_s, component k]   (n-vector)
//     Sigma_s = base_s + sum_k p_sk m_sk m_sk^T - mu_s mu_s^T      base_s = sum_k p_sk ESigma_k
//
// The component means of a block of samples are ONE GEMM with a shared operand ((rows x p') (p' x K n)) and so is base
// ((rows x K) (K x n^2)); what is left is a weighted rank-K update PER SAMPLE — (n x K) (K x n) with no operand shared
// between samples — which as a batch of tiny GEMMs was 60 % of predict's time (35 ms per 1 Mi inputs at n = p = 32, K = 64).
// Here one warp owns one sample: lane a keeps row a of Sigma_s in registers (n <= 32), walks the components, and reads the
// component mean once as its own element (coalesced) and once as a row broadcast through L1 (eight 16-byte loads that
// every lane issues to the same address): 2 K n^2 flops against 4 K n bytes per sample, HBM-bound on the means
// (8.6 GB per 1 Mi samples at K n = 2048) and on the n^2 outputs.
#include "common.cuh"

namespace vbmp {

template <int NP>                      // NP = n rounded up to a multiple of 4, <= 32
__global__ void __launch_bounds__(256) moe_moments_kernel(const float* __restrict__ mean, const float* __restrict__ p,
                                                          const float* __restrict__ base, long long N, int K, int n,
                                                          float* __restrict__ mu, float* __restrict__ Sigma) {
  const int lane = threadIdx.x & 31;
  const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const bool live = lane < n;
  const bool vec = (n & 3) == 0;                      // rows of the means are 16-byte aligned
  for (long long s = w0; s < N; s += nw) {
    const float* ms = mean + (size_t)s * K * n;
    const float* ps = p + (size_t)s * K;
    float acc[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) acc[j] = 0.f;
    float mua = 0.f;
    for (int k = 0; k < K; ++k) {
      const float pk = __ldg(ps + k);
      const float* mk = ms + (size_t)k * n;
      const float ma = live ? __ldg(mk + lane) : 0.f;
      const float w = pk * ma;
      mua += w;
      if (vec) {
#pragma unroll
        for (int j = 0; j < NP; j += 4) {
          if (j < n) {
            const float4 m4 = __ldg(reinterpret_cast<const float4*>(mk + j));
            acc[j] = fmaf(w, m4.x, acc[j]); acc[j + 1] = fmaf(w, m4.y, acc[j + 1]);
            acc[j + 2] = fmaf(w, m4.z, acc[j + 2]); acc[j + 3] = fmaf(w, m4.w, acc[j + 3]);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < NP; ++j)
          if (j < n) acc[j] = fmaf(w, __ldg(mk + j), acc[j]);
      }
    }
    // Sigma[a][j] = base[a][j] + acc[j] - mu_a mu_j
    if (live) mu[(size_t)s * n + lane] = mua;
    float* So = Sigma + ((size_t)s * n + lane) * n;
    const float* Bo = base ? base + ((size_t)s * n + lane) * n : nullptr;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const float muj = __shfl_sync(0xffffffffu, mua, j);
      if (live && j < n) So[j] = (Bo ? Bo[j] : 0.f) + acc[j] - mua * muj;
    }
  }
}

int launch_moe_moments(const float* mean, const float* p, const float* base, long long N, int K, int n, float* mu, float* Sigma,
                       cudaStream_t st) {
  if (N < 0 || K < 1 || n < 1 || n > 32) { set_error("moe_moments: bad shape N=%lld K=%d n=%d (n <= 32)", N, K, n); return VBMP_ERR_SHAPE; }
  if (N == 0) return VBMP_OK;
  const int np = (n + 3) / 4 * 4;
  long long blocks = (N + 7) / 8;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  switch (np) {
#define MM_CASE(V) case V: moe_moments_kernel<V><<<(unsigned)blocks, 256, 0, st>>>(mean, p, base, N, K, n, mu, Sigma); break;
    MM_CASE(4) MM_CASE(8) MM_CASE(12) MM_CASE(16) MM_CASE(20) MM_CASE(24) MM_CASE(28) MM_CASE(32)
#undef MM_CASE
    default: set_error("moe_moments: n=%d", n); return VBMP_ERR_SHAPE;
  }
  return check_launch("moe_moments");
}

}  // namespace vbmp
