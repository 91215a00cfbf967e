// K6: HMM forward-backward in log space (models/HMM.py:72-105, HMM.forward_backward_logits) — the first "next" row of
// SURVEY.md §8f.  The reference runs a Python loop over T with an (S,K,K) logsumexp per step (~10 tiny kernels per
// step, twice); here one warp owns one sequence for the whole recursion, lane = hidden state (K <= 32):
//   forward   a_t[j]  = lse_i(a_{t-1}[i] + A[i][j] + l_t[j]),  logZ = lse_j a_{T-1}[j]
//   backward  xi_ij   = (f_t[i] + A[i][j] - lse_i'(f_t[i'] + A[i'][j])) + b_{t+1}[j],  b_t[i] = lse_j xi_ij,
//             SEzz   += exp(xi - lse_ij xi);   initial step with pi_0 in place of f_t;   p_t = softmax(b_t / ptemp)
// Each lane keeps column j and row i of the (log) transition matrix and its row of SEzz in registers; values move
// between lanes by shuffles.  The filtered values a_t are staged in the output buffer p and overwritten by p_t.
// Sequence s uses parameter group s % G (batches of HMMs).  fp32 throughout, as in the reference.
#include "common.cuh"

namespace vbmp {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// exp for the K-term inner sums: ex2.approx (relative error ~2^-22), arguments are <= 0
__device__ __forceinline__ float fexp(float x) { return exp2f(x * 1.4426950408889634f); }
__device__ __forceinline__ float warp_lse(float v) {        // logsumexp over the lanes (-inf lanes contribute nothing)
  const float m = warp_max(v);
  return m + logf(warp_sum(expf(v - m)));
}

template <int KP>
__global__ void __launch_bounds__(128, 4) hmm_fb_kernel(const float* __restrict__ logits, const float* __restrict__ trans,
                                                     const float* __restrict__ init, int T, long long S, int G, int K,
                                                     float inv_ptemp, float* __restrict__ p, float* __restrict__ SEzz,
                                                     float* __restrict__ SEz0, float* __restrict__ logZ) {
  const int lane = threadIdx.x & 31;
  const long long s = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= S) return;
  const int g = (int)(s % G);
  const float NEG = -INFINITY;
  const bool live = lane < K;
  const float* A = trans + (size_t)g * K * K;
  float trc[KP], trr[KP], zz[KP];
#pragma unroll
  for (int i = 0; i < KP; ++i) {
    trc[i] = (live && i < K) ? A[i * K + lane] : NEG;       // column `lane`
    trr[i] = (live && i < K) ? A[lane * K + i] : NEG;       // row `lane`
    zz[i] = 0.f;
  }
  const float pi0 = live ? init[(size_t)g * K + lane] : NEG;
  const size_t stride = (size_t)S * K;                       // elements between consecutive time steps
  const float* lg = logits + (size_t)s * K + lane;
  float* pb = p + (size_t)s * K + lane;

  // ---- forward filter (the observation logits of step t+4 are fetched while step t runs: consecutive steps of a
  //      sequence are S*K floats apart, so every step is a fresh DRAM access that must not sit on the recursion's chain)
  float a = pi0;                                             // plays a_{-1}
  float lbuf[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) lbuf[u] = (live && u < T) ? lg[(size_t)u * stride] : NEG;
  for (int t0 = 0; t0 < T; t0 += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u;
      if (t >= T) break;
      const float l = lbuf[u];
      lbuf[u] = (live && t + 4 < T) ? lg[(size_t)(t + 4) * stride] : NEG;
      float x[KP], mq[4] = {NEG, NEG, NEG, NEG};
#pragma unroll
      for (int i = 0; i < KP; ++i) {
        x[i] = (__shfl_sync(0xffffffffu, a, i) + trc[i]) + l;
        mq[i & 3] = fmaxf(mq[i & 3], x[i]);                  // four short chains instead of one long one
      }
      const float m = fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3]));
      float sq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < KP; ++i) sq[i & 3] += fexp(x[i] - m);
      const float sum = (sq[0] + sq[1]) + (sq[2] + sq[3]);
      a = live ? m + logf(sum) : NEG;
      if (live) pb[(size_t)t * stride] = a;
    }
  }
  const float lz = warp_lse(a);
  if (lane == 0) logZ[s] = lz;

  // ---- backward smoother
  float b = a - lz;                                          // smoothed log-marginal of step t+1 (lane = state)
  auto emit = [&](int t, float bv) {                         // p_t = softmax(b_t / ptemp)
    const float mm = warp_max(bv);
    const float e = live ? expf((bv - mm) * inv_ptemp) : 0.f;
    const float z = warp_sum(e);
    if (live) pb[(size_t)t * stride] = e / z;
  };
  emit(T - 1, b);
  float fbuf[4];                                             // filtered values of steps t, t-1, t-2, t-3 (prefetched)
#pragma unroll
  for (int u = 0; u < 4; ++u) fbuf[u] = (live && T - 2 - u >= 0) ? pb[(size_t)(T - 2 - u) * stride] : NEG;
  for (int t0 = T - 2; t0 >= -1; t0 -= 4) {
#pragma unroll
   for (int u = 0; u < 4; ++u) {
    const int t = t0 - u;
    if (t < -1) break;
    // f = filtered log-marginal of step t (normalised by logZ); t = -1 is the initial-state step (pi_0, HMM.py:94-98)
    const float f = (t >= 0) ? (live ? fbuf[u] - lz : NEG) : pi0;
    fbuf[u] = (live && t - 4 >= 0) ? pb[(size_t)(t - 4) * stride] : NEG;
    // lane j: norm_j = lse_i(f_i + A_ij)
    float x[KP], mq[4] = {NEG, NEG, NEG, NEG};
#pragma unroll
    for (int i = 0; i < KP; ++i) {
      x[i] = __shfl_sync(0xffffffffu, f, i) + trc[i];
      mq[i & 3] = fmaxf(mq[i & 3], x[i]);
    }
    const float m = fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3]));
    float sq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < KP; ++i) sq[i & 3] += fexp(x[i] - m);
    const float sum = (sq[0] + sq[1]) + (sq[2] + sq[3]);
    const float c = live ? b - (m + logf(sum)) : NEG;        // c_j = b_{t+1}[j] - norm_j
    // lane i: row_i = lse_j(A_ij + c_j),  b_t[i] = f_i + row_i
    mq[0] = mq[1] = mq[2] = mq[3] = NEG;
#pragma unroll
    for (int j = 0; j < KP; ++j) {
      x[j] = trr[j] + __shfl_sync(0xffffffffu, c, j);
      mq[j & 3] = fmaxf(mq[j & 3], x[j]);
    }
    const float m2 = fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3]));
    sq[0] = sq[1] = sq[2] = sq[3] = 0.f;
#pragma unroll
    for (int j = 0; j < KP; ++j) { x[j] = fexp(x[j] - m2); sq[j & 3] += x[j]; }
    const float sum2 = (sq[0] + sq[1]) + (sq[2] + sq[3]);
    const float bn = live ? f + (m2 + logf(sum2)) : NEG;
    const float ltot = warp_lse(bn);                         // lse over all (i,j) of xi
    const float w = live ? expf(f + m2 - ltot) : 0.f;        // exp(xi_ij - ltot) = e_ij * w_i
#pragma unroll
    for (int j = 0; j < KP; ++j) zz[j] = fmaf(x[j], w, zz[j]);
    if (t >= 0) {
      b = bn;
      emit(t, b);
    } else {
      const float l0 = warp_lse(bn);                         // SEz0 = softmax_i(lse_j xi_ij)
      if (live) SEz0[(size_t)s * K + lane] = expf(bn - l0);
    }
   }
  }
  if (live) {
    float* out = SEzz + ((size_t)s * K + lane) * K;
#pragma unroll
    for (int j = 0; j < KP; ++j)
      if (j < K) out[j] = zz[j];
  }
}

int launch_hmm_fb(const float* logits, const float* trans, const float* init, int T, long long S, int G, int K, float ptemp,
                  float* p, float* SEzz, float* SEz0, float* logZ, cudaStream_t st) {
  if (T < 1 || S < 1 || G < 1 || K < 1 || K > 32 || !(ptemp > 0.f)) {
    set_error("hmm_fb: bad shape T=%d S=%lld G=%d K=%d ptemp=%g (K <= 32 states)", T, S, G, K, (double)ptemp);
    return VBMP_ERR_SHAPE;
  }
  const int wpb = 4;
  const unsigned grid = (unsigned)((S + wpb - 1) / wpb);
  const float ip = 1.0f / ptemp;
  if (K <= 8) hmm_fb_kernel<8><<<grid, wpb * 32, 0, st>>>(logits, trans, init, T, S, G, K, ip, p, SEzz, SEz0, logZ);
  else if (K <= 16) hmm_fb_kernel<16><<<grid, wpb * 32, 0, st>>>(logits, trans, init, T, S, G, K, ip, p, SEzz, SEz0, logZ);
  else hmm_fb_kernel<32><<<grid, wpb * 32, 0, st>>>(logits, trans, init, T, S, G, K, ip, p, SEzz, SEz0, logZ);
  return check_launch("hmm_fb");
}

}  // namespace vbmp
