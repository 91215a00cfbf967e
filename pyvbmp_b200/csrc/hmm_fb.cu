// K6: HMM forward-backward in log space (models/HMM.py:72-105, HMM.forward_backward_logits) — the first "next" row of
// SURVEY.md §8f.  The reference runs a Python loop over T with an (S,K,K) logsumexp per step (~10 tiny kernels per
// step, twice); here one warp owns one sequence for the whole recursion, lane = hidden state (K <= 32):
//   forward   a_t[j]  = lse_i(a_{t-1}[i] + A[i][j] + l_t[j]),  logZ = lse_j a_{T-1}[j]
//   backward  xi_ij   = (f_t[i] + A[i][j] - lse_i'(f_t[i'] + A[i'][j])) + b_{t+1}[j],  b_t[i] = lse_j xi_ij,
//             SEzz   += exp(xi - lse_ij xi);   initial step with pi_0 in place of f_t;   p_t = softmax(b_t / ptemp)
// Each lane keeps column j and row i of the (log) transition matrix and its row of SEzz in registers; values move
// between lanes by shuffles.  The filtered values a_t are staged in the output buffer p and overwritten by p_t.
// Sequence s uses parameter group s % G (batches of HMMs).  fp32 throughout, as in the reference.
#include <cstdlib>
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
__global__ void __launch_bounds__(128, 4) hmm_fb_log_kernel(const float* __restrict__ logits, const float* __restrict__ trans,
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

// ---- scaled linear-domain variant (ptemp == 1) ---------------------------------------------------------------------
// The log-space recursion above spends ~K exponentials per lane per step on each of its three K-term logsumexps (MUFU
// bound: 8.2 ms at 4096 sequences x T = 1024, K = 32).  The same quantities in the probability domain with per-step
// renormalisation need K FMAs per matrix-vector product and ONE exponential per lane per step (the emission term):
//   forward   u_j = (sum_i ah_i A_ij) 2^((l_t[j] - max_j l_t) log2 e),  c_t = sum_j u_j,  ah <- u / c_t,
//             logZ = log sum_i pi_i + sum_t (log c_t + max_j l_t)                       (ah = normalised filtered marginal)
//   backward  n_j = sum_i f_i A_ij,  c_j = g_j / n_j,  x_ij = A_ij c_j,  g'_i = f_i sum_j x_ij,  tot = sum_i g'_i,
//             SEzz_ij += x_ij f_i / tot,  g <- g' / tot = p_t                           (g = smoothed marginal)
// which is models/HMM.py:72-105 term by term: xi_ij = f_i A_ij / n_j g_j, fw[t] = lse_j xi, SEzz += exp(xi - lse_ij xi),
// p = softmax(fw).  Vectors cross lanes through a 128-byte shared-memory line per warp (one store + KP/4 broadcast
// 16-byte loads instead of KP shuffles).  States whose probability underflows fp32 here are below e^-87 in the
// reference's p as well.  logZ is accumulated in fp64 (T terms of size |l|).  ptemp != 1 raises the smoothed marginals
// to 1/ptemp, where an underflowed state could matter: those calls keep the log-space kernel.
// (K = 32: three register arrays of 32 — column and row of A, the lane's row of SEzz — do not fit 128 registers; three
// resident blocks of 4 warps instead of four avoid the spills)
template <int KP>
__global__ void __launch_bounds__(128, KP == 32 ? 3 : 4) hmm_fb_lin_kernel(const float* __restrict__ logits, const float* __restrict__ trans,
                                                         const float* __restrict__ init, int T, long long S, int G, int K,
                                                         float* __restrict__ p, float* __restrict__ SEzz,
                                                         float* __restrict__ SEz0, float* __restrict__ logZ) {
  __shared__ __align__(16) float sh[4][2][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long s = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
  if (s >= S) return;
  const int g = (int)(s % G);
  const bool live = lane < K;
  const float* Atr = trans + (size_t)g * K * K;
  constexpr float L2E = 1.4426950408889634f;
  float ac[KP];
#pragma unroll
  for (int i = 0; i < KP; ++i) ac[i] = (live && i < K) ? exp2f(Atr[i * K + lane] * L2E) : 0.f;   // column `lane` of A = exp(log transition)
  const float pi0 = live ? exp2f(init[(size_t)g * K + lane] * L2E) : 0.f;
  float* va = sh[wib][0];                                    // vector being broadcast (filtered marginal / f)
  float* vc = sh[wib][1];                                    // second vector (c)
  const size_t stride = (size_t)S * K;
  const float* lg = logits + (size_t)s * K + lane;
  float* pb = p + (size_t)s * K + lane;
  const float NEG = -INFINITY;

  // matrix-vector product against the vector in `v` (shared): sum_i v[i] m[i], four partial sums
  auto dot = [&](const float* v, const float (&m)[KP]) {
    float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
    for (int i = 0; i < KP; i += 4) {
      const float4 x = *reinterpret_cast<const float4*>(v + i);
      q0 = fmaf(x.x, m[i], q0); q1 = fmaf(x.y, m[i + 1], q1); q2 = fmaf(x.z, m[i + 2], q2); q3 = fmaf(x.w, m[i + 3], q3);
    }
    return (q0 + q1) + (q2 + q3);
  };

  // ---- forward filter
  const float s0 = warp_sum(pi0);
  float ah = pi0 / s0;
  double lzacc = (double)logf(s0);
  float lbuf[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) lbuf[u] = (live && u < T) ? lg[(size_t)u * stride] : NEG;
  // emission factor of the NEXT step, computed one step ahead so its max-reduction is off the recursion's chain
  float mx = warp_max(lbuf[0]);
  float em = live ? exp2f((lbuf[0] - mx) * L2E) : 0.f;
  for (int t0 = 0; t0 < T; t0 += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u;
      if (t >= T) break;
      const float e_t = em, mx_t = mx;
      lbuf[u] = (live && t + 4 < T) ? lg[(size_t)(t + 4) * stride] : NEG;
      if (t + 1 < T) {
        const float ln = lbuf[(u + 1) & 3];
        mx = warp_max(ln);
        em = live ? exp2f((ln - mx) * L2E) : 0.f;
      }
      va[lane] = ah;
      __syncwarp();
      const float uj = dot(va, ac) * e_t;
      __syncwarp();
      const float c = warp_sum(uj);
      ah = uj / c;
      lzacc += (double)(logf(c) + mx_t);
      if (live) pb[(size_t)t * stride] = ah;
    }
  }
  if (lane == 0) logZ[s] = (float)lzacc;
  __syncwarp();

  // ---- backward smoother (row `lane` of A and this lane's row of SEzz only live from here on: register budget)
  float ar[KP], zz[KP];
#pragma unroll
  for (int i = 0; i < KP; ++i) {
    ar[i] = (live && i < K) ? exp2f(Atr[lane * K + i] * L2E) : 0.f;
    zz[i] = 0.f;
  }
  float gm = ah;                                             // smoothed marginal of step t+1 (= filtered at T-1)
  float fbuf[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) fbuf[u] = (live && T - 2 - u >= 0) ? pb[(size_t)(T - 2 - u) * stride] : 0.f;
  for (int t0 = T - 2; t0 >= -1; t0 -= 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 - u;
      if (t < -1) break;
      const float f = (t >= 0) ? fbuf[u] : pi0;              // t = -1: the initial-state step (HMM.py:94-98)
      fbuf[u] = (live && t - 4 >= 0) ? pb[(size_t)(t - 4) * stride] : 0.f;
      va[lane] = f;
      __syncwarp();
      const float nj = dot(va, ac);                          // lane j: sum_i f_i A_ij
      const float cj = nj > 0.f ? gm / nj : 0.f;
      vc[lane] = cj;
      __syncwarp();
      const float gn = f * dot(vc, ar);                      // lane i: f_i sum_j A_ij c_j
      const float inv = 1.f / warp_sum(gn);
      const float w = f * inv;
#pragma unroll
      for (int j = 0; j < KP; j += 4) {                      // SEzz_ij += A_ij c_j f_i / tot (c re-read: 32 fewer live registers)
        const float4 c4 = *reinterpret_cast<const float4*>(vc + j);
        zz[j] = fmaf(ar[j] * w, c4.x, zz[j]); zz[j + 1] = fmaf(ar[j + 1] * w, c4.y, zz[j + 1]);
        zz[j + 2] = fmaf(ar[j + 2] * w, c4.z, zz[j + 2]); zz[j + 3] = fmaf(ar[j + 3] * w, c4.w, zz[j + 3]);
      }
      __syncwarp();
      if (t >= 0) {
        gm = gn * inv;
        if (live) pb[(size_t)t * stride] = gm;
      } else if (live) {
        SEz0[(size_t)s * K + lane] = gn * inv;
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
  static const int force_log = [] { const char* e = getenv("VBMP_HMM_LOG"); return e ? atoi(e) : 0; }();
  if (ptemp == 1.0f && !force_log) {
    if (K <= 8) hmm_fb_lin_kernel<8><<<grid, wpb * 32, 0, st>>>(logits, trans, init, T, S, G, K, p, SEzz, SEz0, logZ);
    else if (K <= 16) hmm_fb_lin_kernel<16><<<grid, wpb * 32, 0, st>>>(logits, trans, init, T, S, G, K, p, SEzz, SEz0, logZ);
    else hmm_fb_lin_kernel<32><<<grid, wpb * 32, 0, st>>>(logits, trans, init, T, S, G, K, p, SEzz, SEz0, logZ);
    return check_launch("hmm_fb_lin");
  }
  if (K <= 8) hmm_fb_log_kernel<8><<<grid, wpb * 32, 0, st>>>(logits, trans, init, T, S, G, K, ip, p, SEzz, SEz0, logZ);
  else if (K <= 16) hmm_fb_log_kernel<16><<<grid, wpb * 32, 0, st>>>(logits, trans, init, T, S, G, K, ip, p, SEzz, SEz0, logZ);
  else hmm_fb_log_kernel<32><<<grid, wpb * 32, 0, st>>>(logits, trans, init, T, S, G, K, ip, p, SEzz, SEz0, logZ);
  return check_launch("hmm_fb");
}

}  // namespace vbmp
