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
// bound: 8.2 ms at 4096 sequences x T = 1024, K = 32).  The same quantities in the probability domain need K FMAs per
// matrix-vector product and ONE exponential per lane per step (the emission term):
//   forward   a_t[j] = (sum_i a_{t-1}[i] A_ij) e_t[j] / d_t,   e_t[j] = 2^((l_t[j] - max_j l_t) log2 e),  d_t = sum_j a_{t-1}[j]
//             logZ = sum_t (log d_t + max_j l_t) + log sum_j a_{T-1}[j]
//             (a_t is the filtered marginal up to a warp-uniform scale: the divisor is the PREVIOUS step's sum, so the
//             reduction runs beside the matrix-vector product instead of behind it; sum_j a_t[j] is the one-step
//             likelihood ratio, in (0, 1])
//   backward  n_j = sum_i f_i A_ij,  c_j = g_j / n_j,  g'_i = f_i sum_j A_ij c_j,  tot = sum_i g'_i,
//             SEzz_ij += A_ij c_j f_i / tot,  p_t = g' / tot,  g <- g'            (g = smoothed marginal of step t+1)
// which is models/HMM.py:72-105 term by term: xi_ij = f_i A_ij / n_j g_j, fw[t] = lse_j xi, SEzz += exp(xi - lse_ij xi),
// p = softmax(fw); like the reference the carried g' is not renormalised (its sum is 1 up to rounding), so only
// c_j -> matrix-vector product -> f_i * . sits on the recursion's dependency chain; n_j (a function of the stored
// filtered values only), tot, p_t and the SEzz update are off it.  A_ij is factored out of the SEzz sum over time.
// Vectors cross lanes through a 128-byte shared-memory line per warp (one store + KP/4 broadcast 16-byte loads instead
// of KP shuffles); products run as packed fp32x2 FMAs.  States whose probability underflows fp32 here are below e^-87 in
// the reference's p as well.  logZ is accumulated in fp64 (T terms of size |l|).  ptemp != 1 raises the smoothed
// marginals to 1/ptemp, where an underflowed state could matter: those calls keep the log-space kernel.
// (K = 32: three register arrays of 32 — column and row of A, the lane's row of SEzz — do not fit 128 registers; three
// resident blocks of 4 warps instead of four avoid the spills)
template <int KP>
__global__ void __launch_bounds__(128, KP == 32 ? 3 : 4) hmm_fb_lin_kernel(const float* __restrict__ logits,
                                                                        const float* __restrict__ trans,
                                                                        const float* __restrict__ init, int T, long long S,
                                                                        int G, int K, float* __restrict__ p,
                                                                        float* __restrict__ SEzz, float* __restrict__ SEz0,
                                                                        float* __restrict__ logZ) {
  constexpr int PF = KP == 32 ? 6 : 8;                       // prefetch distance (steps) of the per-step global loads
  constexpr int H = KP / 2;
  __shared__ __align__(16) float sh[4][3][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long s = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
  if (s >= S) return;
  const int g = (int)(s % G);
  const bool live = lane < K;
  const float* Atr = trans + (size_t)g * K * K;
  constexpr float L2E = 1.4426950408889634f;
  float2 ac[H];                                              // column `lane` of A = exp(log transition), packed pairs
#pragma unroll
  for (int i = 0; i < H; ++i)
    ac[i] = make_float2((live && 2 * i < K) ? exp2f(Atr[(2 * i) * K + lane] * L2E) : 0.f,
                        (live && 2 * i + 1 < K) ? exp2f(Atr[(2 * i + 1) * K + lane] * L2E) : 0.f);
  const float pi0 = live ? exp2f(init[(size_t)g * K + lane] * L2E) : 0.f;
  float* va = sh[wib][0];
  float* vb = sh[wib][1];
  float* vc = sh[wib][2];
  const size_t stride = (size_t)S * K;
  const float* lg = logits + (size_t)s * K + lane;
  float* pb = p + (size_t)s * K + lane;
  const float NEG = -INFINITY;

  // sum_i v[i] m[i] against a vector in shared memory: KP/4 broadcast 16-byte loads, packed FMAs, two chains
  auto dot = [&](const float* v, const float2 (&m)[H]) {
    float2 q0 = make_float2(0.f, 0.f), q1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < KP; i += 4) {
      const float4 x = *reinterpret_cast<const float4*>(v + i);
      q0 = __ffma2_rn(make_float2(x.x, x.y), m[i / 2], q0);
      q1 = __ffma2_rn(make_float2(x.z, x.w), m[i / 2 + 1], q1);
    }
    return (q0.x + q0.y) + (q1.x + q1.y);
  };

  // ---- forward filter
  float at = pi0;
  float d = warp_sum(pi0);                                   // divisor of the coming step = sum of the current vector
  double lzacc = 0.0;
  float lbuf[PF];
#pragma unroll
  for (int u = 0; u < PF; ++u) lbuf[u] = (live && u < T) ? lg[(size_t)u * stride] : NEG;
  // emission factor of the NEXT step, computed one step ahead so its max-reduction is off the recursion's chain
  float mx = warp_max(lbuf[0]);
  float em = live ? exp2f((lbuf[0] - mx) * L2E) : 0.f;
  for (int t0 = 0; t0 < T; t0 += PF) {
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int t = t0 + u;
      if (t >= T) break;
      const float e_t = em, mx_t = mx;
      lbuf[u] = (live && t + PF < T) ? lg[(size_t)(t + PF) * stride] : NEG;
      if (t + 1 < T) {
        const float ln = lbuf[(u + 1) % PF];
        mx = warp_max(ln);
        em = live ? exp2f((ln - mx) * L2E) : 0.f;
      }
      float* v = (u & 1) ? vb : va;                          // alternate lines: one __syncwarp per step
      v[lane] = at;
      __syncwarp();
      const float n = dot(v, ac);
      lzacc += (double)(__logf(d) + mx_t);
      at = n * (e_t * __frcp_rn(d));
      d = warp_sum(at);                                      // used by the NEXT step, beside its matrix-vector product
      if (live) pb[(size_t)t * stride] = at;
    }
  }
  if (lane == 0) logZ[s] = (float)(lzacc + (double)logf(d));
  __syncwarp();

  // ---- backward smoother (row `lane` of A and this lane's row of SEzz only live from here on: register budget)
  float2 ar[H], zz[H];
#pragma unroll
  for (int i = 0; i < H; ++i) {
    ar[i] = make_float2((live && 2 * i < K) ? exp2f(Atr[lane * K + 2 * i] * L2E) : 0.f,
                        (live && 2 * i + 1 < K) ? exp2f(Atr[lane * K + 2 * i + 1] * L2E) : 0.f);
    zz[i] = make_float2(0.f, 0.f);
  }
  float gm = at / d;                                         // smoothed marginal of step t+1 (= filtered, normalised, at T-1)
  if (live) pb[(size_t)(T - 1) * stride] = gm;
  float fbuf[PF];
#pragma unroll
  for (int u = 0; u < PF; ++u) fbuf[u] = (live && T - 2 - u >= 0) ? pb[(size_t)(T - 2 - u) * stride] : 0.f;
  // n_j of the coming step, one step ahead (it depends on the stored filtered values only)
  float fcur = (T >= 2) ? fbuf[0] : pi0;
  va[lane] = fcur;
  __syncwarp();
  float ncur = dot(va, ac);
  for (int t0 = T - 2; t0 >= -1; t0 -= PF) {
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int t = t0 - u;
      if (t < -1) break;
      const float f = fcur, nj = ncur;                       // t = -1: the initial-state step (f = pi_0, HMM.py:94-98)
      fbuf[u] = (live && t - PF >= 0) ? pb[(size_t)(t - PF) * stride] : 0.f;
      // ---- the recursion's chain: c -> row -> g'
      const float cj = nj > 0.f ? gm / nj : 0.f;
      vc[lane] = cj;
      // (off the chain) next step's f and n_j; lines va / vb alternate so one barrier covers both stores
      float fnext = 0.f;
      float* v = (u & 1) ? va : vb;
      if (t >= 0) {
        fnext = (t >= 1) ? fbuf[(u + 1) % PF] : pi0;
        v[lane] = fnext;
      }
      __syncwarp();
      const float gn = f * dot(vc, ar);                      // lane i: f_i sum_j A_ij c_j
      if (t >= 0) ncur = dot(v, ac);
      // ---- off the chain: normaliser, outputs, expected transition counts
      const float inv = __frcp_rn(warp_sum(gn));
      const float w = f * inv;
      const float2 w2 = make_float2(w, w);
#pragma unroll
      for (int j = 0; j < KP; j += 4) {                      // zz_ij += c_j f_i / tot (A_ij is applied once, at the end)
        const float4 c4 = *reinterpret_cast<const float4*>(vc + j);
        zz[j / 2] = __ffma2_rn(w2, make_float2(c4.x, c4.y), zz[j / 2]);
        zz[j / 2 + 1] = __ffma2_rn(w2, make_float2(c4.z, c4.w), zz[j / 2 + 1]);
      }
      __syncwarp();                                          // vc is rewritten by the next step
      if (t >= 0) {
        if (live) pb[(size_t)t * stride] = gn * inv;
        gm = gn;
        fcur = fnext;
      } else if (live) {
        SEz0[(size_t)s * K + lane] = gn * inv;
      }
    }
  }
  if (live) {
    float* out = SEzz + ((size_t)s * K + lane) * K;
#pragma unroll
    for (int j = 0; j < KP; ++j)
      if (j < K) out[j] = ((j & 1) ? zz[j / 2].y * ar[j / 2].y : zz[j / 2].x * ar[j / 2].x);
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
