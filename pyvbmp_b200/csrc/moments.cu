// Mixture-of-experts predictive moments: the per-sample part of MixtureofLinearTransforms.predict
// (transforms/MixtureofLinearTransforms.py:100-106)
//
//     mu_s    = sum_k p_sk m_sk                                   m_sk = E[y | x_s, component k]   (n-vector)
//     Sigma_s = base_s + sum_k p_sk m_sk m_sk^T - mu_s mu_s^T      base_s = sum_k p_sk ESigma_k
//
// The component means of a block of samples are ONE GEMM with a shared operand ((rows x p') (p' x K n)) and so is base
// ((rows x K) (K x n^2)); what is left is a weighted rank-K update PER SAMPLE — (n x K) (K x n) with no operand shared
// between samples — which as a batch of tiny GEMMs was 60 % of predict's time (35 ms per 1 Mi inputs at n = p = 32, K = 64).
// Here one warp owns one sample and its 32 lanes tile Sigma_s in 4 x 8 register blocks (lane = (row group of 4, column
// group of 8), n <= 32, n % 4 == 0): per component a lane loads the 4 + 8 mean entries of its block's rows and columns —
// 48 bytes, 1.5 KB per warp — and does 32 FMAs.  What bounds such a kernel is the load path's delivery rate into registers
// (128 B per cycle per SM), not the flops: the first version kept a full ROW of Sigma_s per lane and therefore every lane
// read the whole mean vector of every component (132 bytes per lane, 4.2 KB per warp and component — 19.9 ms per 1 Mi
// samples at n = 32, K = 64, measured).  Other n take that simpler row-per-lane kernel.
#include <cstdlib>
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

// lane = (rg, cg): rows 4 rg .. 4 rg + 3 (rg = lane / 4), columns 8 cg .. 8 cg + 7 (cg = lane % 4) of the 32 x 32 frame.
// The means of a sample (K n floats, 8 KB at K = 64, n = 32) are staged through shared memory with cp.async, one chunk of
// <= MM_KC components ahead per warp: read straight from global memory the same arithmetic ran at 0.9 TB/s (19 ms per 1 Mi
// samples) — a warp had two 128-byte rows in flight, far too few bytes to cover DRAM latency.
constexpr int MM_KC = 64;              // components per staged chunk
constexpr int MM_WARPS = 4;            // warps per CTA (2 x MM_KC x 32 floats = 16 KB of staging per warp)

__device__ __forceinline__ void mm_cp_async16(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void mm_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N_> __device__ __forceinline__ void mm_cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N_)); }

__global__ void __launch_bounds__(MM_WARPS * 32) moe_moments_tile_kernel(const float* __restrict__ mean, const float* __restrict__ p,
                                                                         const float* __restrict__ base, long long N, int K, int n,
                                                                         float* __restrict__ mu, float* __restrict__ Sigma) {
  extern __shared__ __align__(16) float mm_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long w0 = (long long)blockIdx.x * MM_WARPS + wib;
  const long long nw = (long long)gridDim.x * MM_WARPS;
  const int r0 = (lane >> 2) * 4, c0 = (lane & 3) * 8;
  const bool rok = r0 < n, cok0 = c0 < n, cok1 = c0 + 4 < n;         // n % 4 == 0: a float4 is all inside or all outside
  float* buf = mm_smem + (size_t)wib * 2 * MM_KC * n;
  const int nch = (K + MM_KC - 1) / MM_KC;
  const long long nsamp = w0 < N ? (N - w0 + nw - 1) / nw : 0;       // samples of this warp
  const long long items = nsamp * nch;
  auto prefetch = [&](long long t) {
    const long long s = w0 + (t / nch) * nw;
    const int kc0 = (int)(t % nch) * MM_KC;
    const int cnt = min(MM_KC, K - kc0) * n;                         // floats, a multiple of 4
    const float* src = mean + ((size_t)s * K + kc0) * n;
    float* dst = buf + (size_t)(t & 1) * MM_KC * n;
    for (int e = lane * 4; e < cnt; e += 128) mm_cp_async16(dst + e, src + e);
    mm_cp_async_commit();
  };
  float2 acc[4][4];                      // 4 rows x 8 columns as packed pairs: one FFMA2 per row and column pair
  float4 mur, muc0, muc1;
  if (items > 0) prefetch(0);
  for (long long t = 0; t < items; ++t) {
    const long long s = w0 + (t / nch) * nw;
    const int ch = (int)(t % nch), kc0 = ch * MM_KC, kn = min(MM_KC, K - kc0);
    if (t + 1 < items) { prefetch(t + 1); mm_cp_async_wait<1>(); } else { mm_cp_async_wait<0>(); }
    __syncwarp();
    if (ch == 0) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[a][j] = make_float2(0.f, 0.f);
      mur = make_float4(0.f, 0.f, 0.f, 0.f); muc0 = mur; muc1 = mur;
    }
    const float* mb_ = buf + (size_t)(t & 1) * MM_KC * n;
    const float* ps = p + (size_t)s * K + kc0;
#pragma unroll 4
    for (int k = 0; k < kn; ++k) {
      const float pk = __ldg(ps + k);
      const float* mk = mb_ + (size_t)k * n;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 mr = rok ? *reinterpret_cast<const float4*>(mk + r0) : z;
      const float4 ma = cok0 ? *reinterpret_cast<const float4*>(mk + c0) : z;
      const float4 mb = cok1 ? *reinterpret_cast<const float4*>(mk + c0 + 4) : z;
      const float w[4] = {pk * mr.x, pk * mr.y, pk * mr.z, pk * mr.w};
      const float2 c[4] = {make_float2(ma.x, ma.y), make_float2(ma.z, ma.w), make_float2(mb.x, mb.y), make_float2(mb.z, mb.w)};
      mur.x += w[0]; mur.y += w[1]; mur.z += w[2]; mur.w += w[3];
      muc0.x = fmaf(pk, ma.x, muc0.x); muc0.y = fmaf(pk, ma.y, muc0.y); muc0.z = fmaf(pk, ma.z, muc0.z); muc0.w = fmaf(pk, ma.w, muc0.w);
      muc1.x = fmaf(pk, mb.x, muc1.x); muc1.y = fmaf(pk, mb.y, muc1.y); muc1.z = fmaf(pk, mb.z, muc1.z); muc1.w = fmaf(pk, mb.w, muc1.w);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[a][j] = __ffma2_rn(make_float2(w[a], w[a]), c[j], acc[a][j]);
    }
    __syncwarp();                                                      // the buffer is refilled two items on
    if (ch != nch - 1) continue;
    if (rok && (lane & 3) == 0) *reinterpret_cast<float4*>(mu + (size_t)s * n + r0) = mur;
    const float mrow[4] = {mur.x, mur.y, mur.z, mur.w};
    const float mcol[8] = {muc0.x, muc0.y, muc0.z, muc0.w, muc1.x, muc1.y, muc1.z, muc1.w};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (!rok) break;
      float* So = Sigma + ((size_t)s * n + r0 + a) * n + c0;
      const float* Bo = base ? base + ((size_t)s * n + r0 + a) * n + c0 : nullptr;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (!(h ? cok1 : cok0)) continue;
        float4 b4 = Bo ? *reinterpret_cast<const float4*>(Bo + 4 * h) : make_float4(0.f, 0.f, 0.f, 0.f);
        b4.x += acc[a][2 * h].x - mrow[a] * mcol[4 * h];
        b4.y += acc[a][2 * h].y - mrow[a] * mcol[4 * h + 1];
        b4.z += acc[a][2 * h + 1].x - mrow[a] * mcol[4 * h + 2];
        b4.w += acc[a][2 * h + 1].y - mrow[a] * mcol[4 * h + 3];
        *reinterpret_cast<float4*>(So + 4 * h) = b4;
      }
    }
  }
}

// ---- n = 16 / 32: the per-sample rank-K update as a tensor-core SYRK -------------------------------------------------
// Sigma_s - base_s + mu_s mu_s^T = M^T diag(p) M with M = the sample's (K x n) means: per warp and sample one
// (n x K) (K x n) product on mma.sync.m16n8k8 (TF32, 3-term split: fp32-grade).  Both operands are the SAME staged rows
// of M — the A fragment (scaled by p) and the B fragment of a K-step need exactly the elements M[8 ks + t (+4)][g + 8 c],
// c = 0 .. n/8 - 1 — so a K-step costs n/4 four-byte shared-memory loads per lane (row stride 40 floats: conflict free)
// instead of the 3 sixteen-byte loads PER COMPONENT of the register-tiled kernel above, and no FMAs at all; mu comes out
// of one extra column tile whose B operand is the constant e_0.  What is left is the kernel's traffic (means in, base in,
// Sigma out).
constexpr int MQ_KC = 32;              // components per staged chunk
constexpr int MQ_WARPS = 4;
constexpr int MQ_RS = 40;              // row stride (floats) of a staged chunk

__device__ __forceinline__ void mq_split(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mq_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int NN>                      // n = 16 or 32
__global__ void __launch_bounds__(MQ_WARPS * 32) moe_moments_mma_kernel(const float* __restrict__ mean, const float* __restrict__ p,
                                                                        const float* __restrict__ base, long long N, int K,
                                                                        float* __restrict__ mu, float* __restrict__ Sigma) {
  constexpr int MT = NN / 16, NT = NN / 8, NC = NN / 8;          // row tiles, column tiles, distinct columns per lane
  extern __shared__ __align__(16) float mq_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const long long w0 = (long long)blockIdx.x * MQ_WARPS + wib;
  const long long nw = (long long)gridDim.x * MQ_WARPS;
  float* buf = mq_smem + (size_t)wib * (2 * MQ_KC * MQ_RS + 32);
  float* mus = buf + 2 * MQ_KC * MQ_RS;                           // the sample's mu, for the final mu mu^T
  const int nch = (K + MQ_KC - 1) / MQ_KC;
  const long long nsamp = w0 < N ? (N - w0 + nw - 1) / nw : 0;
  const long long items = nsamp * nch;
  auto prefetch = [&](long long it) {
    const long long s = w0 + (it / nch) * nw;
    const int kc0 = (int)(it % nch) * MQ_KC;
    const int rows = min(MQ_KC, K - kc0);
    const float* src = mean + ((size_t)s * K + kc0) * NN;
    float* dst = buf + (size_t)(it & 1) * MQ_KC * MQ_RS;
    for (int e = lane; e < MQ_KC * (NN / 4); e += 32) {             // 16-byte pieces; rows past K are zero filled
      const int r = e / (NN / 4), c4 = (e % (NN / 4)) * 4;
      if (r < rows) mm_cp_async16(dst + r * MQ_RS + c4, src + (size_t)r * NN + c4);
      else *reinterpret_cast<float4*>(dst + r * MQ_RS + c4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    mm_cp_async_commit();
  };
  float acc[MT][NT][4], am[MT][4];
  const uint32_t one = (g == 0) ? 0x3f800000u : 0u;                // B operand of the mu tile: column 0 = 1
  if (items > 0) prefetch(0);
  for (long long it = 0; it < items; ++it) {
    const long long s = w0 + (it / nch) * nw;
    const int ch = (int)(it % nch), kc0 = ch * MQ_KC;
    if (it + 1 < items) { prefetch(it + 1); mm_cp_async_wait<1>(); } else { mm_cp_async_wait<0>(); }
    __syncwarp();
    if (ch == 0) {
#pragma unroll
      for (int a = 0; a < MT; ++a) {
#pragma unroll
        for (int b = 0; b < NT; ++b)
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[a][b][v] = 0.f;
#pragma unroll
        for (int v = 0; v < 4; ++v) am[a][v] = 0.f;
      }
    }
    const float* mb_ = buf + (size_t)(it & 1) * MQ_KC * MQ_RS;
    const float pl = (kc0 + lane < K) ? __ldg(p + (size_t)s * K + kc0 + lane) : 0.f;     // lane k holds p of component kc0 + k
#pragma unroll
    for (int ks = 0; ks < MQ_KC / 8; ++ks) {
      const float p0 = __shfl_sync(0xffffffffu, pl, 8 * ks + t), p1 = __shfl_sync(0xffffffffu, pl, 8 * ks + t + 4);
      // M[8 ks + t][g + 8 c], M[8 ks + t + 4][g + 8 c]: the B fragments as they are, the A fragments scaled by p
      uint32_t bh[NC][2], bl[NC][2], ah[NC][2], al[NC][2];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const float m0 = mb_[(8 * ks + t) * MQ_RS + g + 8 * c], m1 = mb_[(8 * ks + t + 4) * MQ_RS + g + 8 * c];
        mq_split(m0, bh[c][0], bl[c][0]);
        mq_split(m1, bh[c][1], bl[c][1]);
        mq_split(p0 * m0, ah[c][0], al[c][0]);
        mq_split(p1 * m1, ah[c][1], al[c][1]);
      }
#pragma unroll
      for (int a = 0; a < MT; ++a) {
        // rows g + 16 a (c = 2a) and g + 8 + 16 a (c = 2a + 1): a0 = (g, t), a1 = (g + 8, t), a2 = (g, t + 4), a3 = (g + 8, t + 4)
        const uint32_t fh[4] = {ah[2 * a][0], ah[2 * a + 1][0], ah[2 * a][1], ah[2 * a + 1][1]};
        const uint32_t fl[4] = {al[2 * a][0], al[2 * a + 1][0], al[2 * a][1], al[2 * a + 1][1]};
#pragma unroll
        for (int b = 0; b < NT; ++b) {
          mq_mma(acc[a][b], fl, bh[b][0], bh[b][1]);              // small terms first
          mq_mma(acc[a][b], fh, bl[b][0], bl[b][1]);
          mq_mma(acc[a][b], fh, bh[b][0], bh[b][1]);
        }
        mq_mma(am[a], fl, one, one);
        mq_mma(am[a], fh, one, one);
      }
    }
    __syncwarp();                                                  // the buffer is refilled two items on
    if (ch != nch - 1) continue;
    // mu[i]: column 0 of the mu tile lives in the lanes with t == 0 (c0: row g, c2: row g + 8)
    if (t == 0) {
#pragma unroll
      for (int a = 0; a < MT; ++a) { mus[16 * a + g] = am[a][0]; mus[16 * a + g + 8] = am[a][2]; }
    }
    __syncwarp();
    if (lane < NN) mu[(size_t)s * NN + lane] = mus[lane];
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = 16 * a + g + 8 * h;
        const float mi = mus[i];
#pragma unroll
        for (int b = 0; b < NT; ++b) {
          const int j = 8 * b + 2 * t;
          const size_t o = ((size_t)s * NN + i) * NN + j;
          float2 r = base ? *reinterpret_cast<const float2*>(base + o) : make_float2(0.f, 0.f);
          r.x += acc[a][b][2 * h] - mi * mus[j];
          r.y += acc[a][b][2 * h + 1] - mi * mus[j + 1];
          *reinterpret_cast<float2*>(Sigma + o) = r;
        }
      }
    __syncwarp();                                                  // mus is rewritten by the next sample
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
  const bool aligned = ((size_t)mean % 16 == 0) && ((size_t)Sigma % 16 == 0) && ((size_t)mu % 16 == 0) && (!base || (size_t)base % 16 == 0);
  static const int use_mma = [] { const char* e = getenv("VBMP_MOE_MMA"); return e ? atoi(e) : 1; }();
  if ((n == 32 || n == 16) && aligned && use_mma) {
    const size_t smem = (size_t)MQ_WARPS * (2 * MQ_KC * MQ_RS + 32) * sizeof(float);    // 41.5 KB: five CTAs per SM
    long long tb = (N + MQ_WARPS - 1) / MQ_WARPS;
    const long long tcap = (long long)num_sms() * 5 * 4;
    if (tb > tcap) tb = tcap;
    if (n == 32) {
      cudaFuncSetAttribute(moe_moments_mma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      moe_moments_mma_kernel<32><<<(unsigned)tb, MQ_WARPS * 32, smem, st>>>(mean, p, base, N, K, mu, Sigma);
    } else {
      cudaFuncSetAttribute(moe_moments_mma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      moe_moments_mma_kernel<16><<<(unsigned)tb, MQ_WARPS * 32, smem, st>>>(mean, p, base, N, K, mu, Sigma);
    }
    return check_launch("moe_moments_mma");
  }
  if ((n & 3) == 0 && aligned) {
    const size_t smem = (size_t)MM_WARPS * 2 * MM_KC * n * sizeof(float);       // 64 KB at n = 32: three CTAs per SM
    long long tb = (N + MM_WARPS - 1) / MM_WARPS;
    const long long tcap = (long long)num_sms() * 3 * 4;
    if (tb > tcap) tb = tcap;
    cudaFuncSetAttribute(moe_moments_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    moe_moments_tile_kernel<<<(unsigned)tb, MM_WARPS * 32, smem, st>>>(mean, p, base, N, K, n, mu, Sigma);
    return check_launch("moe_moments_tile");
  }
  switch (np) {
#define MM_CASE(V) case V: moe_moments_kernel<V><<<(unsigned)blocks, 256, 0, st>>>(mean, p, base, N, K, n, mu, Sigma); break;
    MM_CASE(4) MM_CASE(8) MM_CASE(12) MM_CASE(16) MM_CASE(20) MM_CASE(24) MM_CASE(28) MM_CASE(32)
#undef MM_CASE
    default: set_error("moe_moments: n=%d", n); return VBMP_ERR_SHAPE;
  }
  return check_launch("moe_moments");
}

}  // namespace vbmp
