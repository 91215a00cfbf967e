// K3 (general-shape CUDA-core variant): responsibility-weighted Gram statistics
//
//   Gram[g,k] = sum_n r[n,pg[g],k] * [z;1][z;1]^T ,   z = z_{n,xg[g]}  (D+1 x D+1, row-major)
//
// whose blocks are the reference's SExx / SEx / N (dists/NormalInverseWishart.py:80-84) and
// SExx / SEyx / SEyy / SEx / SEy / N (transforms/MatrixNormalWishart.py:185-202) with z = [x; y].
//
// A CTA owns CK components x one sample split; each thread owns an 8x8 tile of one component's
// Gram matrix.  fp32 accumulators are flushed into a second level every 256 samples and a split
// never spans more than 65 536 samples, so no accumulation chain exceeds 256 terms before the
// fixed-order fp64 reduction over splits (long fp32 chains miss 1e-4).
#include "common.cuh"

namespace vbmp {


__device__ inline void cp_async4_zfill(void* smem, const void* gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ inline void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ inline void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

#define GRAM_SC 64      // samples per shared-memory chunk
#define GRAM_FL 4       // chunks per first-level flush (256 samples)

__global__ void __launch_bounds__(256, 1) gram_simt_kernel(GramArgs a) {
  extern __shared__ __align__(16) float smf[];
  const int tid = threadIdx.x;
  const int Dp = a.Dp, IT = Dp >> 3, TPC = IT * IT, CK = 256 / TPC;
  const int D = a.d0 + a.d1, D1 = D + 1;
  float* Zc = smf;                           // [2][SC][Dp]
  float* Rc = Zc + 2 * GRAM_SC * Dp;         // [2][SC][CK]
  const int g = blockIdx.z, split = blockIdx.y;
  const int k0 = blockIdx.x * CK;
  const int cs = tid / TPC, t = tid % TPC, it = t / IT, jt = t % IT;
  const int k = k0 + cs;
  const int xgi = a.xg ? a.xg[g] : 0, pgi = a.pg ? a.pg[g] : 0;
  const long long nb = (long long)split * a.S_per;
  long long ne = nb + a.S_per; if (ne > a.N) ne = a.N;
  const int half = Dp >> 1;
  const int i_lo = 4 * it, i_hi = half + 4 * it, j_lo = 4 * jt, j_hi = half + 4 * jt;

  auto load_chunk = [&](long long c0, int buf) {
    float* zd = Zc + buf * GRAM_SC * Dp;
    for (int e = tid; e < GRAM_SC * Dp; e += 256) {
      const int s = e / Dp, i = e % Dp;
      const long long n = c0 + s;
      const bool ok = (n < ne) && (i < D);
      const float* src = a.z0;
      if (ok) {
        const long long row = n * a.GX + xgi;
        src = (i < a.d0) ? a.z0 + row * a.d0 + i : a.z1 + row * a.d1 + (i - a.d0);
      }
      cp_async4_zfill(zd + e, src, ok);
    }
    float* rd = Rc + buf * GRAM_SC * CK;
    for (int e = tid; e < GRAM_SC * CK; e += 256) {
      const int s = e / CK, c = e % CK;
      const long long n = c0 + s;
      const bool ok = (n < ne) && (k0 + c < a.K);
      if (a.p) {
        const float* src = ok ? a.p + ((n * a.GP + pgi) * a.K + k0 + c) : a.p;
        cp_async4_zfill(rd + e, src, ok);
      } else {
        rd[e] = ok ? 1.f : 0.f;
      }
    }
    cp_commit();
  };

  float acc1[8][8], acc2[8][8], sx1[8], sx2[8];
  float n1 = 0.f, n2 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sx1[i] = 0.f; sx2[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc1[i][j] = 0.f; acc2[i][j] = 0.f; }
  }

  const long long nchunks = (ne > nb) ? (ne - nb + GRAM_SC - 1) / GRAM_SC : 0;
  if (nchunks > 0) load_chunk(nb, 0);
  for (long long c = 0; c < nchunks; ++c) {
    const int buf = (int)(c & 1);
    if (c + 1 < nchunks) { load_chunk(nb + (c + 1) * GRAM_SC, buf ^ 1); cp_wait<1>(); } else { cp_wait<0>(); }
    __syncthreads();
    const float* zc = Zc + buf * GRAM_SC * Dp;
    const float* rc = Rc + buf * GRAM_SC * CK + cs;
#pragma unroll 2
    for (int s = 0; s < GRAM_SC; ++s) {
      const float r = rc[s * CK];
      const float4 a0 = *reinterpret_cast<const float4*>(zc + s * Dp + i_lo);
      const float4 a1 = *reinterpret_cast<const float4*>(zc + s * Dp + i_hi);
      const float4 b0 = *reinterpret_cast<const float4*>(zc + s * Dp + j_lo);
      const float4 b1 = *reinterpret_cast<const float4*>(zc + s * Dp + j_hi);
      const float av[8] = {r * a0.x, r * a0.y, r * a0.z, r * a0.w, r * a1.x, r * a1.y, r * a1.z, r * a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc1[i][j] = fmaf(av[i], bv[j], acc1[i][j]);
      if (jt == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sx1[i] += av[i];
        if (it == 0) n1 += r;
      }
    }
    if ((c % GRAM_FL) == GRAM_FL - 1 || c + 1 == nchunks) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sx2[i] += sx1[i]; sx1[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc2[i][j] += acc1[i][j]; acc1[i][j] = 0.f; }
      }
      n2 += n1; n1 = 0.f;
    }
    __syncthreads();
  }

  if (k < a.K) {
    float* out = a.part + (((size_t)split * a.G + g) * a.K + k) * (size_t)D1 * D1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = (i < 4 ? i_lo + i : i_hi + i - 4);
      if (row >= D) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = (j < 4 ? j_lo + j : j_hi + j - 4);
        if (col < D) out[(size_t)row * D1 + col] = acc2[i][j];
      }
      if (jt == 0) { out[(size_t)row * D1 + D] = sx2[i]; out[(size_t)D * D1 + row] = sx2[i]; }
    }
    if (it == 0 && jt == 0) out[(size_t)D * D1 + D] = n2;
  }
}

// gram[e] = sum_s part[s][e] in fp64, fixed order (deterministic run to run).
__global__ void gram_reduce_kernel(const float* __restrict__ part, int splits, size_t per, float* __restrict__ gram) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= per) return;
  double acc = 0.0;
  for (int s = 0; s < splits; ++s) acc += (double)part[(size_t)s * per + e];
  gram[e] = (float)acc;
}

static size_t gram_simt_smem(int Dp) {
  const int IT = Dp / 8, CK = 256 / (IT * IT);
  return (size_t)(2 * GRAM_SC * Dp + 2 * GRAM_SC * CK) * sizeof(float);
}

int gram_simt_plan(long long N, int G, int K, int Dp, long long* S_per, int* splits) {
  const int IT = Dp / 8, CK = 256 / (IT * IT);
  const long long ctas_per_split = (long long)G * cdiv(K, CK);
  long long want = (148 * 4 + ctas_per_split - 1) / ctas_per_split;      // ~4 waves of CTAs
  const long long min_for_chain = (N + 65535) / 65536;                   // <= 65 536 samples per split
  if (want < min_for_chain) want = min_for_chain;
  // >= 256 samples per split (one first-level block).  This bound only binds for small N with few components, where the
  // grid is tiny anyway: at N = 10 000, d = 2, K = 20 the former 1024 gave 10 CTAs (0.30 ms, the longest kernel of an EM
  // iteration there); the partials stay small (splits x G x K x (D+1)^2 floats) because want is capped at ~4 waves
  const long long max_useful = (N + 255) / 256;
  if (want > max_useful) want = max_useful;
  if (want < 1) want = 1;
  long long sp = (N + want - 1) / want;
  sp = ((sp + GRAM_SC - 1) / GRAM_SC) * GRAM_SC;
  if (sp < GRAM_SC) sp = GRAM_SC;
  *S_per = sp;
  *splits = (int)((N + sp - 1) / sp);
  if (*splits < 1) *splits = 1;
  return 0;
}

int launch_gram_simt(const GramArgs& a, cudaStream_t st) {
  const int IT = a.Dp / 8, CK = 256 / (IT * IT);
  const size_t smem = gram_simt_smem(a.Dp);
  cudaFuncSetAttribute(gram_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((unsigned)cdiv(a.K, CK), (unsigned)a.splits, (unsigned)a.G);
  gram_simt_kernel<<<grid, 256, smem, st>>>(a);
  return check_launch("gram_simt");
}

int launch_gram_reduce(const float* part, int splits, size_t per, float* gram, cudaStream_t st) {
  gram_reduce_kernel<<<(unsigned)((per + 255) / 256), 256, 0, st>>>(part, splits, per, gram);
  return check_launch("gram_reduce");
}

}  // namespace vbmp
