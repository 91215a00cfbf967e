// Row GEMM with fp32-grade accuracy on the warp-level tensor-core path:  C[N x M] (+)= A[N x Kd] B[Kd x M] (+ bias[M])
//
// The products of the expectation-input E-step and of predict that have one operand shared by all samples but are too
// skinny for the tcgen05 kernels' tiling:
//   * component means of predict            (rows x p') (p' x K n)      transforms/MixtureofLinearTransforms.py:100-102
//   * sum_k p_k ESigma_k of predict         (rows x K)  (K x n^2)       :103
//   * covariance terms of Elog_like_given_pX_pY  (rows x n^2) (n^2 x K), (rows x p^2) (p^2 x K)
//                                                                        transforms/MatrixNormalWishart.py:236-247
// N is the sample axis (millions), Kd and M are small (33 .. 2048): these calls are bound by the bytes of A and C, and an
// fp32 SGEMM on the CUDA cores (what torch dispatches to without TF32) runs them at 0.85 - 2.5 TB/s-equivalent — 10.1 ms
// for predict's means per 1 Mi inputs.  Here every product is the 3-term TF32 split (hi*hi + hi*lo + lo*hi, error ~2^-21,
// the same scheme as the TF32 variants of K2 / K3) issued as mma.sync.m16n8k8: a CTA owns 128 rows x 128 columns, 8
// warps of 32 x 64, A and B staged through shared memory with cp.async in reduction chunks of 32, two stages.
// mma.sync cannot reach tcgen05 throughput, but for these shapes the tensor work (3 x 2 N Kd M flop) is already below the
// time it takes to stream A and C.
#include "common.cuh"

namespace vbmp {

constexpr int RG_BM = 128, RG_BN = 128, RG_KC = 32;
constexpr int RG_AS = RG_KC + 4;       // row stride of the A tile (floats): fragment loads hit 32 distinct banks
constexpr int RG_BS = RG_BN + 8;       // row stride of the B tile

__device__ __forceinline__ void rg_cp_async16(void* smem, const void* gmem, bool ok) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = ok ? 16 : 0;                                        // src-size 0: the 16 bytes are zero filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void rg_cp_async4(void* smem, const void* gmem, bool ok) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = ok ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(sa), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void rg_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N_> __device__ __forceinline__ void rg_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N_)); }

__device__ __forceinline__ void rg_split(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void rg_mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// A (N, lda) row-major, B (Kd, ldb) row-major, C (N, ldc) row-major; VEC: lda, ldb multiples of 4 and 16-byte aligned bases
template <bool VEC>
__global__ void __launch_bounds__(256) rowgemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                      const float* __restrict__ bias, float* __restrict__ C, int ldc,
                                                      long long N, int Kd, int M, int accumulate) {
  extern __shared__ __align__(16) float rg_smem[];
  float* As = rg_smem;                                   // [2][RG_BM][RG_AS]
  float* Bs = rg_smem + 2 * RG_BM * RG_AS;               // [2][RG_KC][RG_BS]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;               // warp tile: rows 32 wm .., columns 64 wn ..
  const long long row0 = (long long)blockIdx.x * RG_BM;
  const int col0 = blockIdx.y * RG_BN;
  const int nk = (Kd + RG_KC - 1) / RG_KC;

  auto load = [&](int kc, int buf) {
    const int k0 = kc * RG_KC;
    float* as = As + buf * RG_BM * RG_AS;
    float* bs = Bs + buf * RG_KC * RG_BS;
    if (VEC) {
      for (int e = tid; e < RG_BM * (RG_KC / 4); e += 256) {            // A tile: 128 rows x 8 float4
        const int r = e >> 3, c4 = (e & 7) * 4;
        const bool ok = row0 + r < N && k0 + c4 < Kd;                   // Kd % 4 == 0 on this path
        rg_cp_async16(as + r * RG_AS + c4, A + (size_t)(ok ? row0 + r : 0) * lda + (ok ? k0 + c4 : 0), ok);
      }
      for (int e = tid; e < RG_KC * (RG_BN / 4); e += 256) {            // B tile: 32 rows x 32 float4
        const int r = e >> 5, c4 = (e & 31) * 4;
        const bool ok = k0 + r < Kd && col0 + c4 < M;                   // M % 4 == 0 on this path
        rg_cp_async16(bs + r * RG_BS + c4, B + (size_t)(ok ? k0 + r : 0) * ldb + (ok ? col0 + c4 : 0), ok);
      }
    } else {
      for (int e = tid; e < RG_BM * RG_KC; e += 256) {
        const int r = e / RG_KC, c = e % RG_KC;
        const bool ok = row0 + r < N && k0 + c < Kd;
        rg_cp_async4(as + r * RG_AS + c, A + (size_t)(ok ? row0 + r : 0) * lda + (ok ? k0 + c : 0), ok);
      }
      for (int e = tid; e < RG_KC * RG_BN; e += 256) {
        const int r = e / RG_BN, c = e % RG_BN;
        const bool ok = k0 + r < Kd && col0 + c < M;
        rg_cp_async4(bs + r * RG_BS + c, B + (size_t)(ok ? k0 + r : 0) * ldb + (ok ? col0 + c : 0), ok);
      }
    }
    rg_commit();
  };

  float acc[2][8][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[i][j][v] = 0.f;

  load(0, 0);
  for (int kc = 0; kc < nk; ++kc) {
    if (kc + 1 < nk) { load(kc + 1, (kc + 1) & 1); rg_wait<1>(); } else { rg_wait<0>(); }
    __syncthreads();
    const float* as = As + (kc & 1) * RG_BM * RG_AS + (wm * 32) * RG_AS;
    const float* bs = Bs + (kc & 1) * RG_KC * RG_BS + wn * 64;
#pragma unroll
    for (int ks = 0; ks < RG_KC; ks += 8) {
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float* ap = as + (i * 16 + g) * RG_AS + ks + t;
        rg_split(ap[0], ah[i][0], al[i][0]);
        rg_split(ap[8 * RG_AS], ah[i][1], al[i][1]);
        rg_split(ap[4], ah[i][2], al[i][2]);
        rg_split(ap[8 * RG_AS + 4], ah[i][3], al[i][3]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float* bp = bs + (ks + t) * RG_BS + j * 8 + g;
        uint32_t bh[2], bl[2];
        rg_split(bp[0], bh[0], bl[0]);
        rg_split(bp[4 * RG_BS], bh[1], bl[1]);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          rg_mma(acc[i][j], al[i], bh);          // small terms first
          rg_mma(acc[i][j], ah[i], bl);
          rg_mma(acc[i][j], ah[i], bh);
        }
      }
    }
    __syncthreads();
  }
  // epilogue: (+ bias) (+ C) -> C;  c0,c1: row g, columns 2t, 2t+1;  c2,c3: row g + 8
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long r = row0 + wm * 32 + i * 16 + g + 8 * h;
      if (r >= N) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = col0 + wn * 64 + j * 8 + 2 * t;
        if (c >= M) continue;
        float v0 = acc[i][j][2 * h], v1 = acc[i][j][2 * h + 1];
        float* cp = C + (size_t)r * ldc + c;
        if (bias) { v0 += bias[c]; if (c + 1 < M) v1 += bias[c + 1]; }
        if (accumulate) { v0 += cp[0]; if (c + 1 < M) v1 += cp[1]; }
        if (c + 1 < M && ((ldc & 1) == 0)) *reinterpret_cast<float2*>(cp) = make_float2(v0, v1);
        else { cp[0] = v0; if (c + 1 < M) cp[1] = v1; }
      }
    }
}

int launch_rowgemm(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc, long long N, int Kd,
                   int M, int accumulate, cudaStream_t st) {
  if (N < 0 || Kd < 1 || M < 1 || lda < Kd || ldb < M || ldc < M) {
    set_error("rowgemm: bad shape N=%lld Kd=%d M=%d lda=%d ldb=%d ldc=%d", N, Kd, M, lda, ldb, ldc);
    return VBMP_ERR_SHAPE;
  }
  if (N == 0) return VBMP_OK;
  const size_t smem = (size_t)(2 * RG_BM * RG_AS + 2 * RG_KC * RG_BS) * sizeof(float);
  dim3 grid((unsigned)((N + RG_BM - 1) / RG_BM), (unsigned)((M + RG_BN - 1) / RG_BN));
  const bool vec = (lda % 4 == 0) && (ldb % 4 == 0) && (Kd % 4 == 0) && (M % 4 == 0) && ((size_t)A % 16 == 0) && ((size_t)B % 16 == 0);
  if (vec) {
    cudaFuncSetAttribute(rowgemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    rowgemm_kernel<true><<<grid, 256, smem, st>>>(A, lda, B, ldb, bias, C, ldc, N, Kd, M, accumulate);
  } else {
    cudaFuncSetAttribute(rowgemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    rowgemm_kernel<false><<<grid, 256, smem, st>>>(A, lda, B, ldb, bias, C, ldc, N, Kd, M, accumulate);
  }
  return check_launch("rowgemm");
}

}  // namespace vbmp
