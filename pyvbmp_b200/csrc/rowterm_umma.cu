// Sample-major product with a LONG reduction and few columns on the 5th-gen tensor cores:
//
//   C[n][k] = (accumulate ? C[n][k] : 0) + alpha * sum_f A[n][f] B[f][k]        N rows (millions), F >= 32, K <= 256
//
// The covariance terms of the expectation-input E-step, MatrixNormalWishart.Elog_like_given_pX_pY
// (transforms/MatrixNormalWishart.py:236-247):  -1/2 tr(Sigma_y,n E[invSigma_k])  and  -1/2 tr(Sigma_x,n E[X^T invU X]_k)
// with A = the flattened per-sample covariances (N x n^2, 4 KB per sample at n = 32), B = the flattened K-sized
// expectations.  The call is bound by the bytes of A (one pass over HBM); the 3-term TF32 split keeps fp32-grade accuracy.
//
//   * CTA = 128 rows (TMEM lanes) x all K columns, persistent over row tiles, two CTAs per SM when K <= 64
//     (TMEM 256 columns each: two or three 64-column A stages + two accumulators; <= 48 KB of shared memory).  192 threads: warp 0 bulk-copy producer (B chunks), warp 1 MMA
//     issuer, warps 2-5 workers.
//   * A never touches shared memory.  A worker warp reads its 32 rows in chunks of 32 features straight into registers
//     with 16-byte loads laid out like the m16n8 fragment (thread t: rows t/4 and t/4 + 8, four consecutive features
//     4 (t%4) .. of each 16 — 64 contiguous bytes per row and instruction, every sector used once), two chunks ahead of the
//     one it is processing (register double buffering: the loads of ~64 KB per SM in flight cover the HBM latency),
//     splits each value hi + lo (hi = TF32 rounded to nearest, lo = x - hi exact) and writes both images to TENSOR
//     MEMORY with tcgen05.st.16x256b — the A operand of the MMAs.  The 16-byte loads permute the features of a chunk
//     (register pair i of the fragment holds features 16 (i/2) + 4 (t%4) + 2 (i%2) + {0,1}); the packed B uses the same
//     permutation, and a sum does not care.
//   * B (F x K) is split and packed once per call by rowterm_pack_kernel into the K-major / no-swizzle core-matrix
//     layout, one 32-feature chunk = [K-step (4)][hi | lo][16-byte chunk (2)][column (Kc)][4 tf32] = Kc x 256 bytes, ONE
//     bulk copy per pipeline stage (3 stages).
//   * per chunk 4 K-steps x 3 terms of M = 128 x N = Kc x K = 8.  The tensor core TRUNCATES on every fp32 accumulate
//     (tools/gram_bias.py: -1.7e-8 of the running sum per MMA), and a row's reduction is F / 8 K-steps long, so the two
//     small terms (lo*hi, hi*lo: 2^-11 of the result, their own truncation is invisible) go to a SECOND accumulator and
//     only hi*hi accumulates into the large running sum: a third of the bias of one accumulator, -0.7e-6 of the result at
//     F = 1024 (K <= 128; beyond that one accumulator).  After the tile's last chunk the workers read both (thread = row),
//     add them, apply alpha / accumulate and write 16-byte pieces of their own row of C.
#include "common.cuh"
#include "umma.cuh"

namespace vbmp {
using namespace umma;

constexpr int RT_THREADS = 192;
constexpr int RT_NST = 3;            // most pipeline stages: A images in TMEM (64 columns each), B chunks in shared memory
constexpr int RT_FC = 32;            // features per chunk

// feature (within a chunk) held by TMEM column kk of the A images / row kk of the packed B
__host__ __device__ constexpr int rt_perm(int kk) {
  const int i = kk >> 3, c = (kk >> 1) & 3, e = kk & 1;
  return 16 * (i >> 1) + 4 * c + 2 * (i & 1) + e;
}

__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// 16-byte read-only load that the compiler may not sink towards its first use: the loads of chunk gc + 2 must be in flight
// while chunk gc is processed (asm volatile keeps its order relative to the barrier waits)
__device__ __forceinline__ float4 rt_ldg16(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// B (F, ldb) row-major -> packed chunks; columns >= K and features >= F are zero
__global__ void rowterm_pack_kernel(const float* __restrict__ B, int ldb, int F, int K, int Kc, uint8_t* __restrict__ Bp) {
  const int chunk = blockIdx.x;
  float* out = reinterpret_cast<float*>(Bp + (size_t)chunk * Kc * 256);
  for (int o = threadIdx.x; o < Kc * 32; o += blockDim.x) {
    const int n = o / 32, kk = o % 32;
    const int f = chunk * RT_FC + rt_perm(kk);
    const float v = (f < F && n < K) ? B[(size_t)f * ldb + n] : 0.f;
    uint32_t hi, lo;
    split_tf32(v, hi, lo);
    const int ks = kk >> 3, k8 = kk & 7;
    // [ks][img][k-chunk (2)][row n][4]
    const size_t base = (size_t)ks * 2 * Kc * 8;                         // floats per K-step: 2 images x Kc x 8
    const size_t e = (size_t)(k8 >> 2) * Kc * 4 + (size_t)n * 4 + (k8 & 3);
    out[base + e] = __uint_as_float(hi);
    out[base + (size_t)Kc * 8 + e] = __uint_as_float(lo);
  }
}

struct RtSmem {
  uint64_t bfull[RT_NST], afull[RT_NST], empty[RT_NST];
  uint64_t dfull, dempty;
  uint32_t tmem_base;
};

template <int TCOLS>
__global__ void __launch_bounds__(RT_THREADS) rowterm_umma_kernel(const float* __restrict__ A, int lda, const uint8_t* __restrict__ Bp,
                                                                  float* __restrict__ C, int ldc, long long N, int F, int K, int Kc,
                                                                  float alpha, int accumulate, int ntiles, int nst, int dual) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int stageB = Kc * 256;
  uint8_t* bst = smem_raw;
  RtSmem* S = reinterpret_cast<RtSmem*>(bst + nst * stageB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nch = (F + RT_FC - 1) / RT_FC;
  if (tid == 0) {
    for (int s = 0; s < RT_NST; ++s) { mbar_init(&S->bfull[s], 1); mbar_init(&S->afull[s], 4); mbar_init(&S->empty[s], 1); }
    mbar_init(&S->dfull, 1); mbar_init(&S->dempty, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TCOLS>(&S->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = S->tmem_base;
  // TMEM columns: A stage s at 64 s (hi 32 columns, lo 32 columns); accumulator of hi*hi at 64 nst, of the small terms
  // (dual) at 64 nst + Kc
  const int RT_ACOLS = nst * 64;
  const int my_tiles = (ntiles > (int)blockIdx.x) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const long long total = (long long)my_tiles * nch;          // chunks this CTA processes, in order

  if (warp == 0) {
    // ================= producer: the packed B chunk of every (tile, chunk) =================
    int s = 0; uint32_t ph = 0;
    for (long long gc = 0; gc < total; ++gc) {
      const int c = (int)(gc % nch);
      mbar_wait(&S->empty[s], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&S->bfull[s], (uint32_t)stageB);
        bulk_g2s(bst + (size_t)s * stageB, Bp + (size_t)c * stageB, (uint32_t)stageB, &S->bfull[s]);
      }
      __syncwarp();
      if (++s == nst) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = idesc_tf32(128, Kc);
    const uint32_t dcol = tm + RT_ACOLS, dsml = dual ? dcol + Kc : dcol;
    int s = 0; uint32_t ph = 0;
    for (int t = 0; t < my_tiles; ++t) {
      if (t > 0) mbar_wait(&S->dempty, (uint32_t)(t - 1) & 1u);           // the workers have read tile t-1's accumulator
      for (int c = 0; c < nch; ++c) {
        mbar_wait(&S->bfull[s], ph);
        mbar_wait(&S->afull[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_hi = tm + s * 64, a_lo = a_hi + 32;
          const uint32_t sb = smem_u32(bst + (size_t)s * stageB);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t b_hi = smem_desc(sb + ks * 2 * Kc * 32, Kc * 16, 128);
            const uint64_t b_lo = smem_desc(sb + (ks * 2 + 1) * Kc * 32, Kc * 16, 128);
            const uint32_t fresh = !(c == 0 && ks == 0);
            mma_tf32_ts(dsml, a_lo + ks * 8, b_hi, idesc, fresh);                      // small terms first
            mma_tf32_ts(dsml, a_hi + ks * 8, b_lo, idesc, 1);
            mma_tf32_ts(dcol, a_hi + ks * 8, b_hi, idesc, dual ? fresh : 1);
          }
          mma_commit(&S->empty[s]);
          if (c == nch - 1) mma_commit(&S->dfull);
        }
        __syncwarp();
        if (++s == nst) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // ================= workers: A rows -> registers -> split -> TMEM; epilogue =================
    const int w = warp & 3;                                    // TMEM lanes 32 w .. 32 w + 31 (a warp reaches its own quarter only)
    const int r4 = lane >> 2, c4 = lane & 3;
    const bool vecC = (ldc % 4 == 0) && ((size_t)C % 16 == 0);
    // fragment loads of global chunk gc: v[hf][j][rw] = 16 bytes of row 32 w + 16 hf + 8 rw + r4 at features 16 j + 4 c4
    auto load = [&](long long gc, float4 (&v)[2][2][2]) {
      const int t = (int)(gc / nch), c = (int)(gc % nch);
      const long long row0 = ((long long)blockIdx.x + (long long)t * gridDim.x) * 128 + 32 * w + r4;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf)
#pragma unroll
        for (int rw = 0; rw < 2; ++rw) {
          const long long row = row0 + 16 * hf + 8 * rw;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int f = c * RT_FC + 16 * j + 4 * c4;
            v[hf][j][rw] = (row < N && f < F) ? rt_ldg16(A + (size_t)row * lda + f) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
    };
    // split and store chunk registers into A stage s: register pair i of the 16x256b fragment = (j = i / 2, half i % 2)
    auto put = [&](int s, const float4 (&v)[2][2][2]) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int rw = 0; rw < 2; ++rw) {
            const float4 q = v[hf][i >> 1][rw];
            const float x0 = (i & 1) ? q.z : q.x, x1 = (i & 1) ? q.w : q.y;
            const uint32_t h0 = (__float_as_uint(x0) + 0x1000u) & 0xffffe000u, h1 = (__float_as_uint(x1) + 0x1000u) & 0xffffe000u;
            hi[4 * i + 2 * rw] = h0; hi[4 * i + 2 * rw + 1] = h1;
            lo[4 * i + 2 * rw] = __float_as_uint(x0 - __uint_as_float(h0));
            lo[4 * i + 2 * rw + 1] = __float_as_uint(x1 - __uint_as_float(h1));
          }
        const uint32_t ad = tm + ((uint32_t)(32 * w + 16 * hf) << 16) + s * 64;
        tmem_st_16x256b_x4(ad, hi);
        tmem_st_16x256b_x4(ad + 32, lo);
      }
    };
    float4 b0[2][2][2], b1[2][2][2], b2[2][2][2];            // chunks gc, gc + 1, gc + 2 (rotating)
    if (total > 0) load(0, b0);
    if (total > 1) load(1, b1);
    int s = 0; uint32_t ph = 0;
    auto step = [&](long long gc, float4 (&cur)[2][2][2], float4 (&nxt2)[2][2][2]) {
      if (gc + 2 < total) load(gc + 2, nxt2);
      mbar_wait(&S->empty[s], ph ^ 1);
      tc_fence_after();
      put(s, cur);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&S->afull[s]);
      if (++s == nst) { s = 0; ph ^= 1u; }
      const int c = (int)(gc % nch);
      if (c == nch - 1) {
        // ---- epilogue of this tile: thread = row
        const int t = (int)(gc / nch);
        mbar_wait(&S->dfull, (uint32_t)t & 1u);
        tc_fence_after();
        const long long row = ((long long)blockIdx.x + (long long)t * gridDim.x) * 128 + 32 * w + lane;
        const uint32_t dcol = tm + ((uint32_t)(32 * w) << 16) + RT_ACOLS;
        for (int k0 = 0; k0 < Kc; k0 += 16) {
          float y[16];
          tmem_ld16(dcol + k0, y);
          if (dual) {
            float ys[16];
            tmem_ld16(dcol + Kc + k0, ys);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) y[j] += ys[j];
          } else {
            tmem_wait_ld();
          }
          if (row < N) {
            float* cp = C + (size_t)row * ldc + k0;
            if (vecC && k0 + 16 <= K) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                float4 o = make_float4(alpha * y[j], alpha * y[j + 1], alpha * y[j + 2], alpha * y[j + 3]);
                if (accumulate) {
                  const float4 old = *reinterpret_cast<const float4*>(cp + j);
                  o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                *reinterpret_cast<float4*>(cp + j) = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (k0 + j < K) cp[j] = (accumulate ? cp[j] : 0.f) + alpha * y[j];
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S->dempty);
      }
    };
    for (long long gc = 0; gc < total; gc += 3) {
      step(gc, b0, b2);
      if (gc + 1 < total) step(gc + 1, b1, b0);
      if (gc + 2 < total) step(gc + 2, b2, b1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TCOLS>(tm);
}

// ---- host side -------------------------------------------------------------------------------------------
static int rt_kc(int K) { return (K + 15) / 16 * 16; }
bool rowterm_umma_supported(long long N, int F, int K, int lda) {
  return N >= 128 && F >= RT_FC && (F % 4 == 0) && (lda % 4 == 0) && lda >= F && K >= 1 && K <= 256;
}
size_t rowterm_umma_workspace_bytes(int F, int K) {
  return 256 + (size_t)((F + RT_FC - 1) / RT_FC) * rt_kc(K) * 256;
}

int launch_rowterm_umma(const float* A, int lda, const float* B, int ldb, float* C, int ldc, long long N, int F, int K,
                        float alpha, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!rowterm_umma_supported(N, F, K, lda) || ((size_t)A % 16) || ldb < K || ldc < K) {
    set_error("rowterm: unsupported shape N=%lld F=%d K=%d lda=%d ldb=%d ldc=%d (needs N >= 128, F >= 32, F %% 4 == 0, lda %% 4 == 0, "
              "K <= 256, 16-byte aligned A)", N, F, K, lda, ldb, ldc);
    return VBMP_ERR_UNSUPPORTED;
  }
  if (ws_bytes < rowterm_umma_workspace_bytes(F, K)) { set_error("rowterm: workspace too small"); return VBMP_ERR_WORKSPACE; }
  const int Kc = rt_kc(K), nch = (F + RT_FC - 1) / RT_FC;
  uint8_t* Bp = (uint8_t*)(((size_t)ws + 255) / 256 * 256);
  rowterm_pack_kernel<<<nch, 256, 0, st>>>(B, ldb, F, K, Kc, Bp);
  int rc = check_launch("rowterm_pack");
  if (rc) return rc;
  const int ntiles = (int)((N + 127) / 128);
  const int dual = Kc <= 128;                                // second accumulator for the small terms
  const bool small = 2 * 64 + 2 * Kc <= 256;                 // two CTAs per SM share tensor memory (K <= 64)
  int nst = ((small ? 256 : 512) - (dual ? 2 : 1) * Kc) / 64;
  if (nst > RT_NST) nst = RT_NST;
  const size_t smem = (size_t)nst * Kc * 256 + sizeof(RtSmem) + 64;
  const int per_sm = small ? 2 : 1;
  const int grid = ntiles < num_sms() * per_sm ? ntiles : num_sms() * per_sm;
  if (small) {
    cudaFuncSetAttribute(rowterm_umma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    rowterm_umma_kernel<256><<<grid, RT_THREADS, smem, st>>>(A, lda, Bp, C, ldc, N, F, K, Kc, alpha, accumulate, ntiles, nst, dual);
  } else {
    cudaFuncSetAttribute(rowterm_umma_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    rowterm_umma_kernel<512><<<grid, RT_THREADS, smem, st>>>(A, lda, Bp, C, ldc, N, F, K, Kc, alpha, accumulate, ntiles, nst, dual);
  }
  return check_launch("rowterm_umma");
}

}  // namespace vbmp
