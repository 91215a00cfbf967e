// Sample-major product with a SHORT reduction and many columns on the 5th-gen tensor cores:
//
//   C[n][j] = (accumulate ? C[n][j] : 0) + sum_k A[n][k] B[k][j] (+ bias[j])        N rows (millions), Kd <= 64, M columns
//
// The two shared-operand products of MixtureofLinearTransforms.predict (transforms/MixtureofLinearTransforms.py:100-103):
// the component means  (rows x p') (p' x K n)  and  sum_k p_k ESigma_k = (rows x K) (K x n^2).  Their reduction lengths are
// 33 and 64 and their outputs 2048 and 1024 columns wide (8 / 4 KB per row): the calls are bound by the bytes they WRITE.
// The warp-level kernel (rowgemm.cu, mma.sync) ran them at 1.5 - 1.8 TB/s, tensor bound; here
//
//   * CTA = 128 rows (TMEM lanes), persistent over row tiles, one per SM.  192 threads: warp 0 bulk-copy producer (B
//     chunks), warp 1 MMA issuer, warps 2-5 workers (thread = row).
//   * the A tile lives in TENSOR MEMORY: a worker loads its row (Kd floats, one extra column of ones when there is a
//     bias), splits hi + lo (TF32) and writes both images with tcgen05.st; two A buffers, so the rows of tile t + 1 are
//     loaded (into registers, at the start of tile t) and stored under the MMAs of tile t.
//   * B (+ bias as row Kd) is split and packed once per call into 128-column chunks in the K-major / no-swizzle
//     core-matrix layout, [K-step][hi | lo][16-byte chunk (2)][column (128)][4 tf32]: ONE bulk copy per pipeline stage.
//   * per chunk Kp / 8 K-steps x 3 terms of M = 128 x N = 128 x K = 8 into one of two 128-column accumulators; the workers
//     read the other (tcgen05.ld, thread = row), transpose it through a padded shared-memory tile (row stride 132 floats:
//     conflict free both ways) and write C with every warp instruction covering 512 contiguous bytes of one row.  (Thread
//     = row stores straight from the registers touch 32 rows per instruction: the first version was bound by the load /
//     store unit's address divergence and no faster than the mma.sync kernel.)
#include "common.cuh"
#include "umma.cuh"

namespace vbmp {
using namespace umma;

constexpr int RW_THREADS = 192;
constexpr int RW_NC = 128;           // columns per chunk
constexpr int RW_MAXST = 3;
constexpr int RW_KMAX = 64;          // padded reduction length (TMEM columns of one A image)
constexpr int RW_DCOL = 256;         // TMEM: A buffer b at 128 b (hi at +0, lo at +64), accumulator d at 256 + 128 d
constexpr int RW_TS = 132;           // row stride (floats) of the output tile in shared memory
constexpr int RW_TILE_BYTES = 128 * RW_TS * 4;

// B (Kd, ldb) row-major (+ bias) -> packed chunks; rows >= Kd (+1) and columns >= M are zero
__global__ void rowwide_pack_kernel(const float* __restrict__ B, int ldb, const float* __restrict__ bias, int Kd, int Kp, int M,
                                    uint8_t* __restrict__ Bp) {
  const int chunk = blockIdx.x;
  float* out = reinterpret_cast<float*>(Bp + (size_t)chunk * Kp * 1024);
  for (int o = threadIdx.x; o < RW_NC * Kp; o += blockDim.x) {
    const int n = o / Kp, kk = o % Kp;
    const int col = chunk * RW_NC + n;
    float v = 0.f;
    if (col < M) v = kk < Kd ? B[(size_t)kk * ldb + col] : ((kk == Kd && bias) ? bias[col] : 0.f);
    uint32_t hi, lo;
    split_tf32(v, hi, lo);
    const int ks = kk >> 3, k8 = kk & 7;
    const size_t base = (size_t)ks * 2 * RW_NC * 8;                          // floats per K-step: 2 images x 128 x 8
    const size_t e = (size_t)(k8 >> 2) * RW_NC * 4 + (size_t)n * 4 + (k8 & 3);
    out[base + e] = __uint_as_float(hi);
    out[base + (size_t)RW_NC * 8 + e] = __uint_as_float(lo);
  }
}

struct RwSmem {
  uint64_t bfull[RW_MAXST], bempty[RW_MAXST];
  uint64_t afull[2], aempty[2];
  uint64_t dfull[2], dempty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(RW_THREADS, 1) rowwide_umma_kernel(const float* __restrict__ A, int lda, const uint8_t* __restrict__ Bp,
                                                                     float* __restrict__ C, int ldc, long long N, int Kd, int Kp,
                                                                     int M, int has_bias, int accumulate, int ntiles, int nst) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int stageB = Kp * 1024;
  uint8_t* bst = smem_raw;
  float* tile = reinterpret_cast<float*>(bst + (size_t)nst * stageB);                 // [128][RW_TS]
  RwSmem* S = reinterpret_cast<RwSmem*>(bst + (size_t)nst * stageB + RW_TILE_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nch = (M + RW_NC - 1) / RW_NC, nks = Kp >> 3;
  if (tid == 0) {
    for (int s = 0; s < RW_MAXST; ++s) { mbar_init(&S->bfull[s], 1); mbar_init(&S->bempty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&S->afull[b], 4); mbar_init(&S->aempty[b], 1);
      mbar_init(&S->dfull[b], 1); mbar_init(&S->dempty[b], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&S->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = S->tmem_base;
  const int my_tiles = (ntiles > (int)blockIdx.x) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    // ================= producer: the packed B chunk of every (tile, chunk) =================
    int s = 0; uint32_t ph = 0;
    for (int t = 0; t < my_tiles; ++t)
      for (int c = 0; c < nch; ++c) {
        mbar_wait(&S->bempty[s], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&S->bfull[s], (uint32_t)stageB);
          bulk_g2s(bst + (size_t)s * stageB, Bp + (size_t)c * stageB, (uint32_t)stageB, &S->bfull[s]);
        }
        __syncwarp();
        if (++s == nst) { s = 0; ph ^= 1u; }
      }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = idesc_tf32(128, RW_NC);
    int s = 0; uint32_t ph = 0;
    uint32_t g = 0;                                        // global chunk counter: accumulator g % 2, its (g / 2)-th use
    for (int t = 0; t < my_tiles; ++t) {
      const int ab = t & 1;
      mbar_wait(&S->afull[ab], (uint32_t)(t >> 1) & 1u);
      const uint32_t a_hi = tm + ab * 128, a_lo = a_hi + 64;
      for (int c = 0; c < nch; ++c, ++g) {
        const uint32_t d = g & 1u;
        mbar_wait(&S->bfull[s], ph);
        mbar_wait(&S->dempty[d], ((g >> 1) & 1u) ^ 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t dcol = tm + RW_DCOL + d * RW_NC;
          const uint32_t sb = smem_u32(bst + (size_t)s * stageB);
          for (int ks = 0; ks < nks; ++ks) {
            const uint64_t b_hi = smem_desc(sb + (uint32_t)(ks * 2) * 4096u, RW_NC * 16, 128);
            const uint64_t b_lo = smem_desc(sb + (uint32_t)(ks * 2 + 1) * 4096u, RW_NC * 16, 128);
            mma_tf32_ts(dcol, a_lo + ks * 8, b_hi, idesc, ks != 0);          // small terms first
            mma_tf32_ts(dcol, a_hi + ks * 8, b_lo, idesc, 1);
            mma_tf32_ts(dcol, a_hi + ks * 8, b_hi, idesc, 1);
          }
          mma_commit(&S->bempty[s]);
          mma_commit(&S->dfull[d]);
          if (c == nch - 1) mma_commit(&S->aempty[ab]);    // the tile's MMAs have read this A buffer
        }
        __syncwarp();
        if (++s == nst) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // ================= workers: thread = row =================
    const int q = warp & 3;                                // TMEM lanes 32 q .. (a warp reaches its own quarter only)
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int rloc = q * 32 + lane;
    const bool vecA = (lda % 4 == 0) && (Kd % 4 == 0) && ((size_t)A % 16 == 0);
    const bool vecC = (ldc % 4 == 0) && (M % 4 == 0) && ((size_t)C % 16 == 0);
    float v[RW_KMAX];
    auto load_row = [&](int t) {                           // this thread's row of tile t into registers (zeros past Kd / N)
      const long long row = ((long long)blockIdx.x + (long long)t * gridDim.x) * 128 + rloc;
      const bool ok = t < my_tiles && row < N;
      const float* ar = A + (size_t)(ok ? row : 0) * lda;
#pragma unroll
      for (int k = 0; k < RW_KMAX; k += 4) {
        if (vecA) {
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok && k < Kd) x = __ldg(reinterpret_cast<const float4*>(ar + k));
          v[k] = x.x; v[k + 1] = x.y; v[k + 2] = x.z; v[k + 3] = x.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) v[k + j] = (ok && k + j < Kd) ? __ldg(ar + k + j) : 0.f;
        }
      }
      if (has_bias) {                                      // column Kd of the A image meets the bias row of B
#pragma unroll
        for (int k = 0; k < RW_KMAX; ++k) if (k == Kd) v[k] = ok ? 1.f : 0.f;
      }
    };
    auto put_row = [&](int t) {                            // split and store v into A buffer t % 2
      const int ab = t & 1;
      if (t >= 2) mbar_wait(&S->aempty[ab], (uint32_t)((t >> 1) - 1) & 1u);
      tc_fence_after();
      const uint32_t ad = tm + lane_base + ab * 128;
#pragma unroll
      for (int k0 = 0; k0 < RW_KMAX; k0 += 8) {
        if (k0 < Kp) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            hi[j] = (__float_as_uint(v[k0 + j]) + 0x1000u) & 0xffffe000u;
            lo[j] = __float_as_uint(v[k0 + j] - __uint_as_float(hi[j]));
          }
          tmem_st8(ad + k0, hi);
          tmem_st8(ad + 64 + k0, lo);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&S->afull[ab]);
    };
    if (my_tiles > 0) { load_row(0); put_row(0); }
    uint32_t g = 0;
    for (int t = 0; t < my_tiles; ++t) {
      const long long row = ((long long)blockIdx.x + (long long)t * gridDim.x) * 128 + rloc;
      if (t + 1 < my_tiles) load_row(t + 1);               // in flight under this tile's first epilogue
      for (int c = 0; c < nch; ++c, ++g) {
        const uint32_t d = g & 1u;
        mbar_wait(&S->dfull[d], (g >> 1) & 1u);
        tc_fence_after();
        const uint32_t dcol = tm + lane_base + RW_DCOL + d * RW_NC;
        float* trow = tile + (size_t)rloc * RW_TS;
#pragma unroll
        for (int h = 0; h < 4; ++h) {                      // 32 columns at a time: accumulator -> this thread's row of the tile
          float y[32];
          tmem_ld32(dcol + 32 * h, y);
          tmem_wait_ld();
          if (h == 3) {                                    // the accumulator is free again
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S->dempty[d]);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(trow + 32 * h + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");     // the four worker warps: tile complete
        // copy out: warp q takes rows q, q + 4, ...; a lane takes 4 consecutive columns, so one instruction = 512 bytes of a row
        const int col = c * RW_NC + 4 * lane;
        const long long trow0 = ((long long)blockIdx.x + (long long)t * gridDim.x) * 128;
        if (col < M) {
#pragma unroll 4
          for (int rr = 0; rr < 32; ++rr) {
            const int r = rr * 4 + q;
            const long long orow = trow0 + r;
            if (orow >= N) break;
            const float4 o4 = *reinterpret_cast<const float4*>(tile + (size_t)r * RW_TS + 4 * lane);
            float* cp = C + (size_t)orow * ldc + col;
            if (vecC) {
              float4 o = o4;
              if (accumulate) {
                const float4 old = *reinterpret_cast<const float4*>(cp);
                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
              }
              *reinterpret_cast<float4*>(cp) = o;
            } else {
              const float ov[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (col + j < M) cp[j] = (accumulate ? cp[j] : 0.f) + ov[j];
            }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");     // the tile may be overwritten
        if (c == 0 && t + 1 < my_tiles) put_row(t + 1);    // the next tile's A image, under this tile's MMAs
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tm);
}

// ---- host side -------------------------------------------------------------------------------------------
static int rw_kp(int Kd, bool bias) { return (Kd + (bias ? 1 : 0) + 7) / 8 * 8; }
bool rowwide_umma_supported(long long N, int Kd, int M, bool bias) {
  return N >= 128 && Kd >= 1 && rw_kp(Kd, bias) <= RW_KMAX && M >= 64;
}
size_t rowwide_umma_workspace_bytes(int Kd, int M, bool bias) {
  return 256 + (size_t)((M + RW_NC - 1) / RW_NC) * rw_kp(Kd, bias) * 1024;
}

int launch_rowwide_umma(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc, long long N, int Kd,
                        int M, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!rowwide_umma_supported(N, Kd, M, bias != nullptr) || lda < Kd || ldb < M || ldc < M) {
    set_error("rowwide: unsupported shape N=%lld Kd=%d M=%d lda=%d ldb=%d ldc=%d", N, Kd, M, lda, ldb, ldc);
    return VBMP_ERR_UNSUPPORTED;
  }
  if (ws_bytes < rowwide_umma_workspace_bytes(Kd, M, bias != nullptr)) { set_error("rowwide: workspace too small"); return VBMP_ERR_WORKSPACE; }
  const int Kp = rw_kp(Kd, bias != nullptr), nch = (M + RW_NC - 1) / RW_NC;
  uint8_t* Bp = (uint8_t*)(((size_t)ws + 255) / 256 * 256);
  rowwide_pack_kernel<<<nch, 256, 0, st>>>(B, ldb, bias, Kd, Kp, M, Bp);
  int rc = check_launch("rowwide_pack");
  if (rc) return rc;
  const int ntiles = (int)((N + 127) / 128);
  int nst = (int)((226 * 1024 - RW_TILE_BYTES - sizeof(RwSmem)) / ((size_t)Kp * 1024));
  if (nst > RW_MAXST) nst = RW_MAXST;
  if (nst > nch) nst = nch < 1 ? 1 : nch;
  if (nst < 1) nst = 1;
  const size_t smem = (size_t)nst * Kp * 1024 + RW_TILE_BYTES + sizeof(RwSmem) + 64;
  const int grid = ntiles < num_sms() ? ntiles : num_sms();
  cudaFuncSetAttribute(rowwide_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  rowwide_umma_kernel<<<grid, RW_THREADS, smem, st>>>(A, lda, Bp, C, ldc, N, Kd, Kp, M, bias ? 1 : 0, accumulate, ntiles, nst);
  return check_launch("rowwide_umma");
}

}  // namespace vbmp
