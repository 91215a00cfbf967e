// K2 (tcgen05 variant): whitened quadratic forms on the 5th-gen tensor cores + fused responsibility softmax.
//
//   l[n,k] = cst[k] - 1/2 || W_k^T z_n - m_k ||^2          (NIW/MNW.Elog_like; dists/NormalInverseWishart.py:91-97,
//                                                           transforms/MatrixNormalWishart.py:219-232)
//   mode 1: p = exp(l - logZ_n), logZ_n, NA_k, sum_n logZ_n (Mixture.update_assignments dists/Mixture.py:38-45,
//                                                           MixtureofLinearTransforms.update_assignments :34-41)
//
// Design (numbers from tools/umma_probe.cu on a B200, see umma.cuh):
//   * persistent CTA per SM, 256-sample tiles (two 128-row halves).  The sample tile is the A operand and
//     lives in TENSOR MEMORY as split-precision TF32 (z = hi + lo): with A in TMEM an M=128 x N=64 x K=8
//     MMA issues every N/2 cycles, whereas a shared-memory A costs 32 extra cycles of operand fetch.
//   * the whitening factors are the B operand: one 64-column group (64/DP components) per pipeline stage,
//     pre-split into hi / lo TF32 and pre-arranged in the K-major core-matrix layout by a pack kernel, so a
//     stage is ONE cp.async.bulk.  Each stage is used by both halves (6 MMAs per K-step:
//     hi*hi + lo*hi + hi*lo per half), which halves the L2 -> shared-memory stream per flop.
//   * W_k is upper triangular, so for DP = 64 K-step ks only feeds output columns >= 16*(ks/2): the MMA is
//     issued with N = 64 - 16*(ks/2) on the column suffix and only that suffix of B is stored / copied
//     (62.5 % of the dense work).
//   * accumulators D[half][buf] (64 fp32 columns each) are double buffered in TMEM; 8 epilogue warps (one
//     thread per sample row) read them with tcgen05.ld, subtract m_k, square-reduce, keep an online
//     logsumexp and write the logits.  For mode 1 the tile's rows are then normalised in place (the re-read
//     hits L2) with coalesced float4 accesses, accumulating NA per CTA in a fixed order (deterministic).
#include "common.cuh"
#include "umma.cuh"

namespace vbmp {
using namespace umma;

constexpr int EU_THREADS = 320;     // warp 0: bulk-copy producer, warp 1: MMA issuer, warps 2..9: workers
constexpr int EU_TILE = 256;
constexpr int EU_NSTAGE = 8;
constexpr int EU_MAXK = 512;

template <int DP>
struct EuCfg {
  static constexpr int CG = 64 / DP;          // components per 64-column MMA group
  static constexpr int KS = DP / 8;           // K-steps (8 TF32 each)
  static constexpr bool TRI = (DP == 64);
  __host__ __device__ static constexpr int n0(int ks) { return TRI ? 16 * (ks / 2) : 0; }
  __host__ __device__ static constexpr int nn(int ks) { return 64 - n0(ks); }
  __host__ __device__ static constexpr int blk_off(int ks) {
    int o = 0;
    for (int i = 0; i < ks; ++i) o += nn(i) * 32;
    return o;
  }
  static constexpr int WB = blk_off(KS);      // bytes of one operand image (hi or lo) of a group
  static constexpr int GB = 2 * WB;           // hi image then lo image
  static constexpr int REC = GB + 64 * 4 + 16; // + m of the group's components (64 floats) + their cst (<= 4 floats)
  static constexpr int STAGE = (REC + 127) / 128 * 128;
};

// ---- pack: W (C, DP, DP) fp32 row-major [i][j] -> per group [hi | lo] images in UMMA K-major core-matrix layout
// B[n][k] = W_c[i = k][j],  n = cl*DP + j (cl = component within the group); K-step block ks holds rows
// n >= n0(ks) as [chunk (2)][row][4 floats], i = 8 ks + 4 chunk + e.
template <int DP>
__global__ void estep_pack_kernel(const float* __restrict__ W, const float* __restrict__ m, const float* __restrict__ cst,
                                  int K, uint8_t* __restrict__ Wp) {
  using C = EuCfg<DP>;
  const int g = blockIdx.x;
  float* hi = reinterpret_cast<float*>(Wp + (size_t)g * C::REC);
  float* lo = reinterpret_cast<float*>(Wp + (size_t)g * C::REC + C::WB);
  float* mo = reinterpret_cast<float*>(Wp + (size_t)g * C::REC + C::GB);
  for (int o = threadIdx.x; o < 68; o += blockDim.x) {
    if (o < 64) { const int c = g * C::CG + o / DP; mo[o] = c < K ? m[(size_t)c * DP + o % DP] : 0.f; }
    else { const int c = g * C::CG + (o - 64); mo[o] = (o - 64 < C::CG && c < K) ? cst[c] : 0.f; }
  }
  for (int o = threadIdx.x; o < C::WB / 4; o += blockDim.x) {
    int ks = 0, rem = o * 4;
    while (ks + 1 < C::KS && rem >= C::nn(ks) * 32) { rem -= C::nn(ks) * 32; ++ks; }
    const int nn = C::nn(ks);
    const int ch = rem / (nn * 16);
    const int r = (rem % (nn * 16)) / 16, e = (rem % 16) / 4;
    const int n = C::n0(ks) + r, cl = n / DP, j = n % DP, i = 8 * ks + 4 * ch + e;
    const int c = g * C::CG + cl;
    const float v = (c < K) ? W[((size_t)c * DP + i) * DP + j] : 0.f;
    uint32_t h, l;
    split_tf32(v, h, l);
    hi[o] = __uint_as_float(h);
    lo[o] = __uint_as_float(l);
  }
}

struct EuSmem {
  uint64_t full[EU_NSTAGE], empty[EU_NSTAGE];
  uint64_t tfull[2][2], tempty[2][2];
  uint64_t afull;
  uint32_t tmem_base;
  double red[8];
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int DP, int MODE>
__global__ void __launch_bounds__(EU_THREADS, 1)
estep_umma_kernel(EstepArgs a, const uint8_t* __restrict__ Wp, int ntiles, int ngroups) {
  using C = EuCfg<DP>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* stages = smem_raw;                                             // EU_NSTAGE * STAGE
  EuSmem* S = reinterpret_cast<EuSmem*>(stages + EU_NSTAGE * C::STAGE);
  float* lz = reinterpret_cast<float*>(S + 1);                            // [256]
  float* colsum = lz + EU_TILE;                                           // [8][K]     (mode 1)
  double* NAacc = reinterpret_cast<double*>(colsum + 8 * a.K);            // [K]        (mode 1)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = a.K;

  if (tid == 0) {
    for (int s = 0; s < EU_NSTAGE; ++s) { mbar_init(&S->full[s], 1); mbar_init(&S->empty[s], 1 + 256); }
    for (int h = 0; h < 2; ++h)
      for (int b = 0; b < 2; ++b) { mbar_init(&S->tfull[h][b], 1); mbar_init(&S->tempty[h][b], 128); }
    mbar_init(&S->afull, 256);
    fence_barrier_init();
  }
  if (MODE == 1) for (int k = tid; k < K; k += EU_THREADS) NAacc[k] = 0.0;
  if (warp == 1) tmem_alloc<512>(&S->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = S->tmem_base;
  // TMEM columns: A hi/lo of half h at h*2*DP (+DP for lo); D[h][buf] at 256 + (2h+buf)*64
  const int my_tiles = (ntiles > (int)blockIdx.x) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    // ================= producer: one bulk copy per 64-column group =================
    long long it = 0;
    for (int t = 0; t < my_tiles; ++t)
      for (int g = 0; g < ngroups; ++g, ++it) {
        const int s = (int)(it % EU_NSTAGE);
        const uint32_t n = (uint32_t)(it / EU_NSTAGE);
        mbar_wait(&S->empty[s], (n & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&S->full[s], C::REC);
          bulk_g2s(stages + (size_t)s * C::STAGE, Wp + (size_t)g * C::REC, C::REC, &S->full[s]);
        }
        __syncwarp();
      }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    long long it = 0;
    for (int t = 0; t < my_tiles; ++t) {
      mbar_wait(&S->afull, t & 1);
      tc_fence_after();
      for (int g = 0; g < ngroups; ++g, ++it) {
        const int s = (int)(it % EU_NSTAGE);
        const uint32_t n = (uint32_t)(it / EU_NSTAGE);
        const int buf = (int)(it & 1);
        const uint32_t nb = (uint32_t)(it >> 1);
        mbar_wait(&S->full[s], n & 1);
        const uint32_t sbase = smem_u32(stages + (size_t)s * C::STAGE);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(&S->tempty[h][buf], (nb & 1) ^ 1);
          tc_fence_after();
          const uint32_t dcol = tm + 256 + (2 * h + buf) * 64;
          const uint32_t a_hi = tm + h * 2 * DP, a_lo = a_hi + DP;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < C::KS; ++ks) {
              const int nn = C::nn(ks), n0 = C::n0(ks);
              const uint32_t idesc = idesc_tf32(128, nn);
              const uint64_t b_hi = smem_desc(sbase + C::blk_off(ks), nn * 16, 128);
              const uint64_t b_lo = smem_desc(sbase + C::WB + C::blk_off(ks), nn * 16, 128);
              mma_tf32_ts(dcol + n0, a_lo + ks * 8, b_hi, idesc, ks > 0);   // small terms first
              mma_tf32_ts(dcol + n0, a_hi + ks * 8, b_lo, idesc, 1);
              mma_tf32_ts(dcol + n0, a_hi + ks * 8, b_hi, idesc, 1);
            }
            mma_commit(&S->tfull[h][buf]);
            if (h == 1) mma_commit(&S->empty[s]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================= workers: A tile -> TMEM, epilogue, normalisation =================
    const int w8 = warp - 2, h = w8 >> 2, q = warp & 3;
    const int wtid = tid - 64;                                  // 0..255
    const int rloc = h * 128 + q * 32 + lane;                   // row within the tile
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int D = a.d0 + a.d1;
    double lzsum = 0.0;
    long long it = 0;
    for (int t = 0; t < my_tiles; ++t) {
      const long long tile = (long long)blockIdx.x + (long long)t * gridDim.x;
      const long long row = tile * EU_TILE + rloc;
      const bool valid = row < a.N;
      // ---- this thread's sample row, split into TF32 hi / lo, into TMEM (previous tile's MMAs on this half
      //      are complete: we waited on tfull of its last group)
      {
        const uint32_t a_hi = tm + lane_base + h * 2 * DP, a_lo = a_hi + DP;
        const float* r0 = a.z0 + (size_t)(valid ? row : 0) * a.d0;
        const float* r1 = a.z1 ? a.z1 + (size_t)(valid ? row : 0) * a.d1 : nullptr;
        const bool vec = valid && (a.d0 % 4 == 0) && (a.d1 % 4 == 0);
#pragma unroll
        for (int c0 = 0; c0 < DP; c0 += 16) {
          float v[16];
          if (vec) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const int f = c0 + j;
              float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
              if (f < a.d0) x = __ldg(reinterpret_cast<const float4*>(r0 + f));
              else if (f < D) x = __ldg(reinterpret_cast<const float4*>(r1 + (f - a.d0)));
              v[j] = x.x; v[j + 1] = x.y; v[j + 2] = x.z; v[j + 3] = x.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int f = c0 + j;
              v[j] = !valid ? 0.f : (f < a.d0 ? __ldg(r0 + f) : (f < D ? __ldg(r1 + (f - a.d0)) : 0.f));
            }
          }
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) split_tf32(v[j], hi[j], lo[j]);
          tmem_st16(a_hi + c0, hi);
          tmem_st16(a_lo + c0, lo);
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&S->afull);
      }
      float mx = -INFINITY, sm = 0.f;
      float l4[4];
      for (int g = 0; g < ngroups; ++g, ++it) {
        const int buf = (int)(it & 1);
        const uint32_t nb = (uint32_t)(it >> 1);
        const int s = (int)(it % EU_NSTAGE);
        const float4* mstage = reinterpret_cast<const float4*>(stages + (size_t)s * C::STAGE + C::GB);
        mbar_wait(&S->tfull[h][buf], nb & 1);
        mbar_wait(&S->full[s], (uint32_t)(it / EU_NSTAGE) & 1);   // already complete; acquires the stage's m / cst
        tc_fence_after();
        float y[64];
        const uint32_t dcol = tm + lane_base + 256 + (2 * h + buf) * 64;
        tmem_ld32(dcol, y);
        tmem_ld32(dcol + 32, y + 32);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(&S->tempty[h][buf]);
#pragma unroll
        for (int cl = 0; cl < C::CG; ++cl) {
          const int c = g * C::CG + cl;
          if (c < K) {
            const float4* mp = mstage + cl * (DP / 4);
            float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
            for (int j = 0; j < DP; j += 4) {
              const float4 mm = mp[j >> 2];
              const float r0_ = y[cl * DP + j] - mm.x, r1_ = y[cl * DP + j + 1] - mm.y;
              const float r2_ = y[cl * DP + j + 2] - mm.z, r3_ = y[cl * DP + j + 3] - mm.w;
              q0 = fmaf(r0_, r0_, q0); q1 = fmaf(r1_, r1_, q1); q2 = fmaf(r2_, r2_, q2); q3 = fmaf(r3_, r3_, q3);
            }
            const float l = reinterpret_cast<const float*>(mstage + 16)[cl] - 0.5f * ((q0 + q1) + (q2 + q3));
            if (MODE == 1) {          // online logsumexp with one exp per component
              const float e = expf(-fabsf(l - mx));
              sm = (l > mx) ? fmaf(sm, e, 1.f) : sm + e;
              mx = fmaxf(mx, l);
            }
            l4[c & 3] = l;
            if ((c & 3) == 3 && valid)
              *reinterpret_cast<float4*>(a.out + (size_t)row * K + (c - 3)) = make_float4(l4[0], l4[1], l4[2], l4[3]);
          }
        }
        mbar_arrive(&S->empty[s]);                  // done with the stage's m / cst (the MMA commit is the other arrival)
      }
      if (MODE == 1) {
        const float v = valid ? mx + logf(sm) : 0.f;
        lz[rloc] = v;
        if (valid) { a.logZn[row] = v; lzsum += (double)v; }
        named_bar_sync(1, 256);
        // ---- normalise the tile's rows in place: p = exp(l - logZ_n); column sums -> NA
        const long long row0 = tile * EU_TILE;
        const long long rem = a.N - row0;
        const int rows = rem < EU_TILE ? (int)rem : EU_TILE;
        const int K4 = K >> 2;
        float cs[EU_MAXK / 128][4];
#pragma unroll
        for (int u = 0; u < EU_MAXK / 128; ++u) { cs[u][0] = cs[u][1] = cs[u][2] = cs[u][3] = 0.f; }
        for (int r = w8; r < rows; r += 32) {           // four rows in flight per warp (memory-level parallelism)
          float4 x[4][EU_MAXK / 128];
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const int rr = r + 8 * v;
            const float4* prow = reinterpret_cast<const float4*>(a.out + (size_t)(row0 + (rr < rows ? rr : r)) * K);
#pragma unroll
            for (int u = 0; u < EU_MAXK / 128; ++u) {
              const int c4 = lane + 32 * u;
              if (c4 < K4) x[v][u] = __ldcg(prow + c4);
            }
          }
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const int rr = r + 8 * v;
            if (rr < rows) {
              const float lzr = lz[rr];
              float4* prow = reinterpret_cast<float4*>(a.out + (size_t)(row0 + rr) * K);
#pragma unroll
              for (int u = 0; u < EU_MAXK / 128; ++u) {
                const int c4 = lane + 32 * u;
                if (c4 < K4) {
                  float4 y = x[v][u];
                  y.x = expf(y.x - lzr); y.y = expf(y.y - lzr); y.z = expf(y.z - lzr); y.w = expf(y.w - lzr);
                  prow[c4] = y;
                  cs[u][0] += y.x; cs[u][1] += y.y; cs[u][2] += y.z; cs[u][3] += y.w;
                }
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < EU_MAXK / 128; ++u) {
          const int c4 = lane + 32 * u;
          if (c4 < K4) *reinterpret_cast<float4*>(colsum + (size_t)w8 * K + 4 * c4) = make_float4(cs[u][0], cs[u][1], cs[u][2], cs[u][3]);
        }
        named_bar_sync(1, 256);
        for (int k = wtid; k < K; k += 256) {
          float s = 0.f;
#pragma unroll
          for (int w = 0; w < 8; ++w) s += colsum[(size_t)w * K + k];      // fixed order
          NAacc[k] += (double)s;
        }
        // (the next tile's colsum / lz writes happen after its first named barrier)
      }
    }
    if (MODE == 1) {
      // sum of logZ_n over this CTA's rows (fixed order: warp shuffle tree, then warps in order)
      for (int o = 16; o > 0; o >>= 1) lzsum += __shfl_xor_sync(0xffffffffu, lzsum, o);
      if (lane == 0) S->red[w8] = lzsum;
      named_bar_sync(1, 256);
      if (wtid == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += S->red[w];
        a.logZ_part[blockIdx.x] = s;
      }
      for (int k = wtid; k < K; k += 256) a.NA_part[(size_t)blockIdx.x * K + k] = (float)NAacc[k];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tm);
}

// ---- host side -------------------------------------------------------------------------------------------
static int eu_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static size_t eu_group_bytes(int Dp) { return Dp == 64 ? EuCfg<64>::REC : (Dp == 32 ? EuCfg<32>::REC : EuCfg<16>::REC); }
static int eu_cg(int Dp) { return 64 / Dp; }
static size_t eu_align(size_t x) { return (x + 255) / 256 * 256; }

bool estep_umma_supported(long long N, int GX, int G, int K, int Dp, int d0, int d1) {
  (void)d0; (void)d1;
  return G == 1 && GX == 1 && (Dp == 16 || Dp == 32 || Dp == 64) && (K % 4 == 0) && K <= EU_MAXK && N >= 256;
}

size_t estep_umma_workspace_bytes(long long N, int G, int K, int Dp, int mode) {
  (void)mode;
  if (!estep_umma_supported(N, 1, G, K, Dp, 1, 0)) return 0;
  const int ngroups = (K + eu_cg(Dp) - 1) / eu_cg(Dp);
  const size_t ctas = 512;   // upper bound on the persistent grid
  return eu_align((size_t)ngroups * eu_group_bytes(Dp)) + eu_align(ctas * K * sizeof(float)) + eu_align(ctas * sizeof(double)) + 256;
}

int launch_estep_reduce(const float*, const double*, int nb, int G, int K, float* NA, float* logZ, cudaStream_t);

template <int DP>
static int eu_launch(EstepArgs a, int mode, uint8_t* Wp, float* NA_part, double* logZ_part, float* NA, float* logZ,
                     cudaStream_t st) {
  using C = EuCfg<DP>;
  const int ngroups = (a.K + C::CG - 1) / C::CG;
  estep_pack_kernel<DP><<<ngroups, 256, 0, st>>>(a.W, a.m, a.cst, a.K, Wp);
  int rc = check_launch("estep_pack");
  if (rc) return rc;
  const int ntiles = (int)((a.N + EU_TILE - 1) / EU_TILE);
  const int grid = ntiles < eu_num_sms() ? ntiles : eu_num_sms();
  const size_t smem = (size_t)EU_NSTAGE * C::STAGE + sizeof(EuSmem) + EU_TILE * sizeof(float) +
                      (mode == 1 ? (size_t)8 * a.K * sizeof(float) + (size_t)a.K * sizeof(double) : 0) + 64;
  a.NA_part = NA_part;
  a.logZ_part = logZ_part;
  if (mode == 0) {
    cudaFuncSetAttribute(estep_umma_kernel<DP, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    estep_umma_kernel<DP, 0><<<grid, EU_THREADS, smem, st>>>(a, Wp, ntiles, ngroups);
  } else {
    cudaFuncSetAttribute(estep_umma_kernel<DP, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    estep_umma_kernel<DP, 1><<<grid, EU_THREADS, smem, st>>>(a, Wp, ntiles, ngroups);
  }
  rc = check_launch("estep_umma");
  if (rc) return rc;
  if (mode == 1) rc = launch_estep_reduce(NA_part, logZ_part, grid, 1, a.K, NA, logZ, st);
  return rc;
}

int launch_estep_umma(const EstepArgs& a, int mode, void* ws, size_t ws_bytes, float* NA, float* logZ, cudaStream_t st) {
  const size_t need = estep_umma_workspace_bytes(a.N, a.G, a.K, a.Dp, mode);
  if (ws_bytes < need) { set_error("estep_umma: workspace too small (%zu < %zu)", ws_bytes, need); return VBMP_ERR_WORKSPACE; }
  const int ngroups = (a.K + eu_cg(a.Dp) - 1) / eu_cg(a.Dp);
  char* p = (char*)eu_align((size_t)ws);
  uint8_t* Wp = (uint8_t*)p;
  p += eu_align((size_t)ngroups * eu_group_bytes(a.Dp));
  float* NA_part = (float*)p;
  p += eu_align((size_t)512 * a.K * sizeof(float));
  double* logZ_part = (double*)p;
  switch (a.Dp) {
    case 64: return eu_launch<64>(a, mode, Wp, NA_part, logZ_part, NA, logZ, st);
    case 32: return eu_launch<32>(a, mode, Wp, NA_part, logZ_part, NA, logZ, st);
    case 16: return eu_launch<16>(a, mode, Wp, NA_part, logZ_part, NA, logZ, st);
  }
  set_error("estep_umma: Dp=%d not supported", a.Dp);
  return VBMP_ERR_UNSUPPORTED;
}

}  // namespace vbmp
