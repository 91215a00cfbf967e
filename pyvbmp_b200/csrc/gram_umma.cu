// K3 (tcgen05 variant): responsibility-weighted Gram statistics as ONE tensor-core contraction over samples.
//
//   Gram[k] = sum_n r[n,k] [z_n;1][z_n;1]^T     (SExx / SEx / N of NormalInverseWishart.raw_update,
//                                                dists/NormalInverseWishart.py:80-84; SExx / SEyx / SEyy / SEx /
//                                                SEy / N of MatrixNormalWishart.raw_update with z = [x;y],
//                                                transforms/MatrixNormalWishart.py:185-202)
//
// The matrix is symmetric, so only the P = (D+1)(D+2)/2 products phi_n[p] = zt_n[i_p] zt_n[j_p] (i <= j,
// zt = [z;1]) are needed:   G[k][p] = sum_n r[n,k] phi_n[p]   is a (K x N) x (N x P) GEMM whose reduction
// dimension is the SAMPLE axis.  That is half the flops of forming (r_k o Z)^T Z per component and, unlike
// it, needs no per-component rescaling of the sample tile: phi is shared by every component.
//
//   * CTA task = (128-component block) x (NPB <= 224 pair columns) x (sample split).  A operand = R^T
//     (lanes = components, K = samples) written to TENSOR MEMORY by the workers as split TF32 (hi, lo);
//     B operand = phi^T generated on the fly in shared memory (K-major core-matrix layout, hi / lo).
//     D1 (first-level accumulator) and D2 (second level) both live in TMEM: D1 is folded into D2 every
//     512 samples so no fp32 chain is longer than that, at zero memory traffic.
//   * raw R / Z chunks (16 samples) are brought in by TMA tiled loads (one box per operand, zero-filled past
//     the last row) into a 4-deep ring (warp 0), the MMAs are
//     issued by warp 1 (3 split-precision terms x 2 K-steps per chunk), 8 worker warps split / multiply.
//   * per-split partials are reduced in a fixed order in fp64 by gram_pair_reduce_kernel (deterministic).
#include <cstdlib>
#include "common.cuh"
#include "umma.cuh"

namespace vbmp {
using namespace umma;

constexpr int GU_THREADS = 320;
constexpr int GU_SC = 16;            // samples per chunk (2 K-steps)
constexpr int GU_NR = 4;             // raw ring depth
constexpr int GU_NPMAX = 224;        // pair columns per CTA (D1 + D2 = 448 TMEM columns, A buffers = 64)
constexpr int GU_FL = 32;            // chunks per first-level accumulation block (512 samples)
constexpr int GU_CB = 128;           // components per CTA

struct GuArgs {
  int d0, d1;
  long long N; int K;
  int npb, NPB, P, ncb, splits;      // pair blocks, pairs per block, total pairs, component blocks, sample splits
  long long S_per;                   // samples per split (multiple of GU_SC)
  int Kp, PP;                        // padded partial dims: Kp = ncb*128, PP = npb*NPB
  int kcb;                           // columns of the R box = min(128, K)
  int FL;                            // chunks per first-level accumulation block
  float* part;                       // [splits][Kp][PP]
};

struct GuSmem {
  uint64_t rfull[GU_NR], rempty[GU_NR];
  uint64_t bfull[2], bempty[2];
  uint64_t dfull, dempty;
  uint32_t tmem_base;
  float consts[2];                   // {1, 0}: the padded "1" feature and the zero used by padding pair columns
};

#ifndef VBMP_GRAM_ROUND_LO
#define VBMP_GRAM_ROUND_LO 1
#endif
__device__ __forceinline__ void split_fast(float x, uint32_t& hi, uint32_t& lo) {
  // hi = round-to-nearest (ties away) TF32 of x, lo = round-to-nearest TF32 of the (exact) remainder.
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
#if VBMP_GRAM_ROUND_LO
  lo = (__float_as_uint(x - __uint_as_float(hi)) + 0x1000u) & 0xffffe000u;
#else
  lo = __float_as_uint(x - __uint_as_float(hi));     // the tensor core ignores the low 13 bits (probe: truncation)
#endif
}

__global__ void __launch_bounds__(GU_THREADS, 1)
gram_umma_kernel(const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmZ0,
                 const __grid_constant__ CUtensorMap tmZ1, GuArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int D = a.d0 + a.d1;
  // carve: raw ring [NR][R 16 x kcb floats | Z0 16 x d0 | Z1 16 x d1], B stages [2][hi NPB*64 B | lo NPB*64 B]
  const int rawR = GU_SC * a.kcb * 4, rawZ0 = GU_SC * a.d0 * 4, rawZ1 = GU_SC * a.d1 * 4;
  const int rawB = (rawR + rawZ0 + rawZ1 + 127) / 128 * 128;
  const int stageB = 2 * a.NPB * 64;
  uint8_t* raw = smem_raw;
  uint8_t* bst = raw + GU_NR * rawB;
  GuSmem* S = reinterpret_cast<GuSmem*>(bst + 2 * stageB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // task decode: blockIdx.x = (split * ncb + cb) * npb + pb
  const int pb = blockIdx.x % a.npb;
  const int cb = (blockIdx.x / a.npb) % a.ncb;
  const int split = blockIdx.x / (a.npb * a.ncb);
  const long long nb = (long long)split * a.S_per;
  long long ne = nb + a.S_per; if (ne > a.N) ne = a.N;
  const int nchunks = ne > nb ? (int)((ne - nb + GU_SC - 1) / GU_SC) : 0;
  const int NPB = a.NPB, FL = a.FL;

  if (tid == 0) {
    for (int s = 0; s < GU_NR; ++s) { mbar_init(&S->rfull[s], 1); mbar_init(&S->rempty[s], 256); }
    for (int s = 0; s < 2; ++s) { mbar_init(&S->bfull[s], 256); mbar_init(&S->bempty[s], 1); }
    mbar_init(&S->dfull, 1); mbar_init(&S->dempty, 256);
    S->consts[0] = 1.f; S->consts[1] = 0.f;
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmR); tma_prefetch_desc(&tmZ0); if (a.d1 > 0) tma_prefetch_desc(&tmZ1); }
  if (warp == 1) tmem_alloc<512>(&S->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = S->tmem_base;
  // TMEM columns: D1 [0,224), D2 [224,448), A buffers at 448 + 32*buf (hi 16 columns, lo 16 columns)

  if (warp == 0) {
    // ================= producer: one TMA box per operand per 16-sample chunk (rows past N are zero filled) ====
    const uint32_t bytes = (uint32_t)(rawR + rawZ0 + rawZ1);
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % GU_NR;
      mbar_wait(&S->rempty[s], ((c / GU_NR) & 1) ^ 1);
      if (elect_one()) {
        const int r0 = (int)(nb + (long long)c * GU_SC);
        uint8_t* dst = raw + (size_t)s * rawB;
        mbar_arrive_expect_tx(&S->rfull[s], bytes);
        tma_load_2d(dst, &tmR, cb * GU_CB, r0, &S->rfull[s]);
        tma_load_2d(dst + rawR, &tmZ0, 0, r0, &S->rfull[s]);
        if (a.d1 > 0) tma_load_2d(dst + rawR + rawZ0, &tmZ1, 0, r0, &S->rfull[s]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = idesc_tf32(128, NPB);
    for (int c = 0; c < nchunks; ++c) {
      const int st = c & 1;
      mbar_wait(&S->bfull[st], (c >> 1) & 1);
      const bool first = (c % FL) == 0;
      if (first && c > 0) mbar_wait(&S->dempty, ((c / FL) - 1) & 1);
      tc_fence_after();
      __syncwarp();
      if (elect_one()) {
        const uint32_t sbase = smem_u32(bst + (size_t)st * stageB);
        const uint32_t a_hi = tm + 448 + st * 32, a_lo = a_hi + 16;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t b_hi = smem_desc(sbase + ks * 2 * NPB * 16, NPB * 16, 128);
          const uint64_t b_lo = smem_desc(sbase + NPB * 64 + ks * 2 * NPB * 16, NPB * 16, 128);
          mma_tf32_ts(tm, a_lo + ks * 8, b_hi, idesc, !(first && ks == 0));
          mma_tf32_ts(tm, a_hi + ks * 8, b_lo, idesc, 1);
          mma_tf32_ts(tm, a_hi + ks * 8, b_hi, idesc, 1);
        }
        mma_commit(&S->bempty[st]);
        if ((c % FL) == FL - 1 || c == nchunks - 1) mma_commit(&S->dfull);
      }
      __syncwarp();
    }
  } else {
    // ================= workers =================
    const int w8 = warp - 2, q = warp & 3, sh = w8 >> 2;
    const int wtid = tid - 64;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int comp = q * 32 + lane;                         // component (TMEM lane) this thread feeds
    const bool comp_ok = comp < a.kcb && cb * GU_CB + comp < a.K;
    // pair owned by this thread in the phi generation (row-major upper triangle over D+1 features); the two
    // factors are read through (base, per-slot stride, per-sample stride) triples so the inner loop is branch free
    const uint8_t* bi; const uint8_t* bj; int sli, slj, sti, stj;
    {
      const int pg_ = pb * NPB + wtid;
      const bool pair_ok = (wtid < NPB) && (pg_ < a.P);
      int pi = 0, pj = 0;
      if (pair_ok) {
        int rem = pg_, i = 0, len = D + 1;
        while (rem >= len) { rem -= len; ++i; --len; }
        pi = i; pj = i + rem;
      }
      auto setup = [&](int f, const uint8_t*& b, int& sl, int& stv) {
        if (!pair_ok) { b = reinterpret_cast<const uint8_t*>(&S->consts[1]); sl = 0; stv = 0; }
        else if (f < a.d0) { b = raw + rawR + f * 4; sl = rawB; stv = a.d0 * 4; }
        else if (f < D) { b = raw + rawR + rawZ0 + (f - a.d0) * 4; sl = rawB; stv = a.d1 * 4; }
        else { b = reinterpret_cast<const uint8_t*>(&S->consts[0]); sl = 0; stv = 0; }
      };
      setup(pi, bi, sli, sti);
      setup(pj, bj, slj, stj);
    }
    int nflush = 0;
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % GU_NR, st = c & 1;
      mbar_wait(&S->rfull[s], (c / GU_NR) & 1);
      mbar_wait(&S->bempty[st], ((c >> 1) & 1) ^ 1);
      tc_fence_after();
      // ---- A operand: r[s][comp] for this thread's 8 samples, split, into TMEM
      {
        const float* rawr = reinterpret_cast<const float*>(raw + (size_t)s * rawB) + (sh * 8) * a.kcb + comp;
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float r = comp_ok ? rawr[u * a.kcb] : 0.f;
          split_fast(r, hi[u], lo[u]);
        }
        const uint32_t ad = tm + lane_base + 448 + st * 32 + sh * 8;
        tmem_st8(ad, hi);
        tmem_st8(ad + 16, lo);
      }
      // ---- B operand: phi[s][pair] = zt[s][i] * zt[s][j] for 16 samples, split, K-major core-matrix layout
      if (wtid < NPB) {
        uint8_t* bh = bst + (size_t)st * stageB + (size_t)wtid * 16;
        uint8_t* bl = bh + NPB * 64;
        const uint8_t* zi = bi + s * sli;
        const uint8_t* zj = bj + s * slj;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int sl = qd * 4 + u;
            const float v = *reinterpret_cast<const float*>(zi + sl * sti) * *reinterpret_cast<const float*>(zj + sl * stj);
            split_fast(v, hi[u], lo[u]);
          }
          *reinterpret_cast<uint4*>(bh + (size_t)qd * NPB * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(bl + (size_t)qd * NPB * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      mbar_arrive(&S->rempty[s]);
      fence_proxy_async();
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&S->bfull[st]);

      const bool last = (c == nchunks - 1);
      if ((c % FL) == FL - 1 || last) {
        // ---- fold D1 into D2 (or, at the end, write D1 + D2 to this split's partial)
        mbar_wait(&S->dfull, nflush & 1);
        tc_fence_after();
        const bool firstf = (nflush == 0);
        float* prow = a.part + ((size_t)split * a.Kp + (size_t)cb * GU_CB + comp) * a.PP + (size_t)pb * NPB;
        for (int c0 = sh * 112; c0 < sh * 112 + 112 && c0 < NPB; c0 += 16) {
          float v1[16], v2[16];
          tmem_ld16(tm + lane_base + c0, v1);
          if (!firstf) tmem_ld16(tm + lane_base + 224 + c0, v2);
          tmem_wait_ld();
          if (!firstf) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v1[j] += v2[j];
          }
          if (last) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(prow + c0 + j) = make_float4(v1[j], v1[j + 1], v1[j + 2], v1[j + 3]);
          } else {
            tmem_st16(tm + lane_base + 224 + c0, reinterpret_cast<const uint32_t*>(v1));
          }
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&S->dempty);
        ++nflush;
      }
    }
    if (nchunks == 0) {      // empty split: contribute zeros
      float* prow = a.part + ((size_t)split * a.Kp + (size_t)cb * GU_CB + comp) * a.PP + (size_t)pb * NPB;
      for (int c0 = sh * 112; c0 < sh * 112 + 112 && c0 < NPB; ++c0) prow[c0] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tm);
}

// gram[k][i][j] = gram[k][j][i] = sum_split part[split][k][pair(i,j)], fp64, fixed order.
__global__ void gram_pair_reduce_kernel(const float* __restrict__ part, int splits, int K, int Kp, int PP, int D1,
                                        float* __restrict__ gram) {
  const int P = D1 * (D1 + 1) / 2;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)K * P) return;
  const int k = (int)(e / P);
  int rem = (int)(e % P), i = 0, len = D1;
  while (rem >= len) { rem -= len; ++i; --len; }
  const int j = i + rem;
  double acc = 0.0;
  for (int s = 0; s < splits; ++s) acc += (double)part[((size_t)s * Kp + k) * PP + (e % P)];
  const float v = (float)acc;
  gram[((size_t)k * D1 + i) * D1 + j] = v;
  gram[((size_t)k * D1 + j) * D1 + i] = v;
}

// ---- host side -------------------------------------------------------------------------------------------
static int gu_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static void gu_plan(long long N, int K, int D, int sms, GuArgs* g) {
  const int D1 = D + 1;
  g->P = D1 * (D1 + 1) / 2;
  g->npb = (g->P + GU_NPMAX - 1) / GU_NPMAX;
  g->NPB = ((g->P + g->npb - 1) / g->npb + 15) / 16 * 16;
  if (g->NPB < 16) g->NPB = 16;
  g->ncb = (K + GU_CB - 1) / GU_CB;
  const int tasks = g->npb * g->ncb;
  long long sp = sms / tasks;
  const long long maxsp = (N + 2047) / 2048;
  if (sp > maxsp) sp = maxsp;
  if (sp < 1) sp = 1;
  long long per = (N + sp - 1) / sp;
  per = (per + GU_SC - 1) / GU_SC * GU_SC;
  g->S_per = per;
  g->splits = (int)((N + per - 1) / per);
  if (g->splits < 1) g->splits = 1;
  g->Kp = g->ncb * GU_CB;
  g->PP = g->npb * g->NPB;
}

bool gram_umma_supported(long long N, int GX, int GP, int G, int K, int Dp, int d0, int d1, bool has_p) {
  const int D = d0 + d1;
  return has_p && G == 1 && GX == 1 && GP == 1 && Dp >= 16 && Dp <= 64 && D <= 64 && (K % 4 == 0) && (d0 % 4 == 0) &&
         (d1 % 4 == 0) && N >= 2048;
}

size_t gram_umma_workspace_bytes(long long N, int G, int K, int d0, int d1, int Dp) {
  if (!gram_umma_supported(N, 1, 1, G, K, Dp, d0, d1, true)) return 0;
  GuArgs g{};
  gu_plan(N, K, d0 + d1, 512, &g);       // upper bound on the SM count -> upper bound on the number of splits
  return (size_t)g.splits * g.Kp * g.PP * sizeof(float) + 512;
}

static int gu_fl() {
  static int fl = 0;
  if (fl == 0) {
    const char* e = getenv("VBMP_GRAM_FL");     // tuning knob: chunks (16 samples) per first-level accumulation block
    fl = e ? atoi(e) : GU_FL;
    if (fl < 1) fl = GU_FL;
  }
  return fl;
}

int launch_gram_umma(const GramArgs& a, float* gram, void* ws, size_t ws_bytes, cudaStream_t st) {
  GuArgs g{};
  g.d0 = a.d0; g.d1 = a.d1; g.N = a.N; g.K = a.K;
  const int D = a.d0 + a.d1;
  gu_plan(a.N, a.K, D, gu_num_sms(), &g);
  g.kcb = a.K < GU_CB ? a.K : GU_CB;
  g.FL = gu_fl();
  const size_t need = (size_t)g.splits * g.Kp * g.PP * sizeof(float) + 512;
  if (ws_bytes < need) { set_error("gram_umma: workspace too small (%zu < %zu)", ws_bytes, need); return VBMP_ERR_WORKSPACE; }
  g.part = (float*)(((size_t)ws + 255) / 256 * 256);
  CUtensorMap tmR, tmZ0, tmZ1;
  int e = make_tmap_2d(&tmR, a.p, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.K, (uint32_t)g.kcb, GU_SC);
  if (!e) e = make_tmap_2d(&tmZ0, a.z0, (uint64_t)a.d0, (uint64_t)a.N, (uint64_t)a.d0, (uint32_t)a.d0, GU_SC);
  if (!e && a.d1 > 0) e = make_tmap_2d(&tmZ1, a.z1, (uint64_t)a.d1, (uint64_t)a.N, (uint64_t)a.d1, (uint32_t)a.d1, GU_SC);
  if (a.d1 == 0) tmZ1 = tmZ0;
  if (e) { set_error("gram_umma: cuTensorMapEncodeTiled failed (%d)", e); return VBMP_ERR_CUDA; }
  const int rawB = (GU_SC * g.kcb * 4 + GU_SC * a.d0 * 4 + GU_SC * a.d1 * 4 + 127) / 128 * 128;
  const size_t smem = (size_t)GU_NR * rawB + (size_t)2 * 2 * g.NPB * 64 + sizeof(GuSmem) + 64;
  cudaFuncSetAttribute(gram_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = g.splits * g.ncb * g.npb;
  gram_umma_kernel<<<grid, GU_THREADS, smem, st>>>(tmR, tmZ0, tmZ1, g);
  int rc = check_launch("gram_umma");
  if (rc) return rc;
  const int D1 = D + 1;
  const long long tot = (long long)a.K * g.P;
  gram_pair_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g.part, g.splits, a.K, g.Kp, g.PP, D1, gram);
  return check_launch("gram_pair_reduce");
}

}  // namespace vbmp
