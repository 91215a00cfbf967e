// K2 (tcgen05 variant): whitened quadratic forms on the 5th-gen tensor cores + fused responsibility softmax.
//
//   l[n,k] = cst[k] - 1/2 || W_k^T z_n - m_k ||^2          (NIW/MNW.Elog_like; dists/NormalInverseWishart.py:91-97,
//                                                           transforms/MatrixNormalWishart.py:219-232)
//   mode 1: p = exp(l - logZ_n), logZ_n, NA_k, sum_n logZ_n (Mixture.update_assignments dists/Mixture.py:38-45,
//                                                           MixtureofLinearTransforms.update_assignments :34-41)
//
// Design (numbers from tools/umma_probe.cu on a B200, see umma.cuh):
//   * persistent CTA per SM, 256-sample tiles (two 128-row halves).  The sample tile is the A operand and lives in TENSOR
//     MEMORY as two split-precision images (z = hi + lo): with A in TMEM the MMA issue rate is set by N alone, whereas a
//     shared-memory A costs 32 extra cycles of operand fetch per MMA.
//   * the whitening factors are the B operand: one 128-column group (CG = 128/DP components) per pipeline stage,
//     pre-split into two images and pre-arranged in the K-major core-matrix layout by a pack kernel, so a stage (plus the
//     group's -m block, cst and scales) is ONE cp.async.bulk.  Each stage is used by both halves (hi*hi + lo*hi + hi*lo
//     per half and K-step), which halves the L2 -> shared-memory stream per flop.
//   * W_k is upper triangular.  The group's columns are interleaved in 8-column blocks, n = (j/8)*(8 CG) + cl*8 + j%8,
//     so the columns K-step ks can reach (j >= KSTEP ks) are the contiguous suffix n >= KSTEP CG ks: the MMA is issued
//     with N = 128 - KSTEP CG ks on that suffix and only that suffix of B is stored / copied (DP = 64: 62.5 % of the
//     dense work with KSTEP = 16, 56 % with KSTEP = 8).
//   * 128-column fp32 accumulators in TMEM, rotating over the (group, half) sequence (two with TF32 operands, three
//     with fp16 ones); the two MMA-issuing warps (one per half) take strict turns, so the epilogue of one half runs
//     under the MMAs of the other.  The 8 epilogue warps (one thread per sample row) read D with tcgen05.ld, subtract m,
//     square-reduce with packed fp32x2 FFMA2, keep an online logsumexp and write the logits.  TF32 variant: each burst
//     starts with one TF32 MMA of a small shared-memory A block of ones against [-m_hi; -m_lo], which initialises the
//     accumulator to -m; the fp16 variant does the same with the row's scale in the A block.  That MMA is 1 of the 8.5
//     full-width MMA slots of a burst and the kernel runs at the (power-capped) tensor-pipe rate, so folding -m into the
//     epilogue's rescaling FFMA2 instead looked like 12 % — but the values have to come from shared memory, and both
//     ways of reading them (EU_MFOLD below) load the shared-memory pipe, which the MMAs' operand fetch already keeps
//     busy, by more than the MMA they save.
//   * mode 1: a 12th warp normalises tile t (p = exp(l - logZ_n), in place, L2 hits, coalesced float4, ex2.approx) while
//     the MMAs of tile t + 1 run, keeps NA of its fixed columns in a fixed order (deterministic) and, on request, also
//     writes the responsibilities pre-split into the Gram kernel's fp16 operand images (EstepArgs::rpack).
//
// Operand precision (template parameter F16, the default; VBMP_ESTEP_PREC=tf32 selects the other):
//   * TF32: z and W are split hi + lo into TF32 (3 MMAs with K = 8 per K-step; the tensor core truncates lo to 10 bits,
//     product error ~2^-21).
//   * FP16: each sample row is scaled by 2^sh_n (row maximum -> [2^10, 2^11)) and each component's W by 2^t_k (same
//     normalisation), both EXACT; then z' = a + b and W' = A + B with a, b, A, B fp16 (a = rn(z'), b = rn(z' - a): 22
//     significant bits, product error ~2^-22, so it is at least as accurate as the TF32 split) and the same three terms
//     run as kind::f16 MMAs with K = 16, which issue at twice the TF32 rate and halve the operand bytes (stage 21 KB
//     instead of 41 KB, A 64 instead of 128 TMEM columns per half).  The -m fold stays a TF32 MMA whose A block holds
//     2^sh_n per row and whose B block holds -m 2^t_k; the epilogue multiplies by 2^-(sh_n + t_k) (exact) before squaring,
//     so nothing can overflow that would not overflow in the unscaled form.  Zero rows get sh = 0; sh and t are clamped to
//     +-60.  Before that, feature i of z is multiplied by 2^e_i and row i of every W_k by 2^-e_i (e_i = exponent of the
//     largest |W_k[i][.]| over all components, estep_rowscale_kernel): the product is unchanged, and features measured in
//     very different units (tests: 10 decades apart) all land inside the fp16 window — W's rows carry 1/sigma_i, so this
//     is a whitening by the narrowest component.
#include <cstdlib>
#include <type_traits>
#include <cuda_fp16.h>
#include "common.cuh"
#include "umma.cuh"

namespace vbmp {
using namespace umma;

#ifndef EU_MFOLD
// fp16 variant, where -m is folded in.  Measured at cfg2, in the EM loop (profiles/r02_summary.md, "E-step: where -m goes"):
//   0 = by a TF32 MMA that initialises the accumulator                                       13.4 ms   <- built
//   1 = in the epilogue on m16n8-style fragments (tcgen05.ld.16x256b, quad reduce-scatter)   16.1 ms
//   (in the epilogue with one row per thread and warp-broadcast loads of all 128 values      15.3 ms)
//   2 = nowhere (wrong results; the ceiling of any MMA-free fold)                            12.1 ms
#define EU_MFOLD 0
#endif
constexpr int EU_THREADS = 384;     // warp 0: bulk-copy producer, warps 1-2: MMA issuers (one per half), warps 3..10: workers,
                                    // warp 11: normaliser (mode 1)
constexpr int EU_TILE = 256;
constexpr int EU_MAXSTAGE = 8;     // the launcher picks the number of stages that fits (5 at DP = 64, K <= 256)
constexpr int EU_MAXK = 512;
constexpr int EU_N = 128;           // columns per group
constexpr int EU_FS = 128;          // feature-scale table: fs[i] = 2^e_i, fs[EU_FS + i] = 2^-e_i

template <int DP, bool F16>
struct EuCfg {
  static constexpr int CG = EU_N / DP;        // components per 128-column MMA group
  static constexpr int KSTEP = F16 ? 16 : 8;  // reduction length of one MMA
  static constexpr int KS = DP / KSTEP;       // K-steps
  static constexpr int ACOLS = F16 ? DP / 2 : DP;   // TMEM columns of one A image (hi or lo) of a half
  // accumulators: the (group, half) pairs rotate through NBUF 128-column buffers.  With fp16 operands a half's MMAs take
  // ~630 cycles, less than the commit -> tcgen05.ld -> release round trip of the other half's accumulator, so a third
  // buffer (the A images only need 128 columns up to DP = 64) gives that round trip two bursts to complete.  DP = 128
  // (fp16 operands only): the A images of the two halves take 4 x 64 = 256 columns, which leaves two accumulators.
  static constexpr int DCOL0 = 4 * ACOLS > 128 ? 4 * ACOLS : 128;
  static constexpr int NBUF = (512 - DCOL0) / EU_N >= 3 ? 3 : 2;
  static_assert(DCOL0 + NBUF * EU_N <= 512, "TMEM budget");
  __host__ __device__ static constexpr int n0(int ks) { return ks * CG * KSTEP; }
  __host__ __device__ static constexpr int nn(int ks) { return EU_N - n0(ks); }
  __host__ __device__ static constexpr int blk_off(int ks) {
    int o = 0;
    for (int i = 0; i < ks; ++i) o += nn(i) * 32;
    return o;
  }
  static constexpr int WB = blk_off(KS);      // bytes of one operand image (hi or lo) of a group
  static constexpr int GB = 2 * WB;           // hi image then lo image
  // the "-m" block.  TF32: one K-step of a B operand (k = 0: -m_hi, k = 1: -m_lo, rest 0) for the MMA that initialises the
  // accumulator.  fp16: the plain fp32 values -m[n] in column order, subtracted in the epilogue (see the kernel header).
  static constexpr int MB = (F16 && EU_MFOLD) ? EU_N * 4 : EU_N * 32;
  static constexpr int REC = GB + MB + 64;    // + cst of the CG components (floats 0..7) and 2^-t_k (floats 8..15)
  static constexpr int STAGE = (REC + 127) / 128 * 128;
  // component / feature of column n
  __host__ __device__ static constexpr int col_cl(int n) { return (n % (CG * 8)) / 8; }
  __host__ __device__ static constexpr int col_j(int n) { return (n / (CG * 8)) * 8 + n % 8; }
  // descriptor of the -m block relative to a stage at shared address 0
  __host__ __device__ static constexpr uint64_t desc_m() {
    return (uint64_t)(GB >> 4) | ((uint64_t)((EU_N * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
  }
  // descriptor of K-step block ks (lo = 0/1) relative to a stage at shared address 0 (add stage_addr >> 4)
  __host__ __device__ static constexpr uint64_t desc0(int ks, int lo) {
    return (uint64_t)((blk_off(ks) + lo * WB) >> 4) | ((uint64_t)((nn(ks) * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) |
           (1ull << 46);
  }
};

// ---- pack: W (C, DP, DP) fp32 row-major [i][j] -> per group [hi | lo | m | cst, 2^-t] record; B[n][k] = W_c[i = k][j]
// with (c, j) = column n as above; K-step block ks holds rows n >= n0(ks) as [chunk (2)][row][16 bytes]: 4 TF32 or
// 8 fp16 per chunk, i = KSTEP ks + (KSTEP / 2) chunk + e.
// fp16 operands: fs[i] = 2^e_i (applied to feature i of the samples), fs[EU_FS + i] = 2^-e_i (applied to row i of every W_k)
template <int DP>
__global__ void estep_rowscale_kernel(const float* __restrict__ W, int K, float* __restrict__ fs) {
  __shared__ uint32_t mx;
  const int i = blockIdx.x;
  if (threadIdx.x == 0) mx = 0u;
  __syncthreads();
  uint32_t v = 0u;
  for (int o = threadIdx.x; o < K * DP; o += blockDim.x)
    v = max(v, __float_as_uint(fabsf(W[((size_t)(o / DP) * DP + i) * DP + o % DP])) & 0x7f800000u);
  atomicMax(&mx, v);
  __syncthreads();
  if (threadIdx.x == 0) {
    int e = 0;
    if (mx != 0u) { e = (int)(mx >> 23) - 127; e = e > 60 ? 60 : (e < -60 ? -60 : e); }
    fs[i] = __uint_as_float((uint32_t)(127 + e) << 23);
    fs[EU_FS + i] = __uint_as_float((uint32_t)(127 - e) << 23);
  }
}

template <int DP, bool F16>
__global__ void estep_pack_kernel(const float* __restrict__ W, const float* __restrict__ m, const float* __restrict__ cst,
                                  int K, const float* __restrict__ fs, uint8_t* __restrict__ Wp) {
  using C = EuCfg<DP, F16>;
  const int g = blockIdx.x;
  __shared__ uint32_t wmax[8];
  __shared__ float wsc[8];              // 2^t_k of the group's components (1 for TF32)
  if (threadIdx.x < 8) wmax[threadIdx.x] = 0u;
  __syncthreads();
  if (F16) {
    for (int o = threadIdx.x; o < C::CG * DP * DP; o += blockDim.x) {
      const int c = g * C::CG + o / (DP * DP);
      if (c < K)
        atomicMax(&wmax[o / (DP * DP)],
                  __float_as_uint(fabsf(W[(size_t)c * DP * DP + o % (DP * DP)] * fs[EU_FS + (o % (DP * DP)) / DP])) & 0x7f800000u);
    }
    __syncthreads();
  }
  if (threadIdx.x < 8) {
    int t = 0;
    if (F16 && wmax[threadIdx.x] != 0u) {
      t = 10 - ((int)(wmax[threadIdx.x] >> 23) - 127);
      t = t > 60 ? 60 : (t < -60 ? -60 : t);
    }
    wsc[threadIdx.x] = __uint_as_float((uint32_t)(127 + t) << 23);
  }
  __syncthreads();
  // -m as a K-major operand block [chunk (2)][column n (128)][4 floats]: k = 0 holds -m_hi, k = 1 holds -m_lo.  It meets a
  // constant A block of ones (k = 0, 1; 2^sh_n for fp16) in ONE extra MMA that initialises the accumulator to -m, so y - m
  // comes out of the tensor core and the epilogue only squares: the epilogue's -m loads were ~900 shared-memory wavefronts
  // per group, as much as the MMA operand fetch, and the shared-memory pipe is what both compete for.
  float* mo = reinterpret_cast<float*>(Wp + (size_t)g * C::REC + C::GB);
  float* co = reinterpret_cast<float*>(Wp + (size_t)g * C::REC + C::GB + C::MB);
  if (F16 && EU_MFOLD) {
    // fp16 operands: the epilogue computes (acc 2^-(sh_n + t_k)) - m with one FFMA2 per column pair, so the record holds
    // the unscaled -m in column order (512 bytes, read with warp-broadcast 16-byte loads)
    for (int n = threadIdx.x; n < EU_N; n += blockDim.x) {
      const int c = g * C::CG + C::col_cl(n);
      mo[n] = c < K ? -m[(size_t)c * DP + C::col_j(n)] : 0.f;
    }
  }
  for (int o = threadIdx.x; o < ((F16 && EU_MFOLD) ? 0 : EU_N * 8); o += blockDim.x) {
    const int ch = o / (EU_N * 4), n = (o % (EU_N * 4)) / 4, e = o % 4;
    float v = 0.f;
    if (ch == 0 && e < 2) {
      const int cl = C::col_cl(n), c = g * C::CG + cl;
      const float mv = c < K ? m[(size_t)c * DP + C::col_j(n)] * wsc[cl] : 0.f;
      uint32_t h, l;
      split_tf32(mv, h, l);
      v = e == 0 ? -__uint_as_float(h) : -__uint_as_float(l);
    }
    mo[o] = v;
  }
  for (int o = threadIdx.x; o < 16; o += blockDim.x) {
    const int cl = o & 7, c = g * C::CG + cl;
    const bool ok = cl < C::CG && c < K;
    co[o] = o < 8 ? (ok ? cst[c] : 0.f) : (ok ? 1.f / wsc[cl] : 1.f);
  }
  if (F16) {
    __half* hi = reinterpret_cast<__half*>(Wp + (size_t)g * C::REC);
    __half* lo = reinterpret_cast<__half*>(Wp + (size_t)g * C::REC + C::WB);
    for (int o = threadIdx.x; o < C::WB / 2; o += blockDim.x) {
      int ks = 0, rem = o * 2;
      while (ks + 1 < C::KS && rem >= C::nn(ks) * 32) { rem -= C::nn(ks) * 32; ++ks; }
      const int nn = C::nn(ks);
      const int ch = rem / (nn * 16);
      const int r = (rem % (nn * 16)) / 16, e = (rem % 16) / 2;
      const int n = C::n0(ks) + r, cl = C::col_cl(n), j = C::col_j(n), i = 16 * ks + 8 * ch + e;
      const int c = g * C::CG + cl;
      const float v = (c < K) ? W[((size_t)c * DP + i) * DP + j] * fs[EU_FS + i] * wsc[cl] : 0.f;     // both factors: powers of 2
      const __half a = __float2half_rn(v);
      hi[o] = a;
      lo[o] = __float2half_rn(v - __half2float(a));
    }
  } else {
    float* hi = reinterpret_cast<float*>(Wp + (size_t)g * C::REC);
    float* lo = reinterpret_cast<float*>(Wp + (size_t)g * C::REC + C::WB);
    for (int o = threadIdx.x; o < C::WB / 4; o += blockDim.x) {
      int ks = 0, rem = o * 4;
      while (ks + 1 < C::KS && rem >= C::nn(ks) * 32) { rem -= C::nn(ks) * 32; ++ks; }
      const int nn = C::nn(ks);
      const int ch = rem / (nn * 16);
      const int r = (rem % (nn * 16)) / 16, e = (rem % 16) / 4;
      const int n = C::n0(ks) + r, cl = C::col_cl(n), j = C::col_j(n), i = 8 * ks + 4 * ch + e;
      const int c = g * C::CG + cl;
      const float v = (c < K) ? W[((size_t)c * DP + i) * DP + j] : 0.f;
      uint32_t h, l;
      split_tf32(v, h, l);
      hi[o] = __uint_as_float(h);
      lo[o] = __uint_as_float(l);
    }
  }
}

struct EuSmem {
  uint64_t full[EU_MAXSTAGE], empty[EU_MAXSTAGE];
  uint64_t tfull[3], tempty[3];
  uint64_t afull;
  uint64_t turn[2];           // issue token between the two MMA warps
  uint64_t ndone[2], nfree[2];  // mode 1: tile's logits + logZ_n complete (workers -> normaliser) / lz buffer free again
  uint32_t tmem_base;
  alignas(16) float fsc[EU_FS];  // fp16 operands: 2^e_i per feature
  double red[8];
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// position in a ring of n slots: slot index and the parity of the number of completed laps (no division in the loops)
struct Ring {
  int s = 0;
  uint32_t ph = 0;
  __device__ __forceinline__ void next(int n) { if (++s == n) { s = 0; ph ^= 1u; } }
};
// accumulator-buffer rotation: (group, half) pair number seq = 2 it + h uses buffer seq % NBUF for the (seq / NBUF)-th time
template <int NBUF>
struct BufRing {
  uint32_t buf, ph = 0;
  __device__ __forceinline__ explicit BufRing(int h) : buf(h % NBUF), ph(h / NBUF) {}
  __device__ __forceinline__ void next() { buf += 2; if (buf >= NBUF) { buf -= NBUF; ph ^= 1u; } }
};

// AB2: TWO sets of A images in tensor memory.  The workers then build the images of tile t + 1 before the LAST epilogue
// of tile t instead of behind it, and the MMA warps run on into the next tile while that epilogue is still going: with
// one set the tensor pipe idles from the last burst of a tile until the next images are stored (~8.7 K cycles per tile:
// 5 % of the kernel at K = 256, 20 % at K = 64, where a tile is only 64 bursts long).  The second set costs 4 ACOLS
// columns: nothing up to DP = 32, the third accumulator buffer at DP = 64 — where that loss outweighs the overlap
// (measured, see eu_launch) —, impossible at DP = 128.
template <int DP, int MODE, bool F16, bool AB2>
__global__ void __launch_bounds__(EU_THREADS, 1)
estep_umma_kernel(EstepArgs a, const uint8_t* __restrict__ Wp, const float* __restrict__ fs, int ntiles, int ngroups,
                  int nstage, long long nsb) {
  using C = EuCfg<DP, F16>;
  constexpr int ASZ = 4 * C::ACOLS;                                        // columns of one set of A images (both halves)
  constexpr int DCOL0 = AB2 ? (2 * ASZ > 128 ? 2 * ASZ : 128) : C::DCOL0;
  constexpr int NBUF = AB2 ? ((512 - DCOL0) / EU_N >= 3 ? 3 : 2) : C::NBUF;
  static_assert(DCOL0 + NBUF * EU_N <= 512 && NBUF >= 2, "TMEM budget");
  constexpr int ONES_BYTES = AB2 ? 16384 : 8192;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* stages = smem_raw;                                             // EU_NSTAGE * STAGE
  float* ones = reinterpret_cast<float*>(stages + nstage * C::STAGE);     // per A set and half: A block [chunk (2)][row (128)][4]
                                                                          // with 1 (TF32) or 2^sh_row (fp16) at k = 0, 1
  EuSmem* S = reinterpret_cast<EuSmem*>(stages + nstage * C::STAGE + ONES_BYTES);
  float* lz = reinterpret_cast<float*>(S + 1);                            // [2][256] (tile parity)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = a.K;
  const int ldo = (MODE == 0 && a.ldo > 0) ? a.ldo : K;      // row stride of the logits (mode 0 only)

  if (tid == 0) {
    for (int s = 0; s < nstage; ++s) { mbar_init(&S->full[s], 1); mbar_init(&S->empty[s], 2 + 256); }
    for (int b = 0; b < 3; ++b) { mbar_init(&S->tfull[b], 1); mbar_init(&S->tempty[b], 128); }
    mbar_init(&S->afull, 256);
    mbar_init(&S->turn[0], 1); mbar_init(&S->turn[1], 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&S->ndone[b], 256); mbar_init(&S->nfree[b], 1); }
    fence_barrier_init();
  }
  for (int o = tid; o < ONES_BYTES / 4; o += EU_THREADS) ones[o] = ((o & 1023) < 512 && (o & 3) < 2) ? 1.f : 0.f;
  if (F16 && tid < EU_FS) S->fsc[tid] = tid < DP ? fs[tid] : 1.f;
  fence_proxy_async();
  if (warp == 1) tmem_alloc<512>(&S->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = S->tmem_base;
  // TMEM columns: A hi/lo of half h at h*2*ACOLS (+ACOLS for lo); accumulator buffer b at DCOL0 + 128 b
  const int my_tiles = (ntiles > (int)blockIdx.x) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    // ================= producer: one bulk copy per 64-column group =================
    Ring rg;
    for (int t = 0; t < my_tiles; ++t)
      for (int g = 0; g < ngroups; ++g, rg.next(nstage)) {
        const int s = rg.s;
        mbar_wait(&S->empty[s], rg.ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&S->full[s], C::REC);
          bulk_g2s(stages + (size_t)s * C::STAGE, Wp + (size_t)g * C::REC, C::REC, &S->full[s]);
        }
        __syncwarp();
      }
  } else if (warp <= 2) {
    // ================= MMA issuers: warp 1 drives half 0, warp 2 half 1 =================
    // (one issuer per half: the barrier waits / descriptor set-up of one half overlap the other half's MMAs).  The two
    // take strict turns through a token: if their bursts interleaved in the tensor-pipe FIFO both halves would finish
    // together and both epilogue round trips would be exposed; in turn order half 0's epilogue runs under half 1's MMAs.
    const int h = warp - 1;
    Ring rg;
    BufRing<NBUF> br(h);
    uint32_t itp = 0;                      // parity of the group counter
    for (int t = 0; t < my_tiles; ++t) {
      const int ab = AB2 ? (t & 1) : 0;
      const uint32_t a_hi = tm + ab * ASZ + h * 2 * C::ACOLS, a_lo = a_hi + C::ACOLS;
      const uint64_t ones_desc = smem_desc(smem_u32(ones + ab * 2048 + h * 1024), 128 * 16, 128);
      mbar_wait(&S->afull, t & 1);
      tc_fence_after();
      for (int g = 0; g < ngroups; ++g, rg.next(nstage), br.next(), itp ^= 1u) {
        const int s = rg.s;
        mbar_wait(&S->full[s], rg.ph);
        const uint64_t sb = (uint64_t)((smem_u32(stages) + (uint32_t)s * C::STAGE) >> 4);
        const uint32_t buf = br.buf;
        const uint32_t dcol = tm + DCOL0 + buf * EU_N;
        mbar_wait(&S->tempty[buf], br.ph ^ 1);
        mbar_wait(&S->turn[h], h == 0 ? (itp ^ 1) : itp);     // my turn
        tc_fence_after();
        if (elect_one()) {
          if (!(F16 && EU_MFOLD)) mma_tf32_ss(dcol, ones_desc, C::desc_m() + sb, idesc_tf32(128, EU_N), 0);   // D = -m
#pragma unroll
          for (int ks = 0; ks < C::KS; ++ks) {
            const uint64_t b_hi = C::desc0(ks, 0) + sb, b_lo = C::desc0(ks, 1) + sb;
            if (F16) {
              const uint32_t idesc = idesc_f16(128, C::nn(ks));
              // K-step 0 spans all 128 columns: its first MMA overwrites the accumulator
              mma_f16_ts(dcol + C::n0(ks), a_lo + ks * 8, b_hi, idesc, (EU_MFOLD && ks == 0) ? 0 : 1);       // small terms first
              mma_f16_ts(dcol + C::n0(ks), a_hi + ks * 8, b_lo, idesc, 1);
              mma_f16_ts(dcol + C::n0(ks), a_hi + ks * 8, b_hi, idesc, 1);
            } else {
              const uint32_t idesc = idesc_tf32(128, C::nn(ks));
              mma_tf32_ts(dcol + C::n0(ks), a_lo + ks * 8, b_hi, idesc, 1);
              mma_tf32_ts(dcol + C::n0(ks), a_hi + ks * 8, b_lo, idesc, 1);
              mma_tf32_ts(dcol + C::n0(ks), a_hi + ks * 8, b_hi, idesc, 1);
            }
          }
          mma_commit(&S->tfull[buf]);
          mma_commit(&S->empty[s]);
          mbar_arrive(&S->turn[h ^ 1]);                 // pass the token
        }
        __syncwarp();
      }
    }
  } else if (warp <= 10) {
    // ================= workers: A tile -> TMEM, epilogue =================
    const int w8 = warp - 3, h = w8 >> 2, q = warp & 3;
    const int wtid = tid - 96;                                  // 0..255
    const int rloc = h * 128 + q * 32 + lane;                   // row within the tile
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int D = a.d0 + a.d1;
    double lzsum = 0.0;
    Ring rg;
    BufRing<NBUF> br(h);
    // ---- this thread's sample row of tile tt, split into hi / lo, into the A images of set tt % 2 (AB2) in tensor memory;
    //      returns 2^-sh of the row (fp16 operands).  The MMAs that last read that set are complete: AB2 — those of tile
    //      tt - 2 (this thread has seen tfull of every group of tile tt - 1 but the last); otherwise those of tile tt - 1.
    auto load_A = [&](int tt) -> float {
      const long long row = ((long long)blockIdx.x + (long long)tt * gridDim.x) * EU_TILE + rloc;
      const bool valid = row < a.N;
      const int ab = AB2 ? (tt & 1) : 0;
      float rs = 1.f;
        const uint32_t a_hi = tm + lane_base + ab * ASZ + h * 2 * C::ACOLS, a_lo = a_hi + C::ACOLS;
        const float* r0 = a.z0 + (size_t)(valid ? row : 0) * a.d0;
        const float* r1 = a.z1 ? a.z1 + (size_t)(valid ? row : 0) * a.d1 : nullptr;
        const bool vec = valid && (a.d0 % 4 == 0) && (a.d1 % 4 == 0);
        auto load16 = [&](int c0, float* v) {
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
        };
        if (F16) {
          float v[DP];
          float mxa = 0.f;
#pragma unroll
          for (int c0 = 0; c0 < DP; c0 += 16) {
            load16(c0, v + c0);
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 f4 = *reinterpret_cast<const float4*>(&S->fsc[c0 + j]);
              v[c0 + j] *= f4.x; v[c0 + j + 1] *= f4.y; v[c0 + j + 2] *= f4.z; v[c0 + j + 3] *= f4.w;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) mxa = fmaxf(mxa, fabsf(v[c0 + j]));
          }
          int sh = 0;
          if (mxa > 0.f) {
            sh = 10 - ((int)(__float_as_uint(mxa) >> 23) - 127);
            sh = sh > 60 ? 60 : (sh < -60 ? -60 : sh);
          }
          const float sc = __uint_as_float((uint32_t)(127 + sh) << 23);
          rs = __uint_as_float((uint32_t)(127 - sh) << 23);
#pragma unroll
          for (int c0 = 0; c0 < DP; c0 += 16) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float x0 = v[c0 + 2 * j] * sc, x1 = v[c0 + 2 * j + 1] * sc;
              const __half2 ah = __floats2half2_rn(x0, x1);
              const float2 af = __half22float2(ah);
              const __half2 bh = __floats2half2_rn(x0 - af.x, x1 - af.y);
              hi[j] = *reinterpret_cast<const uint32_t*>(&ah);
              lo[j] = *reinterpret_cast<const uint32_t*>(&bh);
            }
            tmem_st8(a_hi + c0 / 2, hi);
            tmem_st8(a_lo + c0 / 2, lo);
          }
          if (!EU_MFOLD) {
            // the row's scale into the A block of the -m MMA (k = 0 meets -m_hi, k = 1 meets -m_lo)
            *reinterpret_cast<float2*>(ones + ab * 2048 + h * 1024 + (q * 32 + lane) * 4) = make_float2(sc, sc);
            fence_proxy_async();
          }
        } else {
#pragma unroll
          for (int c0 = 0; c0 < DP; c0 += 16) {
            float v[16];
            load16(c0, v);
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) split_tf32(v[j], hi[j], lo[j]);
            tmem_st16(a_hi + c0, hi);
            tmem_st16(a_lo + c0, lo);
          }
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&S->afull);
      return rs;
    };
    float rs_next = 1.f;
    if (AB2 && my_tiles > 0) rs_next = load_A(0);
    for (int t = 0; t < my_tiles; ++t) {
      const long long tile = (long long)blockIdx.x + (long long)t * gridDim.x;
      const long long row = tile * EU_TILE + rloc;
      const bool valid = row < a.N;
      const float rs = AB2 ? rs_next : load_A(t);
      float mx = -INFINITY, sm = 0.f;
      float l4[4];
      // the row whose logits this thread finishes, stores and normalises by: its own TMEM lane, or (EU_MFOLD == 1) the row
      // the quad reduce-scatter leaves it with; rs4 = scales of the four rows it holds fragments of
      constexpr bool QUAD = F16 && EU_MFOLD == 1;
      const int rloc_o = QUAD ? h * 128 + q * 32 + 16 * ((lane >> 1) & 1) + 8 * (lane & 1) + (lane >> 2) : rloc;
      const long long row_o = tile * EU_TILE + rloc_o;
      const bool valid_o = row_o < a.N;
      float rs4[4];
#pragma unroll
      for (int rho = 0; rho < 4; ++rho)
        rs4[rho] = QUAD ? __shfl_sync(0xffffffffu, rs, 16 * (rho >> 1) + 8 * (rho & 1) + (lane >> 2)) : rs;
      for (int g = 0; g < ngroups; ++g, rg.next(nstage), br.next()) {
        const int s = rg.s;
        const float* cstage = reinterpret_cast<const float*>(stages + (size_t)s * C::STAGE + C::GB + C::MB);
        if (AB2 && g == ngroups - 1 && t + 1 < my_tiles) rs_next = load_A(t + 1);   // under the tile's last bursts
        mbar_wait(&S->tfull[br.buf], br.ph);      // the MMAs read the stage, so its bulk copy (m, cst too) has landed
        tc_fence_after();
        const uint32_t dcol = tm + lane_base + DCOL0 + br.buf * EU_N;
        // one logit: online logsumexp with one exp per component (ex2.approx: rel. error 2^-22), four logits per store
        auto emit = [&](int c, float l, bool ok, long long orow) {
          if (MODE == 1) {
            const float e = exp2f(-1.44269504f * fabsf(l - mx));
            sm = (l > mx) ? fmaf(sm, e, 1.f) : sm + e;
            mx = fmaxf(mx, l);
          }
          l4[c & 3] = l;
          if ((c & 3) == 3 && ok)
            *reinterpret_cast<float4*>(a.out + (size_t)orow * ldo + (c - 3)) = make_float4(l4[0], l4[1], l4[2], l4[3]);
        };
        float cs[C::CG], cv[C::CG];
        if constexpr (F16 && EU_MFOLD == 1) {
          // ---- -m in the epilogue, accumulator read as m16n8-style fragments (tcgen05.ld.16x256b): a thread holds two
          // columns of every 8-column block for FOUR rows, so it needs only a quarter of the group's m values — 16
          // 8-byte loads whose four distinct addresses per warp make one shared-memory wavefront each (with one row
          // per thread every thread needs all 128 values: 4 bytes per wavefront, ~1000 wavefronts per group, more than
          // the MMAs' own operand fetch — measured 15.3 against 13.4 ms).  The four lanes of a quad then reduce-scatter
          // their partial sums (3 shuffles per component) and lane t ends up owning row 16 (t/2 % 2) + 8 (t % 2) + t / 4
          // of the warp's 32.
          float y[EU_N];
          tmem_ld_16x256b_x16(dcol, y);
          tmem_ld_16x256b_x16(dcol + (16u << 16), y + 64);
#pragma unroll
          for (int cl = 0; cl < C::CG; ++cl) { cv[cl] = cstage[cl]; cs[cl] = cstage[8 + cl]; }
          tmem_wait_ld();
          tc_fence_before();
          mbar_arrive(&S->tempty[br.buf]);
          const float2* mneg = reinterpret_cast<const float2*>(stages + (size_t)s * C::STAGE + C::GB) + (lane & 3);
          float2 qa[4][C::CG];
          constexpr int NF = C::CG <= 2 ? C::CG : 1;          // scale products kept in registers (few components per group)
          float f[4][NF];
#pragma unroll
          for (int rho = 0; rho < 4; ++rho) {
#pragma unroll
            for (int cl = 0; cl < C::CG; ++cl) qa[rho][cl] = make_float2(0.f, 0.f);
#pragma unroll
            for (int cl = 0; cl < NF; ++cl) f[rho][cl] = rs4[rho] * cs[cl];
          }
#pragma unroll
          for (int b = 0; b < 16; ++b) {                      // 8-column block b: component b % CG
            const int cl = b % C::CG;
            const float2 mm = mneg[4 * b];                    // -m of columns 8 b + 2 (lane % 4) + {0, 1}
#pragma unroll
            for (int rho = 0; rho < 4; ++rho) {               // row 16 (rho / 2) + 8 (rho % 2) + lane / 4
              const int i = 64 * (rho >> 1) + 4 * b + 2 * (rho & 1);
              // y = (W^T z) 2^(sh_n + t_k): one FFMA2 undoes the scaling (exact) and subtracts m, one squares and accumulates
              const float ff = C::CG <= 2 ? f[rho][cl % NF] : rs4[rho] * cs[cl];
              const float2 r = __ffma2_rn(make_float2(y[i], y[i + 1]), make_float2(ff, ff), mm);
              qa[rho][cl] = __ffma2_rn(r, r, qa[rho][cl]);
            }
          }
          const bool b0 = lane & 1, b1 = lane & 2;
#pragma unroll
          for (int cl = 0; cl < C::CG; ++cl) {
            const int c = g * C::CG + cl;
            if (c < K) {                                      // warp uniform
              const float v0 = qa[0][cl].x + qa[0][cl].y, v1 = qa[1][cl].x + qa[1][cl].y;
              const float v2 = qa[2][cl].x + qa[2][cl].y, v3 = qa[3][cl].x + qa[3][cl].y;
              float ka = (b0 ? v1 : v0) + __shfl_xor_sync(0xffffffffu, b0 ? v0 : v1, 1);
              float kb = (b0 ? v3 : v2) + __shfl_xor_sync(0xffffffffu, b0 ? v2 : v3, 1);
              const float qq = (b1 ? kb : ka) + __shfl_xor_sync(0xffffffffu, b1 ? ka : kb, 2);
              emit(c, cv[cl] - 0.5f * qq, valid_o, row_o);
            }
          }
        } else {
        // all 128 columns into registers, then hand the accumulator straight back to the MMA warp: the arithmetic
        // below overlaps the next MMAs
        float y[EU_N];
        tmem_ld32(dcol, y);
        tmem_ld32(dcol + 32, y + 32);
        tmem_ld32(dcol + 64, y + 64);
        tmem_ld32(dcol + 96, y + 96);
#pragma unroll
        for (int cl = 0; cl < C::CG; ++cl) { cv[cl] = cstage[cl]; cs[cl] = cstage[8 + cl]; }
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(&S->tempty[br.buf]);
        // packed fp32x2 accumulators (FFMA2), NQ independent chains per component
        constexpr int NQ = C::CG >= 4 ? 1 : 4 / C::CG;
        float2 qa[C::CG][NQ];
#pragma unroll
        for (int cl = 0; cl < C::CG; ++cl)
#pragma unroll
          for (int u = 0; u < NQ; ++u) qa[cl][u] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < EU_N; j += 2) {                   // y already holds W^T z - m (EU_MFOLD == 2: W^T z)
          const int cl = C::col_cl(j);                        // columns j, j+1 belong to the same component
          const int u = (j / (8 * C::CG)) % NQ;
          float2 r = make_float2(y[j], y[j + 1]);
          if (F16) { const float f = rs * cs[cl]; r = __fmul2_rn(r, make_float2(f, f)); }   // 2^-(sh_n + t_k), exact
          qa[cl][u] = __ffma2_rn(r, r, qa[cl][u]);
        }
#pragma unroll
        for (int cl = 0; cl < C::CG; ++cl) {
          const int c = g * C::CG + cl;
          if (c < K) {
            float2 qq = qa[cl][0];
#pragma unroll
            for (int u = 1; u < NQ; ++u) qq = __fadd2_rn(qq, qa[cl][u]);
            emit(c, cv[cl] - 0.5f * (qq.x + qq.y), valid, row);
          }
        }
        }
        mbar_arrive(&S->empty[s]);                  // done with the stage's cst (the two MMA commits are the other arrivals)
      }
      if (MODE == 1) {
        const float v = valid_o ? mx + logf(sm) : 0.f;
        if (t >= 2) mbar_wait(&S->nfree[t & 1], ((uint32_t)(t >> 1) & 1) ^ 1);    // tile t-2 has been normalised
        lz[(t & 1) * EU_TILE + rloc_o] = v;
        if (valid_o) { a.logZn[row_o] = v; lzsum += (double)v; }
        mbar_arrive(&S->ndone[t & 1]);              // release: this thread's logits and logZ_n of the tile
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
    }
  } else if (MODE == 1) {
    // ================= normaliser (mode 1): p = exp(l - logZ_n) of tile t, written while the MMAs of tile t+1 run ==========
    // One warp re-reads the tile's logits (L2 hits), normalises them in place with coalesced float4 accesses, 4 rows in
    // flight, and keeps the column sums NA of its fixed columns in registers (float per tile, double across tiles: a
    // fixed order, so NA is deterministic).  Inline in the epilogue warps this pass sat on their critical path.
    constexpr int NU = EU_MAXK / 128;                      // float4 columns per lane (K = 512)
    const int K4 = K >> 2;
    double* na = reinterpret_cast<double*>(lz + 2 * EU_TILE);       // [K] running column sums (each lane owns its columns)
    for (int k = lane; k < K; k += 32) na[k] = 0.0;
    __syncwarp();
    // 16 float4 loads in flight per lane: RB rows of NUA float4 columns each (NUA = columns this K needs)
    auto ex2 = [](float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; };
    // a.rpack != nullptr (host: K <= 256): also emit the responsibilities as the Gram kernel's fp16 A-operand images,
    // r 2^14 = a + b per 8-sample chunk and component (gram_umma.cu, GuArgs::rp) — every value is in a register here, and
    // a lane's four components are 64 contiguous bytes of the image, a warp's 2 KB.
    auto norm_tile = [&](auto nua_c, int t) {
      constexpr int NUA = decltype(nua_c)::value, RB = 16 / NUA;
      const long long row0 = ((long long)blockIdx.x + (long long)t * gridDim.x) * EU_TILE;
      const long long rem = a.N - row0;
      const int rows = rem < EU_TILE ? (int)rem : EU_TILE;
      const bool pack = NUA <= 2 && a.rpack != nullptr;
      const int rows_p = pack ? min(EU_TILE, (rows + 31) & ~31) : rows;      // the images cover whole 32-sample chunks
      const float* lzp = lz + (t & 1) * EU_TILE;
      // K <= 64 (K4 <= 16): a row is at most 16 float4 wide, so the warp splits into 32 / gsz lane groups and group j takes
      // the rows [r + j RB, r + (j + 1) RB) of every batch — with one group half of the lanes (or more) idled, and at
      // K = 64 the ONE normaliser warp, not the tensor pipe, set the kernel's pace
      const int gsz = (NUA == 1 && K4 <= 16) ? (K4 <= 1 ? 1 : K4 <= 2 ? 2 : K4 <= 4 ? 4 : K4 <= 8 ? 8 : 16) : 32;
      const int ngrp = 32 / gsz, lin = lane % gsz, roff = (lane / gsz) * RB;
      float4* base = reinterpret_cast<float4*>(a.out) + (size_t)row0 * K4 + lin;      // 32-bit offsets from here on
      bool cok[NUA];
#pragma unroll
      for (int u = 0; u < NUA; ++u) cok[u] = lin + 32 * u < K4;
      float4 cs[NUA];
#pragma unroll
      for (int u = 0; u < NUA; ++u) cs[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int rb = 0; rb < rows_p; rb += RB * ngrp) {
        const int r = rb + roff;                             // this lane group's rows of the batch
        float4 x[RB][NUA];
#pragma unroll
        for (int v = 0; v < RB; ++v) {
          const float4* pr = base + min(r + v, rows - 1) * K4;
#pragma unroll
          for (int u = 0; u < NUA; ++u) x[v][u] = cok[u] ? __ldcg(pr + 32 * u) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int v = 0; v < RB; ++v) {
          // exp(l - logZ_n) = 2^((l - logZ_n) log2 e) with ex2.approx (rel. error 2^-22): four instructions per element.
          // The full-range expf with its operand set-up came to ~50 per element, more than one warp can issue per tile.
          const bool rok = r + v < rows;
          const float lzr = lzp[min(r + v, rows - 1)];
          float4* pw = base + (r + v) * K4;
#pragma unroll
          for (int u = 0; u < NUA; ++u) {
            float4 y = x[v][u];
            y.x = ex2((y.x - lzr) * 1.44269504f); y.y = ex2((y.y - lzr) * 1.44269504f);
            y.z = ex2((y.z - lzr) * 1.44269504f); y.w = ex2((y.w - lzr) * 1.44269504f);
            if (rok && cok[u]) {
              __stcs(pw + 32 * u, y);          // streaming (evict-first): p is not re-read by this kernel, the logits of
                                               // the tiles still in flight should stay in L2 until they are overwritten
              cs[u].x += y.x; cs[u].y += y.y; cs[u].z += y.z; cs[u].w += y.w;
            }
            x[v][u] = rok ? y : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        if (pack && r < rows_p) {
          constexpr float RS = 16384.f;                                   // GU_RSH = 14
#pragma unroll
          for (int c = 0; c < RB / 8; ++c) {
            const long long ch = ((row0 + r) >> 3) + c;                   // global 8-sample chunk
#pragma unroll
            for (int u = 0; u < NUA; ++u) {
              if (cok[u]) {
                const int comp0 = 4 * (lin + 32 * u);
                uint8_t* rec = a.rpack + ((size_t)(comp0 >> 7) * nsb + (size_t)(ch >> 1)) * 8192 + (size_t)(ch & 1) * 2048 +
                               (size_t)(comp0 & 127) * 16;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  uint32_t hi[4], lo[4];
#pragma unroll
                  for (int w = 0; w < 4; ++w) {
                    const float4 f0 = x[8 * c + 2 * w][u], f1 = x[8 * c + 2 * w + 1][u];
                    const float x0 = (j == 0 ? f0.x : j == 1 ? f0.y : j == 2 ? f0.z : f0.w) * RS;
                    const float x1 = (j == 0 ? f1.x : j == 1 ? f1.y : j == 2 ? f1.z : f1.w) * RS;
                    const __half2 ah = __floats2half2_rn(x0, x1);
                    const float2 af = __half22float2(ah);
                    const __half2 bh = __floats2half2_rn(x0 - af.x, x1 - af.y);
                    hi[w] = *reinterpret_cast<const uint32_t*>(&ah);
                    lo[w] = *reinterpret_cast<const uint32_t*>(&bh);
                  }
                  __stcs(reinterpret_cast<uint4*>(rec + j * 16), make_uint4(hi[0], hi[1], hi[2], hi[3]));
                  __stcs(reinterpret_cast<uint4*>(rec + 4096 + j * 16), make_uint4(lo[0], lo[1], lo[2], lo[3]));
                }
              }
            }
          }
        }
      }
      // lane groups hold partial column sums of the same columns: add them in a fixed order, group 0 keeps the total
      for (int o = gsz; o < 32; o <<= 1) {
#pragma unroll
        for (int u = 0; u < NUA; ++u) {
          cs[u].x += __shfl_xor_sync(0xffffffffu, cs[u].x, o); cs[u].y += __shfl_xor_sync(0xffffffffu, cs[u].y, o);
          cs[u].z += __shfl_xor_sync(0xffffffffu, cs[u].z, o); cs[u].w += __shfl_xor_sync(0xffffffffu, cs[u].w, o);
        }
      }
#pragma unroll
      for (int u = 0; u < NUA; ++u) {
        const int c4 = lin + 32 * u;
        if (c4 < K4 && lane < gsz) {
          na[4 * c4] += (double)cs[u].x; na[4 * c4 + 1] += (double)cs[u].y;
          na[4 * c4 + 2] += (double)cs[u].z; na[4 * c4 + 3] += (double)cs[u].w;
        }
      }
    };
    for (int t = 0; t < my_tiles; ++t) {
      mbar_wait(&S->ndone[t & 1], (uint32_t)(t >> 1) & 1);
      if (K4 <= 32) norm_tile(std::integral_constant<int, 1>{}, t);
      else if (K4 <= 64) norm_tile(std::integral_constant<int, 2>{}, t);
      else norm_tile(std::integral_constant<int, 4>{}, t);
      __syncwarp();
      if (lane == 0) mbar_arrive(&S->nfree[t & 1]);
    }
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int c4 = lane + 32 * u;
      if (c4 < K4)
        reinterpret_cast<float4*>(a.NA_part + (size_t)blockIdx.x * K)[c4] =
            make_float4((float)na[4 * c4], (float)na[4 * c4 + 1], (float)na[4 * c4 + 2], (float)na[4 * c4 + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tm);
}

// ---- host side -------------------------------------------------------------------------------------------
static int eu_num_sms() { return num_sms(); }

// the TF32 record is the larger one: the workspace is sized for it whichever precision runs
static size_t eu_group_bytes(int Dp) {
  return Dp == 128 ? EuCfg<128, true>::REC : Dp == 64 ? EuCfg<64, false>::REC : (Dp == 32 ? EuCfg<32, false>::REC : EuCfg<16, false>::REC);
}
static bool eu_use_f16() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VBMP_ESTEP_PREC");
    v = (e && (e[0] == 't' || e[0] == 'T')) ? 0 : 1;
  }
  return v == 1;
}
static int eu_cg(int Dp) { return EU_N / Dp; }
static size_t eu_align(size_t x) { return (x + 255) / 256 * 256; }

static bool eu_use_f16();
bool estep_umma_supported(long long N, int GX, int G, int K, int Dp, int d0, int d1) {
  (void)d0; (void)d1;
  // Dp = 128 exists with fp16 operands only (the TF32 images of two 128-row halves alone would fill tensor memory)
  return G == 1 && GX == 1 && (Dp == 16 || Dp == 32 || Dp == 64 || (Dp == 128 && eu_use_f16())) && (K % 4 == 0) &&
         K <= EU_MAXK && N >= 256;
}

bool gram_rpack_usable();
// K2 can hand K3 the pre-split weights when both run their fp16 tensor-core variants and one normaliser batch covers
// whole 8-sample chunks of every component (K <= 256)
bool estep_umma_can_pack(long long N, int GX, int G, int K, int Dp, int d0, int d1, int mode) {
  return mode == 1 && K <= 256 && eu_use_f16() && gram_rpack_usable() && estep_umma_supported(N, GX, G, K, Dp, d0, d1);
}

size_t estep_umma_workspace_bytes(long long N, int G, int K, int Dp, int mode) {
  (void)mode;
  if (!estep_umma_supported(N, 1, G, K, Dp, 1, 0)) return 0;
  const int ngroups = (K + eu_cg(Dp) - 1) / eu_cg(Dp);
  const size_t ctas = 512;   // upper bound on the persistent grid
  return eu_align((size_t)ngroups * eu_group_bytes(Dp)) + 2 * EU_FS * sizeof(float) + eu_align(ctas * K * sizeof(float)) + eu_align(ctas * sizeof(double)) + 256;
}

int launch_estep_reduce(const float*, const double*, int nb, int G, int K, float* NA, float* logZ, cudaStream_t);

template <int DP, bool F16>
static int eu_launch(EstepArgs a, int mode, uint8_t* Wp, float* fs, float* NA_part, double* logZ_part, float* NA, float* logZ,
                     cudaStream_t st) {
  using C = EuCfg<DP, F16>;
  const int ngroups = (a.K + C::CG - 1) / C::CG;
  if (F16) {
    estep_rowscale_kernel<DP><<<DP, 256, 0, st>>>(a.W, a.K, fs);
    int rc0 = check_launch("estep_rowscale");
    if (rc0) return rc0;
  }
  estep_pack_kernel<DP, F16><<<ngroups, 256, 0, st>>>(a.W, a.m, a.cst, a.K, fs, Wp);
  int rc = check_launch("estep_pack");
  if (rc) return rc;
  const int ntiles = (int)((a.N + EU_TILE - 1) / EU_TILE);
  const int grid = ntiles < eu_num_sms() ? ntiles : eu_num_sms();
  const long long nsb = (a.N + 31) / 32 * 2;               // 16-sample blocks per component block of the weight images
  if (!F16) a.rpack = nullptr;
  // two sets of A images (AB2) where tensor memory has room for them beside three accumulators: up to DP = 32 (cfg4's
  // E-step 0.72 -> 0.70 ms).  At DP = 64 the second set costs the third accumulator buffer, which loses more than the
  // overlap gains (cfg3, K = 64: 7.05 -> 7.21 ms; cfg2: 13.5 -> 13.8 ms): off unless VBMP_ESTEP_AB2 = 1 (tuning; 0 = never)
  static const int ab2_env = [] { const char* e = getenv("VBMP_ESTEP_AB2"); return e ? atoi(e) : -1; }();
  constexpr bool AB2_OK = F16 && DP <= 64;
  bool ab2 = AB2_OK && ngroups >= 2 && DP <= 32;
  if (ab2_env == 0) ab2 = false;
  if (ab2_env == 1) ab2 = AB2_OK && ngroups >= 2;
  const size_t fixed = (ab2 ? 16384 : 8192) + sizeof(EuSmem) + 2 * EU_TILE * sizeof(float) + (mode == 1 ? (size_t)a.K * sizeof(double) : 0) + 64;
  int nstage = (int)((227 * 1024 - fixed) / C::STAGE);
  if (nstage > EU_MAXSTAGE) nstage = EU_MAXSTAGE;
  if (nstage < 2) { set_error("estep_umma: shared memory too small for K=%d", a.K); return VBMP_ERR_UNSUPPORTED; }
  const size_t smem = (size_t)nstage * C::STAGE + fixed;
  a.NA_part = NA_part;
  a.logZ_part = logZ_part;
#define EU_LAUNCH(MODE_, AB_)                                                                                              \
  do {                                                                                                                    \
    cudaFuncSetAttribute(estep_umma_kernel<DP, MODE_, F16, AB_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    estep_umma_kernel<DP, MODE_, F16, AB_><<<grid, EU_THREADS, smem, st>>>(a, Wp, fs, ntiles, ngroups, nstage, nsb);      \
  } while (0)
  if constexpr (AB2_OK) {
    if (ab2) { if (mode == 0) EU_LAUNCH(0, true); else EU_LAUNCH(1, true); }
    else { if (mode == 0) EU_LAUNCH(0, false); else EU_LAUNCH(1, false); }
  } else {
    if (mode == 0) EU_LAUNCH(0, false); else EU_LAUNCH(1, false);
  }
#undef EU_LAUNCH
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
  float* fs = (float*)p;                                    // feature scales of the fp16 operands: [2^e_i | 2^-e_i]
  p += 2 * EU_FS * sizeof(float);
  float* NA_part = (float*)p;
  p += eu_align((size_t)512 * a.K * sizeof(float));
  double* logZ_part = (double*)p;
  if (eu_use_f16()) {
    switch (a.Dp) {
      case 128: return eu_launch<128, true>(a, mode, Wp, fs, NA_part, logZ_part, NA, logZ, st);
      case 64: return eu_launch<64, true>(a, mode, Wp, fs, NA_part, logZ_part, NA, logZ, st);
      case 32: return eu_launch<32, true>(a, mode, Wp, fs, NA_part, logZ_part, NA, logZ, st);
      case 16: return eu_launch<16, true>(a, mode, Wp, fs, NA_part, logZ_part, NA, logZ, st);
    }
  } else {
    switch (a.Dp) {
      case 64: return eu_launch<64, false>(a, mode, Wp, fs, NA_part, logZ_part, NA, logZ, st);
      case 32: return eu_launch<32, false>(a, mode, Wp, fs, NA_part, logZ_part, NA, logZ, st);
      case 16: return eu_launch<16, false>(a, mode, Wp, fs, NA_part, logZ_part, NA, logZ, st);
    }
  }
  set_error("estep_umma: Dp=%d not supported", a.Dp);
  return VBMP_ERR_UNSUPPORTED;
}

}  // namespace vbmp
