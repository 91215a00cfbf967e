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
//   * CTA task = (128-component block) x (NPB <= 224 pair columns) x (sample split); with an even number of component
//     blocks two CTAs pair up (cta_group::2, M = 256, NPB <= 192 shared between them; the last pair block is only as
//     wide as the pairs it holds).  A operand = R^T (lanes = components, K = samples) in TENSOR MEMORY as two
//     split-precision images; B operand = phi^T generated on the fly in shared memory (K-major core-matrix layout,
//     two images).  D1 (first-level accumulator) and D2 (second level) both live in TMEM: D1 is folded into D2 (fp32
//     round-to-nearest adds on the CUDA cores) every 16 chunks, at zero memory traffic.  The tensor core TRUNCATES
//     on every fp32 accumulate: measured bias -1.0e-7 of the running sum per 16-sample TF32 chunk
//     (tools/gram_bias.py), i.e. -1.6e-6 at 16 chunks per block, -1.3e-5 at 128; hence the short blocks
//     (SURVEY.md Appendix F.2 anticipated this).
//   * warp 0 brings the chunk operands into a 6-deep ring (TF32: TMA tiled loads of the raw R / Z rows, zero-filled past
//     the last row; fp16: two bulk copies of the pre-split weight images and the transposed, pre-scaled samples),
//     warp 1 issues the MMAs (3 split-precision terms x 2 K-steps per chunk), two sets of 8 worker warps alternate
//     chunks: copy / split the A images into TMEM, multiply and split phi into the B stage.
//   * per-split partials are reduced in a fixed order in fp64 by gram_pair_reduce_kernel (deterministic).
//
// Operand precision (template parameter F16; the default, VBMP_GRAM_PREC=tf32 selects the TF32 split only):
//   * FP16 split: r' = r 2^14 and zt'_i = zt_i 2^u_i (u_i puts the column maximum of |z_i| in [2^6, 2^7), found by a
//     column-maximum pre-pass over Z) are EXACT rescalings; r' = a + b and phi' = zt'_i zt'_j = A + B with a, b, A, B fp16
//     (22 significant bits, at least as accurate as the TF32 split) and the three terms run as kind::f16 MMAs with K = 16:
//     twice the TF32 rate and half the operand bytes through shared memory, which is the kernel's limiter.  Chunks are
//     32 samples, so a stage has the same footprint as a 16-sample TF32 stage.  The reduce kernel multiplies by
//     2^-(14 + u_i + u_j) (exact).  Weights outside fp16 range after scaling (|r| > 3.99; responsibilities are <= 1)
//     raise a device flag; the TF32 kernel, launched right behind, returns at once unless the flag is set, in which case
//     it recomputes the partials — no host synchronisation, no silent loss of accuracy.
#include <algorithm>
#include <cstdlib>
#include <cuda_fp16.h>
#include "common.cuh"
#include "umma.cuh"

namespace vbmp {
using namespace umma;

constexpr int GU_THREADS = 576;      // warp 0 producer, warp 1 MMA issuer, two sets of 8 worker warps (even / odd chunks)
__host__ __device__ constexpr int gu_sc(bool f16) { return f16 ? 32 : 16; }
constexpr int GU_RSH = 14;           // fp16 variant: responsibilities are scaled by 2^14
constexpr float GU_RMAX = 3.99f;     // ... and must stay below 65504 / 2^14
#ifndef GU_NR_V
#define GU_NR_V 6
#endif
constexpr int GU_NR = GU_NR_V;       // raw ring depth (the launcher takes fewer slots when the chunk records of a wide
                                     // feature vector, D > 64, would not fit shared memory: GuArgs::nr)
// pipeline depth (B stages in shared memory = A buffers in TMEM) and pair columns per MMA.  TMEM budget: D1 + D2 = 2 NPMAX
// columns + NSTG A buffers of hi + lo = 32 NSTG columns <= 512.  A CTA pair needs 4 stages to hide the cross-CTA barrier
// round trips (192 columns); a single CTA is best with 2 stages and 224 columns.
// (measured at cfg2, whole call: 4 stages x 192 columns 10.5 ms, 5 x 176 10.8 ms, 6 x 160 11.2 ms)
#ifndef GU_PAIR_NSTG
#define GU_PAIR_NSTG 4
#define GU_PAIR_NPMAX 192
#endif
#ifndef GU_ACP
#define GU_ACP 1                     // fp16 variant: the weight images go from the raw ring to tensor memory with tcgen05.cp
#endif                               // (issued by the MMA warp) instead of through the workers' registers
constexpr int GU_MAXSTG = 6;
__host__ __device__ constexpr int gu_nstg(bool pair) { return pair ? GU_PAIR_NSTG : 2; }
__host__ __device__ constexpr int gu_npmax(bool pair) { return pair ? GU_PAIR_NPMAX : 224; }
static_assert(2 * GU_PAIR_NPMAX + 32 * GU_PAIR_NSTG <= 512, "TMEM budget of a CTA pair");
constexpr int GU_FL = 16;            // chunks per first-level accumulation block (256 samples)
constexpr int GU_CB = 128;           // components per CTA

struct GuArgs {
  int d0, d1;
  long long N; int K;
  int npb, NPB, P, ncb, splits;      // pair blocks, pair columns of the widest block, total pairs, component blocks, sample splits
  int nr;                            // raw ring slots in use, 2 <= nr <= GU_NR
  int diag;                          // 1: only the pairs (i, i) and (i, D) — the statistics of the diagonal-precision nodes
  int lin;                           // 1: "pair" p is column p of z0 itself (phi = z0[n][p]): weighted column sums, see launch_wsum_umma
  int zw;                            // TF32 variant: floats per row of the z0 box in the raw ring (d0; lin: the widest column block)
  int wbase, wextra;                 // block pb holds 16 (wbase + (pb < wextra)) pair columns starting at 16 (pb wbase + min(pb, wextra))
  long long S_per;                   // samples per split (multiple of GU_SC)
  int Kp, PP;                        // padded partial dims: Kp = ncb*128, PP = P rounded up to 16
  int kcb;                           // columns of the R box = min(128, K)
  int FL;                            // chunks per first-level accumulation block
  float* part;                       // [splits][Kp][PP]
  // fp16 variant: cmax[0..D) = bit patterns of the column maxima of |z_i| (head of the sample image, see vbmp_gram_zpack);
  // flag[0] != 0 = "operands outside the fp16 window, recompute with TF32 operands" (weights beyond range: set by
  // gram_rsplit_kernel; a component whose samples sit below the resolution of a feature's scale: set by the first reduce).
  // The TF32 kernel gets flag too (nullptr = run always).
  const uint32_t* cmax;
  uint32_t* flag;
  // fp16 variant: the weights pre-split by gram_rsplit_kernel: record (component block cb, 16-sample block sb) at
  // ((cb * nsb + sb) * 8192) bytes = [hi | lo] x [8-sample chunk (2)][component (128)][8 fp16], i.e. the TMEM image of
  // the A operand of one K-step: lane = component, 16 bytes per chunk
  const uint8_t* rp;
  long long nsb;                     // 16-sample blocks per component block (N rounded up to 32, / 16)
  // fp16 variant: zt = [z;1] scaled by 2^u_i and TRANSPOSED per 32-sample chunk by gram_zprep_kernel: record of zrec bytes
  // = [feature 0..D-1 | constant 2^6 | zeros][GU_ZS floats], so a pair thread reads its two factors for 16 samples with
  // eight 16-byte loads (row stride 36 floats: conflict free) instead of 32 scalar ones
  const uint8_t* zt;
  int zrec;
};
constexpr int GU_HDR_WORDS = 256;    // words of the sample image's header (column maxima) and of the per-call flag block
constexpr int GU_RREC = 8192;
constexpr int GU_ZS = 36;            // floats per feature row of a transposed 32-sample chunk (32 + 4 padding)
__host__ __device__ constexpr int gu_zrec(int D) { return (D + 2) * GU_ZS * 4; }

// exponent u_i of the exact feature scale 2^u_i (column maximum -> [2^6, 2^7)); 0 for an all-zero column
__host__ __device__ inline int gu_feat_exp(uint32_t maxbits) {
  const int e = (int)((maxbits >> 23) & 0xff);
  if (e == 0) return 0;
  int u = 6 - (e - 127);
  return u > 60 ? 60 : (u < -60 ? -60 : u);
}

struct GuSmem {
  uint64_t rfull[GU_NR], rempty[GU_NR];
  uint64_t bfull[GU_MAXSTG], bempty[GU_MAXSTG];
  uint64_t dfull, dempty;
  uint32_t tmem_base;
  float consts[2];                   // {1, 0}: the padded "1" feature and the zero used by padding pair columns
};

// Pair p of the symmetric (D+1) x (D+1) Gram matrix over zt = [z;1]:  p < D(D+1)/2 walks the upper triangle of the
// D x D block row by row; the last D+1 pairs are (i, D), i = 0..D (the SEx column and N).  Keeping the pairs that
// involve the constant feature at the end lets every other warp read both factors with one compile-time stride.
__host__ __device__ inline void gu_pair(int p, int D, int* i, int* j) {
  const int T = D * (D + 1) / 2;
  if (p >= T) { *i = p - T; *j = D; return; }
  int rem = p, r = 0, len = D;
  while (rem >= len) { rem -= len; ++r; --len; }
  *i = r; *j = r + rem;
}

// diag = 1 (NormalGamma.raw_update, dists/NormalGamma.py:58-73: SExx_i = sum r x_i^2, SEx_i = sum r x_i, N = sum r): the
// 2 D + 1 pairs (i, i), i < D, then (i, D), i <= D — the same kernel on 3 % of the pair columns, HBM-bound.
__host__ __device__ inline void gu_pair_any(int p, int D, int diag, int* i, int* j) {
  if (!diag) { gu_pair(p, D, i, j); return; }
  if (p < D) { *i = p; *j = p; } else { *i = p - D; *j = D; }
}
__host__ __device__ inline int gu_npairs(int D, int diag) { return diag ? 2 * D + 1 : (D + 1) * (D + 2) / 2; }

// Split x = hi + lo for the 3-term TF32 product: hi = TF32(x) rounded to nearest; lo = x - hi is exact in fp32 and
// needs no rounding of its own because the tensor core ignores the low 13 mantissa bits of a TF32 operand
// (tools/umma_probe.cu: "operand low-bit handling: truncated").  Feeding x itself as hi (read as trunc(x)) would
// save two integer ops but adds ~30 % to the bias and triples its spread (tools/gram_bias.py).
__device__ __forceinline__ void split_fast(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
// two values at a time: the fp32 subtraction runs as one packed FADD2
__device__ __forceinline__ void split_fast2(float2 x, uint32_t& hi0, uint32_t& hi1, uint32_t& lo0, uint32_t& lo1) {
  hi0 = (__float_as_uint(x.x) + 0x1000u) & 0xffffe000u;
  hi1 = (__float_as_uint(x.y) + 0x1000u) & 0xffffe000u;
  const float2 l = __fadd2_rn(x, make_float2(-__uint_as_float(hi0), -__uint_as_float(hi1)));
  lo0 = __float_as_uint(l.x); lo1 = __float_as_uint(l.y);
}

// SF = compile-time row stride (floats) of the raw Z chunk when d0 == SF and (d1 == 0 or d1 == d0), else 0 (generic);
// R128 = the R box is 128 columns wide (K >= 128).
// PAIR: two CTAs (a cluster) take the two component blocks of a 256-component slab and share the phi block through
// cta_group::2 MMAs (M = 256): each CTA generates and holds HALF of the pair columns, the hardware feeds both tensor
// cores from both halves, so phi generation and the B-operand reads per SM halve (the shared-memory pipe is the limiter).
template <int SF, bool R128, bool PAIR, bool F16>
__global__ void __launch_bounds__(GU_THREADS, 1)
gram_umma_kernel(const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmZ0,
                 const __grid_constant__ CUtensorMap tmZ1, GuArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int SC = gu_sc(F16);                          // samples per chunk
  // TF32 variant behind an fp16 launch: only needed when that launch met weights outside fp16 range (uniform exit, before
  // any barrier or allocation, for both CTAs of a pair)
  if (!F16 && a.flag != nullptr && a.flag[0] == 0u) return;
  const int D = a.d0 + a.d1;
  // carve: raw ring [NR][R 16 x kcb floats | Z0 16 x d0 | Z1 16 x d1], B stages [2][hi NPB*64 B | lo NPB*64 B]
  const int kcb = R128 ? 128 : a.kcb;
  const int rawR = F16 ? 2 * GU_RREC : SC * kcb * 4, rawZ0 = F16 ? a.zrec : SC * a.zw * 4, rawZ1 = F16 ? 0 : SC * a.d1 * 4;
  const int rawB = (rawR + rawZ0 + rawZ1 + 127) / 128 * 128;
  constexpr int GU_NSTG = gu_nstg(PAIR), GU_NPMAX = gu_npmax(PAIR);
  const int stageB = 2 * (PAIR ? a.NPB / 2 : a.NPB) * 64;    // sized for a full-width pair block
  uint8_t* raw = smem_raw;
  const int nr = a.nr;
  uint8_t* bst = raw + nr * rawB;
  GuSmem* S = reinterpret_cast<GuSmem*>(bst + GU_NSTG * stageB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // task decode: cluster (or CTA) index = (split * ncbp + cbp) * npb + pb; a pair covers component blocks 2 cbp, 2 cbp + 1
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int task = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int ncbp = PAIR ? a.ncb / 2 : a.ncb;
  const int pb = task % a.npb;
  const int cb = PAIR ? 2 * ((task / a.npb) % ncbp) + (int)rank : (task / a.npb) % ncbp;
  const int split = task / (a.npb * ncbp);
  const long long nb = (long long)split * a.S_per;
  long long ne = nb + a.S_per; if (ne > a.N) ne = a.N;
  const int nchunks = ne > nb ? (int)((ne - nb + SC - 1) / SC) : 0;
  // the pair columns (P rounded up to 16) are dealt out in 16-column units as evenly as they go (cfg2: 3 blocks of 192
  // and 9 of 176 = 2160 columns for 2145 pairs): the blocks of one sample range then run at nearly the same rate and
  // stay within L2 reach of each other — one narrow last block (11 x 192 + 48) finished its rows four times faster
  // and its re-reads of the weight images all missed (+4 GB of DRAM reads per launch at cfg2)
  const int NPB = 16 * (a.wbase + (pb < a.wextra ? 1 : 0)), FL = a.FL;
  const int poff = 16 * (pb * a.wbase + min(pb, a.wextra));
  const int NH = PAIR ? NPB / 2 : NPB;                     // pair columns generated (and held as B rows) by this CTA

  if (tid == 0) {
    // one arrival per worker warp; in a pair the leader's bfull / dempty collect both CTAs' warps
    // rempty: the 8 worker warps of the consuming set (+ the commit behind the tcgen05.cp of the weight images)
    for (int s = 0; s < GU_NR; ++s) { mbar_init(&S->rfull[s], 1); mbar_init(&S->rempty[s], (F16 && GU_ACP) ? 9 : 8); }
    for (int s = 0; s < GU_NSTG; ++s) { mbar_init(&S->bfull[s], PAIR ? 16 : 8); mbar_init(&S->bempty[s], 1); }
    mbar_init(&S->dfull, 1); mbar_init(&S->dempty, PAIR ? 32 : 16);
    S->consts[0] = 1.f; S->consts[1] = 0.f;
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmR); tma_prefetch_desc(&tmZ0); if (a.d1 > 0) tma_prefetch_desc(&tmZ1); }
  if (warp == 1) { if (PAIR) tmem_alloc2<512>(&S->tmem_base); else tmem_alloc<512>(&S->tmem_base); }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tm = S->tmem_base;
  // TMEM columns: D1 [0,NPMAX), D2 [NPMAX,2 NPMAX), A buffers at 2 NPMAX + 32*buf (hi 16 columns, lo 16 columns)
  constexpr int ACOL = 2 * GU_NPMAX;

  if (warp == 0) {
    // ================= producer: one TMA box per operand per 16-sample chunk (rows past N are zero filled) ====
    const uint32_t bytes = (uint32_t)(rawR + rawZ0 + rawZ1);
    int s = 0;
    uint32_t rph = 0;                                     // slot and lap parity of the raw ring (no division in the loop)
    for (int c = 0; c < nchunks; ++c, s = (s + 1 == nr ? 0 : s + 1), rph ^= (s == 0 ? 1u : 0u)) {
      mbar_wait(&S->rempty[s], rph ^ 1);
      if (elect_one()) {
        const int r0 = (int)(nb + (long long)c * SC);
        uint8_t* dst = raw + (size_t)s * rawB;
        mbar_arrive_expect_tx(&S->rfull[s], bytes);
        if (F16) {
          bulk_g2s(dst, a.rp + ((size_t)cb * a.nsb + (size_t)(r0 >> 4)) * GU_RREC, 2 * GU_RREC, &S->rfull[s]);
          bulk_g2s(dst + rawR, a.zt + (size_t)(r0 >> 5) * a.zrec, (uint32_t)a.zrec, &S->rfull[s]);
        } else {
          tma_load_2d(dst, &tmR, cb * GU_CB, r0, &S->rfull[s]);
          tma_load_2d(dst + rawR, &tmZ0, a.lin ? poff : 0, r0, &S->rfull[s]);    // lin: only this block's columns
          if (a.d1 > 0) tma_load_2d(dst + rawR + rawZ0, &tmZ1, 0, r0, &S->rfull[s]);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer (the leader CTA's in a pair) =================
    if (!PAIR || rank == 0) {
      const uint32_t idesc = F16 ? idesc_f16(PAIR ? 256 : 128, NPB) : idesc_tf32(PAIR ? 256 : 128, NPB);
      const uint64_t dstep = (uint64_t)((2 * NH * 16) >> 4);                 // one K-step = two 16-byte chunks
      const uint64_t d_hi0 = smem_desc(smem_u32(bst), NH * 16, 128), d_lo0 = d_hi0 + (uint64_t)((NH * 64) >> 4);
      int fc = 0, nflush = 0;
      int rs = 0;                                            // raw ring slot of chunk c
      for (int c = 0; c < nchunks; ++c, rs = (rs + 1 == nr ? 0 : rs + 1)) {
        const int st = c % GU_NSTG;
        mbar_wait(&S->bfull[st], (c / GU_NSTG) & 1);
        const bool first = (fc == 0);
        if (first && c > 0) mbar_wait(&S->dempty, (nflush - 1) & 1);
        tc_fence_after();
        __syncwarp();
        const bool flush = (++fc == FL) || (c == nchunks - 1);
        if (elect_one()) {
          const uint64_t sofs = (uint64_t)((st * stageB) >> 4);
          const uint32_t a_hi = tm + ACOL + st * 32, a_lo = a_hi + 16;
          if (F16 && GU_ACP) {
            // A operand: the chunk's pre-split weight images are K-major core matrices as they lie in the raw ring
            // ([hi | lo] x [16-byte K-chunk (2)][component (128)], two 16-sample records): four 128-lane x 8-column copies,
            // in order ahead of the MMAs that read them (and behind the MMAs that read this stage four chunks ago)
            const uint32_t rsa = smem_u32(raw + (size_t)rs * rawB);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t dh = smem_desc(rsa + ks * GU_RREC, 2048, 128), dl = smem_desc(rsa + ks * GU_RREC + 4096, 2048, 128);
              if (PAIR) { tmem_cp2_128x256b(a_hi + ks * 8, dh); tmem_cp2_128x256b(a_lo + ks * 8, dl); }
              else { tmem_cp_128x256b(a_hi + ks * 8, dh); tmem_cp_128x256b(a_lo + ks * 8, dl); }
            }
            // the copies have read the raw slot (this arrives with them, one chunk ahead of the chunk's own MMAs)
            if (PAIR) mma2_commit(&S->rempty[rs], 3); else mma_commit(&S->rempty[rs]);
          }
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t b_hi = d_hi0 + sofs + ks * dstep, b_lo = d_lo0 + sofs + ks * dstep;
            if (PAIR && F16) {
              mma2_f16_ts(tm, a_lo + ks * 8, b_hi, idesc, !(first && ks == 0));
              mma2_f16_ts(tm, a_hi + ks * 8, b_lo, idesc, 1);
              mma2_f16_ts(tm, a_hi + ks * 8, b_hi, idesc, 1);
            } else if (F16) {
              mma_f16_ts(tm, a_lo + ks * 8, b_hi, idesc, !(first && ks == 0));
              mma_f16_ts(tm, a_hi + ks * 8, b_lo, idesc, 1);
              mma_f16_ts(tm, a_hi + ks * 8, b_hi, idesc, 1);
            } else if (PAIR) {
              mma2_tf32_ts(tm, a_lo + ks * 8, b_hi, idesc, !(first && ks == 0));
              mma2_tf32_ts(tm, a_hi + ks * 8, b_lo, idesc, 1);
              mma2_tf32_ts(tm, a_hi + ks * 8, b_hi, idesc, 1);
            } else {
              mma_tf32_ts(tm, a_lo + ks * 8, b_hi, idesc, !(first && ks == 0));
              mma_tf32_ts(tm, a_hi + ks * 8, b_lo, idesc, 1);
              mma_tf32_ts(tm, a_hi + ks * 8, b_hi, idesc, 1);
            }
          }
          if (PAIR) { mma2_commit(&S->bempty[st], 3); if (flush) mma2_commit(&S->dfull, 3); }
          else { mma_commit(&S->bempty[st]); if (flush) mma_commit(&S->dfull); }

        }
        __syncwarp();
        if (flush) { fc = 0; ++nflush; }
      }
    }
  } else {
    // ================= workers =================
    // two worker sets alternate chunks (set = chunk parity = B stage / A buffer), so the per-chunk latency chain
    // (barrier wait -> loads -> split -> stores -> fences -> arrive) of one set hides behind the other's
    const int set = (warp - 2) >> 3;
    const int w8 = (warp - 2) & 7, q = warp & 3, sh = w8 >> 2;
    const int wtid = (tid - 64) & 255;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int comp = q * 32 + lane;                         // component (TMEM lane) this thread feeds
    const bool comp_ok = comp < kcb && cb * GU_CB + comp < a.K;
    // pair owned by this thread in the phi generation; factors are read through (base, per-slot stride, per-sample
    // stride) triples so the inner loop is branch free.  Warps whose 32 pairs all avoid the constant feature and the
    // padding use the compile-time stride SF instead.
    const uint8_t* bi; const uint8_t* bj; int sli, slj, sti, stj;
    bool plain;
    // fp16: a chunk has two 16-sample K-steps; when 2 NH <= 256 the threads split them (thread = pair x K-step), else
    // every pair thread does both
    const bool split_ks = F16 && 2 * NH <= 256;
    const int pslot = split_ks ? wtid % NH : wtid;            // pair column (B row) generated by this thread
    const int ks0 = split_ks ? wtid / NH : 0, ksn = F16 ? (split_ks ? 1 : 2) : 1;
    const bool gen = split_ks ? wtid < 2 * NH : wtid < NH;
    {
      const int pg_ = poff + (int)rank * NH + pslot;
      const bool pair_ok = gen && (pg_ < a.P);
      int pi = 0, pj = 0;
      if (pair_ok && !a.lin) gu_pair_any(pg_, D, a.diag, &pi, &pj);
      auto setup = [&](int f, const uint8_t*& b, int& sl, int& stv) {
        if (!pair_ok) { b = reinterpret_cast<const uint8_t*>(&S->consts[1]); sl = 0; stv = 0; }
        else if (f < a.d0) { b = raw + rawR + f * 4; sl = rawB; stv = a.d0 * 4; }
        else if (f < D) { b = raw + rawR + rawZ0 + (f - a.d0) * 4; sl = rawB; stv = a.d1 * 4; }
        else { b = reinterpret_cast<const uint8_t*>(&S->consts[0]); sl = 0; stv = 0; }
      };
      setup(pi, bi, sli, sti);
      setup(pj, bj, slj, stj);
      if (!F16 && a.lin) {       // weighted column sums: phi = z0[n][column] * 1, the box holds this block's columns only
        if (pair_ok) { bi = raw + rawR + ((int)rank * NH + pslot) * 4; sli = rawB; sti = a.zw * 4; }
        bj = reinterpret_cast<const uint8_t*>(&S->consts[pair_ok ? 0 : 1]); slj = 0; stj = 0;
      }
      plain = SF > 0 && __all_sync(0xffffffffu, pair_ok && pj < D);
      if (F16) {       // transposed, pre-scaled chunk: row f of the record; padding pairs read the zero row D + 1
        bi = raw + rawR + (pair_ok ? pi : D + 1) * (GU_ZS * 4); sli = rawB;
        bj = raw + rawR + (pair_ok ? pj : D + 1) * (GU_ZS * 4); slj = rawB;
      }
    }
    const int fls = __ffs(FL) - 1;                       // FL is a power of two
    // barriers the workers signal: the leader CTA's (rank 0) in a pair, this CTA's own otherwise
    uint32_t bfull_addr[GU_NSTG];
#pragma unroll
    for (int i = 0; i < GU_NSTG; ++i) bfull_addr[i] = PAIR ? mapa_u32(&S->bfull[i], 0) : smem_u32(&S->bfull[i]);
    const uint32_t dempty_addr = PAIR ? mapa_u32(&S->dempty, 0) : smem_u32(&S->dempty);
    // ---- fold D1 into D2 (or, for the last block, write D1 + D2 to this split's partial).  BOTH worker sets fold every
    // block, a quarter of the columns per warp, and each set first generates its next chunk: the fold can only start when
    // the block's last MMAs have completed, one chunk time after its last operands were published, so that chunk is
    // free — and when the fold ends the first two chunks of the next block are already waiting for the MMA warp.
    // (One set folding all columns right behind its last chunk left the tensor pipe idle for the fold plus the
    // generation of the chunk after it: ~11 % of the kernel.)  Every set observes every phase of dfull in order, and
    // phase b + 1 cannot complete before all warps arrived on dempty for block b, so the parity wait is unambiguous.
    const int nblocks = (nchunks + FL - 1) >> fls;
    auto fold = [&](int b) {
      mbar_wait(&S->dfull, (uint32_t)b & 1u);
      tc_fence_after();
      const bool firstf = (b == 0), last = (b == nblocks - 1);
      constexpr int QW = (GU_NPMAX / 4 + 15) / 16 * 16;
      const int cbeg = (2 * sh + set) * QW, cend = min(cbeg + QW, NPB);
      float* prow = a.part + ((size_t)split * a.Kp + (size_t)cb * GU_CB + comp) * a.PP + (size_t)poff;
      for (int c0 = cbeg; c0 < cend; c0 += 16) {
        float v1[16], v2[16];
        tmem_ld16(tm + lane_base + c0, v1);
        if (!firstf) tmem_ld16(tm + lane_base + GU_NPMAX + c0, v2);
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
          tmem_st16(tm + lane_base + GU_NPMAX + c0, reinterpret_cast<const uint32_t*>(v1));
        }
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(dempty_addr);
    };
    int nf = 0;                                            // next block this set folds
    int s = set;                                           // raw ring position of chunk c (nr >= 2): slot, lap parity
    uint32_t rph = 0;
    for (int c = set; c < nchunks; c += 2, s += 2) {
      if (s >= nr) { s -= nr; rph ^= 1u; }
      const int st = c % GU_NSTG;
      mbar_wait(&S->rfull[s], rph);
      mbar_wait(&S->bempty[st], ((c / GU_NSTG) & 1) ^ 1);
      tc_fence_after();
      // ---- A operand: r[s][comp] for this thread's 8 (fp16: 16) samples, split, into TMEM
      if (F16 && GU_ACP) {
        // (copied by the MMA warp, see above)
      } else if (F16) {
        // pre-split by gram_rsplit_kernel: this thread's 16 samples are two 16-byte chunks of hi and two of lo
        const uint8_t* rr = raw + (size_t)s * rawB + (size_t)sh * GU_RREC + (size_t)comp * 16;
        const uint4 h0 = *reinterpret_cast<const uint4*>(rr), h1 = *reinterpret_cast<const uint4*>(rr + 2048);
        const uint4 l0 = *reinterpret_cast<const uint4*>(rr + 4096), l1 = *reinterpret_cast<const uint4*>(rr + 6144);
        const uint32_t hi[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        const uint32_t lo[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
        const uint32_t ad = tm + lane_base + ACOL + st * 32 + sh * 8;
        tmem_st8(ad, hi);
        tmem_st8(ad + 16, lo);
      } else {
        const float* rawr = reinterpret_cast<const float*>(raw + (size_t)s * rawB) + (sh * 8) * kcb + comp;
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int u = 0; u < 8; u += 2) {
          const float2 r = make_float2(comp_ok ? rawr[u * kcb] : 0.f, comp_ok ? rawr[(u + 1) * kcb] : 0.f);
          split_fast2(r, hi[u], hi[u + 1], lo[u], lo[u + 1]);
        }
        const uint32_t ad = tm + lane_base + ACOL + st * 32 + sh * 8;
        tmem_st8(ad, hi);
        tmem_st8(ad + 16, lo);
      }
      // ---- B operand: phi[s][pair] = zt[s][i] * zt[s][j], split, K-major core-matrix layout
      if (F16) {
        if (gen) {
          const uint8_t* zi = bi + s * sli;
          const uint8_t* zj = bj + s * slj;
          for (int kk = 0; kk < ksn; ++kk) {
            const int ks = ks0 + kk;
            uint8_t* bh = bst + (size_t)st * stageB + (size_t)ks * (2 * NH * 16) + (size_t)pslot * 16;
            uint8_t* bl = bh + NH * 64;
            // the eight 16-byte loads first (the stores below may not be reordered above them)
            float4 av[4], bv[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              av[v] = *reinterpret_cast<const float4*>(zi + ks * 64 + v * 16);
              bv[v] = *reinterpret_cast<const float4*>(zj + ks * 64 + v * 16);
            }
#pragma unroll
            for (int qd = 0; qd < 2; ++qd) {                  // 8 samples = one 16-byte K-chunk of fp16
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float4 a4 = av[qd * 2 + (u >> 1)], b4 = bv[qd * 2 + (u >> 1)];
                // factors carry exact power-of-two scales, so this is the reference's fp32 product times 2^(u_i + u_j)
                const float2 x = (u & 1) ? __fmul2_rn(make_float2(a4.z, a4.w), make_float2(b4.z, b4.w))
                                         : __fmul2_rn(make_float2(a4.x, a4.y), make_float2(b4.x, b4.y));
                const __half2 ah = __floats2half2_rn(x.x, x.y);
                const float2 af = __half22float2(ah);
                const __half2 bh2 = __floats2half2_rn(x.x - af.x, x.y - af.y);
                hi[u] = *reinterpret_cast<const uint32_t*>(&ah);
                lo[u] = *reinterpret_cast<const uint32_t*>(&bh2);
              }
              *reinterpret_cast<uint4*>(bh + (size_t)qd * NH * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(bl + (size_t)qd * NH * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
        }
      } else if (wtid < NH) {
        uint8_t* bh = bst + (size_t)st * stageB + (size_t)wtid * 16;
        uint8_t* bl = bh + NH * 64;
        const uint8_t* zi = bi + s * sli;
        const uint8_t* zj = bj + s * slj;
        // all 32 loads first (the stores below may not be reordered above them), then multiply / split / store
        float av[16], bv[16];
        if (plain) {
#pragma unroll
          for (int sl = 0; sl < 16; ++sl) {
            av[sl] = *reinterpret_cast<const float*>(zi + sl * (SF * 4));
            bv[sl] = *reinterpret_cast<const float*>(zj + sl * (SF * 4));
          }
        } else {
#pragma unroll
          for (int sl = 0; sl < 16; ++sl) {
            av[sl] = *reinterpret_cast<const float*>(zi + sl * sti);
            bv[sl] = *reinterpret_cast<const float*>(zj + sl * stj);
          }
        }
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int u = 0; u < 4; u += 2) {
            const int sl = qd * 4 + u;
            split_fast2(__fmul2_rn(make_float2(av[sl], av[sl + 1]), make_float2(bv[sl], bv[sl + 1])), hi[u], hi[u + 1],
                        lo[u], lo[u + 1]);
          }
          *reinterpret_cast<uint4*>(bh + (size_t)qd * NH * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(bl + (size_t)qd * NH * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      // one arrival per warp: every lane fences its own writes, the warp converges, lane 0 publishes
      fence_proxy_async();
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&S->rempty[s]);
        uint32_t ba = bfull_addr[0];
#pragma unroll
        for (int i = 1; i < GU_NSTG; ++i) ba = (st == i) ? bfull_addr[i] : ba;
        mbar_arrive_cluster(ba);
      }

      // this set has just generated a chunk of the block after nf: fold nf
      while (nf < nblocks - 1 && c >= ((nf + 1) << fls)) fold(nf++);
    }
    while (nf < nblocks) fold(nf++);
    if (nchunks == 0 && set == 0) {      // empty split: contribute zeros
      float* prow = a.part + ((size_t)split * a.Kp + (size_t)cb * GU_CB + comp) * a.PP + (size_t)poff;
      for (int c0 = sh * (GU_NPMAX / 2); c0 < (sh + 1) * (GU_NPMAX / 2) && c0 < NPB; ++c0) prow[c0] = 0.f;
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();      // no CTA of a pair may exit while its peer can still signal it
  if (warp == 1) { if (PAIR) tmem_dealloc2<512>(tm); else tmem_dealloc<512>(tm); }
}


// ---- K <= 64 components: operand roles swapped ----------------------------------------------------------------------
// With K <= 64 the kernel above fills at most half of the 128 accumulator lanes (lanes = components): cfg3 (K = 64) spent
// 13.3 of its 22.4 ms there, cfg4 (K = 32) a quarter of the lanes.  Here the PAIRS are the MMA's M dimension and the
// components its N:   D[pair][k] += phi[pair][n] r[n][k]   with
//   * A = phi^T in TENSOR MEMORY (lane = pair, two fp16 per column: one K-step of 16 samples = 8 columns), written by the
//     worker thread that owns the pair — tcgen05.st straight from the registers it multiplied in, no shared-memory stage;
//   * B = the pre-split weight images AS THEY ARE in the raw ring: [hi | lo][8-sample chunk][component][8 fp16] is
//     the K-major core-matrix layout already (SBO = 128 B between 8-component groups, LBO = Kp x 16 B between the two
//     8-sample chunks of a K-step); only the first Kp = K rounded up to 16 components of each piece are copied in;
//   * a CTA owns GS_J = 2 blocks of 128 pairs — the first and the second pairs of 128 COUPLES that share a factor, see
//     gs_couple — (D1 and D2 of both: 4 Kp <= 256 columns; A stages: 4 x 2 x 32 = 256), so
//     a weight stage is used by 12 MMAs with N = Kp.  MMA time per 32-sample chunk ~ 12 x 0.57 Kp cycles for 256 pairs
//     against 6 x 0.57 x 224 for 224 pairs: 2.1x less at K = 64, 4.3x at K = 32.
// Everything else — the sample image, the balanced sample splits, the two-level fp32 accumulation with round-to-nearest
// folds every GU_FL chunks, the fixed-order fp64 reduce with the resolution check and the TF32 fallback — is shared.
// Couples.  A worker thread of the swapped-role kernel owns TWO pairs that share their first factor, (i, j1) in the CTA's
// pair block 0 and (i, j2) in block 1 (same tensor-memory lane, different columns): three factor rows are loaded for two
// products instead of four — the kernel is paced by exactly those shared-memory loads.  Row i of the upper triangle
// (j = i .. D, the constant feature last) is cut into couples (i, j), (i, j + 1); an odd row ends in a couple without a
// second pair.  The partial sums are written at the pairs' canonical indices (gu_pair_any's order), so the reduce kernel
// does not know about couples.  diag: couple c = ((c, c), (c, D)).
__host__ __device__ inline int gs_ncouples(int D, int diag) {
  if (diag) return D + 1;
  int n = 0;
  for (int m = 1; m <= D + 1; ++m) n += (m + 1) / 2;
  return n;
}
__host__ __device__ inline int gs_pair_index(int i, int j, int D, int diag) {      // inverse of gu_pair_any
  if (diag) return (i == j && i < D) ? i : D + i;
  if (j == D) return D * (D + 1) / 2 + i;
  return i * D - i * (i - 1) / 2 + (j - i);
}
__host__ __device__ inline void gs_couple(int c, int D, int diag, int* i, int* j1, int* j2) {
  if (diag) { *i = c; *j1 = c < D ? c : D; *j2 = c < D ? D : -1; return; }
  int r = 0, rem = c;
  for (;; ++r) {
    const int nc = (D + 1 - r + 1) / 2;                // couples of row r (D + 1 - r pairs)
    if (rem < nc) break;
    rem -= nc;
  }
  *i = r; *j1 = r + 2 * rem; *j2 = (*j1 + 1 <= D) ? *j1 + 1 : -1;
}

constexpr int GS_J = 2;              // pair blocks of 128 per CTA
constexpr int GS_NSTG = 4;           // A (phi) stages in tensor memory
constexpr int GS_KMAX = 64;

struct GsSmem {
  uint64_t rfull[GU_NR], rempty[GU_NR];
  uint64_t bfull[GS_NSTG], bempty[GS_NSTG];
  uint64_t dfull, dempty;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(GU_THREADS, 1) gram_swap_kernel(GuArgs a, int Kp, int ncpl) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int SC = 32;
  const int D = a.d0 + a.d1;
  const int rawR = 8 * Kp * 16;                        // 2 K-steps x [hi c0 | hi c1 | lo c0 | lo c1] x Kp components x 16 B
  const int rawB = (rawR + a.zrec + 127) / 128 * 128;
  const int nr = a.nr;
  uint8_t* raw = smem_raw;
  GsSmem* S = reinterpret_cast<GsSmem*>(raw + (size_t)nr * rawB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntask = (ncpl + 127) / 128;                // 128 couples = two blocks of 128 pairs per CTA
  const int task = (int)blockIdx.x % ntask, split = (int)blockIdx.x / ntask;
  const long long nb = (long long)split * a.S_per;
  long long ne = nb + a.S_per; if (ne > a.N) ne = a.N;
  const int nchunks = ne > nb ? (int)((ne - nb + SC - 1) / SC) : 0;
  const int FL = a.FL;
  constexpr int jn = GS_J;                             // block 0: the couples' first pairs, block 1: their second pairs

  if (tid == 0) {
    for (int s = 0; s < GU_NR; ++s) { mbar_init(&S->rfull[s], 1); mbar_init(&S->rempty[s], 9); }   // 8 worker warps + the MMAs
    for (int s = 0; s < GS_NSTG; ++s) { mbar_init(&S->bfull[s], 8); mbar_init(&S->bempty[s], 1); }
    mbar_init(&S->dfull, 1); mbar_init(&S->dempty, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&S->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = S->tmem_base;
  // TMEM columns: D1 of block j at 64 j, D2 at 128 + 64 j, A stage (st, j) at 256 + 32 (2 st + j): hi 16 columns, lo 16
  constexpr int D2COL = 128, ACOL = 256;

  if (warp == 0) {
    // ================= producer: per chunk 8 slices of the weight images + the transposed sample chunk =================
    const uint32_t bytes = (uint32_t)(rawR + a.zrec);
    const uint32_t piece = (uint32_t)Kp * 16;
    int s = 0;
    uint32_t rph = 0;
    for (int c = 0; c < nchunks; ++c, s = (s + 1 == nr ? 0 : s + 1), rph ^= (s == 0 ? 1u : 0u)) {
      mbar_wait(&S->rempty[s], rph ^ 1);
      if (elect_one()) {
        const long long r0 = nb + (long long)c * SC;
        uint8_t* dst = raw + (size_t)s * rawB;
        mbar_arrive_expect_tx(&S->rfull[s], bytes);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint8_t* rec = a.rp + (size_t)((r0 >> 4) + ks) * GU_RREC;          // component block 0: K <= 64
#pragma unroll
          for (int pc = 0; pc < 4; ++pc)
            bulk_g2s(dst + (size_t)(ks * 4 + pc) * piece, rec + (size_t)pc * 2048, piece, &S->rfull[s]);
        }
        bulk_g2s(dst + rawR, a.zt + (size_t)(r0 >> 5) * a.zrec, (uint32_t)a.zrec, &S->rfull[s]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = idesc_f16(128, Kp);
    const uint32_t piece = (uint32_t)Kp * 16;
    int s = 0, fc = 0, nflush = 0;
    uint32_t rph = 0;
    for (int c = 0; c < nchunks; ++c, s = (s + 1 == nr ? 0 : s + 1), rph ^= (s == 0 ? 1u : 0u)) {
      const int st = c % GS_NSTG;
      mbar_wait(&S->rfull[s], rph);                          // the weight slices are read by the tensor core itself
      mbar_wait(&S->bfull[st], (c / GS_NSTG) & 1);
      const bool first = (fc == 0);
      if (first && c > 0) mbar_wait(&S->dempty, (nflush - 1) & 1);
      tc_fence_after();
      __syncwarp();
      const bool flush = (++fc == FL) || (c == nchunks - 1);
      if (elect_one()) {
        const uint32_t sbase = smem_u32(raw + (size_t)s * rawB);
        for (int j = 0; j < jn; ++j) {
          const uint32_t dcol = tm + j * 64;
          const uint32_t a_hi = tm + ACOL + (st * GS_J + j) * 32, a_lo = a_hi + 16;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t b_hi = smem_desc(sbase + (uint32_t)(ks * 4) * piece, piece, 128);
            const uint64_t b_lo = smem_desc(sbase + (uint32_t)(ks * 4 + 2) * piece, piece, 128);
            mma_f16_ts(dcol, a_lo + ks * 8, b_hi, idesc, !(first && ks == 0));
            mma_f16_ts(dcol, a_hi + ks * 8, b_lo, idesc, 1);
            mma_f16_ts(dcol, a_hi + ks * 8, b_hi, idesc, 1);
          }
        }
        mma_commit(&S->bempty[st]);
        mma_commit(&S->rempty[s]);
        if (flush) mma_commit(&S->dfull);
      }
      __syncwarp();
      if (flush) { fc = 0; ++nflush; }
    }
  } else {
    // ================= workers: thread = couple (tensor-memory lane), two sets of 8 warps alternate chunks ================
    // within a set, warps 0-3 generate K-step 0 of both pair blocks and fold block 0, warps 4-7 K-step 1 and block 1
    const int set = (warp - 2) >> 3, w8 = (warp - 2) & 7;
    const int ksel = w8 >> 2, q = warp & 3;                   // K-step / folded block, lane quarter (a warp reaches lanes 32 (warp % 4) ..)
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int prow = q * 32 + lane;
    const int cidx = task * 128 + prow;                       // couple
    const bool c_ok = cidx < ncpl;
    int ci = 0, cj1 = 0, cj2 = -1;
    if (c_ok) gs_couple(cidx, D, a.diag, &ci, &cj1, &cj2);
    const bool has2 = c_ok && cj2 >= 0;
    const uint8_t* bi = raw + rawR + (size_t)(c_ok ? ci : D + 1) * (GU_ZS * 4) + ksel * 64;      // row D + 1 of the record: zeros
    const uint8_t* bj1 = raw + rawR + (size_t)(c_ok ? cj1 : D + 1) * (GU_ZS * 4) + ksel * 64;
    const uint8_t* bj2 = raw + rawR + (size_t)(has2 ? cj2 : D + 1) * (GU_ZS * 4) + ksel * 64;
    // the row of the partials this thread writes when it folds block ksel: the canonical index of that pair
    const int pout = !c_ok ? -1 : (ksel == 0 ? gs_pair_index(ci, cj1, D, a.diag) : (has2 ? gs_pair_index(ci, cj2, D, a.diag) : -1));
    const int fls = __ffs(FL) - 1;
    // phi of 16 samples (two 16-byte K-chunks of fp16) for one pair: multiply, split, store both images
    auto gen = [&](const float4 (&av)[4], const float4 (&bv)[4], uint32_t ad) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float4 a4 = av[u >> 1], b4 = bv[u >> 1];
        // factors carry exact power-of-two scales, so this is the reference's fp32 product times 2^(u_i + u_j)
        const float2 x = (u & 1) ? __fmul2_rn(make_float2(a4.z, a4.w), make_float2(b4.z, b4.w))
                                 : __fmul2_rn(make_float2(a4.x, a4.y), make_float2(b4.x, b4.y));
        const __half2 ah = __floats2half2_rn(x.x, x.y);
        const float2 af = __half22float2(ah);
        const __half2 bh2 = __floats2half2_rn(x.x - af.x, x.y - af.y);
        hi[u] = *reinterpret_cast<const uint32_t*>(&ah);
        lo[u] = *reinterpret_cast<const uint32_t*>(&bh2);
      }
      tmem_st8(ad + ksel * 8, hi);
      tmem_st8(ad + 16 + ksel * 8, lo);
    };
    int s = set;
    uint32_t rph = 0;
    for (int c = set; c < nchunks; c += 2, s += 2) {
      if (s >= nr) { s -= nr; rph ^= 1u; }
      const int st = c % GS_NSTG;
      const int nflush = c >> fls;
      mbar_wait(&S->rfull[s], rph);
      mbar_wait(&S->bempty[st], ((c / GS_NSTG) & 1) ^ 1);
      tc_fence_after();
      {
        const size_t so = (size_t)s * rawB;
        float4 av[4], b1[4], b2[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          av[v] = *reinterpret_cast<const float4*>(bi + so + v * 16);
          b1[v] = *reinterpret_cast<const float4*>(bj1 + so + v * 16);
          b2[v] = *reinterpret_cast<const float4*>(bj2 + so + v * 16);
        }
        const uint32_t ad = tm + lane_base + ACOL + (st * GS_J) * 32;
        gen(av, b1, ad);
        gen(av, b2, ad + 32);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&S->rempty[s]); mbar_arrive(&S->bfull[st]); }

      const bool last = (c == nchunks - 1);
      if (((c + 1) & (FL - 1)) == 0 || last) {
        // ---- fold D1 into D2 (or, at the end, write D1 + D2 to this split's partial): thread = pair row of block ksel
        if (nflush > 0) mbar_wait(&S->dfull, (nflush - 1) & 1);
        mbar_wait(&S->dfull, nflush & 1);
        tc_fence_after();
        const bool firstf = (nflush == 0);
        {
          // (the tensor-memory loads / stores are warp-collective: every lane runs them, only the global stores of
          // lanes without a pair in this block are skipped)
          float* prw = a.part + ((size_t)split * a.PP + (size_t)(pout >= 0 ? pout : 0)) * Kp;
          for (int c0 = 0; c0 < Kp; c0 += 16) {
            float v1[16], v2[16];
            tmem_ld16(tm + lane_base + ksel * 64 + c0, v1);
            if (!firstf) tmem_ld16(tm + lane_base + D2COL + ksel * 64 + c0, v2);
            tmem_wait_ld();
            if (!firstf) {
#pragma unroll
              for (int u = 0; u < 16; ++u) v1[u] += v2[u];
            }
            if (last) {
              if (pout >= 0) {
#pragma unroll
                for (int u = 0; u < 16; u += 4)
                  *reinterpret_cast<float4*>(prw + c0 + u) = make_float4(v1[u], v1[u + 1], v1[u + 2], v1[u + 3]);
              }
            } else {
              tmem_st16(tm + lane_base + D2COL + ksel * 64 + c0, reinterpret_cast<const uint32_t*>(v1));
            }
          }
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S->dempty);
      }
    }
    if (nchunks == 0 && set == 0 && pout >= 0) {      // empty split: contribute zeros
      float* prw = a.part + ((size_t)split * a.PP + (size_t)pout) * Kp;
      for (int c0 = 0; c0 < Kp; ++c0) prw[c0] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tm);
}

// gram[k][i][j] = gram[k][j][i] = sum_split part[split][k][pair(i,j)], fp64, fixed order.
//   MODE 0: partials of the TF32 kernel (raw units), unconditional.
//   MODE 1: partials of the fp16 kernel: they carry the factor 2^(14 + u_i + u_j).  The threads of the diagonal pairs
//           also check that the fp16 window resolved every component: mean_k(z'_i^2) (z' = z 2^u_i, column maximum in
//           [2^6, 2^7)) below 2^-6 means component k's samples sit more than ~9 binades under feature i's scale, where
//           the low piece of phi' = z'_i z'_j falls into fp16's subnormals (absolute floor 2^-25) and the statistic
//           would lose bits the fp32 reference keeps (one huge outlier in a column, a tight cluster inside a wide data
//           range).  Such a component raises flag[0]; the TF32 kernel behind this one then recomputes the partials.
//   MODE 2: runs after the conditional TF32 kernel; returns at once unless flag[0] is set, else as MODE 0.
template <int MODE>
__global__ void gram_pair_reduce_kernel(const float* __restrict__ part, int splits, int K, int Kp, int PP, int D1,
                                        const uint32_t* __restrict__ cmax, uint32_t* __restrict__ flag,
                                        float* __restrict__ gram, int diag, int swapKp = 0) {
  // swapKp > 0: partials of gram_swap_kernel, part[split][pair (PP)][component (swapKp)]
  if (MODE == 2 && flag[0] == 0u) return;
  const int P = gu_npairs(D1 - 1, diag);
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)K * P) return;
  const int k = (int)(e / P), p = (int)(e % P);
  int i, j;
  gu_pair_any(p, D1 - 1, diag, &i, &j);
  double acc = 0.0;
  if (swapKp > 0) { for (int s = 0; s < splits; ++s) acc += (double)part[((size_t)s * PP + p) * swapKp + k]; }
  else { for (int s = 0; s < splits; ++s) acc += (double)part[((size_t)s * Kp + k) * PP + p]; }
  if (MODE == 1) {
    const int ui = gu_feat_exp(i == D1 - 1 ? 0x3f800000u : cmax[i]), uj = gu_feat_exp(j == D1 - 1 ? 0x3f800000u : cmax[j]);
    // column i needs the check only if it holds a nonzero value below 2^-3 after scaling (column maximum in [2^6, 2^7))
    const bool small_values = cmax[i < D1 - 1 ? i : 0] != 0u && cmax[VBMP_MAX_D + (i < D1 - 1 ? i : 0)] != 0u &&
        ldexpf(__uint_as_float(0x7f800000u - cmax[VBMP_MAX_D + (i < D1 - 1 ? i : 0)]), ui) < 0.125f;
    if (i == j && i < D1 - 1 && small_values) {
      double accn = 0.0;                                   // the (1, 1) pair: 2^(14 + 12) sum_n r[n,k]
      if (swapKp > 0) { for (int s = 0; s < splits; ++s) accn += (double)part[((size_t)s * PP + (P - 1)) * swapKp + k]; }
      else { for (int s = 0; s < splits; ++s) accn += (double)part[((size_t)s * Kp + k) * PP + (P - 1)]; }
      const double nk = ldexp(accn, -(GU_RSH + 12));
      if (nk >= 4.0 && acc * 4096.0 < accn * 0.015625) atomicOr(flag, 2u);     // mean z'^2 < 2^-6 over >= 4 samples' weight
    }
    acc = ldexp(acc, -(GU_RSH + ui + uj));
  }
  const float v = (float)acc;
  gram[((size_t)k * D1 + i) * D1 + j] = v;
  gram[((size_t)k * D1 + j) * D1 + i] = v;
}

// column maxima of |z| (bit patterns, which order like the values for non-negative floats) into hdr[col0 + c]
__global__ void __launch_bounds__(256) gram_colmax_kernel(const float* __restrict__ z, int d, long long N, int col0,
                                                          uint32_t* __restrict__ hdr) {
  // hdr[c] = bits of max |z_c|; hdr[VBMP_MAX_D + c] = 0x7f800000 - bits of the smallest NONZERO |z_c| (0 = none seen):
  // the resolution check of the first reduce only applies to columns that hold nonzero values far below their maximum —
  // exact zeros (one-hot / sparse / integer-coded features) are exact in any format and must not trigger the fallback
  __shared__ uint32_t smax[VBMP_MAX_D], smin[VBMP_MAX_D];
  if (threadIdx.x < VBMP_MAX_D) { smax[threadIdx.x] = 0u; smin[threadIdx.x] = 0u; }
  __syncthreads();
  const int d4 = d >> 2;                                         // d % 4 == 0
  const long long T = ((long long)gridDim.x * blockDim.x) / d4 * d4;   // threads that keep a fixed column group
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < T) {
    const long long n4 = N * d4;
    float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
    const float BIG = __uint_as_float(0x7f800000u);
    float n0 = BIG, n1 = BIG, n2 = BIG, n3 = BIG;
    const float4* z4 = reinterpret_cast<const float4*>(z);
    for (long long e = g; e < n4; e += T) {
      const float4 v = __ldg(z4 + e);
      const float a0 = fabsf(v.x), a1 = fabsf(v.y), a2 = fabsf(v.z), a3 = fabsf(v.w);
      m0 = fmaxf(m0, a0); m1 = fmaxf(m1, a1); m2 = fmaxf(m2, a2); m3 = fmaxf(m3, a3);
      n0 = a0 > 0.f ? fminf(n0, a0) : n0; n1 = a1 > 0.f ? fminf(n1, a1) : n1;
      n2 = a2 > 0.f ? fminf(n2, a2) : n2; n3 = a3 > 0.f ? fminf(n3, a3) : n3;
    }
    const int c = (int)(g % d4) * 4;
    atomicMax(&smax[c], __float_as_uint(m0)); atomicMax(&smax[c + 1], __float_as_uint(m1));
    atomicMax(&smax[c + 2], __float_as_uint(m2)); atomicMax(&smax[c + 3], __float_as_uint(m3));
    atomicMax(&smin[c], 0x7f800000u - __float_as_uint(n0)); atomicMax(&smin[c + 1], 0x7f800000u - __float_as_uint(n1));
    atomicMax(&smin[c + 2], 0x7f800000u - __float_as_uint(n2)); atomicMax(&smin[c + 3], 0x7f800000u - __float_as_uint(n3));
  }
  __syncthreads();
  if (threadIdx.x < d && smax[threadIdx.x] != 0u) {
    atomicMax(&hdr[col0 + threadIdx.x], smax[threadIdx.x]);
    atomicMax(&hdr[VBMP_MAX_D + col0 + threadIdx.x], smin[threadIdx.x]);
  }
}

// fp16 variant pre-pass: r' = r 2^14 = a + b (a = rn_fp16(r'), b = rn_fp16(r' - a)) written once as the TMEM images the
// Gram CTAs copy in (see GuArgs::rp).  Without it every pair-block CTA (12 at d = 64) repeated this split on the same
// weights, which was 60 % of the worker instructions.  Thread = (8-sample chunk, component): 8 coalesced reads down a
// column of p, one 16-byte store each to hi and lo.  Weights beyond fp16 range raise the fallback flag.
__global__ void __launch_bounds__(256) gram_rsplit_kernel(const float* __restrict__ p, long long N, int K, int ncb, long long nsb,
                                                          uint8_t* __restrict__ rp, uint32_t* __restrict__ flag) {
  const long long total = nsb * 2 * (long long)ncb * GU_CB;           // (chunk of 8 samples) x padded component
  const float rsc = (float)(1 << GU_RSH);
  float rmax = 0.f;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int kp = (int)(e % (ncb * GU_CB));
    const long long ch = e / (ncb * GU_CB);                            // global 8-sample chunk
    const int cb = kp / GU_CB, comp = kp % GU_CB;
    const long long n0 = ch * 8;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long n = n0 + 2 * u;
      const float r0 = (kp < K && n < N) ? __ldg(p + (size_t)n * K + kp) : 0.f;
      const float r1 = (kp < K && n + 1 < N) ? __ldg(p + (size_t)(n + 1) * K + kp) : 0.f;
      rmax = fmaxf(rmax, fmaxf(fabsf(r0), fabsf(r1)));                 // (a NaN weight propagates through the MMA by itself)
      const float x0 = r0 * rsc, x1 = r1 * rsc;
      const __half2 ah = __floats2half2_rn(x0, x1);
      const float2 af = __half22float2(ah);
      const __half2 bh = __floats2half2_rn(x0 - af.x, x1 - af.y);
      hi[u] = *reinterpret_cast<const uint32_t*>(&ah);
      lo[u] = *reinterpret_cast<const uint32_t*>(&bh);
    }
    uint8_t* rec = rp + ((size_t)cb * nsb + (size_t)(ch >> 1)) * GU_RREC + (size_t)(ch & 1) * 2048 + (size_t)comp * 16;
    *reinterpret_cast<uint4*>(rec) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(rec + 4096) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  if (__any_sync(0xffffffffu, rmax > GU_RMAX) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}

// fp16 variant pre-pass: zt = [z0 | z1 | 1] scaled by the exact feature scales 2^u_i and transposed per 32-sample chunk
// (see GuArgs::zt).  One block iteration per chunk: coalesced row reads into a shared tile, feature-major 16-byte writes.
__global__ void __launch_bounds__(256) gram_zprep_kernel(const float* __restrict__ z0, int d0, const float* __restrict__ z1,
                                                         int d1, long long N, const uint32_t* __restrict__ hdr,
                                                         uint8_t* __restrict__ zt, int zrec) {
  __shared__ float tile[VBMP_MAX_D][33];         // [feature][sample], stride 33: both phases are bank-conflict free
  __shared__ float fs[VBMP_MAX_D];
  const int D = d0 + d1;
  if (threadIdx.x < D) fs[threadIdx.x] = __uint_as_float((uint32_t)(127 + gu_feat_exp(hdr[threadIdx.x])) << 23);
  __syncthreads();
  const long long nch = (N + 31) / 32;
  const int sidx = threadIdx.x >> 3, c0 = threadIdx.x & 7;       // 8 threads per sample row, float4 columns c0 + 8 h
  const int q0 = d0 >> 2, q1 = d1 >> 2;
  for (long long ch = blockIdx.x; ch < nch; ch += gridDim.x) {
    const long long n = ch * 32 + sidx;
    const bool ok = n < N;
#pragma unroll
    for (int h = 0; h < VBMP_MAX_D / 32; ++h) {
      const int c4 = c0 + 8 * h;
      if (c4 < q0) {
        const float4 v = ok ? __ldg(reinterpret_cast<const float4*>(z0 + (size_t)n * d0) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        const int f = 4 * c4;
        tile[f][sidx] = v.x * fs[f]; tile[f + 1][sidx] = v.y * fs[f + 1];
        tile[f + 2][sidx] = v.z * fs[f + 2]; tile[f + 3][sidx] = v.w * fs[f + 3];
      }
      if (c4 < q1) {
        const float4 v = ok ? __ldg(reinterpret_cast<const float4*>(z1 + (size_t)n * d1) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        const int f = d0 + 4 * c4;
        tile[f][sidx] = v.x * fs[f]; tile[f + 1][sidx] = v.y * fs[f + 1];
        tile[f + 2][sidx] = v.z * fs[f + 2]; tile[f + 3][sidx] = v.w * fs[f + 3];
      }
    }
    __syncthreads();
    float* rec = reinterpret_cast<float*>(zt + (size_t)ch * zrec);
    for (int e = threadIdx.x; e < (D + 2) * 8; e += 256) {              // (feature row, group of 4 samples)
      const int f = e >> 3, s4 = (e & 7) * 4;
      float4 o;
      if (f < D) o = make_float4(tile[f][s4], tile[f][s4 + 1], tile[f][s4 + 2], tile[f][s4 + 3]);
      else if (f == D) {           // the constant feature: 1 * 2^6 on real rows, 0 on the zero-filled tail
        const long long m = ch * 32 + s4;
        o = make_float4(m < N ? 64.f : 0.f, m + 1 < N ? 64.f : 0.f, m + 2 < N ? 64.f : 0.f, m + 3 < N ? 64.f : 0.f);
      } else o = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(rec + f * GU_ZS + s4) = o;
    }
    __syncthreads();
  }
}

// ---- host side -------------------------------------------------------------------------------------------
static int gu_num_sms() { return num_sms(); }

static bool gu_pair_mode(int K) {
  static const int pair_ok = [] { const char* e = getenv("VBMP_GRAM_PAIR"); return e ? atoi(e) : 1; }();
  return pair_ok && (((K + GU_CB - 1) / GU_CB) % 2 == 0);     // CTA pairs share the phi block (cta_group::2)
}

static void gu_plan(long long N, int K, int D, int sms, GuArgs* g, int diag = 0, int lin_cols = 0) {
  const int GU_NPMAX = gu_npmax(gu_pair_mode(K));
  g->diag = diag;
  g->lin = lin_cols > 0;
  g->P = lin_cols > 0 ? lin_cols : gu_npairs(D, diag);
  static const int npb_cap = [] { const char* e = getenv("VBMP_GRAM_NPB"); return e ? atoi(e) : 0; }();   // tuning: pair columns per block
  const int npmax = (npb_cap >= 16 && npb_cap <= GU_NPMAX) ? npb_cap / 16 * 16 : GU_NPMAX;
  // 16-column units dealt out evenly: the first wextra blocks are one unit wider than the rest
  const int units = (g->P + 15) / 16;
  g->npb = (units * 16 + npmax - 1) / npmax;
  g->wbase = units / g->npb;
  g->wextra = units % g->npb;
  g->NPB = 16 * (g->wbase + (g->wextra ? 1 : 0));
  g->ncb = (K + GU_CB - 1) / GU_CB;
  const int tasks = g->npb * g->ncb;
  // sample splits: a multiple of the number of task groups that fit the SMs at once, with at most `cap` rows per CTA.
  // The pair-block CTAs of one (split, component block) stream the same weight rows and are launched back to back, so
  // their re-reads hit L2 as long as they do not drift apart; short tasks bound the drift.  ncu, cfg2, fp16 variant:
  // 13.4 GB of DRAM reads per launch with 128 Ki-row tasks, 8.0 GB at 64 Ki, 6.6 GB at 32 Ki (5.5 GB is one pass over
  // the packed weights and samples), 5.8 GB at 16 Ki where the per-task prologue starts to cost time; 32 Ki is also
  // the fastest.  rounds of `sms` co-resident CTAs needed at <= cap rows per CTA; then as many splits as fill those
  // rounds exactly (cfg2: 24 tasks, 21 rounds of 148 -> 129 splits), so the last round is not a partial one.
  // VBMP_GRAM_ROWS overrides the cap (tuning).
  static const long long cap = [] { const char* e = getenv("VBMP_GRAM_ROWS"); long long v = e ? atoll(e) : 0; return v >= 2048 ? v : 32768; }();
  long long rounds = ((long long)N * tasks + (long long)sms * cap - 1) / ((long long)sms * cap);
  if (rounds < 1) rounds = 1;
  long long sp = rounds * sms / tasks;
  const long long maxsp = (N + 2047) / 2048;
  if (sp > maxsp) sp = maxsp;
  if (sp < 1) sp = 1;
  long long per = (N + sp - 1) / sp;
  per = (per + 31) / 32 * 32;                  // whole chunks of either variant (16 / 32 samples)
  g->S_per = per;
  g->splits = (int)((N + per - 1) / per);
  if (g->splits < 1) g->splits = 1;
  g->Kp = g->ncb * GU_CB;
  g->PP = units * 16;
}

static size_t gu_al(size_t x) { return (x + 255) / 256 * 256; }

// ---- plan of the swapped-role kernel (K <= GS_KMAX): blocks of 128 pairs, GS_J per CTA, same split rule as gu_plan
struct GsPlan { int Kp, nblk, ncpl, ntask, splits, PP; long long S_per; };
static bool gs_enabled() {
  static const int on = [] { const char* e = getenv("VBMP_GRAM_SWAP"); return e ? atoi(e) : 1; }();
  return on != 0;
}
static GsPlan gs_plan(long long N, int K, int D, int diag, int sms) {
  GsPlan q{};
  q.Kp = (K + 15) / 16 * 16;
  const int P = gu_npairs(D, diag);
  q.nblk = (P + 127) / 128;
  q.ncpl = gs_ncouples(D, diag);                     // a CTA takes 128 couples = two blocks of 128 pairs
  q.ntask = (q.ncpl + 127) / 128;
  q.PP = q.nblk * 128;                               // rows of the partials: canonical pair index
  const long long cap = 32768;
  long long rounds = ((long long)N * q.ntask + (long long)sms * cap - 1) / ((long long)sms * cap);
  if (rounds < 1) rounds = 1;
  long long sp = rounds * sms / q.ntask;
  const long long maxsp = (N + 2047) / 2048;
  if (sp > maxsp) sp = maxsp;
  if (sp < 1) sp = 1;
  long long per = (N + sp - 1) / sp;
  per = (per + 31) / 32 * 32;
  q.S_per = per;
  q.splits = (int)((N + per - 1) / per);
  if (q.splits < 1) q.splits = 1;
  return q;
}
static long long gu_nsb(long long N) { return (N + 31) / 32 * 2; }
static size_t gu_rp_bytes(long long N, int ncb) { return (size_t)ncb * (size_t)gu_nsb(N) * GU_RREC; }
static size_t gu_zt_bytes(long long N, int D) { return (size_t)((N + 31) / 32) * gu_zrec(D) + 256; }

static bool gu_use_f16();
// bytes of the pre-split weight images of (N, K): what gram_rsplit_kernel writes and vbmp_estep_rpack may write instead
size_t gram_rpack_bytes(long long N, int K) { return gu_rp_bytes(N, (K + GU_CB - 1) / GU_CB); }
bool gram_rpack_usable() { return gu_use_f16(); }
// bytes of the sample image of (N, D): [column maxima, GU_HDR_WORDS words | transposed, pre-scaled 32-sample chunks]
size_t gram_zpack_bytes(long long N, int D) { return GU_HDR_WORDS * sizeof(uint32_t) + gu_al(gu_zt_bytes(N, D)); }

bool gram_umma_supported(long long N, int GX, int GP, int G, int K, int Dp, int d0, int d1, bool has_p) {
  const int D = d0 + d1;
  // D > 64 (Dp = 128) exists with fp16 operands only: the TF32 kernel stays its on-device fallback
  return has_p && G == 1 && GX == 1 && GP == 1 && Dp >= 16 && Dp <= 128 && D <= 128 && (D <= 64 || gu_use_f16()) &&
         (K % 4 == 0) && (d0 % 4 == 0) && (d1 % 4 == 0) && N >= 2048;
}

// [flag block | per-split partials | weight images unless handed in | sample image unless handed in]
size_t gram_umma_workspace_bytes(long long N, int G, int K, int d0, int d1, int Dp, bool has_rpack, bool has_zpack) {
  if (!gram_umma_supported(N, 1, 1, G, K, Dp, d0, d1, true)) return 0;
  GuArgs g{};
  gu_plan(N, K, d0 + d1, gu_num_sms(), &g);   // same plan as the launch
  size_t part = (size_t)g.splits * g.Kp * g.PP * sizeof(float);
  if (K <= GS_KMAX) {                         // the swapped-role kernel's partials share the region
    const GsPlan q = gs_plan(N, K, d0 + d1, 0, gu_num_sms());
    const size_t ps = (size_t)q.splits * q.PP * q.Kp * sizeof(float);
    if (ps > part) part = ps;
  }
  return 512 + GU_HDR_WORDS * sizeof(uint32_t) + gu_al(part) +
         (has_rpack ? 0 : gu_al(gu_rp_bytes(N, g.ncb))) + (has_zpack ? 0 : gram_zpack_bytes(N, d0 + d1));
}

static bool gu_use_f16() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VBMP_GRAM_PREC");
    v = (e && (e[0] == 't' || e[0] == 'T')) ? 0 : 1;
  }
  return v == 1;
}

static int gu_fl() {
  static int fl = 0;
  if (fl == 0) {
    const char* e = getenv("VBMP_GRAM_FL");     // tuning knob: chunks (16 samples) per first-level accumulation block
    fl = e ? atoi(e) : GU_FL;
    if (fl < 1 || (fl & (fl - 1))) fl = GU_FL;   // power of two
  }
  return fl;
}

// lin_lds > 0 (TF32 variant only): weighted column sums of z0 (N, d0) with row stride lin_lds, see launch_wsum_umma
template <bool F16>
static int gu_launch_main(const GramArgs& a, GuArgs g, bool pair, cudaStream_t st, int lin_lds = 0) {
  constexpr int SC = gu_sc(F16);
  const int D = a.d0 + a.d1;
  (void)D;
  g.zw = lin_lds > 0 ? g.NPB : a.d0;
  CUtensorMap tmR, tmZ0, tmZ1;
  int e = make_tmap_2d(&tmR, a.p, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.K, (uint32_t)g.kcb, SC);
  if (!e) e = make_tmap_2d(&tmZ0, a.z0, (uint64_t)a.d0, (uint64_t)a.N, (uint64_t)(lin_lds > 0 ? lin_lds : a.d0), (uint32_t)g.zw, SC);
  if (!e && a.d1 > 0) e = make_tmap_2d(&tmZ1, a.z1, (uint64_t)a.d1, (uint64_t)a.N, (uint64_t)a.d1, (uint32_t)a.d1, SC);
  if (a.d1 == 0) tmZ1 = tmZ0;
  if (e) { set_error("gram_umma: cuTensorMapEncodeTiled failed (%d)", e); return VBMP_ERR_CUDA; }
  const int NH = pair ? g.NPB / 2 : g.NPB;
  const int rawB = (F16 ? 2 * GU_RREC + g.zrec : SC * g.kcb * 4 + SC * g.zw * 4 + SC * a.d1 * 4) / 128 * 128 + 128;
  const size_t stages = (size_t)gu_nstg(pair) * 2 * NH * 64 + sizeof(GuSmem) + 64;
  int nr = (int)((227 * 1024 - stages) / rawB);
  if (nr > GU_NR) nr = GU_NR;
  if (nr < 2) { set_error("gram_umma: D=%d does not fit shared memory", D); return VBMP_ERR_UNSUPPORTED; }
  g.nr = nr;
  const size_t smem = (size_t)nr * rawB + stages;
  const int grid = g.splits * g.ncb * g.npb;
  const bool same = (a.d1 == 0 || a.d1 == a.d0);
  const int sf = same && lin_lds == 0 && (a.d0 == 64 || a.d0 == 32 || a.d0 == 16) ? a.d0 : 0;
  const bool r128 = g.kcb == GU_CB;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(GU_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaSuccess;
#define GU_LAUNCH(SF, R, P)                                                                                            \
  do {                                                                                                                 \
    cudaFuncSetAttribute(gram_umma_kernel<SF, R, P, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
    le = cudaLaunchKernelEx(&cfg, gram_umma_kernel<SF, R, P, F16>, tmR, tmZ0, tmZ1, g);                                \
  } while (0)
  if (pair) {                                               // K >= 256: the R box is always 128 wide
    if (sf == 64) GU_LAUNCH(64, true, true);
    else if (sf == 32) GU_LAUNCH(32, true, true);
    else if (sf == 16) GU_LAUNCH(16, true, true);
    else GU_LAUNCH(0, true, true);
  } else if (sf == 64 && r128) GU_LAUNCH(64, true, false);
  else if (sf == 64) GU_LAUNCH(64, false, false);
  else if (sf == 32 && r128) GU_LAUNCH(32, true, false);
  else if (sf == 32) GU_LAUNCH(32, false, false);
  else if (sf == 16) GU_LAUNCH(16, false, false);
  else if (r128) GU_LAUNCH(0, true, false);
  else GU_LAUNCH(0, false, false);
#undef GU_LAUNCH
  if (le != cudaSuccess) { set_error("gram_umma launch: %s", cudaGetErrorString(le)); return VBMP_ERR_CUDA; }
  return check_launch(F16 ? "gram_umma_f16" : "gram_umma");
}

// column maxima -> exact power-of-two feature scales; then the transposed, pre-scaled sample chunks (fp16 variant).
// zpack = [GU_HDR_WORDS words | image]: what vbmp_gram_zpack hands back to the caller, who may keep it for as long as the
// samples do not change (the rows of an EM run are the same every iteration, dists/Mixture.py:54-62).
int launch_gram_zpack(const float* z0, int d0, const float* z1, int d1, long long N, void* zpack, cudaStream_t st) {
  const int D = d0 + d1;
  uint32_t* cmax = (uint32_t*)zpack;
  if (cudaMemsetAsync(cmax, 0, GU_HDR_WORDS * sizeof(uint32_t), st) != cudaSuccess) {
    set_error("gram_zpack: memset failed"); return VBMP_ERR_CUDA;
  }
  const int cg = gu_num_sms() * 4;
  gram_colmax_kernel<<<cg, 256, 0, st>>>(z0, d0, N, 0, cmax);
  if (d1 > 0) { gram_colmax_kernel<<<cg, 256, 0, st>>>(z1, d1, N, d0, cmax); count_launch(1); }
  int rc = check_launch("gram_colmax");
  if (rc) return rc;
  gram_zprep_kernel<<<gu_num_sms() * 8, 256, 0, st>>>(z0, d0, z1, d1, N, cmax, (uint8_t*)(cmax + GU_HDR_WORDS), gu_zrec(D));
  return check_launch("gram_zprep");
}
bool gram_zpack_usable(long long N, int K, int Dp, int d0, int d1) {
  return gu_use_f16() && gram_umma_supported(N, 1, 1, 1, K, Dp, d0, d1, true);
}

int launch_gram_umma(const GramArgs& a, float* gram, void* ws, size_t ws_bytes, cudaStream_t st) {
  GuArgs g{};
  g.d0 = a.d0; g.d1 = a.d1; g.N = a.N; g.K = a.K;
  const int D = a.d0 + a.d1;
  gu_plan(a.N, a.K, D, gu_num_sms(), &g, a.diag ? 1 : 0);
  g.kcb = a.K < GU_CB ? a.K : GU_CB;
  g.FL = gu_fl();
  const bool f16 = gu_use_f16();
  const bool swap = f16 && a.K <= GS_KMAX && gs_enabled();
  const GsPlan q = gs_plan(a.N, a.K, D, a.diag ? 1 : 0, gu_num_sms());
  size_t part_raw = (size_t)g.splits * g.Kp * g.PP * sizeof(float);
  if (a.K <= GS_KMAX) {
    const GsPlan q0 = gs_plan(a.N, a.K, D, 0, gu_num_sms());        // as the workspace query sized it (diag or not)
    part_raw = std::max(part_raw, (size_t)q0.splits * q0.PP * q0.Kp * sizeof(float));
  }
  const size_t part_bytes = gu_al(part_raw);
  const size_t need = gram_umma_workspace_bytes(a.N, a.G, a.K, a.d0, a.d1, a.Dp, a.rpack != nullptr, a.zpack != nullptr);
  if (ws_bytes < need) { set_error("gram_umma: workspace too small (%zu < %zu)", ws_bytes, need); return VBMP_ERR_WORKSPACE; }
  uint32_t* flag = (uint32_t*)gu_al((size_t)ws);
  g.part = (float*)(flag + GU_HDR_WORDS);
  const bool pair = gu_pair_mode(a.K);
  const int D1 = D + 1;
  if (g.diag && cudaMemsetAsync(gram, 0, (size_t)a.K * D1 * D1 * sizeof(float), st) != cudaSuccess) {
    set_error("gram_umma: memset failed"); return VBMP_ERR_CUDA;       // only the diagonal and the last row / column are written
  }
  const long long tot = (long long)a.K * g.P;
  const unsigned rgrid = (unsigned)((tot + 255) / 256);
  int rc;
  if (!f16) {
    g.flag = nullptr; g.cmax = nullptr;
    rc = gu_launch_main<false>(a, g, pair, st);
    if (rc) return rc;
    gram_pair_reduce_kernel<0><<<rgrid, 256, 0, st>>>(g.part, g.splits, a.K, g.Kp, g.PP, D1, nullptr, nullptr, gram, g.diag);
    return check_launch("gram_pair_reduce");
  }
  if (cudaMemsetAsync(flag, 0, GU_HDR_WORDS * sizeof(uint32_t), st) != cudaSuccess) {
    set_error("gram_umma: memset failed"); return VBMP_ERR_CUDA;
  }
  g.flag = flag;
  g.nsb = gu_nsb(a.N);
  uint8_t* next = (uint8_t*)g.part + part_bytes;
  if (a.rpack) {
    g.rp = a.rpack;                                        // already written by the E-step's normaliser
  } else {
    g.rp = next;
    gram_rsplit_kernel<<<gu_num_sms() * 8, 256, 0, st>>>(a.p, a.N, a.K, g.ncb, g.nsb, next, flag);
    rc = check_launch("gram_rsplit");
    if (rc) return rc;
    next += gu_al(gu_rp_bytes(a.N, g.ncb));
  }
  const uint8_t* zp = a.zpack;
  if (!zp) {                                               // no image handed in: make it in the workspace
    rc = launch_gram_zpack(a.z0, a.d0, a.z1, a.d1, a.N, next, st);
    if (rc) return rc;
    zp = next;
  }
  g.cmax = (const uint32_t*)zp;
  g.zt = zp + GU_HDR_WORDS * sizeof(uint32_t);
  g.zrec = gu_zrec(D);
  // 16 chunks of 32 samples per first-level block: a kind::f16 MMA accumulates 16 samples per step (TF32: 8), so the
  // truncation bias per block (-1.6e-6, tools/gram_bias.py) matches the TF32 variant's 256-sample blocks while the
  // fold, during which the tensor pipe idles, comes half as often
  if (swap) {
    // K <= 64: pairs on the MMA's M dimension, components on N (gram_swap_kernel)
    GuArgs gs = g;
    gs.S_per = q.S_per; gs.splits = q.splits; gs.PP = q.PP;
    const int rawB = (8 * q.Kp * 16 + g.zrec + 127) / 128 * 128;
    int nr = (int)((227 * 1024 - sizeof(GsSmem) - 64) / rawB);
    if (nr > GU_NR) nr = GU_NR;
    gs.nr = nr;
    const size_t smem = (size_t)nr * rawB + sizeof(GsSmem) + 64;
    cudaFuncSetAttribute(gram_swap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    gram_swap_kernel<<<(unsigned)(q.ntask * q.splits), GU_THREADS, smem, st>>>(gs, q.Kp, q.ncpl);
    rc = check_launch("gram_swap");
    if (rc) return rc;
    gram_pair_reduce_kernel<1><<<rgrid, 256, 0, st>>>(g.part, q.splits, a.K, g.Kp, q.PP, D1, g.cmax, flag, gram, g.diag, q.Kp);
    rc = check_launch("gram_pair_reduce");
    if (rc) return rc;
  } else {
    rc = gu_launch_main<true>(a, g, pair, st);
    if (rc) return rc;
    // reduce + resolution check; then the TF32 kernel and its reduce, both of which return at once unless the flag is up
    gram_pair_reduce_kernel<1><<<rgrid, 256, 0, st>>>(g.part, g.splits, a.K, g.Kp, g.PP, D1, g.cmax, flag, gram, g.diag);
    rc = check_launch("gram_pair_reduce");
    if (rc) return rc;
  }
  rc = gu_launch_main<false>(a, g, pair, st);
  if (rc) return rc;
  gram_pair_reduce_kernel<2><<<rgrid, 256, 0, st>>>(g.part, g.splits, a.K, g.Kp, g.PP, D1, nullptr, flag, gram, g.diag);
  return check_launch("gram_pair_reduce2");
}


// ---- weighted column sums  C[k][f] = sum_n p[n][k] S[n][f]  (K x N)(N x F), reduction over the SAMPLE axis ---------------
// The responsibility-weighted sums of flattened covariances in MatrixNormalWishart.update(pX, pY, p)
// (transforms/MatrixNormalWishart.py:150-156: SExx += sum_n p E[xx^T]_n; F = p^2 or n^2 columns, thousands).  This is the
// Gram contraction with phi_n[f] = S[n][f] instead of a product of two features, so it runs on gram_umma_kernel's TF32
// variant in "lin" mode: the column blocks take the place of the pair blocks, each CTA's TMA box covers only its block's
// columns (S is read from HBM once), the weights come through the same box / TMEM path, the two-level accumulation and the
// fixed-order fp64 reduce over the sample splits are shared.  TF32 operands (3-term split): S has no scale structure to
// exploit and the call is bound by the bytes of S.
__global__ void wsum_reduce_kernel(const float* __restrict__ part, int splits, int K, int Kp, int PP, int F, float* __restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)K * F) return;
  const int k = (int)(e / F), f = (int)(e % F);
  double acc = 0.0;
  for (int s = 0; s < splits; ++s) acc += (double)part[((size_t)s * Kp + k) * PP + f];
  out[e] = (float)acc;
}

bool wsum_umma_supported(long long N, int K, int F, int lds) {
  return N >= 2048 && K >= 4 && (K % 4 == 0) && F >= 16 && (lds % 4 == 0) && lds >= F;
}
static void wsum_plan(long long N, int K, int F, GuArgs* g) {
  g->d0 = F; g->d1 = 0; g->N = N; g->K = K;
  gu_plan(N, K, 0, gu_num_sms(), g, 0, F);
  g->kcb = K < GU_CB ? K : GU_CB;
  g->FL = gu_fl();
  g->flag = nullptr; g->cmax = nullptr;
}
size_t wsum_umma_workspace_bytes(long long N, int K, int F) {
  GuArgs g{};
  wsum_plan(N, K, F, &g);
  return 512 + gu_al((size_t)g.splits * g.Kp * g.PP * sizeof(float));
}
int launch_wsum_umma(const float* p, const float* S, int lds, long long N, int K, int F, float* out, void* ws, size_t ws_bytes,
                     cudaStream_t st) {
  if (!wsum_umma_supported(N, K, F, lds) || ((size_t)S % 16) || ((size_t)p % 16)) {
    set_error("wsum: unsupported shape N=%lld K=%d F=%d lds=%d (needs N >= 2048, K %% 4 == 0, lds %% 4 == 0, 16-byte aligned bases)",
              N, K, F, lds);
    return VBMP_ERR_UNSUPPORTED;
  }
  const size_t need = wsum_umma_workspace_bytes(N, K, F);
  if (ws_bytes < need) { set_error("wsum: workspace too small (%zu < %zu)", ws_bytes, need); return VBMP_ERR_WORKSPACE; }
  GuArgs g{};
  wsum_plan(N, K, F, &g);
  g.part = (float*)gu_al((size_t)ws);
  GramArgs a{};
  a.z0 = S; a.z1 = nullptr; a.d0 = F; a.d1 = 0; a.N = N; a.GX = 1; a.p = p; a.GP = 1; a.G = 1; a.K = K; a.Dp = 0;
  int rc = gu_launch_main<false>(a, g, gu_pair_mode(K), st, lds);
  if (rc) return rc;
  const long long tot = (long long)K * F;
  wsum_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g.part, g.splits, K, g.Kp, g.PP, F, out);
  return check_launch("wsum_reduce");
}

}  // namespace vbmp
