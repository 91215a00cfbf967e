// tools/umma_probe.cu — standalone B200 probe for the tcgen05 building blocks in pyvbmp_b200/csrc/umma.cuh.
//   (1) correctness of kind::tf32 MMAs with K-major / SWIZZLE_NONE shared-memory operands (SS) and with the
//       A operand resident in TMEM (TS), for several N and several K-steps (descriptor advance), operands
//       brought in by cp.async.bulk;
//   (2) issue-rate microbenchmarks: cycles per MMA for SS / TS and N in {64,...,256} on one SM and on all SMs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include "../pyvbmp_b200/csrc/umma.cuh"

using namespace umma;


#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

// A: [128][K] row-major, B: [N][K] row-major (both already TF32-representable), D: [128][N]
__global__ void __launch_bounds__(128, 1) probe_kernel(const float* __restrict__ A, const float* __restrict__ Bp,
                                                        float* __restrict__ D, int N, int K, int ts_mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_b, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* As = reinterpret_cast<float*>(smem);                     // K-major core-matrix layout, LBO = 128*16
  float* Bs = reinterpret_cast<float*>(smem + 128 * K * 4);       // Bp is already packed in that layout: LBO = N*16
  if (tid == 0) { mbar_init(&bar_b, 1); mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  // B via bulk copy (packed layout prepared on the host)
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar_b, (uint32_t)(N * K * 4));
    bulk_g2s(Bs, Bp, (uint32_t)(N * K * 4), &bar_b);
  }
  // A: row per thread
  if (!ts_mode) {
    for (int k = 0; k < K; ++k) As[(k >> 2) * (128 * 4) + tid * 4 + (k & 3)] = A[tid * K + k];
    fence_proxy_async();
  } else {
    for (int k0 = 0; k0 < K; k0 += 8) {
      uint32_t r[8];
      for (int j = 0; j < 8; ++j) r[j] = __float_as_uint(A[tid * K + k0 + j]);
      tmem_st8(tm + ((uint32_t)(warp * 32) << 16) + 256 + k0, r);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    mbar_wait(&bar_b, 0);
    const uint32_t idesc = idesc_tf32(128, N);
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint64_t bd = smem_desc(smem_u32(Bs) + ks * 2 * (N * 16), N * 16, 128);
      if (!ts_mode) {
        const uint64_t ad = smem_desc(smem_u32(As) + ks * 2 * (128 * 16), 128 * 16, 128);
        mma_tf32_ss(tm, ad, bd, idesc, ks > 0);
      } else {
        mma_tf32_ts(tm, tm + 256 + ks * 8, bd, idesc, ks > 0);
      }
    }
    mma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_wait_ld();
    for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}

// Issue-rate benchmark: every CTA issues iters*8 MMAs (K=64 worth of K-steps, operands fixed in smem / TMEM).
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, int ts_mode, int alt_d, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  float* As = reinterpret_cast<float*>(smem);                  // 128 x 64
  float* Bs = reinterpret_cast<float*>(smem + 128 * 64 * 4);   // N x 64
  for (int e = tid; e < 128 * 64 + N * 64; e += 128) As[e] = (float)((e * 37) % 7 - 3);
  fence_proxy_async();
  if (tid == 0) { mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  if (ts_mode) {
    uint32_t r[16];
    for (int j = 0; j < 16; ++j) r[j] = __float_as_uint((float)(j - 8));
    for (int c = 0; c < 64; c += 16) tmem_st16(tm + ((uint32_t)(warp * 32) << 16) + 448 + c, r);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = idesc_tf32(128, N);
    const uint64_t bd0 = smem_desc(smem_u32(Bs), N * 16, 128);
    const uint64_t ad0 = smem_desc(smem_u32(As), 128 * 16, 128);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tm + ((alt_d && (it & 1)) ? (uint32_t)N : 0u);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t bd = bd0 + (uint64_t)((ks * 2 * N * 16) >> 4);
        if (!ts_mode) mma_tf32_ss(d, ad0 + (uint64_t)((ks * 2 * 128 * 16) >> 4), bd, idesc, 1);
        else mma_tf32_ts(d, tm + 448 + ks * 8, bd, idesc, 1);
      }
    }
    mma_commit(&bar_mma);
    mbar_wait(&bar_mma, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}


// nchain independent accumulators, interleaved at MMA granularity (is the ~94-cycle floor a dependent-accumulate latency?)
template <bool TS, int NCHAIN>
__global__ void __launch_bounds__(128, 1) chain_kernel(int N, int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  float* As = reinterpret_cast<float*>(smem);
  float* Bs = reinterpret_cast<float*>(smem + 128 * 64 * 4);
  for (int e = tid; e < 128 * 64 + N * 64; e += 128) As[e] = (float)((e * 37) % 7 - 3);
  fence_proxy_async();
  if (tid == 0) { mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  if (warp == 1) {
    const uint32_t idesc = idesc_tf32(128, N);
    const uint64_t bd0 = smem_desc(smem_u32(Bs), N * 16, 128);
    const uint64_t ad0 = smem_desc(smem_u32(As), 128 * 16, 128);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t bd = bd0 + (uint64_t)((ks * 2 * N * 16) >> 4);
#pragma unroll
        for (int c = 0; c < NCHAIN; ++c) {
          const uint32_t d = tm + (uint32_t)(c * N);
          if (elect_one()) {
            if (!TS) mma_tf32_ss(d, ad0 + (uint64_t)((ks * 2 * 128 * 16) >> 4), bd, idesc, 1);
            else mma_tf32_ts(d, tm + 448 + ks * 8, bd, idesc, 1);
          }
        }
      }
    }
    if (elect_one()) mma_commit(&bar_mma);
    __syncwarp();
    mbar_wait(&bar_mma, 0);
    const long long t1 = clock64();
    if ((tid & 31) == 0) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}

template <bool TS, int NCHAIN>
static void run_chain_t(int N, int grid) {
  const int iters = 1000;
  long long* dc; CK(cudaMalloc(&dc, grid * sizeof(long long)));
  const size_t smem = (size_t)(128 + N) * 64 * 4;
  CK(cudaFuncSetAttribute(chain_kernel<TS, NCHAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  chain_kernel<TS, NCHAIN><<<grid, 128, smem>>>(N, 100, dc);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  chain_kernel<TS, NCHAIN><<<grid, 128, smem>>>(N, iters, dc);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> c(grid);
  CK(cudaMemcpy(c.data(), dc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  long long mx = 0; for (auto v : c) if (v > mx) mx = v;
  const double flops = 2.0 * 128 * N * 8 * iters * 8.0 * NCHAIN * grid;
  printf("chain %s N=%3d nchain=%d grid=%3d : %.1f cycles/MMA (ideal %.0f)  %.1f TFLOP/s\n", TS ? "TS" : "SS", N, NCHAIN, grid,
         (double)mx / (iters * 8.0 * NCHAIN), N / 2.0, flops / ms / 1e9);
  cudaFree(dc);
}
static void run_chain(int N, int ts_mode, int nchain, int grid) {
  if (ts_mode) { if (nchain == 1) run_chain_t<true, 1>(N, grid); else if (nchain == 2) run_chain_t<true, 2>(N, grid); else if (nchain == 3) run_chain_t<true, 3>(N, grid); else run_chain_t<true, 4>(N, grid); }
  else { if (nchain == 1) run_chain_t<false, 1>(N, grid); else if (nchain == 2) run_chain_t<false, 2>(N, grid); else if (nchain == 3) run_chain_t<false, 3>(N, grid); else run_chain_t<false, 4>(N, grid); }
}
static float tf32_exact(int v) { return (float)v; }

static int run_case(int N, int K, int ts_mode) {
  std::vector<float> A(128 * K), B(N * K), Bp(N * K), D(128 * N), R(128 * N);
  for (int i = 0; i < 128 * K; ++i) A[i] = tf32_exact((rand() % 17) - 8);
  for (int i = 0; i < N * K; ++i) B[i] = tf32_exact((rand() % 13) - 6);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) Bp[(k >> 2) * (N * 4) + n * 4 + (k & 3)] = B[n * K + k];
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k];
      R[m * N + n] = (float)s;
    }
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, Bp.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bp.data(), Bp.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, D.size() * 4));
  const size_t smem = (size_t)(128 + N) * K * 4;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem>>>(dA, dB, dD, N, K, ts_mode);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0; double maxerr = 0;
  for (int i = 0; i < 128 * N; ++i) { double e = fabs((double)D[i] - R[i]); if (e > maxerr) maxerr = e; if (e > 1e-3) ++bad; }
  printf("probe %s N=%3d K=%3d : %s (mismatches %d / %d, max err %.3g)  D[0..3]=%g %g %g %g ref %g %g %g %g\n",
         ts_mode ? "TS" : "SS", N, K, bad ? "FAIL" : "ok", bad, 128 * N, maxerr, D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3]);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return bad != 0;
}

static void run_rate(int N, int ts_mode, int alt_d, int grid) {
  const int iters = 2000;
  long long* dc; CK(cudaMalloc(&dc, grid * sizeof(long long)));
  const size_t smem = (size_t)(128 + N) * 64 * 4;
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rate_kernel<<<grid, 128, smem>>>(N, 200, ts_mode, alt_d, dc);   // warm
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  rate_kernel<<<grid, 128, smem>>>(N, iters, ts_mode, alt_d, dc);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> c(grid);
  CK(cudaMemcpy(c.data(), dc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  long long mx = 0; for (auto v : c) if (v > mx) mx = v;
  const double per = (double)mx / (iters * 8.0);
  const double flops = 2.0 * 128 * N * 8 * iters * 8.0 * grid;
  printf("rate %s N=%3d alt_d=%d grid=%3d : %.1f cycles/MMA (ideal %.0f), %.3f ms -> %.1f TFLOP/s (tf32 issued)\n",
         ts_mode ? "TS" : "SS", N, alt_d, grid, per, N / 2.0, ms, flops / ms / 1e9);
  cudaFree(dc);
}


// Does kind::tf32 truncate or round fp32 operands whose low 13 mantissa bits are set?
static void run_trunc_test() {
  const int N = 64, K = 8;
  std::vector<float> A(128 * K), B(N * K), Bp(N * K), D(128 * N);
  for (auto& v : A) v = 1.0f + (float)(rand() % 8191) / 8388608.0f * 1.0f + (float)(rand() % 1000) / 1000.0f;   // low bits set
  for (int i = 0; i < N * K; ++i) B[i] = (i % K == (i / K) % K) ? 1.0f : 0.0f;                                  // picks A[m][n % 8]
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) Bp[(k >> 2) * (N * 4) + n * 4 + (k & 3)] = B[n * K + k];
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, Bp.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bp.data(), Bp.size() * 4, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)(128 + N) * K * 4;
  for (int ts = 0; ts < 2; ++ts) {
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, N, K, ts);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    int n_trunc = 0, n_rna = 0, n_exact = 0, n_other = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 8; ++n) {
        const float a = A[m * K + n];
        uint32_t b; memcpy(&b, &a, 4);
        uint32_t t = b & 0xffffe000u, r = (b + 0x1000u) & 0xffffe000u;
        float ft, fr; memcpy(&ft, &t, 4); memcpy(&fr, &r, 4);
        const float d = D[m * N + n];
        if (d == a) ++n_exact; else if (d == ft && ft != fr) ++n_trunc; else if (d == fr && ft != fr) ++n_rna; else if (d == ft) ++n_trunc; else ++n_other;
      }
    printf("operand low-bit handling (%s): exact %d, truncated %d, rounded %d, other %d\n", ts ? "A in TMEM" : "A in smem", n_exact, n_trunc, n_rna, n_other);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

// Replays the E-step MMA sequence of estep_umma.cu (2 halves x 8 K-steps x 3 split terms per 64-column group)
// with static operands: TRI = triangular column-suffix skipping, per-group cycles.
template <bool TRI, int VARIANT>
__global__ void __launch_bounds__(128, 1) eseq_kernel(int groups, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  float* Bs = reinterpret_cast<float*>(smem);
  for (int e = tid; e < 9 * 5120; e += 128) Bs[e] = (float)((e * 37) % 7 - 3);      // 8 stages x 40 KB >= hi+lo images
  fence_proxy_async();
  if (tid == 0) { mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  {
    uint32_t r[16];
    for (int j = 0; j < 16; ++j) r[j] = __float_as_uint((float)(j - 8));
    for (int c = 0; c < 256; c += 16) tmem_st16(tm + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      const uint32_t sbase = smem_u32(smem) + (g & 7) * 20480;
      const int buf = g & 1;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t dcol = tm + 256 + (2 * h + buf) * 64;
        const uint32_t a_hi = tm + h * 128, a_lo = a_hi + 64;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const int n0 = TRI ? 16 * (ks / 2) : 0, nn = 64 - n0;
            int off = 0;
            for (int i = 0; i < ks; ++i) off += (64 - (TRI ? 16 * (i / 2) : 0)) * 32;
            const uint32_t idesc = idesc_tf32(128, nn);
            const uint64_t b_hi = smem_desc(sbase + off, nn * 16, 128);
            const uint64_t b_lo = smem_desc(sbase + 10240 + off, nn * 16, 128);
            if (VARIANT == 0) {
              mma_tf32_ts(dcol + n0, a_lo + ks * 8, b_hi, idesc, ks > 0);
              mma_tf32_ts(dcol + n0, a_hi + ks * 8, b_lo, idesc, 1);
              mma_tf32_ts(dcol + n0, a_hi + ks * 8, b_hi, idesc, 1);
            } else if (VARIANT == 1) {     // same A for consecutive MMAs where possible
              mma_tf32_ts(dcol + n0, a_hi + ks * 8, b_lo, idesc, ks > 0);
              mma_tf32_ts(dcol + n0, a_hi + ks * 8, b_hi, idesc, 1);
              mma_tf32_ts(dcol + n0, a_lo + ks * 8, b_hi, idesc, 1);
            } else {                       // D always written at column offset 0 (alignment test; wrong maths)
              mma_tf32_ts(dcol, a_lo + ks * 8, b_hi, idesc, ks > 0);
              mma_tf32_ts(dcol, a_hi + ks * 8, b_lo, idesc, 1);
              mma_tf32_ts(dcol, a_hi + ks * 8, b_hi, idesc, 1);
            }
          }
        }
        __syncwarp();
      }
    }
    if (elect_one()) mma_commit(&bar_mma);
    __syncwarp();
    mbar_wait(&bar_mma, 0);
    const long long t1 = clock64();
    if ((tid & 31) == 0) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}

template <bool TRI, int VARIANT>
static void run_eseq() {
  const int groups = 4000, grid = 148;
  long long* dc; CK(cudaMalloc(&dc, grid * sizeof(long long)));
  const size_t smem = 8 * 40960;
  CK(cudaFuncSetAttribute(eseq_kernel<TRI, VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * 20480 + 20480)));
  eseq_kernel<TRI, VARIANT><<<grid, 128, 8 * 20480 + 20480>>>(100, dc);
  CK(cudaDeviceSynchronize());
  eseq_kernel<TRI, VARIANT><<<grid, 128, 8 * 20480 + 20480>>>(groups, dc);
  CK(cudaDeviceSynchronize());
  std::vector<long long> c(grid);
  CK(cudaMemcpy(c.data(), dc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  long long mx = 0; for (auto v : c) if (v > mx) mx = v;
  printf("eseq tri=%d variant=%d : %.1f cycles per group (48 MMAs) -> %.1f per MMA\n", (int)TRI, VARIANT, (double)mx / groups, (double)mx / groups / 48.0);
  (void)smem;
  cudaFree(dc);
}

// ---- cta_group::2: a CTA pair computes D[256 x N] = A[256 x K] B[N x K]^T; CTA r holds rows r*128.. of A (smem or TMEM),
//      rows r*N/2.. of B in its shared memory, and receives rows r*128.. of D in its TMEM.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mma2_tf32_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_kernel(const float* __restrict__ A, const float* __restrict__ Bp, float* __restrict__ D, int N, int K, int ts_mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_ready, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int NH = N / 2;
  float* As = reinterpret_cast<float*>(smem);
  float* Bs = reinterpret_cast<float*>(smem + 128 * K * 4);
  if (tid == 0) { mbar_init(&bar_ready, 256); mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  // this CTA's half of B (host-packed per half: [chunk][NH rows][4]) and its 128 rows of A
  for (int e = tid; e < NH * K; e += 128) Bs[e] = Bp[(size_t)rank * NH * K + e];
  const float* Ar = A + (size_t)rank * 128 * K;
  if (!ts_mode) {
    for (int k = 0; k < K; ++k) As[(k >> 2) * (128 * 4) + tid * 4 + (k & 3)] = Ar[tid * K + k];
  } else {
    for (int k0 = 0; k0 < K; k0 += 8) {
      uint32_t r[8];
      for (int j = 0; j < 8; ++j) r[j] = __float_as_uint(Ar[tid * K + k0 + j]);
      tmem_st8(tm + ((uint32_t)(warp * 32) << 16) + 256 + k0, r);
    }
    tmem_wait_st();
  }
  fence_proxy_async();
  tc_fence_before();
  // everyone (both CTAs) arrives on the LEADER's ready barrier
  mbar_arrive_remote(mapa_u32(&bar_ready, 0));
  if (rank == 0 && warp == 1) {
    mbar_wait(&bar_ready, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t idesc = idesc_tf32(256, N);
      for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t bd = smem_desc(smem_u32(Bs) + ks * 2 * (NH * 16), NH * 16, 128);
        if (!ts_mode) mma2_tf32_ss(tm, smem_desc(smem_u32(As) + ks * 2 * (128 * 16), 128 * 16, 128), bd, idesc, ks > 0);
        else mma2_tf32_ts(tm, tm + 256 + ks * 8, bd, idesc, ks > 0);
      }
      mma2_commit_mc(&bar_mma, 3);
    }
    __syncwarp();
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_wait_ld();
    for (int j = 0; j < 16; ++j) D[((size_t)rank * 128 + tid) * N + c0 + j] = v[j];
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512));
}

static int run_pair_case(int N, int K, int ts_mode) {
  const int NH = N / 2;
  std::vector<float> A(256 * K), B(N * K), Bp(N * K), D(256 * N), R(256 * N);
  for (auto& v : A) v = (float)((rand() % 17) - 8);
  for (auto& v : B) v = (float)((rand() % 13) - 6);
  for (int h = 0; h < 2; ++h)
    for (int n = 0; n < NH; ++n)
      for (int k = 0; k < K; ++k) Bp[(size_t)h * NH * K + (k >> 2) * (NH * 4) + n * 4 + (k & 3)] = B[(h * NH + n) * K + k];
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k]; R[m * N + n] = (float)s; }
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, Bp.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bp.data(), Bp.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, D.size() * 4));
  const size_t smem = (size_t)(128 + NH) * K * 4;
  CK(cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pair_kernel<<<2, 128, smem>>>(dA, dB, dD, N, K, ts_mode);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int i = 0; i < 256 * N; ++i) if (fabs((double)D[i] - R[i]) > 1e-3) ++bad;
  printf("pair (cta_group::2) %s N=%3d K=%3d : %s (mismatches %d / %d)  D[0]=%g ref %g  D[128*N]=%g ref %g  D[last]=%g ref %g\n", ts_mode ? "TS" : "SS", N, K,
         bad ? "FAIL" : "ok", bad, 256 * N, D[0], R[0], D[128 * N], R[128 * N], D[256 * N - 1], R[256 * N - 1]);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return bad != 0;
}

// Replays the CURRENT E-step MMA sequence (128-column interleaved groups: N = 128 - 16 ks on the column suffix, 3 split
// terms, A in TMEM, two halves) with static operands.  MODE 0: tri-skip, 1: dense N = 128, 2: tri-skip rounded up to N % 32 == 0,
// 3: tri-skip, one term only (8 MMAs per half)
template <int MODE>
__global__ void __launch_bounds__(128, 1) eseq2_kernel(int groups, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  float* Bs = reinterpret_cast<float*>(smem);
  for (int e = tid; e < 5 * 10240; e += 128) Bs[e] = (float)((e * 37) % 7 - 3);      // 5 stages x 40 KB
  fence_proxy_async();
  if (tid == 0) { mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  {
    uint32_t r[16];
    for (int j = 0; j < 16; ++j) r[j] = __float_as_uint((float)(j - 8));
    for (int c = 0; c < 256; c += 16) tmem_st16(tm + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      const uint32_t sbase = smem_u32(smem) + (g % 5) * 40960;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t dcol = tm + 256 + h * 128;
        const uint32_t a_hi = tm + h * 128, a_lo = a_hi + 64;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            int n0 = 16 * ks, nn = 128 - n0;
            if (MODE == 1) { n0 = 0; nn = 128; }
            if (MODE == 2) { nn = (nn + 31) / 32 * 32; n0 = 128 - nn; }
            int off = 0;
            for (int i = 0; i < ks; ++i) off += (128 - 16 * i) * 32;
            const uint32_t idesc = idesc_tf32(128, nn);
            const uint64_t b_hi = smem_desc(sbase + off, nn * 16, 128);
            const uint64_t b_lo = smem_desc(sbase + 18432 + off, nn * 16, 128);
            if (MODE == 4 && ks == 0) {
              mma_tf32_ss(dcol, smem_desc(sbase + 36864, 128 * 16, 128), smem_desc(sbase + 18432, 128 * 16, 128), idesc_tf32(128, 128), 0);
              mma_tf32_ts(dcol + n0, a_lo + ks * 8, b_hi, idesc, 1);
            } else
            mma_tf32_ts(dcol + n0, a_lo + ks * 8, b_hi, idesc, ks > 0);
            if (MODE != 3) {
              mma_tf32_ts(dcol + n0, a_hi + ks * 8, b_lo, idesc, 1);
              mma_tf32_ts(dcol + n0, a_hi + ks * 8, b_hi, idesc, 1);
            }
          }
        }
        __syncwarp();
      }
    }
    if (elect_one()) mma_commit(&bar_mma);
    __syncwarp();
    mbar_wait(&bar_mma, 0);
    const long long t1 = clock64();
    if ((tid & 31) == 0) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}

template <int MODE>
static void run_eseq2() {
  const int groups = 3000, grid = 148;
  long long* dc; CK(cudaMalloc(&dc, grid * sizeof(long long)));
  const int smem = 5 * 40960;
  CK(cudaFuncSetAttribute(eseq2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  eseq2_kernel<MODE><<<grid, 128, smem>>>(100, dc);
  CK(cudaDeviceSynchronize());
  eseq2_kernel<MODE><<<grid, 128, smem>>>(groups, dc);
  CK(cudaDeviceSynchronize());
  std::vector<long long> c(grid);
  CK(cudaMemcpy(c.data(), dc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  long long mx = 0; for (auto v : c) if (v > mx) mx = v;
  printf("eseq2 mode=%d : %.1f cycles per group (2 components x 256 rows)\n", MODE, (double)mx / groups);
  cudaFree(dc);
}

// ---- kind::f16 (fp16 operands, fp32 accumulate), A in TMEM packed two fp16 per 32-bit column, B K-major in smem:
//      checks the packing order and the K = 16 step.
#include <cuda_fp16.h>
// A: [128][K] fp32 values (exactly representable in fp16), Bp: fp16 packed K-major image [K/8 chunks][N rows][8], D: [128][N]
__global__ void __launch_bounds__(128, 1) f16_kernel(const float* __restrict__ A, const __half* __restrict__ Bp, float* __restrict__ D, int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  __half* Bs = reinterpret_cast<__half*>(smem);
  for (int e = tid; e < N * K; e += 128) Bs[e] = Bp[e];
  fence_proxy_async();
  if (tid == 0) { mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  for (int k0 = 0; k0 < K; k0 += 16) {                        // 16 fp16 = 8 columns; column c = {k = 2c (low), 2c+1 (high)}
    uint32_t r[8];
    for (int j = 0; j < 8; ++j) {
      const __half2 h = __floats2half2_rn(A[tid * K + k0 + 2 * j], A[tid * K + k0 + 2 * j + 1]);
      r[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    tmem_st8(tm + ((uint32_t)(warp * 32) << 16) + 256 + k0 / 2, r);
  }
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = idesc_f16(128, N);
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint64_t bd = smem_desc(smem_u32(Bs) + ks * 2 * (N * 16), N * 16, 128);
        mma_f16_ts(tm, tm + 256 + ks * 8, bd, idesc, ks > 0);
      }
      mma_commit(&bar_mma);
    }
    __syncwarp();
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_wait_ld();
    for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}
static int run_f16_case(int N, int K) {
  std::vector<float> A(128 * K), B(N * K), D(128 * N), R(128 * N);
  std::vector<__half> Bp(N * K);
  for (auto& v : A) v = (float)((rand() % 17) - 8);
  for (auto& v : B) v = (float)((rand() % 13) - 6);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) Bp[(k >> 3) * (N * 8) + n * 8 + (k & 7)] = __float2half(B[n * K + k]);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k]; R[m * N + n] = (float)s; }
  float *dA, *dD; __half* dB;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, Bp.size() * 2)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bp.data(), Bp.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, D.size() * 4));
  f16_kernel<<<1, 128, N * K * 2>>>(dA, dB, dD, N, K);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int i = 0; i < 128 * N; ++i) if (fabs((double)D[i] - R[i]) > 1e-3) ++bad;
  printf("f16 TS N=%3d K=%3d : %s (mismatches %d / %d)  D[0..2]=%g %g %g ref %g %g %g\n", N, K, bad ? "FAIL" : "ok", bad, 128 * N, D[0], D[1], D[2], R[0], R[1], R[2]);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return bad != 0;
}

// Mixed sequence: per half one SS MMA (the -m fold), 8 TF32 TS hi*hi steps (N = 128 - 16 ks) and 2 x 4 fp16 TS correction
// steps (K = 16, N = 128 - 32 ks16) into the same accumulator.  MODE 0: mixed, 1: the fp16 steps only, 2: TF32 steps only
template <int MODE>
__global__ void __launch_bounds__(128, 1) eseq3_kernel(int groups, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t* Bs = reinterpret_cast<uint32_t*>(smem);
  for (int e = tid; e < 5 * 10240; e += 128) Bs[e] = 0;
  fence_proxy_async();
  if (tid == 0) { mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  {
    uint32_t r[16];
    for (int j = 0; j < 16; ++j) r[j] = 0;
    for (int c = 0; c < 256; c += 16) tmem_st16(tm + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      const uint32_t sbase = smem_u32(smem) + (g % 5) * 40960;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t dcol = tm + 256 + h * 128;
        const uint32_t a_hi = tm + h * 128, a_h16 = a_hi + 64, a_l16 = a_hi + 96;
        if (elect_one()) {
          mma_tf32_ss(dcol, smem_desc(sbase + 36864, 128 * 16, 128), smem_desc(sbase + 18432, 128 * 16, 128), idesc_tf32(128, 128), 0);
          if (MODE != 1) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const int n0 = 16 * ks, nn = 128 - n0;
              int off = 0;
              for (int i = 0; i < ks; ++i) off += (128 - 16 * i) * 32;
              mma_tf32_ts(dcol + n0, a_hi + ks * 8, smem_desc(sbase + off, nn * 16, 128), idesc_tf32(128, nn), 1);
            }
          }
          if (MODE != 2) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const int n0 = 32 * ks, nn = 128 - n0;
              int off = 0;
              for (int i = 0; i < ks; ++i) off += (128 - 32 * i) * 32;
              mma_f16_ts(dcol + n0, a_l16 + ks * 8, smem_desc(sbase + 18432 + off, nn * 16, 128), idesc_f16(128, nn), 1);
              mma_f16_ts(dcol + n0, a_h16 + ks * 8, smem_desc(sbase + 18432 + 10240 + off, nn * 16, 128), idesc_f16(128, nn), 1);
            }
          }
        }
        __syncwarp();
      }
    }
    if (elect_one()) mma_commit(&bar_mma);
    __syncwarp();
    mbar_wait(&bar_mma, 0);
    const long long t1 = clock64();
    if ((tid & 31) == 0) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}
template <int MODE>
static void run_eseq3() {
  const int groups = 3000, grid = 148;
  long long* dc; CK(cudaMalloc(&dc, grid * sizeof(long long)));
  const int smem = 5 * 40960;
  CK(cudaFuncSetAttribute(eseq3_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  eseq3_kernel<MODE><<<grid, 128, smem>>>(100, dc);
  CK(cudaDeviceSynchronize());
  eseq3_kernel<MODE><<<grid, 128, smem>>>(groups, dc);
  CK(cudaDeviceSynchronize());
  std::vector<long long> c(grid);
  CK(cudaMemcpy(c.data(), dc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  long long mx = 0; for (auto v : c) if (v > mx) mx = v;
  printf("eseq3 mode=%d : %.1f cycles per group (2 components x 256 rows)\n", MODE, (double)mx / groups);
  cudaFree(dc);
}

// ---- TMEM read throughput: NW warps each re-read their 32 lanes x 128 columns (4 x tcgen05.ld.32x32b.x32) in a loop.
__global__ void __launch_bounds__(512, 1) ldtm_kernel(int iters, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    float y[128];
    tmem_ld32(tm, y); tmem_ld32(tm + 32, y + 32); tmem_ld32(tm + 64, y + 64); tmem_ld32(tm + 96, y + 96);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 128; j += 32) acc += y[j];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + tid] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base_s);
}
static void run_ldtm(int nwarps) {
  const int iters = 2000, grid = 148;
  long long* dc; float* ds;
  CK(cudaMalloc(&dc, grid * sizeof(long long))); CK(cudaMalloc(&ds, grid * 512 * sizeof(float)));
  ldtm_kernel<<<grid, nwarps * 32>>>(10, dc, ds);
  CK(cudaDeviceSynchronize());
  ldtm_kernel<<<grid, nwarps * 32>>>(iters, dc, ds);
  CK(cudaDeviceSynchronize());
  std::vector<long long> c(grid);
  CK(cudaMemcpy(c.data(), dc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  long long mx = 0; for (auto v : c) if (v > mx) mx = v;
  printf("ldtm %2d warps: %.1f B/cycle/SM (%.0f cycles per 128-col pass)\n", nwarps, (double)nwarps * 32 * 128 * 4 * iters / mx, (double)mx / iters);
  cudaFree(dc); cudaFree(ds);
}

int main(int argc, char** argv) {
  srand(1);
  int fails = 0;
  if (argc > 1 && atoi(argv[1]) == 9) { for (int w : {1, 4, 8, 16}) run_ldtm(w); return 0; }
  if (argc > 1 && atoi(argv[1]) == 8) { run_eseq2<4>(); run_eseq3<0>(); run_eseq3<1>(); run_eseq3<2>(); return 0; }
  if (argc > 1 && atoi(argv[1]) == 7) { int f = 0; for (int N : {32, 96, 128}) for (int K : {16, 64}) f += run_f16_case(N, K); printf("f16 probe: %d failing\n", f); return f; }
  if (argc > 1 && atoi(argv[1]) == 6) { run_eseq2<0>(); run_eseq2<1>(); run_eseq2<2>(); run_eseq2<3>(); run_eseq2<4>(); return 0; }
  if (argc > 1 && atoi(argv[1]) == 5) { int f = 0; for (int ts = 0; ts < 2; ++ts) for (int N : {64, 128, 224, 256}) for (int K : {8, 32}) f += run_pair_case(N, K, ts); printf("pair probe: %d failing\n", f); return f; }
  if (argc > 1 && atoi(argv[1]) == 4) { run_eseq<false, 0>(); run_eseq<true, 0>(); run_eseq<true, 1>(); run_eseq<true, 2>(); run_eseq<false, 1>(); return 0; }
  if (argc > 1 && atoi(argv[1]) == 3) { run_trunc_test(); return 0; }
  if (argc > 1 && atoi(argv[1]) == 2) {
    for (int ts = 0; ts < 2; ++ts)
      for (int N : {16, 32, 48, 64, 96, 128, 192})
        for (int nc : {1, 2, 3, 4}) if (nc * N <= 448) run_chain(N, ts, nc, 148);
    return 0;
  }
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {64, 96, 192, 256})
      for (int K : {8, 32, 64}) fails += run_case(N, K, ts);
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {64, 96, 128, 192, 256}) { run_rate(N, ts, 0, 1); }
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {64, 96, 128, 192, 256}) { run_rate(N, ts, 0, 148); run_rate(N, ts, 1, 148); }
  printf("probe done: %d failing cases\n", fails);
  return fails ? 1 : 0;
}
