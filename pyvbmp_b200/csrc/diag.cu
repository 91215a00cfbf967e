// K2d: E-step of the diagonal-precision nodes — a streaming CUDA-core kernel.
//
//   l[n,g,k] = cst[g,k] - 1/2 sum_i tau[g,k,i] (x[n,xg[g],i] - mu[g,k,i])^2
//
// NormalGamma.Elog_like (dists/NormalGamma.py:76-86; the line that counts is :83: -1/2 ((X - mu)^2 gamma.mean()).sum(-1)
// + 1/2 gamma.loggeomean().sum(-1), i.e. tau = alpha / beta and cst = 1/2 sum_i (log alpha_i - log beta_i) [+ log prior]),
// and with mode 1 the responsibilities of a GaussianMixtureModel(isotropic=True) (models/GaussianMixtureModel.py:8-11,
// dists/Mixture.py:38-45).  O(N K d) work: 3 flops per (sample, component, feature) against 4 (d + K) bytes per sample,
// so at d = 64, K = 256 it is bound by the fp32 pipe, not by HBM (2e11 flop vs 5.4 GB per pass); the subtraction is done
// before the square, as in the reference (the expanded form x^2 tau - 2 x tau mu + tau mu^2 would cancel for |mu| >> sigma).
//
// A CTA owns 128 samples of one theta group; components stream through shared memory in tiles of 128 (mu and tau,
// feature-major, cp.async); each of the 256 threads owns an 8-sample x 8-component register tile — components
// 4 tk .. 4 tk + 3 and 64 + 4 tk .. 64 + 4 tk + 3, so that the 16 lanes of a half-warp read consecutive 16-byte words
// (conflict free; 8 consecutive components per lane put lanes l and l + 4 on the same banks: measured 15.6 ms per 4 Mi
// rows at d = 64, K = 256, shared-memory bound) — and walks the features with packed fp32x2 arithmetic: 6 shared 16-byte
// loads per 64 (sample, component) pairs.  Up to d = 64 one parameter tile is resident per CTA and two CTAs share an SM
// (the loads of one hide behind the arithmetic of the other).
// Mode 1 reuses the CUDA-core E-step's epilogue: online logsumexp per sample, logits normalised in place from L2,
// per-CTA partial NA / sum logZ_n reduced in a fixed order (estep_reduce_kernel).
#include <cuda_fp16.h>
#include "common.cuh"

namespace vbmp {

__device__ inline void dg_cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ inline void dg_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ inline void dg_cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

constexpr int DG_TN = 128;     // samples per CTA
constexpr int DG_KT = 128;     // components per shared-memory tile

struct DiagArgs {
  const float* x; long long N; int GX; const int* xg;     // x (N, GX, d)
  const float* mut; const float* taut;                    // (G, d, Kp) FEATURE-major, component-minor, Kp = K rounded up to 128
  const float* cst;                                       // (G, K)
  int G, K, Kp, d, nbuf;                                  // nbuf: parameter tiles in flight (2 while they fit, d <= 64)
  float* out; float* logZn; float* NA_part; double* logZ_part;
  // mode 1, optional: also emit the responsibilities as K3's pre-split fp16 weight images (gram_umma.cu, GuArgs::rp: record
  // (component block, 16-sample block) = [hi | lo][8-sample chunk][component][8 fp16] of r 2^14), as the tcgen05 E-step's
  // normaliser does — the M-step of a NormalGamma mixture then skips its own pass over p (G = 1, K <= 256)
  unsigned char* rpack; long long nsb;
};

// (G, K, d) -> (G, d, Kp), zero padded (tau = 0 makes a padding component's quadratic form vanish)
__global__ void diag_transpose_kernel(const float* __restrict__ mu, const float* __restrict__ tau, int G, int K, int Kp, int d,
                                      float* __restrict__ mut, float* __restrict__ taut) {
  const long long tot = (long long)G * d * Kp;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e % Kp), i = (int)((e / Kp) % d), g = (int)(e / ((long long)Kp * d));
    const bool ok = k < K;
    mut[e] = ok ? mu[((size_t)g * K + k) * d + i] : 0.f;
    taut[e] = ok ? tau[((size_t)g * K + k) * d + i] : 0.f;
  }
}

template <int MODE>
__global__ void __launch_bounds__(256, 2) diag_estep_kernel(DiagArgs a) {
  extern __shared__ __align__(16) float smf[];
  const int tid = threadIdx.x;
  const int d = a.d;
  float* Xs = smf;                                   // [d][DG_TN + 4]   feature-major sample tile
  float* Pb = Xs + (size_t)d * (DG_TN + 4);          // [nbuf][2][d][DG_KT] (buffer, mu | tau)
  float* lz = Pb + (size_t)2 * a.nbuf * d * DG_KT;   // [DG_TN]  (mode 1)
  float* red = lz + DG_TN;                           // [256]    (mode 1)
  const int g = blockIdx.y;
  const long long n0 = (long long)blockIdx.x * DG_TN;
  const int xgi = a.xg ? a.xg[g] : 0;
  const int tk = tid & 15, ts = tid >> 4;            // 16 component groups x 16 sample groups, 8 x 8 register tiles
  const int XS = DG_TN + 4;

  for (int e = tid; e < DG_TN * d; e += 256) {       // coalesced rows -> feature-major tile
    const int s = e / d, i = e % d;
    const long long n = n0 + s;
    Xs[i * XS + s] = (n < a.N) ? a.x[((size_t)n * a.GX + xgi) * d + i] : 0.f;
  }
  const float* mg = a.mut + (size_t)g * d * a.Kp;
  const float* tg = a.taut + (size_t)g * d * a.Kp;
  auto prefetch = [&](int kt, int buf) {
    float* dm = Pb + (size_t)buf * 2 * d * DG_KT;
    float* dt = dm + (size_t)d * DG_KT;
    for (int e = tid * 4; e < d * DG_KT; e += 1024) {
      const int i = e / DG_KT, k = e % DG_KT;
      dg_cp_async16(dm + e, mg + (size_t)i * a.Kp + kt * DG_KT + k);
      dg_cp_async16(dt + e, tg + (size_t)i * a.Kp + kt * DG_KT + k);
    }
    dg_cp_async_commit();
  };

  float mx[8], sm_[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) { mx[s] = -INFINITY; sm_[s] = 0.f; }
  const int nkt = a.Kp / DG_KT;
  const bool dbl = a.nbuf == 2;
  if (dbl) prefetch(0, 0);
  for (int kt = 0; kt < nkt; ++kt) {
    if (!dbl) { prefetch(kt, 0); dg_cp_async_wait<0>(); }
    else if (kt + 1 < nkt) { prefetch(kt + 1, (kt + 1) & 1); dg_cp_async_wait<1>(); }
    else { dg_cp_async_wait<0>(); }
    __syncthreads();
    const float* Mk = Pb + (size_t)(dbl ? (kt & 1) : 0) * 2 * d * DG_KT + tk * 4;
    const float* Tk = Mk + (size_t)d * DG_KT;
    const float* Xr = Xs + ts * 8;
    float2 acc[8][4];                                 // [sample][component pair]
#pragma unroll
    for (int s = 0; s < 8; ++s)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[s][c] = make_float2(0.f, 0.f);
#pragma unroll 2
    for (int i = 0; i < d; ++i) {
      const float4 xa = *reinterpret_cast<const float4*>(Xr + i * XS);
      const float4 xb = *reinterpret_cast<const float4*>(Xr + i * XS + 4);
      const float4 ma = *reinterpret_cast<const float4*>(Mk + i * DG_KT);
      const float4 mb = *reinterpret_cast<const float4*>(Mk + i * DG_KT + 64);
      const float4 ta = *reinterpret_cast<const float4*>(Tk + i * DG_KT);
      const float4 tb = *reinterpret_cast<const float4*>(Tk + i * DG_KT + 64);
      const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
      const float2 nm[4] = {make_float2(-ma.x, -ma.y), make_float2(-ma.z, -ma.w), make_float2(-mb.x, -mb.y), make_float2(-mb.z, -mb.w)};
      const float2 tt[4] = {make_float2(ta.x, ta.y), make_float2(ta.z, ta.w), make_float2(tb.x, tb.y), make_float2(tb.z, tb.w)};
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const float2 xs2 = make_float2(xv[s], xv[s]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float2 df = __fadd2_rn(xs2, nm[c]);               // x - mu first, as the reference does
          acc[s][c] = __ffma2_rn(__fmul2_rn(df, tt[c]), df, acc[s][c]);
        }
      }
    }
    // logits of this tile: 8 samples x 8 components per thread; rows are completed across the 16 tk lanes by the epilogue
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const long long n = n0 + ts * 8 + s;
      const int k0 = kt * DG_KT + tk * 4;                 // this thread's components: k0 .. k0 + 3 and k0 + 64 .. k0 + 67
      float l[8];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int ka = k0 + (c >> 1) * 64 + (c & 1) * 2;
        l[2 * c] = (ka < a.K) ? a.cst[(size_t)g * a.K + ka] - 0.5f * acc[s][c].x : -INFINITY;
        l[2 * c + 1] = (ka + 1 < a.K) ? a.cst[(size_t)g * a.K + ka + 1] - 0.5f * acc[s][c].y : -INFINITY;
      }
      if (n < a.N) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int kb = k0 + 64 * h;
          float* o = a.out + ((size_t)n * a.G + g) * a.K + kb;
          if ((a.K & 3) == 0 && kb + 4 <= a.K) {
            *reinterpret_cast<float4*>(o) = make_float4(l[4 * h], l[4 * h + 1], l[4 * h + 2], l[4 * h + 3]);
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (kb + c < a.K) o[c] = l[4 * h + c];
          }
        }
      }
      if (MODE == 1) {
        float m = l[0];
#pragma unroll
        for (int c = 1; c < 8; ++c) m = fmaxf(m, l[c]);
        if (m > -INFINITY) {
          const float nm_ = fmaxf(mx[s], m);
          float t = 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) t += __expf(l[c] - nm_);
          sm_[s] = sm_[s] * __expf(mx[s] - nm_) + t;
          mx[s] = nm_;
        }
      }
    }
    __syncthreads();
  }

  if (MODE == 1) {
    // combine the 16 component groups of a sample row (lanes tk = 0..15 of a half-warp), then the estep_simt epilogue
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      float m = mx[s], v = sm_[s];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), v2 = __shfl_xor_sync(0xffffffffu, v, o);
        const float nm_ = fmaxf(m, m2);
        v = (m > -INFINITY ? v * __expf(m - nm_) : 0.f) + (m2 > -INFINITY ? v2 * __expf(m2 - nm_) : 0.f);
        m = nm_;
      }
      if (tk == 0) {
        const int sl = ts * 8 + s;
        const long long n = n0 + sl;
        const float lzv = (n < a.N) ? m + logf(v) : 0.f;
        lz[sl] = lzv;
        if (n < a.N) a.logZn[n * a.G + g] = lzv;
      }
    }
    __syncthreads();
    const long long rem = a.N - n0;
    const int rows = rem < DG_TN ? (int)rem : DG_TN;
    int KT = 1; while (KT < a.K && KT < 256) KT <<= 1;
    const int RT = 256 / KT, tx = tid % KT, ty = tid / KT;
    const long long npad = (a.N + 31) / 32 * 32;           // the weight images cover whole 32-sample chunks
    for (int kb = 0; kb < a.K; kb += KT) {
      const int kk = kb + tx;
      float cs = 0.f;
      if (kk < a.K && a.rpack != nullptr) {
        // thread = (component, group of 8 consecutive rows): p as before, plus one 16-byte store each to the hi and lo image
        for (int rg = ty; rg < DG_TN / 8 && n0 + rg * 8 < npad; rg += RT) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float pv[2];
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              const int r = rg * 8 + 2 * u + v;
              pv[v] = 0.f;
              if (r < rows) {
                const size_t ad = ((size_t)(n0 + r) * a.G + g) * a.K + kk;
                pv[v] = expf(a.out[ad] - lz[r]);
                a.out[ad] = pv[v];
                cs += pv[v];
              }
            }
            const float x0 = pv[0] * 16384.f, x1 = pv[1] * 16384.f;
            const __half2 ah = __floats2half2_rn(x0, x1);
            const float2 af = __half22float2(ah);
            const __half2 bh = __floats2half2_rn(x0 - af.x, x1 - af.y);
            hi[u] = *reinterpret_cast<const uint32_t*>(&ah);
            lo[u] = *reinterpret_cast<const uint32_t*>(&bh);
          }
          const long long ch = (n0 + rg * 8) >> 3;
          unsigned char* rec = a.rpack + ((size_t)(kk >> 7) * a.nsb + (size_t)(ch >> 1)) * 8192 + (size_t)(ch & 1) * 2048 +
                               (size_t)(kk & 127) * 16;
          *reinterpret_cast<uint4*>(rec) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(rec + 4096) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      } else if (kk < a.K) {
        for (int r = ty; r < rows; r += RT) {
          const size_t ad = ((size_t)(n0 + r) * a.G + g) * a.K + kk;
          const float p = expf(a.out[ad] - lz[r]);
          a.out[ad] = p;
          cs += p;
        }
      }
      red[ty * KT + tx] = cs;
      __syncthreads();
      if (ty == 0 && kk < a.K) {
        float t = 0.f;
        for (int y = 0; y < RT; ++y) t += red[y * KT + tx];
        a.NA_part[((size_t)blockIdx.x * a.G + g) * a.K + kk] = t;
      }
      __syncthreads();
    }
    double v = 0.0;
    for (int r = tid; r < rows; r += 256) v += (double)lz[r];
    __shared__ double dred[32];
    v = block_sum(v, dred);
    if (tid == 0) a.logZ_part[(size_t)blockIdx.x * a.G + g] = v;
  }
}

int launch_estep_reduce(const float*, const double*, int nb, int G, int K, float* NA, float* logZ, cudaStream_t);

static size_t dg_al(size_t x) { return (x + 255) / 256 * 256; }
static int dg_kp(int K) { return (K + DG_KT - 1) / DG_KT * DG_KT; }

size_t diag_estep_workspace_bytes(long long N, int G, int K, int d, int mode) {
  const size_t nb = (size_t)cdiv(N > 0 ? N : 1, DG_TN);
  size_t b = 256 + 2 * dg_al((size_t)G * d * dg_kp(K) * sizeof(float));
  if (mode == 1) b += dg_al(nb * G * K * sizeof(float)) + dg_al(nb * G * sizeof(double));
  return b;
}

int launch_diag_estep(const float* x, int d, long long N, int GX, const int* xg, const float* mu, const float* tau,
                      const float* cst, int G, int K, int mode, float* out, float* logZn, float* NA, float* logZ,
                      void* ws, size_t ws_bytes, cudaStream_t st, unsigned char* rpack) {
  if (d < 1 || d > VBMP_MAX_D || G < 1 || K < 1 || GX < 1 || N < 0 || (mode != 0 && mode != 1)) {
    set_error("diag_estep: bad shape N=%lld GX=%d G=%d K=%d d=%d mode=%d (d <= %d)", N, GX, G, K, d, mode, VBMP_MAX_D);
    return VBMP_ERR_SHAPE;
  }
  if (mode == 1 && (!NA || !logZ || (N > 0 && !logZn))) { set_error("diag_estep: mode 1 needs logZn, NA, logZ"); return VBMP_ERR_SHAPE; }
  if (N == 0) {
    if (mode == 1) { cudaMemsetAsync(NA, 0, sizeof(float) * G * K, st); cudaMemsetAsync(logZ, 0, sizeof(float) * G, st); }
    return VBMP_OK;
  }
  const size_t need = diag_estep_workspace_bytes(N, G, K, d, mode);
  if (ws_bytes < need) { set_error("diag_estep: workspace too small (%zu < %zu)", ws_bytes, need); return VBMP_ERR_WORKSPACE; }
  const int Kp = dg_kp(K);
  char* p = (char*)dg_al((size_t)ws);
  float* mut = (float*)p; p += dg_al((size_t)G * d * Kp * sizeof(float));
  float* taut = (float*)p; p += dg_al((size_t)G * d * Kp * sizeof(float));
  const int nb = cdiv(N, DG_TN);
  float* NA_part = (float*)p; p += dg_al((size_t)nb * G * K * sizeof(float));
  double* logZ_part = (double*)p;
  const long long tot = (long long)G * d * Kp;
  diag_transpose_kernel<<<(unsigned)((tot + 255) / 256 < 4096 ? (tot + 255) / 256 : 4096), 256, 0, st>>>(mu, tau, G, K, Kp, d, mut, taut);
  int rc = check_launch("diag_transpose");
  if (rc) return rc;
  const int nbuf = 1;                      // one resident parameter tile: two CTAs per SM up to d = 64 (see the header)
  DiagArgs a{x, N, GX, xg, mut, taut, cst, G, K, Kp, d, nbuf, out, logZn, mode == 1 ? NA_part : nullptr,
             mode == 1 ? logZ_part : nullptr, (mode == 1 && G == 1) ? rpack : nullptr, (N + 31) / 32 * 2};
  const size_t smem = ((size_t)d * (DG_TN + 4) + (size_t)2 * nbuf * d * DG_KT + DG_TN + 256) * sizeof(float);
  if (smem > 227 * 1024) { set_error("diag_estep: d=%d needs %zu bytes of shared memory", d, smem); return VBMP_ERR_UNSUPPORTED; }
  dim3 grid((unsigned)nb, (unsigned)G);
  if (mode == 0) {
    cudaFuncSetAttribute(diag_estep_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    diag_estep_kernel<0><<<grid, 256, smem, st>>>(a);
  } else {
    cudaFuncSetAttribute(diag_estep_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    diag_estep_kernel<1><<<grid, 256, smem, st>>>(a);
  }
  rc = check_launch("diag_estep");
  if (rc) return rc;
  if (mode == 1) rc = launch_estep_reduce(NA_part, logZ_part, nb, G, K, NA, logZ, st);
  return rc;
}

}  // namespace vbmp
