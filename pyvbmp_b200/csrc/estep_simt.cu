// K2 (general-shape CUDA-core variant): whitened quadratic forms + responsibility softmax.
//
//   l[n,g,k] = cst[g,k] - 1/2 || W_{g,k}^T z_{n,xg[g]} - m_{g,k} ||^2
//   mode 0: write l                                  (NIW/MNW.Elog_like; HMM obs_logits)
//   mode 1: write p = exp(l - logZ_n), logZ_n, and per-CTA partial NA_k / sum logZ_n
//           (Mixture.update_assignments dists/Mixture.py:38-45, MoLT.update_assignments :34-41)
//
// Register-tiled fp32 GEMM on CUDA cores: a CTA owns TN samples of one theta group g, loops over the K
// components with W_k double-buffered in shared memory through cp.async, each thread owning an
// 8-sample x 8-column accumulator tile.  This is the path for shapes the tcgen05 kernel does not
// take (d not in {32,64,...}, tiny K, G>1); it is hand-written sm_100a code, not a fallback to torch.
#include "common.cuh"

namespace vbmp {


__device__ inline void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ inline void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ inline void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int MODE>
__global__ void __launch_bounds__(256) estep_simt_kernel(EstepArgs a) {
  extern __shared__ __align__(16) float smf[];
  const int tid = threadIdx.x;
  const int Dp = a.Dp, JT = Dp >> 3, SR = 256 / JT, TN = SR * 8, ZS = TN + 4;
  const int D = a.d0 + a.d1;
  float* Zs = smf;                       // [Dp][ZS]  feature-major sample tile
  float* Wb = Zs + (size_t)Dp * ZS;      // [2][Dp*Dp]
  float* mb = Wb + 2 * Dp * Dp;          // [2][Dp]
  float* lz = mb + 2 * Dp;               // [TN]       (mode 1)
  float* red = lz + TN;                  // [256]      (mode 1)
  const int g = blockIdx.y;
  const long long n0 = (long long)blockIdx.x * TN;
  const int xgi = a.xg ? a.xg[g] : 0;
  const int jt = tid % JT, sr = tid / JT;
  const int half_s = TN >> 1, half_j = Dp >> 1;
  // this thread's 8 samples: [4sr,4sr+4) and [TN/2+4sr, ...); 8 columns: [4jt,4jt+4) and [Dp/2+4jt, ...)
  const int s_lo = 4 * sr, s_hi = half_s + 4 * sr, j_lo = 4 * jt, j_hi = half_j + 4 * jt;

  // ---- sample tile -> shared (zero padded in both directions)
  for (int e = tid; e < TN * Dp; e += 256) {
    const int s = e / Dp, i = e % Dp;
    const long long n = n0 + s;
    float v = 0.f;
    if (n < a.N && i < D) {
      const long long row = n * a.GX + xgi;
      v = (i < a.d0) ? a.z0[row * a.d0 + i] : a.z1[row * a.d1 + (i - a.d0)];
    }
    Zs[i * ZS + s] = v;
  }

  const float* Wg = a.W + (size_t)g * a.K * Dp * Dp;
  const float* mg = a.m + (size_t)g * a.K * Dp;
  const float* cg = a.cst + (size_t)g * a.K;
  auto prefetch = [&](int k, int buf) {
    const float* src = Wg + (size_t)k * Dp * Dp;
    float* dst = Wb + buf * Dp * Dp;
    for (int e = tid * 4; e < Dp * Dp; e += 1024) cp_async16(dst + e, src + e);
    if (tid * 4 < Dp) cp_async16(mb + buf * Dp + tid * 4, mg + (size_t)k * Dp + tid * 4);
    cp_async_commit();
  };

  float mx[8], sm_[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) { mx[s] = -INFINITY; sm_[s] = 0.f; }

  prefetch(0, 0);
  for (int k = 0; k < a.K; ++k) {
    if (k + 1 < a.K) { prefetch(k + 1, (k + 1) & 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const float* Wk = Wb + (k & 1) * Dp * Dp;
    const float* mk = mb + (k & 1) * Dp;
    float acc[8][8];
#pragma unroll
    for (int s = 0; s < 8; ++s)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[s][j] = 0.f;
#pragma unroll 2
    for (int i = 0; i < D; ++i) {
      const float4 za = *reinterpret_cast<const float4*>(Zs + i * ZS + s_lo);
      const float4 zb = *reinterpret_cast<const float4*>(Zs + i * ZS + s_hi);
      const float4 wa = *reinterpret_cast<const float4*>(Wk + i * Dp + j_lo);
      const float4 wb = *reinterpret_cast<const float4*>(Wk + i * Dp + j_hi);
      const float z[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
      const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
      for (int s = 0; s < 8; ++s)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[s][j] = fmaf(z[s], w[j], acc[s][j]);
    }
    const float4 ma = *reinterpret_cast<const float4*>(mk + j_lo);
    const float4 mbv = *reinterpret_cast<const float4*>(mk + j_hi);
    const float mm[8] = {ma.x, ma.y, ma.z, ma.w, mbv.x, mbv.y, mbv.z, mbv.w};
    float q[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float r = acc[s][j] - mm[j]; t = fmaf(r, r, t); }
      q[s] = t;
    }
    for (int o = JT >> 1; o > 0; o >>= 1)
#pragma unroll
      for (int s = 0; s < 8; ++s) q[s] += __shfl_xor_sync(0xffffffffu, q[s], o);
    if (jt == 0) {
      const float ck = cg[k];
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const long long n = n0 + (s < 4 ? s_lo + s : s_hi + s - 4);
        if (n < a.N) {
          const float l = ck - 0.5f * q[s];
          a.out[(n * a.G + g) * a.K + k] = l;
          if (MODE == 1) {
            if (l > mx[s]) { sm_[s] = sm_[s] * expf(mx[s] - l) + 1.f; mx[s] = l; }
            else sm_[s] += expf(l - mx[s]);
          }
        }
      }
    }
    __syncthreads();
  }

  if (MODE == 1) {
    if (jt == 0) {
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const int sl = (s < 4 ? s_lo + s : s_hi + s - 4);
        const long long n = n0 + sl;
        const float v = (n < a.N) ? mx[s] + logf(sm_[s]) : 0.f;
        lz[sl] = v;
        if (n < a.N) a.logZn[n * a.G + g] = v;
      }
    }
    __syncthreads();
    const long long rem = a.N - n0;
    const int rows = rem < TN ? (int)rem : TN;
    // normalise this CTA's own (rows x K) slab of logits in place; column sums -> NA partial
    int KT = 1; while (KT < a.K && KT < 256) KT <<= 1;
    const int RT = 256 / KT, tx = tid % KT, ty = tid / KT;
    for (int kb = 0; kb < a.K; kb += KT) {
      const int kk = kb + tx;
      float cs = 0.f;
      if (kk < a.K) {
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

// Fixed-order reduction of the per-CTA partials: NA[g,k] = sum_b NA_part[b,g,k], logZ[g] = sum_b logZ_part[b,g].
__global__ void estep_reduce_kernel(const float* __restrict__ NA_part, const double* __restrict__ logZ_part,
                                    int nb, int GK, int G, float* __restrict__ NA, float* __restrict__ logZ) {
  __shared__ double sred[32][33];
  const int x = threadIdx.x, y = threadIdx.y;
  const int col = blockIdx.x * 32 + x;
  const int ncolNA = GK;
  double acc = 0.0;
  if (col < ncolNA) {
    for (int b = y; b < nb; b += 32) acc += (double)NA_part[(size_t)b * GK + col];
  } else if (col - ncolNA < G) {
    for (int b = y; b < nb; b += 32) acc += logZ_part[(size_t)b * G + (col - ncolNA)];
  }
  sred[y][x] = acc;
  __syncthreads();
  if (y == 0) {
    double t = 0.0;
    for (int i = 0; i < 32; ++i) t += sred[i][x];
    if (col < ncolNA) NA[col] = (float)t;
    else if (col - ncolNA < G) logZ[col - ncolNA] = (float)t;
  }
}

static size_t estep_simt_smem(int Dp) {
  const int JT = Dp / 8, TN = (256 / JT) * 8;
  return (size_t)(Dp * (TN + 4) + 2 * Dp * Dp + 2 * Dp + TN + 256) * sizeof(float);
}

int estep_simt_tile(int Dp) { return (256 / (Dp / 8)) * 8; }

int launch_estep_simt(const EstepArgs& a, int mode, cudaStream_t st) {
  const int TN = estep_simt_tile(a.Dp);
  const size_t smem = estep_simt_smem(a.Dp);
  dim3 grid((unsigned)cdiv(a.N, TN), (unsigned)a.G);
  if (mode == 0) {
    cudaFuncSetAttribute(estep_simt_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    estep_simt_kernel<0><<<grid, 256, smem, st>>>(a);
  } else {
    cudaFuncSetAttribute(estep_simt_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    estep_simt_kernel<1><<<grid, 256, smem, st>>>(a);
  }
  return check_launch("estep_simt");
}

int launch_estep_reduce(const float* NA_part, const double* logZ_part, int nb, int G, int K, float* NA, float* logZ,
                        cudaStream_t st) {
  const int cols = G * K + G;
  estep_reduce_kernel<<<cdiv(cols, 32), dim3(32, 32), 0, st>>>(NA_part, logZ_part, nb, G * K, G, NA, logZ);
  return check_launch("estep_reduce");
}

}  // namespace vbmp
