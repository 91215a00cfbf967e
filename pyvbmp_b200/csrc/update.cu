// K5: replicated natural-parameter updates and KL terms (one CTA per component, fp64 inside).
//   wishart_update : dists/Wishart.py:43-56
//   niw_update     : dists/NormalInverseWishart.py:49-68   (statistics already beta-accumulated by the caller kernel arg)
//   mnw_update     : transforms/MatrixNormalWishart.py:82-141 (no-mask branch)
//   *_kl           : dists/Wishart.py:88-94, dists/NormalInverseWishart.py:99-105, transforms/MatrixNormalWishart.py:206-216
#include "common.cuh"
#include "linalg.cuh"

namespace vbmp {

// Shared tail: given the packed lower triangle of the NEW invU in L (fp64), write invU (fp32, symmetric),
// U = invU^{-1}, logdet(invU).  L is destroyed, Li is scratch.
__device__ inline void wishart_finish(double* L, double* Li, int d, float* invU_out, float* U_out,
                                      float* logdet_out, int* s_info, double* red) {
  const int tid = threadIdx.x, nt = blockDim.x;
  __syncthreads();
  for (int e = tid; e < d * d; e += nt) {
    const int i = e / d, j = e % d;
    invU_out[e] = (float)(j <= i ? L[tri(i, j)] : L[tri(j, i)]);
  }
  const double logdet = chol_packed(L, d, s_info, red);
  tri_inverse_packed(L, Li, d);
  for (int e = tid; e < d * d; e += nt) {
    const int i = e / d, j = e % d;
    if (j <= i) {
      const float u = (float)inv_entry(Li, d, i, j);
      U_out[i * d + j] = u;
      U_out[j * d + i] = u;
    }
  }
  if (tid == 0) *logdet_out = (float)logdet;
}

__global__ void wishart_update_kernel(const float* __restrict__ SExx, const float* __restrict__ N,
                                      const float* __restrict__ invU0, const float* __restrict__ nu0,
                                      const float* __restrict__ invU_old, const float* __restrict__ nu_old,
                                      int d, float lr,
                                      float* __restrict__ invU_new, float* __restrict__ nu_new,
                                      float* __restrict__ U_new, float* __restrict__ logdet_new,
                                      int* __restrict__ info) {
  extern __shared__ double sm[];
  const int c = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int np = d * (d + 1) / 2;
  double* L = sm; double* Li = sm + np; double* red = Li + np;
  __shared__ int s_info;
  if (tid == 0) s_info = 0;
  const size_t o = (size_t)c * d * d;
  const double lrd = lr;
  for (int e = tid; e < d * d; e += nt) {
    const int i = e / d, j = e % d;
    if (j <= i) {
      // symmetrised read of the statistics (they are symmetric up to rounding)
      const double s = 0.5 * ((double)SExx[o + i * d + j] + (double)SExx[o + j * d + i]);
      const double a0 = 0.5 * ((double)invU0[o + i * d + j] + (double)invU0[o + j * d + i]);
      const double ao = 0.5 * ((double)invU_old[o + i * d + j] + (double)invU_old[o + j * d + i]);
      L[tri(i, j)] = lrd * (a0 + s) + (1.0 - lrd) * ao;
    }
  }
  if (tid == 0) nu_new[c] = (float)(lrd * ((double)nu0[c] + (double)N[c]) + (1.0 - lrd) * (double)nu_old[c]);
  wishart_finish(L, Li, d, invU_new + o, U_new + o, logdet_new + c, &s_info, red);
  __syncthreads();
  if (tid == 0 && info) info[c] = s_info;
}

__global__ void niw_update_kernel(const float* __restrict__ SExx, const float* __restrict__ SEx,
                                  const float* __restrict__ N,
                                  const float* __restrict__ lam0, const float* __restrict__ mu0,
                                  const float* __restrict__ invU0, const float* __restrict__ nu0,
                                  const float* __restrict__ lam_old, const float* __restrict__ mu_old,
                                  const float* __restrict__ invU_old, const float* __restrict__ nu_old,
                                  int d, float lr, int fixed_precision,
                                  float* __restrict__ lam_new, float* __restrict__ mu_new,
                                  float* __restrict__ invU_new, float* __restrict__ nu_new,
                                  float* __restrict__ U_new, float* __restrict__ logdet_new,
                                  int* __restrict__ info) {
  extern __shared__ double sm[];
  const int c = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int np = d * (d + 1) / 2;
  double* L = sm; double* Li = sm + np; double* red = Li + np; double* mus = red + 32;  // mus: d
  __shared__ int s_info;
  if (tid == 0) { s_info = 0; if (info) info[c] = 0; }
  const size_t o = (size_t)c * d * d, ov = (size_t)c * d;
  const double lrd = lr, l0 = lam0[c], Nc = N[c];
  const double lam = l0 + Nc;                                   // :61
  for (int i = tid; i < d; i += nt) {
    const double mi = (l0 * (double)mu0[ov + i] + (double)SEx[ov + i]) / lam;   // :62
    mus[i] = mi;
    mu_new[ov + i] = (float)(lrd * mi + (1.0 - lrd) * (double)mu_old[ov + i]);  // :66
  }
  if (tid == 0) lam_new[c] = (float)(lrd * lam + (1.0 - lrd) * (double)lam_old[c]);   // :65
  if (fixed_precision) return;                                  // :67 (uniform across the grid)
  __syncthreads();
  for (int e = tid; e < d * d; e += nt) {
    const int i = e / d, j = e % d;
    if (j <= i) {
      const double s = 0.5 * ((double)SExx[o + i * d + j] + (double)SExx[o + j * d + i]);
      const double S = s + l0 * (double)mu0[ov + i] * (double)mu0[ov + j] - lam * mus[i] * mus[j];   // :63
      const double a0 = 0.5 * ((double)invU0[o + i * d + j] + (double)invU0[o + j * d + i]);
      const double ao = 0.5 * ((double)invU_old[o + i * d + j] + (double)invU_old[o + j * d + i]);
      L[tri(i, j)] = lrd * (a0 + S) + (1.0 - lrd) * ao;          // Wishart.py:53
    }
  }
  if (tid == 0) nu_new[c] = (float)(lrd * ((double)nu0[c] + Nc) + (1.0 - lrd) * (double)nu_old[c]);
  wishart_finish(L, Li, d, invU_new + o, U_new + o, logdet_new + c, &s_info, red);
  __syncthreads();
  if (tid == 0 && info) info[c] = s_info;
}

// MNW update.  SExx (pp x pp), SEyx (n x pp), SEyy (n x n) are the (already beta-accumulated) statistics.
__global__ void mnw_update_kernel(const float* __restrict__ SExx, const float* __restrict__ SEyx,
                                  const float* __restrict__ SEyy, const float* __restrict__ N,
                                  const float* __restrict__ mu0, const float* __restrict__ invV0,
                                  const float* __restrict__ invU0, const float* __restrict__ nu0,
                                  const float* __restrict__ mu_old, const float* __restrict__ invV_old,
                                  const float* __restrict__ invU_old, const float* __restrict__ nu_old,
                                  int n, int pp, float lr, int fixed_precision,
                                  float* __restrict__ mu_new, float* __restrict__ invV_new,
                                  float* __restrict__ V_new, float* __restrict__ logdetV_new,
                                  float* __restrict__ invU_new, float* __restrict__ nu_new,
                                  float* __restrict__ U_new, float* __restrict__ logdetU_new,
                                  int* __restrict__ info) {
  extern __shared__ double sm[];
  const int c = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int nmax = n > pp ? n : pp;
  const int np = nmax * (nmax + 1) / 2;
  double* L = sm; double* Li = sm + np; double* red = Li + np;
  double* B = red + 32;            // n x pp : mu0 invV0 + SEyx, then mu' (solved in place)
  double* B0 = B + n * pp;         // n x pp : copy of the right-hand side (needed for SEyy')
  __shared__ int s_infoV, s_infoU, s_infoV2;
  if (tid == 0) { s_infoV = 0; s_infoU = 0; s_infoV2 = 0; }
  const size_t oV = (size_t)c * pp * pp, oU = (size_t)c * n * n, oM = (size_t)c * n * pp;
  const double lrd = lr;
  // invV' = invV0 + SExx  (:106)
  for (int e = tid; e < pp * pp; e += nt) {
    const int i = e / pp, j = e % pp;
    if (j <= i) L[tri(i, j)] = 0.5 * ((double)invV0[oV + i * pp + j] + (double)invV0[oV + j * pp + i])
                             + 0.5 * ((double)SExx[oV + i * pp + j] + (double)SExx[oV + j * pp + i]);
  }
  // B = mu0 invV0 + SEyx  (:107)
  for (int e = tid; e < n * pp; e += nt) {
    const int r = e / pp, j = e % pp;
    double s = (double)SEyx[oM + e];
    for (int k = 0; k < pp; ++k) s += (double)mu0[oM + r * pp + k] * (double)invV0[oV + k * pp + j];
    B[e] = s; B0[e] = s;
  }
  (void)chol_packed(L, pp, &s_infoV, red);
  // mu'^T = invV'^{-1} B^T  (:108): per row r of B solve L L^T x = b in place
  __syncthreads();
  for (int r = tid; r < n; r += nt) {
    double* b = B + r * pp;
    for (int i = 0; i < pp; ++i) {
      double s = b[i];
      for (int k = 0; k < i; ++k) s -= L[tri(i, k)] * b[k];
      b[i] = s / L[tri(i, i)];
    }
    for (int i = pp - 1; i >= 0; --i) {
      double s = b[i];
      for (int k = i + 1; k < pp; ++k) s -= L[tri(k, i)] * b[k];
      b[i] = s / L[tri(i, i)];
    }
  }
  __syncthreads();
  for (int e = tid; e < n * pp; e += nt)
    mu_new[oM + e] = (float)(lrd * B[e] + (1.0 - lrd) * (double)mu_old[oM + e]);       // :127
  // blended, symmetrised invV (:125-126), its inverse and logdet (:134-135)
  __syncthreads();
  for (int e = tid; e < pp * pp; e += nt) {
    const int i = e / pp, j = e % pp;
    if (j <= i) {
      const double a = 0.5 * ((double)invV0[oV + i * pp + j] + (double)invV0[oV + j * pp + i])
                     + 0.5 * ((double)SExx[oV + i * pp + j] + (double)SExx[oV + j * pp + i]);
      const double ao = 0.5 * ((double)invV_old[oV + i * pp + j] + (double)invV_old[oV + j * pp + i]);
      L[tri(i, j)] = lrd * a + (1.0 - lrd) * ao;
    }
  }
  wishart_finish(L, Li, pp, invV_new + oV, V_new + oV, logdetV_new + c, &s_infoV2, red);
  if (!fixed_precision) {
    // SEyy' = SEyy - mu' invV' mu'^T + mu0 invV0 mu0^T  (:123);  mu' invV' = B0
    __syncthreads();
    for (int e = tid; e < n * n; e += nt) {
      const int i = e / n, j = e % n;
      if (j <= i) {
        double s = 0.5 * ((double)SEyy[oU + i * n + j] + (double)SEyy[oU + j * n + i]);
        double t = 0.0, t2 = 0.0;
        for (int k = 0; k < pp; ++k) { t += B0[i * pp + k] * B[j * pp + k]; t2 += B0[j * pp + k] * B[i * pp + k]; }
        s -= 0.5 * (t + t2);
        double q = 0.0;
        for (int a = 0; a < pp; ++a) {
          double w = 0.0;
          for (int b2 = 0; b2 < pp; ++b2) w += (double)invV0[oV + a * pp + b2] * (double)mu0[oM + j * pp + b2];
          q += (double)mu0[oM + i * pp + a] * w;
        }
        s += q;
        const double a0 = 0.5 * ((double)invU0[oU + i * n + j] + (double)invU0[oU + j * n + i]);
        const double ao = 0.5 * ((double)invU_old[oU + i * n + j] + (double)invU_old[oU + j * n + i]);
        L[tri(i, j)] = lrd * (a0 + s) + (1.0 - lrd) * ao;
      }
    }
    if (tid == 0) nu_new[c] = (float)(lrd * ((double)nu0[c] + (double)N[c]) + (1.0 - lrd) * (double)nu_old[c]);
    wishart_finish(L, Li, n, invU_new + oU, U_new + oU, logdetU_new + c, &s_infoU, red);
  }
  __syncthreads();
  if (tid == 0 && info) info[c] = s_infoV ? s_infoV : (s_infoV2 ? 1000 + s_infoV2 : (s_infoU ? 2000 + s_infoU : 0));
}

// ---------------------------------------------------------------------------------------------------
// expectations / KL
// ---------------------------------------------------------------------------------------------------

// Wishart.ElogdetinvSigma (dists/Wishart.py:82-83) for C matrices of size d
__global__ void wishart_elogdet_kernel(const float* __restrict__ nu, const float* __restrict__ logdet_invU,
                                       int d, float* __restrict__ out) {
  __shared__ double red[32];
  const int c = blockIdx.x;
  const double psi = mv_digamma_block(0.5 * (double)nu[c], d, red);
  if (threadIdx.x == 0) out[c] = (float)(d * M_LN2 - (double)logdet_invU[c] + psi);
}

__device__ inline double wishart_kl_block(const float* invU0, const float* U, double nu0, double nu,
                                          double logdet, double logdet0, int d, double* red) {
  double tr = 0.0;
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) tr += (double)invU0[e] * (double)U[e];
  tr = block_sum(tr, red);
  const double lg0 = mv_lgamma_block(0.5 * nu0, d, red);
  const double lg = mv_lgamma_block(0.5 * nu, d, red);
  const double psi = mv_digamma_block(0.5 * nu, d, red);
  return 0.5 * nu0 * (logdet - logdet0) + 0.5 * nu * tr - 0.5 * nu * d + lg0 - lg + 0.5 * (nu - nu0) * psi;
}

__global__ void wishart_kl_kernel(const float* __restrict__ invU0, const float* __restrict__ U,
                                  const float* __restrict__ nu0, const float* __restrict__ nu,
                                  const float* __restrict__ logdet, const float* __restrict__ logdet0,
                                  int d, float* __restrict__ out) {
  __shared__ double red[32];
  const int c = blockIdx.x;
  const size_t o = (size_t)c * d * d;
  const double kl = wishart_kl_block(invU0 + o, U + o, nu0[c], nu[c], logdet[c], logdet0[c], d, red);
  if (threadIdx.x == 0) out[c] = (float)kl;
}

__global__ void niw_kl_kernel(const float* __restrict__ lam0, const float* __restrict__ lam,
                              const float* __restrict__ mu0, const float* __restrict__ mu,
                              const float* __restrict__ invU0, const float* __restrict__ U,
                              const float* __restrict__ nu0, const float* __restrict__ nu,
                              const float* __restrict__ logdet, const float* __restrict__ logdet0,
                              int d, float* __restrict__ out) {
  __shared__ double red[32];
  const int c = blockIdx.x;
  const size_t o = (size_t)c * d * d, ov = (size_t)c * d;
  double q = 0.0;
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) {
    const int i = e / d, j = e % d;
    q += ((double)mu[ov + i] - (double)mu0[ov + i]) * (double)U[o + e] * ((double)mu[ov + j] - (double)mu0[ov + j]);
  }
  q = block_sum(q, red);
  const double klw = wishart_kl_block(invU0 + o, U + o, nu0[c], nu[c], logdet[c], logdet0[c], d, red);
  if (threadIdx.x == 0) {
    const double l0 = lam0[c], l = lam[c];
    out[c] = (float)(0.5 * d * (l0 / l - 1.0 + log(l / l0)) + 0.5 * l0 * (double)nu[c] * q + klw);
  }
}

__global__ void mnw_kl_kernel(const float* __restrict__ mu0, const float* __restrict__ mu,
                              const float* __restrict__ invV0, const float* __restrict__ V,
                              const float* __restrict__ logdetV, const float* __restrict__ logdetV0,
                              const float* __restrict__ invU0, const float* __restrict__ U,
                              const float* __restrict__ nu0, const float* __restrict__ nu,
                              const float* __restrict__ logdetU, const float* __restrict__ logdetU0,
                              int n, int pp, float* __restrict__ out) {
  extern __shared__ double sm[];
  double* red = sm;                // 32
  double* T = sm + 32;             // n x pp : (nu U) (mu - mu0)
  const int c = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const size_t oV = (size_t)c * pp * pp, oU = (size_t)c * n * n, oM = (size_t)c * n * pp;
  const double nuc = nu[c];
  for (int e = tid; e < n * pp; e += nt) {
    const int i = e / pp, a = e % pp;
    double s = 0.0;
    for (int k = 0; k < n; ++k) s += (double)U[oU + i * n + k] * ((double)mu[oM + k * pp + a] - (double)mu0[oM + k * pp + a]);
    T[e] = nuc * s;
  }
  __syncthreads();
  // tr(invV0 * dm^T T) = sum_{a,b} invV0[a][b] sum_i dm[i][b] T[i][a]
  double q = 0.0, trv = 0.0;
  for (int e = tid; e < pp * pp; e += nt) {
    const int a = e / pp, b = e % pp;
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += ((double)mu[oM + i * pp + b] - (double)mu0[oM + i * pp + b]) * T[i * pp + a];
    q += (double)invV0[oV + e] * s;
    trv += (double)invV0[oV + e] * (double)V[oV + e];
  }
  q = block_sum(q, red);
  trv = block_sum(trv, red);
  const double klw = wishart_kl_block(invU0 + oU, U + oU, nu0[c], nuc, logdetU[c], logdetU0[c], n, red);
  if (tid == 0)
    out[c] = (float)(0.5 * n * ((double)logdetV[c] - (double)logdetV0[c]) - 0.5 * n * pp + 0.5 * n * trv + 0.5 * q + klw);
}

// ---------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------
static size_t packed_smem(int d, int extra_doubles) { return (size_t)(d * (d + 1) + 32 + extra_doubles) * sizeof(double); }

int launch_wishart_update(const float* SExx, const float* N, const float* invU0, const float* nu0,
                          const float* invU_old, const float* nu_old, int C, int d, float lr,
                          float* invU_new, float* nu_new, float* U_new, float* logdet_new, int* info, cudaStream_t st) {
  if (C <= 0) return VBMP_OK;
  if (d < 1 || d > VBMP_MAX_D) { set_error("wishart_update: d=%d out of range", d); return VBMP_ERR_SHAPE; }
  const size_t smem = packed_smem(d, 0);
  cudaFuncSetAttribute(wishart_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  wishart_update_kernel<<<C, 256, smem, st>>>(SExx, N, invU0, nu0, invU_old, nu_old, d, lr, invU_new, nu_new, U_new, logdet_new, info);
  return check_launch("wishart_update");
}

int launch_niw_update(const float* SExx, const float* SEx, const float* N, const float* lam0, const float* mu0,
                      const float* invU0, const float* nu0, const float* lam_old, const float* mu_old,
                      const float* invU_old, const float* nu_old, int C, int d, float lr, int fixed_precision,
                      float* lam_new, float* mu_new, float* invU_new, float* nu_new, float* U_new,
                      float* logdet_new, int* info, cudaStream_t st) {
  if (C <= 0) return VBMP_OK;
  if (d < 1 || d > VBMP_MAX_D) { set_error("niw_update: d=%d out of range", d); return VBMP_ERR_SHAPE; }
  const size_t smem = packed_smem(d, d);
  cudaFuncSetAttribute(niw_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  niw_update_kernel<<<C, 256, smem, st>>>(SExx, SEx, N, lam0, mu0, invU0, nu0, lam_old, mu_old, invU_old, nu_old,
                                          d, lr, fixed_precision, lam_new, mu_new, invU_new, nu_new, U_new, logdet_new, info);
  return check_launch("niw_update");
}

int launch_mnw_update(const float* SExx, const float* SEyx, const float* SEyy, const float* N,
                      const float* mu0, const float* invV0, const float* invU0, const float* nu0,
                      const float* mu_old, const float* invV_old, const float* invU_old, const float* nu_old,
                      int C, int n, int pp, float lr, int fixed_precision,
                      float* mu_new, float* invV_new, float* V_new, float* logdetV_new,
                      float* invU_new, float* nu_new, float* U_new, float* logdetU_new, int* info, cudaStream_t st) {
  if (C <= 0) return VBMP_OK;
  if (n < 1 || pp < 1 || n > VBMP_MAX_D || pp > VBMP_MAX_D) { set_error("mnw_update: n=%d p'=%d out of range", n, pp); return VBMP_ERR_SHAPE; }
  const int nmax = n > pp ? n : pp;
  const size_t smem = packed_smem(nmax, 2 * n * pp);
  if (smem > 227 * 1024) { set_error("mnw_update: n=%d p'=%d needs %zu B of shared memory", n, pp, smem); return VBMP_ERR_UNSUPPORTED; }
  cudaFuncSetAttribute(mnw_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mnw_update_kernel<<<C, 256, smem, st>>>(SExx, SEyx, SEyy, N, mu0, invV0, invU0, nu0, mu_old, invV_old, invU_old, nu_old,
                                          n, pp, lr, fixed_precision, mu_new, invV_new, V_new, logdetV_new,
                                          invU_new, nu_new, U_new, logdetU_new, info);
  return check_launch("mnw_update");
}

int launch_wishart_elogdet(const float* nu, const float* logdet_invU, int C, int d, float* out, cudaStream_t st) {
  if (C <= 0) return VBMP_OK;
  wishart_elogdet_kernel<<<C, 128, 0, st>>>(nu, logdet_invU, d, out);
  return check_launch("wishart_elogdet");
}

int launch_wishart_kl(const float* invU0, const float* U, const float* nu0, const float* nu, const float* logdet,
                      const float* logdet0, int C, int d, float* out, cudaStream_t st) {
  if (C <= 0) return VBMP_OK;
  wishart_kl_kernel<<<C, 256, 0, st>>>(invU0, U, nu0, nu, logdet, logdet0, d, out);
  return check_launch("wishart_kl");
}

int launch_niw_kl(const float* lam0, const float* lam, const float* mu0, const float* mu, const float* invU0,
                  const float* U, const float* nu0, const float* nu, const float* logdet, const float* logdet0,
                  int C, int d, float* out, cudaStream_t st) {
  if (C <= 0) return VBMP_OK;
  niw_kl_kernel<<<C, 256, 0, st>>>(lam0, lam, mu0, mu, invU0, U, nu0, nu, logdet, logdet0, d, out);
  return check_launch("niw_kl");
}

int launch_mnw_kl(const float* mu0, const float* mu, const float* invV0, const float* V, const float* logdetV,
                  const float* logdetV0, const float* invU0, const float* U, const float* nu0, const float* nu,
                  const float* logdetU, const float* logdetU0, int C, int n, int pp, float* out, cudaStream_t st) {
  if (C <= 0) return VBMP_OK;
  const size_t smem = (size_t)(32 + n * pp) * sizeof(double);
  cudaFuncSetAttribute(mnw_kl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mnw_kl_kernel<<<C, 256, smem, st>>>(mu0, mu, invV0, V, logdetV, logdetV0, invU0, U, nu0, nu, logdetU, logdetU0, n, pp, out);
  return check_launch("mnw_kl");
}

}  // namespace vbmp
