// K1: per-component parameter preparation for the E-step (one CTA per component, fp64 inside).
//
// Both conjugate families are reduced to the same whitened form
//     l[n,c] = cst[c] - 1/2 || W_c^T z_n - m_c ||^2 ,      W_c upper triangular (Dp x Dp, zero padded)
// so one E-step kernel serves NormalInverseWishart.Elog_like (dists/NormalInverseWishart.py:91-97,
// z = x) and MatrixNormalWishart.Elog_like (transforms/MatrixNormalWishart.py:219-232, z = [x; y]).
// The factors are derived from the live posterior attributes (invU, nu, mu, lambda / invV) every
// call, because callers overwrite them between calls (models/GaussianMixtureModel.py:14-16).
#include "common.cuh"
#include "linalg.cuh"

namespace vbmp {

// NIW: invU = L L^T,  nu (x-mu)^T U (x-mu) = || sqrt(nu) L^{-1} (x-mu) ||^2
//   W[i][j] = sqrt(nu) Linv[j][i] (i <= j),  m = W^T mu,
//   cst = -d/(2 lambda) + 1/2 (d log2 - logdet invU + psi_d(nu/2)) - d/2 log(2 pi) + logprior
//   (dists/NormalInverseWishart.py:93-94,131-132; dists/Wishart.py:82-83; dists/Dirichlet.py:52-53)
__global__ void niw_prep_kernel(const float* __restrict__ invU, const float* __restrict__ mu,
                                const float* __restrict__ nu, const float* __restrict__ lam,
                                const float* __restrict__ logprior, int d, int Dp,
                                float* __restrict__ W, float* __restrict__ m, float* __restrict__ cst,
                                int* __restrict__ info) {
  extern __shared__ double sm[];
  const int c = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int np = d * (d + 1) / 2;
  double* L = sm;
  double* Li = sm + np;
  double* red = Li + np;           // 32 doubles
  __shared__ int s_info;
  if (tid == 0) s_info = 0;
  const float* A = invU + (size_t)c * d * d;
  for (int e = tid; e < d * d; e += nt) {
    const int i = e / d, j = e % d;
    if (j <= i) L[tri(i, j)] = 0.5 * ((double)A[i * d + j] + (double)A[j * d + i]);
  }
  const double logdet = chol_packed(L, d, &s_info, red);
  tri_inverse_packed(L, Li, d);
  const double nuc = (double)nu[c];
  const double sq = sqrt(nuc);
  float* Wc = W + (size_t)c * Dp * Dp;
  for (int e = tid; e < Dp * Dp; e += nt) {
    const int i = e / Dp, j = e % Dp;
    Wc[e] = (i <= j && j < d) ? (float)(sq * Li[tri(j, i)]) : 0.0f;
  }
  const float* muc = mu + (size_t)c * d;
  for (int j = tid; j < Dp; j += nt) {
    double s = 0.0;
    if (j < d) for (int i = 0; i <= j; ++i) s += Li[tri(j, i)] * (double)muc[i];
    m[(size_t)c * Dp + j] = (float)(sq * s);
  }
  const double psi = mv_digamma_block(0.5 * nuc, d, red);
  if (tid == 0) {
    double v = -0.5 * d / (double)lam[c] + 0.5 * (d * M_LN2 - logdet + psi) - 0.5 * d * log(2.0 * M_PI);
    if (logprior) v += (double)logprior[c];
    cst[c] = (float)v;
    if (info) info[c] = s_info;
  }
}

// MNW with z = [x; y] (D = p + n), x~ = [x; 1] when pad_X:
//   invU = C C^T,   A = sqrt(nu) C^{-1}        -> y-outputs  A (y - M x - b)
//   invV (pad coordinate moved first) = F F^T  -> x-outputs  sqrt(n) (F^{-1} [1; x])_{1..p},  constant n (F^{-1})_{00}^2
//   W (upper triangular):  x-out j<p : W[i][j] = sqrt(n) Fi[j'][i'] (i <= j);
//                          y-out j>=p: W[i<p][j] = -(A M)[j-p][i],  W[i>=p][j] = A[j-p][i-p] (i <= j)
//   m: x-out -sqrt(n) Fi[j'][0] (pad) ; y-out (A b) (pad) ; cst per transforms/MatrixNormalWishart.py:229.
// MatrixNormalGamma (transforms/MatrixNormalGamma.py:219-232: the same expression with a DIAGONAL E[invSigma] =
// diag(alpha / beta), dists/DiagonalWishart.py:44-48): tau != nullptr replaces sqrt(nu) C^{-1} by diag(sqrt(tau)) and
// elogdet[c] replaces d log 2 - logdet invU + psi_d(nu / 2) (the Gamma node's sum_i log alpha_i - log beta_i).
__global__ void mnw_prep_kernel(const float* __restrict__ invU, const float* __restrict__ nu,
                                const float* __restrict__ mu, const float* __restrict__ invV,
                                const float* __restrict__ logprior, int n, int pp, int pad, int Dp,
                                float* __restrict__ W, float* __restrict__ m, float* __restrict__ cst,
                                int* __restrict__ info, const float* __restrict__ tau,
                                const float* __restrict__ elogdet) {
  extern __shared__ double sm[];
  const int c = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int p = pp - pad, D = p + n;
  const int npU = n * (n + 1) / 2, npV = pp * (pp + 1) / 2;
  double* LU = sm;             // C, then reused
  double* LiU = LU + npU;      // C^{-1}
  double* LV = LiU + npU;      // F
  double* LiV = LV + npV;      // F^{-1}
  double* red = LiV + npV;     // 32
  __shared__ int s_info;
  if (tid == 0) s_info = 0;
  if (tau == nullptr) {
    const float* Ac = invU + (size_t)c * n * n;
    for (int e = tid; e < n * n; e += nt) {
      const int i = e / n, j = e % n;
      if (j <= i) LU[tri(i, j)] = 0.5 * ((double)Ac[i * n + j] + (double)Ac[j * n + i]);
    }
  }
  const float* Vc = invV + (size_t)c * pp * pp;
  // permuted index: position 0 <- the pad coordinate (pp-1), position t>0 <- coordinate t-1
  for (int e = tid; e < pp * pp; e += nt) {
    const int i = e / pp, j = e % pp;
    if (j <= i) {
      const int oi = pad ? (i == 0 ? pp - 1 : i - 1) : i;
      const int oj = pad ? (j == 0 ? pp - 1 : j - 1) : j;
      LV[tri(i, j)] = 0.5 * ((double)Vc[oi * pp + oj] + (double)Vc[oj * pp + oi]);
    }
  }
  double logdetU = 0.0;
  if (tau == nullptr) {
    logdetU = chol_packed(LU, n, &s_info, red);
    tri_inverse_packed(LU, LiU, n);
  } else {
    for (int e = tid; e < npU; e += nt) LiU[e] = 0.0;
    __syncthreads();
    for (int j = tid; j < n; j += nt) {
      const double t = (double)tau[(size_t)c * n + j];
      if (!(t > 0.0) && s_info == 0) s_info = j + 1;
      LiU[tri(j, j)] = sqrt(t);
    }
    __syncthreads();
  }
  __shared__ int s_info2;
  if (tid == 0) s_info2 = 0;
  (void)chol_packed(LV, pp, &s_info2, red);
  tri_inverse_packed(LV, LiV, pp);
  const double nuc = tau ? 1.0 : (double)nu[c];
  const double sq = sqrt(nuc), sn = sqrt((double)n);
  const float* muc = mu + (size_t)c * n * pp;     // (n, pp) row-major, bias in the last column when pad
  float* Wc = W + (size_t)c * Dp * Dp;
  for (int e = tid; e < Dp * Dp; e += nt) {
    const int i = e / Dp, j = e % Dp;
    double v = 0.0;
    if (i <= j && j < D) {
      if (j < p) {                 // x-output, x-input
        v = sn * LiV[tri(j + pad, i + pad)];
      } else {
        const int jj = j - p;      // y-output row of A
        if (i >= p) {              // y-input
          v = sq * LiU[tri(jj, i - p)];
        } else {                   // x-input: -(A M)[jj][i] = -sqrt(nu) sum_{k<=jj} Ci[jj][k] mu[k][i]
          double s = 0.0;
          for (int k = 0; k <= jj; ++k) s += LiU[tri(jj, k)] * (double)muc[k * pp + i];
          v = -sq * s;
        }
      }
    }
    Wc[e] = (float)v;
  }
  for (int j = tid; j < Dp; j += nt) {
    double v = 0.0;
    if (pad && j < p) v = -sn * LiV[tri(j + 1, 0)];
    else if (pad && j < D) {
      const int jj = j - p;
      double s = 0.0;
      for (int k = 0; k <= jj; ++k) s += LiU[tri(jj, k)] * (double)muc[k * pp + (pp - 1)];
      v = sq * s;
    }
    m[(size_t)c * Dp + j] = (float)v;
  }
  const double psi = tau ? 0.0 : mv_digamma_block(0.5 * nuc, n, red);
  if (tid == 0) {
    double v = 0.5 * (tau ? (double)elogdet[c] : n * M_LN2 - logdetU + psi) - 0.5 * n * log(2.0 * M_PI);
    if (pad) { const double k0 = LiV[0]; v -= 0.5 * n * k0 * k0; }
    if (logprior) v += (double)logprior[c];
    cst[c] = (float)v;
    if (info) info[c] = s_info ? s_info : (s_info2 ? 1000 + s_info2 : 0);
  }
}

int launch_niw_prep(const float* invU, const float* mu, const float* nu, const float* lam, const float* logprior,
                    int C, int d, int Dp, float* W, float* m, float* cst, int* info, cudaStream_t st) {
  if (C <= 0) return VBMP_OK;
  if (d < 1 || d > VBMP_MAX_D || Dp < d || Dp > VBMP_MAX_D) { set_error("niw_prep: d=%d Dp=%d out of range", d, Dp); return VBMP_ERR_SHAPE; }
  const size_t smem = (size_t)(d * (d + 1) + 32) * sizeof(double);
  cudaFuncSetAttribute(niw_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  niw_prep_kernel<<<C, 256, smem, st>>>(invU, mu, nu, lam, logprior, d, Dp, W, m, cst, info);
  return check_launch("niw_prep");
}

int launch_mnw_prep(const float* invU, const float* nu, const float* mu, const float* invV, const float* logprior,
                    int C, int n, int pp, int pad, int Dp, float* W, float* m, float* cst, int* info, cudaStream_t st,
                    const float* tau, const float* elogdet) {
  if (C <= 0) return VBMP_OK;
  const int D = n + pp - pad;
  if (n < 1 || pp - pad < 1 || D > VBMP_MAX_D || Dp < D || Dp > VBMP_MAX_D) { set_error("mnw_prep: n=%d p'=%d Dp=%d out of range", n, pp, Dp); return VBMP_ERR_SHAPE; }
  const size_t smem = (size_t)(n * (n + 1) + pp * (pp + 1) + 32) * sizeof(double);
  cudaFuncSetAttribute(mnw_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if ((tau == nullptr) != (elogdet == nullptr) || (tau == nullptr && (!invU || !nu))) {
    set_error("mnw_prep: give (invU, nu) or (tau, elogdet)"); return VBMP_ERR_SHAPE;
  }
  mnw_prep_kernel<<<C, 256, smem, st>>>(invU, nu, mu, invV, logprior, n, pp, pad, Dp, W, m, cst, info, tau, elogdet);
  return check_launch("mnw_prep");
}

}  // namespace vbmp
