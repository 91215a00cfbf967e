/* vbmp_b200.h — C ABI of libvbmp_b200.so: the B200 (sm_100a) replacement for pyVBMP's conjugate
 * VB-EM hot path (NIW / MNW E-step + M-step).  Plain pointers and sizes only; every pointer is a
 * DEVICE pointer to contiguous row-major fp32 unless noted; the caller owns every buffer and passes
 * the CUDA stream (cudaStream_t as void*).  All functions return 0 on success and a non-zero
 * VBMP_ERR_* code otherwise; vbmp_last_error() gives the message (thread-local).
 *
 * "C" below is the number of flattened components (all batch dims x extra event dims of the
 * reference object); "G" x "K" = C splits them into theta groups (replica / extra-event dims) and
 * the mixture axis.  Each entry cites the reference method it replaces (paths relative to the
 * pyVBMP tree).
 */
#ifndef VBMP_B200_H
#define VBMP_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define VBMP_ABI_VERSION 2

int vbmp_version(void);
const char* vbmp_last_error(void);
/* kernels this library has launched in this process so far (every launch site counts itself; bench.py's gpu_launches) */
unsigned long long vbmp_launch_count(void);

/* ---- K1: parameter preparation -------------------------------------------------------------------
 * Reduce the posterior to the whitened form  l[n,c] = cst[c] - 1/2 ||W_c^T z_n - m_c||^2 with W_c
 * upper triangular, zero padded to Dp x Dp (Dp in {8,16,32,64,128}, Dp >= feature dim).
 * info[c] != 0 reports a non-SPD matrix (1-based failing column).  logprior may be NULL.
 *
 * vbmp_niw_prep: NormalInverseWishart.Elog_like constants — dists/NormalInverseWishart.py:91-97,
 *   122-123, 131-132; Wishart.EinvSigma / ElogdetinvSigma — dists/Wishart.py:76-77, 82-83.
 * vbmp_mnw_prep: MatrixNormalWishart.Elog_like constants — transforms/MatrixNormalWishart.py:219-232,
 *   419-420, 437-438.  Feature order z = [x (p); y (n)];  pp = p + pad_X.                           */
int vbmp_niw_prep(const float* invU, const float* mu, const float* nu, const float* lambda_mu,
                  const float* logprior, int C, int d, int Dp,
                  float* W, float* m, float* cst, int* info, void* stream);
int vbmp_mnw_prep(const float* invU, const float* nu, const float* mu, const float* invV,
                  const float* logprior, int C, int n, int pp, int pad_X, int Dp,
                  float* W, float* m, float* cst, int* info, void* stream);

/* vbmp_mnw_prep_ex: vbmp_mnw_prep that also serves MatrixNormalGamma.Elog_like (transforms/MatrixNormalGamma.py:219-232),
 *   whose E[invSigma] is the diagonal of a Gamma node (dists/DiagonalWishart.py:44-48): pass tau (C, n) = alpha / beta and
 *   elogdet (C) = sum_i log alpha_i - log beta_i with invU = nu = NULL; with tau = elogdet = NULL it is vbmp_mnw_prep.   */
int vbmp_mnw_prep_ex(const float* invU, const float* nu, const float* mu, const float* invV, const float* logprior,
                     const float* tau, const float* elogdet, int C, int n, int pp, int pad_X, int Dp,
                     float* W, float* m, float* cst, int* info, void* stream);

/* ---- K2: E-step -------------------------------------------------------------------------------------
 * z = [z0 | z1] per sample (z1 may be NULL with d1 = 0); z_i is (N, GX, d_i) contiguous; xg[G] maps a
 * theta group to its data column (NULL -> 0).  out is (N, G, K).
 *   mode 0: out = logits                    NIW/MNW.Elog_like; HMM.obs_logits (models/HMM.py:113-117)
 *   mode 1: out = exp(l - logZ_n), logZn (N,G), NA (G,K) = sum_n out, logZ (G) = sum_n logZn
 *           Mixture.update_assignments (dists/Mixture.py:38-45),
 *           MixtureofLinearTransforms.update_assignments (transforms/MixtureofLinearTransforms.py:34-41)
 * flags bit 0: force the CUDA-core kernel even where the tcgen05 kernel applies (testing).          */
size_t vbmp_estep_workspace_bytes(long long N, int G, int K, int Dp, int mode);
int vbmp_estep(const float* z0, int d0, const float* z1, int d1, long long N, int GX, const int* xg,
               const float* W, const float* m, const float* cst, int G, int K, int Dp, int mode, int flags,
               float* out, float* logZn, float* NA, float* logZ,
               void* workspace, size_t workspace_bytes, void* stream);

/* ---- K3: M-step sufficient statistics ----------------------------------------------------------------
 * gram[g,k] = sum_n p[n,pg[g],k] [z;1][z;1]^T, (D+1)x(D+1) row-major, D = d0 + d1; p is (N, GP, K) or
 * NULL for unit weights.  Blocks = SExx / SEx / N of NormalInverseWishart.raw_update
 * (dists/NormalInverseWishart.py:70-86) and SExx / SEyx / SEyy / SEx / SEy / N of
 * MatrixNormalWishart.raw_update (transforms/MatrixNormalWishart.py:174-204) with z = [x; y].        */
size_t vbmp_gram_workspace_bytes(long long N, int G, int K, int d0, int d1, int Dp);
int vbmp_gram(const float* z0, int d0, const float* z1, int d1, long long N, int GX, const int* xg,
              const float* p, int GP, const int* pg, int G, int K, int Dp, int flags,
              float* gram, void* workspace, size_t workspace_bytes, void* stream);

/* ---- K2 -> K3 hand-over (optional) ---------------------------------------------------------------------
 * Mixture.update runs update_assignments (K2 mode 1) and then raw_update (K3) on the SAME responsibilities
 * (dists/Mixture.py:38-45 -> :47-53).  vbmp_estep_rpack is vbmp_estep that, where the tcgen05 fp16 kernels apply
 * (mode 1, K <= 256), ALSO writes the responsibilities pre-split into the fp16 operand images K3 consumes
 * (vbmp_rpack_bytes(N, K) bytes, caller-owned), saving K3's own pass over p; *packed reports whether it did.
 * vbmp_gram_rpack is vbmp_gram given that buffer (it must hold the images of exactly the p passed; NULL = vbmp_gram). */
size_t vbmp_rpack_bytes(long long N, int K);
int vbmp_estep_rpack(const float* z0, int d0, const float* z1, int d1, long long N, int GX, const int* xg,
                     const float* W, const float* m, const float* cst, int G, int K, int Dp, int mode, int flags,
                     float* out, float* logZn, float* NA, float* logZ,
                     void* workspace, size_t workspace_bytes, void* stream,
                     void* rpack, size_t rpack_bytes, int* packed);
int vbmp_gram_rpack(const float* z0, int d0, const float* z1, int d1, long long N, int GX, const int* xg,
                    const float* p, int GP, const int* pg, int G, int K, int Dp, int flags,
                    float* gram, void* workspace, size_t workspace_bytes, void* stream, const void* rpack);

/* ---- K3 sample image (optional) ---------------------------------------------------------------------------
 * The rows of an EM run are the same every iteration (Mixture.update passes the same X to every update_assignments /
 * update_parms, dists/Mixture.py:54-62; MixtureofLinearTransforms.raw_update likewise, :50-61), so the layout K3's
 * tcgen05 fp16 kernel reads them in — column maxima -> exact power-of-two feature scales, then [z0 | z1 | 1] scaled and
 * transposed per 32-sample chunk — can be made once per data set instead of once per call.  vbmp_gram_zpack writes that
 * image (vbmp_zpack_bytes(N, d0, d1) bytes, caller-owned; *packed = 0 when the kernels that take this shape do not use
 * one); vbmp_gram_ex is vbmp_gram given the weight images (rpack, see above) and / or the sample image (zpack), either
 * may be NULL; its workspace (vbmp_gram_ex_workspace_bytes) omits the images that are handed in.  The caller must pass
 * images made from exactly the p / z0 / z1 of the call.                                                              */
size_t vbmp_zpack_bytes(long long N, int d0, int d1);
int vbmp_gram_zpack(const float* z0, int d0, const float* z1, int d1, long long N, int K, int Dp,
                    void* zpack, size_t zpack_bytes, int* packed, void* stream);
size_t vbmp_gram_ex_workspace_bytes(long long N, int G, int K, int d0, int d1, int Dp, int has_rpack, int has_zpack);
int vbmp_gram_ex(const float* z0, int d0, const float* z1, int d1, long long N, int GX, const int* xg,
                 const float* p, int GP, const int* pg, int G, int K, int Dp, int flags,
                 float* gram, void* workspace, size_t workspace_bytes, void* stream,
                 const void* rpack, const void* zpack);

/* ---- diagonal-precision nodes ------------------------------------------------------------------
 * vbmp_diag_estep: l[n,g,k] = cst[g,k] - 1/2 sum_i tau[g,k,i] (x[n,xg[g],i] - mu[g,k,i])^2 — NormalGamma.Elog_like
 *   (dists/NormalGamma.py:76-86) with tau = gamma.mean() = alpha / beta (dists/Gamma.py:93-94) and cst = 1/2 sum_i
 *   gamma.loggeomean()_i (:102-103) [+ the mixture's log prior]; mode 1 adds the responsibility softmax of
 *   Mixture.update_assignments (dists/Mixture.py:38-45) exactly as vbmp_estep does — GaussianMixtureModel(isotropic=True),
 *   models/GaussianMixtureModel.py:8-11.  x (N, GX, d), mu / tau (G, K, d), cst (G, K), d <= 128; out (N, G, K).
 * The weighted statistics of NormalGamma.raw_update (:58-73: sum r x, sum r x^2, sum r) are vbmp_gram / vbmp_gram_ex with
 *   flags bit 1 set ("diagonal statistics only"): the result is the usual (G, K, D+1, D+1) array with the diagonal, the
 *   last row / column and the corner filled and — from the tensor-core kernel — zeros elsewhere.
 * MatrixNormalGamma (transforms/MatrixNormalGamma.py:87-245) runs on vbmp_mnw_prep_ex + vbmp_estep + vbmp_gram +
 *   vbmp_mnw_update (its invV / mu part) — see vbmp_mnw_prep_ex below.                                              */
size_t vbmp_diag_estep_workspace_bytes(long long N, int G, int K, int d, int mode);
int vbmp_diag_estep(const float* x, int d, long long N, int GX, const int* xg, const float* mu, const float* tau,
                    const float* cst, int G, int K, int mode, float* out, float* logZn, float* NA, float* logZ,
                    void* workspace, size_t workspace_bytes, void* stream);
/* ... and with the K2 -> K3 hand-over of vbmp_estep_rpack (mode 1, G = 1, K <= 256): the weight images for vbmp_gram_ex */
int vbmp_diag_estep_rpack(const float* x, int d, long long N, int GX, const int* xg, const float* mu, const float* tau,
                          const float* cst, int G, int K, int mode, float* out, float* logZn, float* NA, float* logZ,
                          void* workspace, size_t workspace_bytes, void* stream,
                          void* rpack, size_t rpack_bytes, int* packed);

/* ---- K5: natural-parameter updates (replicated; statistics are post-beta) -----------------------------
 * Wishart.ss_update — dists/Wishart.py:43-56.                                                          */
int vbmp_wishart_update(const float* SExx, const float* N, const float* invU_0, const float* nu_0,
                        const float* invU_old, const float* nu_old, int C, int d, float lr,
                        float* invU, float* nu, float* U, float* logdet_invU, int* info, void* stream);
/* NormalInverseWishart.ss_update — dists/NormalInverseWishart.py:49-68 (fixed_precision skips the Wishart). */
int vbmp_niw_update(const float* SExx, const float* SEx, const float* N,
                    const float* lambda_0, const float* mu_0, const float* invU_0, const float* nu_0,
                    const float* lambda_old, const float* mu_old, const float* invU_old, const float* nu_old,
                    int C, int d, float lr, int fixed_precision,
                    float* lambda_mu, float* mu, float* invU, float* nu, float* U, float* logdet_invU,
                    int* info, void* stream);
/* MatrixNormalWishart.ss_update, no-mask branch — transforms/MatrixNormalWishart.py:105-108, 122-135.   */
int vbmp_mnw_update(const float* SExx, const float* SEyx, const float* SEyy, const float* N,
                    const float* mu_0, const float* invV_0, const float* invU_0, const float* nu_0,
                    const float* mu_old, const float* invV_old, const float* invU_old, const float* nu_old,
                    int C, int n, int pp, float lr, int fixed_precision,
                    float* mu, float* invV, float* V, float* logdetinvV,
                    float* invU, float* nu, float* U, float* logdet_invU, int* info, void* stream);

/* ---- expectations / KL (per component, length C) ------------------------------------------------------
 * Wishart.ElogdetinvSigma — dists/Wishart.py:82-83;  Wishart.KLqprior — :88-94;
 * NormalInverseWishart.KLqprior — dists/NormalInverseWishart.py:99-105;
 * MatrixNormalWishart.KLqprior — transforms/MatrixNormalWishart.py:206-216.                             */
int vbmp_wishart_elogdet(const float* nu, const float* logdet_invU, int C, int d, float* out, void* stream);
int vbmp_wishart_kl(const float* invU_0, const float* U, const float* nu_0, const float* nu,
                    const float* logdet_invU, const float* logdet_invU_0, int C, int d, float* out, void* stream);
int vbmp_niw_kl(const float* lambda_0, const float* lambda_mu, const float* mu_0, const float* mu,
                const float* invU_0, const float* U, const float* nu_0, const float* nu,
                const float* logdet_invU, const float* logdet_invU_0, int C, int d, float* out, void* stream);
int vbmp_mnw_kl(const float* mu_0, const float* mu, const float* invV_0, const float* V,
                const float* logdetinvV, const float* logdetinvV_0, const float* invU_0, const float* U,
                const float* nu_0, const float* nu, const float* logdet_invU, const float* logdet_invU_0,
                int C, int n, int pp, float* out, void* stream);

/* ---- K6: HMM forward-backward -----------------------------------------------------------
 * HMM.forward_backward_logits — models/HMM.py:72-105 (called by HMM.update_states :119-132 with the output of
 * obs_logits).  logits (T, S, K): observation log-likelihoods, S sequences (all sample / batch dims but time,
 * flattened; sequence s uses parameter group s % G); trans (G, K, K) = transition.loggeomean(), init (G, K) =
 * initial.loggeomean(); K <= 32.  Outputs: p (T, S, K) posterior state probabilities (softmax of the smoothed
 * log-marginals divided by ptemp), SEzz (S, K, K) expected transition counts incl. the initial step, SEz0 (S, K),
 * logZ (S).  p may not alias logits.                                                                          */
int vbmp_hmm_forward_backward(const float* logits, const float* trans, const float* init, int T, long long S, int G, int K,
                              float ptemp, float* p, float* SEzz, float* SEz0, float* logZ, void* stream);

/* ---- row GEMM with fp32-grade accuracy ----------------------------------------------------------
 * C[N x M] (+)= A[N x Kd] B[Kd x M] (+ bias[M]), all row-major with leading dimensions lda / ldb / ldc (floats); every
 * product is the 3-term TF32 split on the warp-level tensor-core path.  For the sample-major products with one small
 * shared operand: the component means and sum_k p_k ESigma_k of MixtureofLinearTransforms.predict
 * (transforms/MixtureofLinearTransforms.py:100-103) and the covariance terms of MatrixNormalWishart.Elog_like_given_pX_pY
 * (transforms/MatrixNormalWishart.py:236-247).  accumulate != 0 adds to C; bias may be NULL.                            */
int vbmp_rowgemm(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc,
                 long long N, int Kd, int M, int accumulate, void* stream);
/* The same product given a workspace of vbmp_rowgemm_workspace_bytes(Kd, M, has_bias) bytes: shapes with a short reduction
 * and a wide output (Kd (+1 with a bias) <= 64 after padding to 8, M >= 64, N >= 128 — the two products of predict) run on a
 * tcgen05 kernel (A tile in tensor memory, B packed per 128-column chunk, 3-term TF32) that is bound by the bytes of C it
 * writes; every other shape, or workspace == NULL, takes vbmp_rowgemm's kernel.                                          */
size_t vbmp_rowgemm_workspace_bytes(int Kd, int M, int has_bias);
int vbmp_rowgemm_ex(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc,
                    long long N, int Kd, int M, int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* ---- responsibilities from given logits ---------------------------------------------------------------------------------
 * p[n][k] = exp(l[n][k] + colbias[k] - logZ_n), logZ_n = logsumexp_k (l[n][k] + colbias[k]), NA[k] = sum_n p[n][k],
 * logZ = sum_n logZ_n: the responsibility step when the logits do not come out of one vbmp_estep launch —
 * MixtureofLinearTransforms.update_assignments_given_pX_pY (transforms/MixtureofLinearTransforms.py:62-69: vbmp_estep mode 0 on
 * the means, vbmp_rowterm for the covariance terms, then this with colbias = Dirichlet.loggeomean()).  logits (N, K, row
 * stride ldl) and p (N, K, row stride ldp) may be the same buffer; colbias may be NULL.  NA / logZ are reduced in a fixed
 * order (bit-reproducible).                                                                                             */
size_t vbmp_softmax_rows_workspace_bytes(long long N, int K);
int vbmp_softmax_rows(const float* logits, int ldl, const float* colbias, long long N, int K, float* p, int ldp,
                      float* logZn, float* NA, float* logZ, void* workspace, size_t workspace_bytes, void* stream);

/* ---- per-sample covariance terms --------------------------------------------------------------------
 * C[n][k] = (accumulate ? C[n][k] : 0) + alpha * sum_f A[n][f] B[f][k]: the trace terms of
 * MatrixNormalWishart.Elog_like_given_pX_pY (transforms/MatrixNormalWishart.py:236-247), -1/2 tr(Sigma_y,n E[invSigma_k])
 * and -1/2 tr(Sigma_x,n E[X^T invU X]_k), with A (N, F, row stride lda) the flattened per-sample covariances and B (F, K,
 * row stride ldb) the flattened K-sized expectations.  tcgen05 kernel: A goes from HBM through registers into tensor
 * memory (one pass, 3-term TF32 split), B is packed once per call.  Needs N >= 128, F >= 32, F % 4 == 0, lda % 4 == 0,
 * K <= 256, a 16-byte aligned A; anything else returns VBMP_ERR_UNSUPPORTED (vbmp_rowgemm takes every shape).            */
size_t vbmp_rowterm_workspace_bytes(int F, int K);
int vbmp_rowterm(const float* A, int lda, const float* B, int ldb, float* C, int ldc, long long N, int F, int K,
                 float alpha, int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* ---- responsibility-weighted column sums ------------------------------------------------------------
 * out[K x F] = sum_n p[n][k] S[n][f]: the weighted sums of the flattened input / output covariances in
 * MatrixNormalWishart.update(pX, pY, p) (transforms/MatrixNormalWishart.py:150-156, SExx = (pX.EXXT() * p).sum(0) etc.; the
 * mean parts go through vbmp_gram).  p (N, K) and S (N, F, row stride lds floats) row-major fp32; the reduction runs over
 * the sample axis on the tcgen05 Gram kernel (3-term TF32 split, two-level accumulation, fixed-order fp64 reduce over the
 * sample splits: bit-reproducible).  Needs N >= 2048, K % 4 == 0, lds % 4 == 0, 16-byte aligned bases; anything else
 * returns VBMP_ERR_UNSUPPORTED.                                                                                          */
size_t vbmp_wsum_workspace_bytes(long long N, int K, int F);
int vbmp_wsum(const float* p, const float* S, int lds, long long N, int K, int F, float* out, void* workspace,
              size_t workspace_bytes, void* stream);

/* ---- mixture-of-experts predictive moments -------------------------------------------------------
 * The per-sample part of MixtureofLinearTransforms.predict (transforms/MixtureofLinearTransforms.py:100-106):
 *   mu[s] = sum_k p[s,k] mean[s,k,:],   Sigma[s] = base[s] + sum_k p[s,k] mean[s,k,:] mean[s,k,:]^T - mu[s] mu[s]^T
 * with mean (N, K, n) the component means E[y | x_s, k], p (N, K) the gate probabilities, base (N, n, n) = sum_k p[s,k]
 * ESigma_k (or NULL), n <= 32.  One warp per sample.                                                                   */
int vbmp_moe_moments(const float* mean, const float* p, const float* base, long long N, int K, int n,
                     float* mu, float* Sigma, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VBMP_B200_H */
