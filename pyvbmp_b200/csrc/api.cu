// extern "C" entry points of libvbmp_b200.so (see include/vbmp_b200.h for the contract).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include "common.cuh"
#include "../../include/vbmp_b200.h"

namespace vbmp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

// called once behind every kernel launch: counts it (vbmp_launch_count) and turns a launch error into a return code
int check_launch(const char* what) {
  count_launch(1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return VBMP_ERR_CUDA;
  }
  return VBMP_OK;
}

// ---- forward declarations of the launchers (prep.cu, update.cu, estep_simt.cu, gram_simt.cu, *_umma.cu)
int launch_niw_prep(const float*, const float*, const float*, const float*, const float*, int, int, int, float*, float*, float*, int*, cudaStream_t);
int launch_mnw_prep(const float*, const float*, const float*, const float*, const float*, int, int, int, int, int, float*, float*, float*, int*, cudaStream_t, const float* tau = nullptr, const float* elogdet = nullptr);
int estep_simt_tile(int Dp);
int launch_estep_simt(const EstepArgs&, int mode, cudaStream_t);
int launch_estep_reduce(const float*, const double*, int nb, int G, int K, float* NA, float* logZ, cudaStream_t);
int gram_simt_plan(long long N, int G, int K, int Dp, long long* S_per, int* splits);
int launch_gram_simt(const GramArgs&, cudaStream_t);
int launch_gram_reduce(const float* part, int splits, size_t per, float* gram, cudaStream_t);
int launch_wishart_update(const float*, const float*, const float*, const float*, const float*, const float*, int, int, float, float*, float*, float*, float*, int*, cudaStream_t);
int launch_niw_update(const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, int, int, float, int, float*, float*, float*, float*, float*, float*, int*, cudaStream_t);
int launch_mnw_update(const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, int, int, int, float, int, float*, float*, float*, float*, float*, float*, float*, float*, int*, cudaStream_t);
int launch_wishart_elogdet(const float*, const float*, int, int, float*, cudaStream_t);
int launch_wishart_kl(const float*, const float*, const float*, const float*, const float*, const float*, int, int, float*, cudaStream_t);
int launch_niw_kl(const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, int, int, float*, cudaStream_t);
int launch_mnw_kl(const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, int, int, int, float*, cudaStream_t);
int launch_hmm_fb(const float*, const float*, const float*, int, long long, int, int, float, float*, float*, float*, float*, cudaStream_t);
// tcgen05 variants (estep_umma.cu / gram_umma.cu)
bool estep_umma_supported(long long N, int GX, int G, int K, int Dp, int d0, int d1);
size_t estep_umma_workspace_bytes(long long N, int G, int K, int Dp, int mode);
int launch_estep_umma(const EstepArgs&, int mode, void* ws, size_t ws_bytes, float* NA, float* logZ, cudaStream_t);
bool gram_umma_supported(long long N, int GX, int GP, int G, int K, int Dp, int d0, int d1, bool has_p);
size_t gram_umma_workspace_bytes(long long N, int G, int K, int d0, int d1, int Dp, bool has_rpack, bool has_zpack);
size_t gram_zpack_bytes(long long N, int D);
bool gram_zpack_usable(long long N, int K, int Dp, int d0, int d1);
int launch_gram_zpack(const float* z0, int d0, const float* z1, int d1, long long N, void* zpack, cudaStream_t st);
int launch_gram_umma(const GramArgs&, float* gram, void* ws, size_t ws_bytes, cudaStream_t);
size_t diag_estep_workspace_bytes(long long N, int G, int K, int d, int mode);
int launch_diag_estep(const float* x, int d, long long N, int GX, const int* xg, const float* mu, const float* tau,
                      const float* cst, int G, int K, int mode, float* out, float* logZn, float* NA, float* logZ,
                      void* ws, size_t ws_bytes, cudaStream_t st, unsigned char* rpack = nullptr);
int launch_moe_moments(const float* mean, const float* p, const float* base, long long N, int K, int n, float* mu, float* Sigma,
                       cudaStream_t st);
int launch_rowgemm(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc, long long N, int Kd,
                   int M, int accumulate, cudaStream_t st);
size_t softmax_rows_workspace_bytes(long long N, int K);
int launch_softmax_rows(const float* l, int ldl, const float* bias, long long N, int K, float* p, int ldp, float* logZn,
                        float* NA, float* logZ, void* ws, size_t ws_bytes, cudaStream_t st);
bool rowwide_umma_supported(long long N, int Kd, int M, bool bias);
size_t rowwide_umma_workspace_bytes(int Kd, int M, bool bias);
int launch_rowwide_umma(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc, long long N, int Kd,
                        int M, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st);
bool rowterm_umma_supported(long long N, int F, int K, int lda);
size_t rowterm_umma_workspace_bytes(int F, int K);
int launch_rowterm_umma(const float* A, int lda, const float* B, int ldb, float* C, int ldc, long long N, int F, int K,
                        float alpha, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st);
bool wsum_umma_supported(long long N, int K, int F, int lds);
size_t wsum_umma_workspace_bytes(long long N, int K, int F);
int launch_wsum_umma(const float* p, const float* S, int lds, long long N, int K, int F, float* out, void* ws, size_t ws_bytes,
                     cudaStream_t st);
bool gram_rpack_usable();
bool estep_umma_can_pack(long long N, int GX, int G, int K, int Dp, int d0, int d1, int mode);
size_t gram_rpack_bytes(long long N, int K);

static bool valid_dp(int Dp) { return Dp == 8 || Dp == 16 || Dp == 32 || Dp == 64 || Dp == 128; }
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- K that is not a multiple of 4 -----------------------------------------------------------------------------------
// The tensor-core kernels write responsibilities / read weights in 16-byte pieces of a row, so they need K % 4 == 0.  A call
// with any other K that is otherwise inside their window runs on them with K padded to Kq = 4 ceil(K / 4) by INERT
// components (W = 0, m = 0, cst = -1e30: probability exactly 0; weights 0: statistics exactly 0): padded parameters, a
// padded (N x Kq) responsibility / weight image and padded outputs live in the workspace, and one streaming pass
// compacts / expands the rows.  Costs ~3 passes over an (N x K) array, against a ~20x slower CUDA-core kernel.
static int kq_of(int K) { return (K + 3) / 4 * 4; }
__global__ void padk_cst_kernel(const float* __restrict__ cst, int K, int Kq, float* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < Kq) out[k] = k < K ? cst[k] : -1e30f;
}
// dst (N x Kd) <- src (N x Ks): the first min(Kd, Ks) columns of every row, zeros beyond
__global__ void padk_cols_kernel(const float* __restrict__ src, int Ks, float* __restrict__ dst, int Kd, long long N) {
  const long long tot = N * Kd;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    const long long n = e / Kd;
    const int k = (int)(e - n * Kd);
    dst[e] = k < Ks ? src[n * Ks + k] : 0.f;
  }
}
static int padk_cols(const float* src, int Ks, float* dst, int Kd, long long N, cudaStream_t st) {
  const long long tot = N * Kd;
  const unsigned grid = (unsigned)((tot + 255) / 256 < (long long)num_sms() * 16 ? (tot + 255) / 256 : (long long)num_sms() * 16);
  padk_cols_kernel<<<grid, 256, 0, st>>>(src, Ks, dst, Kd, N);
  return check_launch("padk_cols");
}

}  // namespace vbmp

using namespace vbmp;

extern "C" {

int vbmp_version(void) { return VBMP_ABI_VERSION; }
const char* vbmp_last_error(void) { return g_err; }
unsigned long long vbmp_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int vbmp_niw_prep(const float* invU, const float* mu, const float* nu, const float* lambda_mu, const float* logprior,
                  int C, int d, int Dp, float* W, float* m, float* cst, int* info, void* stream) {
  if (!valid_dp(Dp)) { set_error("niw_prep: Dp=%d must be one of 8,16,32,64,128", Dp); return VBMP_ERR_SHAPE; }
  return launch_niw_prep(invU, mu, nu, lambda_mu, logprior, C, d, Dp, W, m, cst, info, (cudaStream_t)stream);
}

int vbmp_mnw_prep(const float* invU, const float* nu, const float* mu, const float* invV, const float* logprior,
                  int C, int n, int pp, int pad_X, int Dp, float* W, float* m, float* cst, int* info, void* stream) {
  if (!valid_dp(Dp)) { set_error("mnw_prep: Dp=%d must be one of 8,16,32,64,128", Dp); return VBMP_ERR_SHAPE; }
  return launch_mnw_prep(invU, nu, mu, invV, logprior, C, n, pp, pad_X ? 1 : 0, Dp, W, m, cst, info, (cudaStream_t)stream);
}

int vbmp_mnw_prep_ex(const float* invU, const float* nu, const float* mu, const float* invV, const float* logprior,
                     const float* tau, const float* elogdet, int C, int n, int pp, int pad_X, int Dp,
                     float* W, float* m, float* cst, int* info, void* stream) {
  if (!valid_dp(Dp)) { set_error("mnw_prep_ex: Dp=%d must be one of 8,16,32,64,128", Dp); return VBMP_ERR_SHAPE; }
  return launch_mnw_prep(invU, nu, mu, invV, logprior, C, n, pp, pad_X ? 1 : 0, Dp, W, m, cst, info, (cudaStream_t)stream,
                         tau, elogdet);
}

// K > 512 (the tcgen05 E-step holds at most 512 components per launch): blocks of 512 components write their columns of the
// logits (mode 0, row stride K), then vbmp_softmax_rows turns them into responsibilities (mode 1)
constexpr int ESTEP_KBLK = 512;
static bool estep_bigk(long long N, int GX, int G, int K, int Dp, int d0, int d1) {
  return K > ESTEP_KBLK && (K % 4) == 0 && estep_umma_supported(N, GX, G, ESTEP_KBLK, Dp, d0, d1);
}
// shapes the tcgen05 E-step takes once K is padded (see kq_of)
static bool estep_padk(long long N, int GX, int G, int K, int Dp, int d0, int d1) {
  return (K % 4) != 0 && (estep_umma_supported(N, GX, G, kq_of(K), Dp, d0, d1) || estep_bigk(N, GX, G, kq_of(K), Dp, d0, d1));
}
static size_t estep_padk_extra(long long N, int K, int Dp) {       // padded W, m, cst, responsibilities, NA
  const size_t Kq = (size_t)kq_of(K);
  return align_up(Kq * Dp * Dp * sizeof(float), 256) + align_up(Kq * Dp * sizeof(float), 256) + 2 * align_up(Kq * sizeof(float), 256) +
         align_up((size_t)N * Kq * sizeof(float), 256);
}

size_t vbmp_estep_workspace_bytes(long long N, int G, int K, int Dp, int mode) {
  if (!valid_dp(Dp) || N < 0) return 0;
  if (estep_padk(N, 1, G, K, Dp, 1, 0)) return estep_padk_extra(N, K, Dp) + vbmp_estep_workspace_bytes(N, G, kq_of(K), Dp, mode) + 256;
  if (estep_bigk(N, 1, G, K, Dp, 1, 0)) {
    const size_t u = estep_umma_workspace_bytes(N, G, ESTEP_KBLK, Dp, 0), sm = softmax_rows_workspace_bytes(N, K);
    return (u > sm ? u : sm) + 512;
  }
  size_t simt = 0;
  if (mode == 1) {
    const size_t nb = (size_t)cdiv(N, estep_simt_tile(Dp));
    simt = align_up(nb * G * K * sizeof(float), 256) + align_up(nb * G * sizeof(double), 256);
  }
  const size_t umma = estep_umma_workspace_bytes(N, G, K, Dp, mode);
  return (simt > umma ? simt : umma) + 256;
}

static int estep_impl(const float* z0, int d0, const float* z1, int d1, long long N, int GX, const int* xg,
                      const float* W, const float* m, const float* cst, int G, int K, int Dp, int mode, int flags,
                      float* out, float* logZn, float* NA, float* logZ,
                      void* workspace, size_t workspace_bytes, void* stream, void* rpack, size_t rpack_bytes, int* packed) {
  cudaStream_t st = (cudaStream_t)stream;
  if (packed) *packed = 0;
  if (!valid_dp(Dp) || d0 < 1 || d1 < 0 || d0 + d1 > Dp || G < 1 || K < 1 || GX < 1 || N < 0 || (mode != 0 && mode != 1)) {
    set_error("estep: bad shape N=%lld GX=%d G=%d K=%d d0=%d d1=%d Dp=%d mode=%d", N, GX, G, K, d0, d1, Dp, mode);
    return VBMP_ERR_SHAPE;
  }
  if (d1 > 0 && !z1 && N > 0) { set_error("estep: z1 is NULL with d1=%d", d1); return VBMP_ERR_SHAPE; }
  // an empty batch (its buffers may be NULL): NA = 0, logZ = 0, nothing else to write (the reference's sums over no rows)
  if (mode == 1 && (!NA || !logZ || (N > 0 && !logZn))) { set_error("estep: mode 1 needs logZn, NA, logZ"); return VBMP_ERR_SHAPE; }
  if (N == 0) {
    if (mode == 1) { cudaMemsetAsync(NA, 0, sizeof(float) * G * K, st); cudaMemsetAsync(logZ, 0, sizeof(float) * G, st); }
    return VBMP_OK;
  }
  if (!(flags & 1) && estep_padk(N, GX, G, K, Dp, d0, d1)) {
    // K % 4 != 0: run the tensor-core kernel on K padded by inert components, then compact the rows
    const int Kq = kq_of(K);
    if (workspace_bytes < vbmp_estep_workspace_bytes(N, G, K, Dp, mode)) {
      set_error("estep: workspace too small (%zu < %zu)", workspace_bytes, vbmp_estep_workspace_bytes(N, G, K, Dp, mode));
      return VBMP_ERR_WORKSPACE;
    }
    char* q = (char*)align_up((size_t)workspace, 256);
    float* Wq = (float*)q; q += align_up((size_t)Kq * Dp * Dp * sizeof(float), 256);
    float* mq = (float*)q; q += align_up((size_t)Kq * Dp * sizeof(float), 256);
    float* cq = (float*)q; q += align_up((size_t)Kq * sizeof(float), 256);
    float* NAq = (float*)q; q += align_up((size_t)Kq * sizeof(float), 256);
    float* outq = (float*)q; q += align_up((size_t)N * Kq * sizeof(float), 256);
    const size_t nW = (size_t)K * Dp * Dp * sizeof(float), nm = (size_t)K * Dp * sizeof(float);
    if (cudaMemsetAsync(Wq, 0, (size_t)Kq * Dp * Dp * sizeof(float), st) != cudaSuccess ||
        cudaMemsetAsync(mq, 0, (size_t)Kq * Dp * sizeof(float), st) != cudaSuccess ||
        cudaMemcpyAsync(Wq, W, nW, cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(mq, m, nm, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
      set_error("estep: padding copies failed"); return VBMP_ERR_CUDA;
    }
    padk_cst_kernel<<<(Kq + 127) / 128, 128, 0, st>>>(cst, K, Kq, cq);
    int rc = check_launch("padk_cst");
    if (rc) return rc;
    rc = estep_impl(z0, d0, z1, d1, N, GX, xg, Wq, mq, cq, G, Kq, Dp, mode, flags, outq, logZn, NAq, logZ, q,
                    workspace_bytes - (size_t)(q - (char*)workspace), stream, rpack, rpack_bytes, packed);
    if (rc) return rc;
    rc = padk_cols(outq, Kq, out, K, N, st);
    if (rc) return rc;
    if (mode == 1 && cudaMemcpyAsync(NA, NAq, (size_t)K * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
      set_error("estep: NA copy failed"); return VBMP_ERR_CUDA;
    }
    return VBMP_OK;
  }
  if (!(flags & 1) && estep_bigk(N, GX, G, K, Dp, d0, d1)) {
    if (workspace_bytes < vbmp_estep_workspace_bytes(N, G, K, Dp, mode)) {
      set_error("estep: workspace too small (%zu < %zu)", workspace_bytes, vbmp_estep_workspace_bytes(N, G, K, Dp, mode));
      return VBMP_ERR_WORKSPACE;
    }
    for (int k0 = 0; k0 < K; k0 += ESTEP_KBLK) {
      const int kb = K - k0 < ESTEP_KBLK ? K - k0 : ESTEP_KBLK;
      EstepArgs b{z0, z1, d0, d1, N, GX, xg, W + (size_t)k0 * Dp * Dp, m + (size_t)k0 * Dp, cst + k0, G, kb, Dp, out + k0, nullptr,
                  nullptr, nullptr};
      b.ldo = K;
      int rc = launch_estep_umma(b, 0, workspace, workspace_bytes, nullptr, nullptr, st);
      if (rc) return rc;
    }
    if (mode == 0) return VBMP_OK;
    return launch_softmax_rows(out, K, nullptr, N, K, out, K, logZn, NA, logZ, workspace, workspace_bytes, st);
  }
  if (workspace_bytes < vbmp_estep_workspace_bytes(N, G, K, Dp, mode) && mode == 1) {
    set_error("estep: workspace too small (%zu < %zu)", workspace_bytes, vbmp_estep_workspace_bytes(N, G, K, Dp, mode));
    return VBMP_ERR_WORKSPACE;
  }
  EstepArgs a{z0, z1, d0, d1, N, GX, xg, W, m, cst, G, K, Dp, out, logZn, nullptr, nullptr};
  if (!(flags & 1) && estep_umma_supported(N, GX, G, K, Dp, d0, d1)) {
    if (rpack && packed && estep_umma_can_pack(N, GX, G, K, Dp, d0, d1, mode)) {
      if (rpack_bytes < gram_rpack_bytes(N, K)) {
        set_error("estep_rpack: buffer too small (%zu < %zu)", rpack_bytes, gram_rpack_bytes(N, K));
        return VBMP_ERR_WORKSPACE;
      }
      a.rpack = (unsigned char*)rpack;
      *packed = 1;
    }
    return launch_estep_umma(a, mode, workspace, workspace_bytes, NA, logZ, st);
  }
  int nb = cdiv(N, estep_simt_tile(Dp));
  if (mode == 1) {
    char* ws = (char*)align_up((size_t)workspace, 256);
    a.NA_part = (float*)ws;
    a.logZ_part = (double*)(ws + align_up((size_t)nb * G * K * sizeof(float), 256));
  }
  int rc = launch_estep_simt(a, mode, st);
  if (rc) return rc;
  if (mode == 1) rc = launch_estep_reduce(a.NA_part, a.logZ_part, nb, G, K, NA, logZ, st);
  return rc;
}

int vbmp_estep(const float* z0, int d0, const float* z1, int d1, long long N, int GX, const int* xg,
               const float* W, const float* m, const float* cst, int G, int K, int Dp, int mode, int flags,
               float* out, float* logZn, float* NA, float* logZ,
               void* workspace, size_t workspace_bytes, void* stream) {
  return estep_impl(z0, d0, z1, d1, N, GX, xg, W, m, cst, G, K, Dp, mode, flags, out, logZn, NA, logZ, workspace,
                    workspace_bytes, stream, nullptr, 0, nullptr);
}

size_t vbmp_diag_estep_workspace_bytes(long long N, int G, int K, int d, int mode) {
  return (N < 0 || G < 1 || K < 1 || d < 1) ? 0 : diag_estep_workspace_bytes(N, G, K, d, mode);
}

int vbmp_diag_estep(const float* x, int d, long long N, int GX, const int* xg, const float* mu, const float* tau,
                    const float* cst, int G, int K, int mode, float* out, float* logZn, float* NA, float* logZ,
                    void* workspace, size_t workspace_bytes, void* stream) {
  return launch_diag_estep(x, d, N, GX, xg, mu, tau, cst, G, K, mode, out, logZn, NA, logZ, workspace, workspace_bytes,
                           (cudaStream_t)stream);
}

int vbmp_diag_estep_rpack(const float* x, int d, long long N, int GX, const int* xg, const float* mu, const float* tau,
                          const float* cst, int G, int K, int mode, float* out, float* logZn, float* NA, float* logZ,
                          void* workspace, size_t workspace_bytes, void* stream, void* rpack, size_t rpack_bytes, int* packed) {
  if (!packed) { set_error("diag_estep_rpack: packed is NULL"); return VBMP_ERR_SHAPE; }
  *packed = 0;
  unsigned char* rp = nullptr;
  // the images are what the tcgen05 fp16 Gram kernel consumes: same window as the dense E-step's hand-over
  if (rpack && mode == 1 && G == 1 && GX == 1 && K <= 256 && (K % 4) == 0 && N >= 2048 && d >= 9 && (d % 4) == 0 &&
      gram_rpack_usable()) {
    if (rpack_bytes < gram_rpack_bytes(N, K)) {
      set_error("diag_estep_rpack: buffer too small (%zu < %zu)", rpack_bytes, gram_rpack_bytes(N, K));
      return VBMP_ERR_WORKSPACE;
    }
    rp = (unsigned char*)rpack;
  }
  int rc = launch_diag_estep(x, d, N, GX, xg, mu, tau, cst, G, K, mode, out, logZn, NA, logZ, workspace, workspace_bytes,
                             (cudaStream_t)stream, rp);
  if (rc == VBMP_OK && rp) *packed = 1;
  return rc;
}

size_t vbmp_rpack_bytes(long long N, int K) { return (N < 0 || K < 1) ? 0 : gram_rpack_bytes(N, K); }

int vbmp_estep_rpack(const float* z0, int d0, const float* z1, int d1, long long N, int GX, const int* xg,
                     const float* W, const float* m, const float* cst, int G, int K, int Dp, int mode, int flags,
                     float* out, float* logZn, float* NA, float* logZ,
                     void* workspace, size_t workspace_bytes, void* stream, void* rpack, size_t rpack_bytes, int* packed) {
  if (!packed) { set_error("estep_rpack: packed is NULL"); return VBMP_ERR_SHAPE; }
  return estep_impl(z0, d0, z1, d1, N, GX, xg, W, m, cst, G, K, Dp, mode, flags, out, logZn, NA, logZ, workspace,
                    workspace_bytes, stream, rpack, rpack_bytes, packed);
}

static bool gram_padk(long long N, int GX, int GP, int G, int K, int Dp, int d0, int d1, bool has_p) {
  return (K % 4) != 0 && gram_umma_supported(N, GX, GP, G, kq_of(K), Dp, d0, d1, has_p);
}
static size_t gram_padk_extra(long long N, int K, int d0, int d1) {      // padded weights and padded statistics
  const size_t Kq = (size_t)kq_of(K), D1 = (size_t)d0 + d1 + 1;
  return align_up((size_t)N * Kq * sizeof(float), 256) + align_up(Kq * D1 * D1 * sizeof(float), 256);
}

static size_t gram_ws_bytes(long long N, int G, int K, int d0, int d1, int Dp, bool has_rpack, bool has_zpack) {
  if (!valid_dp(Dp) || N < 0) return 0;
  if (gram_padk(N, 1, 1, G, K, Dp, d0, d1, true))
    return gram_padk_extra(N, K, d0, d1) + gram_ws_bytes(N, G, kq_of(K), d0, d1, Dp, has_rpack, has_zpack) + 256;
  long long S_per; int splits;
  gram_simt_plan(N > 0 ? N : 1, G, K, Dp, &S_per, &splits);
  const size_t D1 = (size_t)d0 + d1 + 1;
  const size_t simt = (size_t)splits * G * K * D1 * D1 * sizeof(float);
  const size_t umma = gram_umma_workspace_bytes(N, G, K, d0, d1, Dp, has_rpack, has_zpack);
  return (simt > umma ? simt : umma) + 256;
}

size_t vbmp_gram_workspace_bytes(long long N, int G, int K, int d0, int d1, int Dp) {
  return gram_ws_bytes(N, G, K, d0, d1, Dp, false, false);
}

size_t vbmp_gram_ex_workspace_bytes(long long N, int G, int K, int d0, int d1, int Dp, int has_rpack, int has_zpack) {
  return gram_ws_bytes(N, G, K, d0, d1, Dp, has_rpack != 0, has_zpack != 0);
}

size_t vbmp_zpack_bytes(long long N, int d0, int d1) {
  return (N < 0 || d0 < 1 || d1 < 0) ? 0 : gram_zpack_bytes(N, d0 + d1);
}

int vbmp_gram_zpack(const float* z0, int d0, const float* z1, int d1, long long N, int K, int Dp,
                    void* zpack, size_t zpack_bytes, int* packed, void* stream) {
  if (!packed) { set_error("gram_zpack: packed is NULL"); return VBMP_ERR_SHAPE; }
  *packed = 0;
  if (!valid_dp(Dp) || d0 < 1 || d1 < 0 || d0 + d1 > Dp || N < 0 || K < 1 || (d1 > 0 && !z1 && N > 0)) {
    set_error("gram_zpack: bad shape N=%lld K=%d d0=%d d1=%d Dp=%d", N, K, d0, d1, Dp);
    return VBMP_ERR_SHAPE;
  }
  if (!gram_zpack_usable(N, kq_of(K), Dp, d0, d1)) return VBMP_OK;   // the kernels that take this shape do not use an image
  if (zpack_bytes < gram_zpack_bytes(N, d0 + d1)) {
    set_error("gram_zpack: buffer too small (%zu < %zu)", zpack_bytes, gram_zpack_bytes(N, d0 + d1));
    return VBMP_ERR_WORKSPACE;
  }
  int rc = launch_gram_zpack(z0, d0, z1, d1, N, zpack, (cudaStream_t)stream);
  if (rc == VBMP_OK) *packed = 1;
  return rc;
}

static int gram_impl(const float* z0, int d0, const float* z1, int d1, long long N, int GX, const int* xg,
                     const float* p, int GP, const int* pg, int G, int K, int Dp, int flags,
                     float* gram, void* workspace, size_t workspace_bytes, void* stream, const void* rpack,
                     const void* zpack = nullptr) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!valid_dp(Dp) || d0 < 1 || d1 < 0 || d0 + d1 > Dp || G < 1 || K < 1 || GX < 1 || GP < 1 || N < 0) {
    set_error("gram: bad shape N=%lld GX=%d GP=%d G=%d K=%d d0=%d d1=%d Dp=%d", N, GX, GP, G, K, d0, d1, Dp);
    return VBMP_ERR_SHAPE;
  }
  if (d1 > 0 && !z1 && N > 0) { set_error("gram: z1 is NULL with d1=%d", d1); return VBMP_ERR_SHAPE; }
  const size_t D1 = (size_t)d0 + d1 + 1, per = (size_t)G * K * D1 * D1;
  if (N == 0) { cudaMemsetAsync(gram, 0, per * sizeof(float), st); return VBMP_OK; }
  const size_t need = gram_ws_bytes(N, G, K, d0, d1, Dp, rpack != nullptr, zpack != nullptr);
  if (workspace_bytes < need) {
    set_error("gram: workspace too small (%zu < %zu)", workspace_bytes, need);
    return VBMP_ERR_WORKSPACE;
  }
  if (!(flags & 1) && gram_padk(N, GX, GP, G, K, Dp, d0, d1, p != nullptr)) {
    // K % 4 != 0: weights padded with zero columns, statistics of the first K components copied out
    const int Kq = kq_of(K);
    char* q = (char*)align_up((size_t)workspace, 256);
    float* pq = (float*)q; q += align_up((size_t)N * Kq * sizeof(float), 256);
    float* gq = (float*)q; q += align_up((size_t)Kq * D1 * D1 * sizeof(float), 256);
    int rc = padk_cols(p, K, pq, Kq, N, st);
    if (rc) return rc;
    rc = gram_impl(z0, d0, z1, d1, N, GX, xg, pq, GP, pg, G, Kq, Dp, flags, gq, q,
                   workspace_bytes - (size_t)(q - (char*)workspace), stream, rpack, zpack);
    if (rc) return rc;
    if (cudaMemcpyAsync(gram, gq, (size_t)K * D1 * D1 * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
      set_error("gram: copy of the statistics failed"); return VBMP_ERR_CUDA;
    }
    return VBMP_OK;
  }
  GramArgs a{z0, z1, d0, d1, N, GX, xg, p, GP, pg, G, K, Dp, 0, 0, nullptr};
  a.rpack = (const unsigned char*)rpack;
  a.zpack = (const unsigned char*)zpack;
  a.diag = (flags & 2) ? 1 : 0;            // the tcgen05 kernel then forms only the pairs (i, i), (i, D); the CUDA-core one ignores it
  if (!(flags & 1) && gram_umma_supported(N, GX, GP, G, K, Dp, d0, d1, p != nullptr))
    return launch_gram_umma(a, gram, workspace, workspace_bytes, st);
  gram_simt_plan(N, G, K, Dp, &a.S_per, &a.splits);
  a.part = (float*)align_up((size_t)workspace, 256);
  int rc = launch_gram_simt(a, st);
  if (rc) return rc;
  return launch_gram_reduce(a.part, a.splits, per, gram, st);
}

int vbmp_gram(const float* z0, int d0, const float* z1, int d1, long long N, int GX, const int* xg,
              const float* p, int GP, const int* pg, int G, int K, int Dp, int flags,
              float* gram, void* workspace, size_t workspace_bytes, void* stream) {
  return gram_impl(z0, d0, z1, d1, N, GX, xg, p, GP, pg, G, K, Dp, flags, gram, workspace, workspace_bytes, stream, nullptr);
}

int vbmp_gram_rpack(const float* z0, int d0, const float* z1, int d1, long long N, int GX, const int* xg,
                    const float* p, int GP, const int* pg, int G, int K, int Dp, int flags,
                    float* gram, void* workspace, size_t workspace_bytes, void* stream, const void* rpack) {
  return gram_impl(z0, d0, z1, d1, N, GX, xg, p, GP, pg, G, K, Dp, flags, gram, workspace, workspace_bytes, stream, rpack);
}

int vbmp_gram_ex(const float* z0, int d0, const float* z1, int d1, long long N, int GX, const int* xg,
                 const float* p, int GP, const int* pg, int G, int K, int Dp, int flags,
                 float* gram, void* workspace, size_t workspace_bytes, void* stream, const void* rpack, const void* zpack) {
  return gram_impl(z0, d0, z1, d1, N, GX, xg, p, GP, pg, G, K, Dp, flags, gram, workspace, workspace_bytes, stream, rpack, zpack);
}

int vbmp_wishart_update(const float* SExx, const float* N, const float* invU_0, const float* nu_0,
                        const float* invU_old, const float* nu_old, int C, int d, float lr,
                        float* invU, float* nu, float* U, float* logdet_invU, int* info, void* stream) {
  return launch_wishart_update(SExx, N, invU_0, nu_0, invU_old, nu_old, C, d, lr, invU, nu, U, logdet_invU, info, (cudaStream_t)stream);
}

int vbmp_niw_update(const float* SExx, const float* SEx, const float* N,
                    const float* lambda_0, const float* mu_0, const float* invU_0, const float* nu_0,
                    const float* lambda_old, const float* mu_old, const float* invU_old, const float* nu_old,
                    int C, int d, float lr, int fixed_precision,
                    float* lambda_mu, float* mu, float* invU, float* nu, float* U, float* logdet_invU,
                    int* info, void* stream) {
  return launch_niw_update(SExx, SEx, N, lambda_0, mu_0, invU_0, nu_0, lambda_old, mu_old, invU_old, nu_old, C, d, lr,
                           fixed_precision, lambda_mu, mu, invU, nu, U, logdet_invU, info, (cudaStream_t)stream);
}

int vbmp_mnw_update(const float* SExx, const float* SEyx, const float* SEyy, const float* N,
                    const float* mu_0, const float* invV_0, const float* invU_0, const float* nu_0,
                    const float* mu_old, const float* invV_old, const float* invU_old, const float* nu_old,
                    int C, int n, int pp, float lr, int fixed_precision,
                    float* mu, float* invV, float* V, float* logdetinvV,
                    float* invU, float* nu, float* U, float* logdet_invU, int* info, void* stream) {
  return launch_mnw_update(SExx, SEyx, SEyy, N, mu_0, invV_0, invU_0, nu_0, mu_old, invV_old, invU_old, nu_old, C, n, pp,
                           lr, fixed_precision, mu, invV, V, logdetinvV, invU, nu, U, logdet_invU, info, (cudaStream_t)stream);
}

int vbmp_wishart_elogdet(const float* nu, const float* logdet_invU, int C, int d, float* out, void* stream) {
  return launch_wishart_elogdet(nu, logdet_invU, C, d, out, (cudaStream_t)stream);
}

int vbmp_wishart_kl(const float* invU_0, const float* U, const float* nu_0, const float* nu,
                    const float* logdet_invU, const float* logdet_invU_0, int C, int d, float* out, void* stream) {
  return launch_wishart_kl(invU_0, U, nu_0, nu, logdet_invU, logdet_invU_0, C, d, out, (cudaStream_t)stream);
}

int vbmp_niw_kl(const float* lambda_0, const float* lambda_mu, const float* mu_0, const float* mu,
                const float* invU_0, const float* U, const float* nu_0, const float* nu,
                const float* logdet_invU, const float* logdet_invU_0, int C, int d, float* out, void* stream) {
  return launch_niw_kl(lambda_0, lambda_mu, mu_0, mu, invU_0, U, nu_0, nu, logdet_invU, logdet_invU_0, C, d, out, (cudaStream_t)stream);
}

int vbmp_mnw_kl(const float* mu_0, const float* mu, const float* invV_0, const float* V,
                const float* logdetinvV, const float* logdetinvV_0, const float* invU_0, const float* U,
                const float* nu_0, const float* nu, const float* logdet_invU, const float* logdet_invU_0,
                int C, int n, int pp, float* out, void* stream) {
  return launch_mnw_kl(mu_0, mu, invV_0, V, logdetinvV, logdetinvV_0, invU_0, U, nu_0, nu, logdet_invU, logdet_invU_0, C, n, pp, out, (cudaStream_t)stream);
}

int vbmp_moe_moments(const float* mean, const float* p, const float* base, long long N, int K, int n,
                     float* mu, float* Sigma, void* stream) {
  if (!mean || !p || !mu || !Sigma) { set_error("moe_moments: NULL argument"); return VBMP_ERR_SHAPE; }
  return launch_moe_moments(mean, p, base, N, K, n, mu, Sigma, (cudaStream_t)stream);
}

int vbmp_rowgemm(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc,
                 long long N, int Kd, int M, int accumulate, void* stream) {
  if (!A || !B || !C) { set_error("rowgemm: NULL argument"); return VBMP_ERR_SHAPE; }
  return launch_rowgemm(A, lda, B, ldb, bias, C, ldc, N, Kd, M, accumulate, (cudaStream_t)stream);
}

size_t vbmp_softmax_rows_workspace_bytes(long long N, int K) { return (N < 0 || K < 1) ? 0 : softmax_rows_workspace_bytes(N, K); }

int vbmp_softmax_rows(const float* logits, int ldl, const float* colbias, long long N, int K, float* p, int ldp,
                      float* logZn, float* NA, float* logZ, void* workspace, size_t workspace_bytes, void* stream) {
  if (!NA || !logZ || (N > 0 && (!logits || !p || !logZn))) { set_error("softmax_rows: NULL argument"); return VBMP_ERR_SHAPE; }
  return launch_softmax_rows(logits, ldl, colbias, N, K, p, ldp, logZn, NA, logZ, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t vbmp_rowgemm_workspace_bytes(int Kd, int M, int has_bias) {
  return (Kd >= 1 && M >= 1 && rowwide_umma_supported(128, Kd, M, has_bias != 0)) ? rowwide_umma_workspace_bytes(Kd, M, has_bias != 0) : 0;
}

int vbmp_rowgemm_ex(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc,
                    long long N, int Kd, int M, int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  if (!A || !B || !C) { set_error("rowgemm: NULL argument"); return VBMP_ERR_SHAPE; }
  // short reduction, wide output: the tcgen05 kernel (needs the workspace for the packed B); anything else: mma.sync kernel
  if (workspace && N >= 128 && Kd >= 1 && M >= 1 && lda >= Kd && ldb >= M && ldc >= M && rowwide_umma_supported(N, Kd, M, bias != nullptr) &&
      workspace_bytes >= rowwide_umma_workspace_bytes(Kd, M, bias != nullptr))
    return launch_rowwide_umma(A, lda, B, ldb, bias, C, ldc, N, Kd, M, accumulate, workspace, workspace_bytes, (cudaStream_t)stream);
  return launch_rowgemm(A, lda, B, ldb, bias, C, ldc, N, Kd, M, accumulate, (cudaStream_t)stream);
}

size_t vbmp_rowterm_workspace_bytes(int F, int K) { return rowterm_umma_workspace_bytes(F, K); }

int vbmp_rowterm(const float* A, int lda, const float* B, int ldb, float* C, int ldc, long long N, int F, int K,
                 float alpha, int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  if (!A || !B || !C) { set_error("rowterm: NULL argument"); return VBMP_ERR_SHAPE; }
  return launch_rowterm_umma(A, lda, B, ldb, C, ldc, N, F, K, alpha, accumulate, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t vbmp_wsum_workspace_bytes(long long N, int K, int F) { return wsum_umma_workspace_bytes(N, K, F); }

int vbmp_wsum(const float* p, const float* S, int lds, long long N, int K, int F, float* out, void* workspace,
              size_t workspace_bytes, void* stream) {
  if (!p || !S || !out) { set_error("wsum: NULL argument"); return VBMP_ERR_SHAPE; }
  return launch_wsum_umma(p, S, lds, N, K, F, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

int vbmp_hmm_forward_backward(const float* logits, const float* trans, const float* init, int T, long long S, int G, int K,
                              float ptemp, float* p, float* SEzz, float* SEz0, float* logZ, void* stream) {
  if (p == logits) { set_error("hmm_forward_backward: p must not alias logits"); return VBMP_ERR_SHAPE; }
  return launch_hmm_fb(logits, trans, init, T, S, G, K, ptemp, p, SEzz, SEz0, logZ, (cudaStream_t)stream);
}

}  // extern "C"
