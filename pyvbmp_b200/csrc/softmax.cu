// Row softmax with the E-step's outputs:  p[n][k] = exp(l[n][k] + b[k] - logZ_n),  logZ_n = logsumexp_k (l[n][k] + b[k]),
// NA[k] = sum_n p[n][k],  logZ = sum_n logZ_n.
//
// The responsibility step of the paths whose logits are not produced by ONE K2 launch: the expectation-input E-step
// (MixtureofLinearTransforms.update_assignments_given_pX_pY, transforms/MixtureofLinearTransforms.py:62-69: the trace terms
// are added to K2's logits first) and mixtures with more than 512 components (K2 runs per block of components, api.cu).
// HBM-bound: one read and one write of the (N x K) array (the second read of a row hits L1 / L2).  A CTA owns 128 rows, a
// warp one row at a time; lane l owns the columns l, l + 32, ... of every row its warp sees, so the per-warp column sums
// in shared memory need no atomics, and the per-CTA partials are added in a fixed order (estep_reduce_kernel):
// NA and logZ are bit-reproducible.
#include "common.cuh"

namespace vbmp {

constexpr int SM_ROWS = 128;

__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* l, int ldl, const float* __restrict__ bias,   // l and p may alias
                                                           long long N, int K, float* p, int ldp,
                                                           float* __restrict__ logZn, float* __restrict__ NA_part,
                                                           double* __restrict__ logZ_part) {
  extern __shared__ float sm_cs[];                       // [8 warps][K] column sums
  __shared__ double dred[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* cs = sm_cs + (size_t)warp * K;
  for (int k = lane; k < K; k += 32) cs[k] = 0.f;
  const long long r0 = (long long)blockIdx.x * SM_ROWS;
  double lzs = 0.0;
  constexpr float L2E = 1.4426950408889634f;
  for (int r = warp; r < SM_ROWS; r += 8) {
    const long long n = r0 + r;
    if (n >= N) break;
    const float* row = l + (size_t)n * ldl;
    float mx = -INFINITY;
    for (int k = lane; k < K; k += 32) mx = fmaxf(mx, row[k] + (bias ? bias[k] : 0.f));
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s += exp2f((row[k] + (bias ? bias[k] : 0.f) - mx) * L2E);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float lz = mx + logf(s);
    float* prow = p + (size_t)n * ldp;
    for (int k = lane; k < K; k += 32) {
      const float v = exp2f((row[k] + (bias ? bias[k] : 0.f) - lz) * L2E);
      prow[k] = v;
      cs[k] += v;
    }
    if (lane == 0) { logZn[n] = lz; lzs += (double)lz; }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += 256) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm_cs[(size_t)w * K + k];
    NA_part[(size_t)blockIdx.x * K + k] = t;
  }
  if (lane == 0) dred[warp] = lzs;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += dred[w];
    logZ_part[blockIdx.x] = t;
  }
}

int launch_estep_reduce(const float*, const double*, int nb, int G, int K, float* NA, float* logZ, cudaStream_t);

static size_t sm_al(size_t x) { return (x + 255) / 256 * 256; }
size_t softmax_rows_workspace_bytes(long long N, int K) {
  const size_t nb = (size_t)((N + SM_ROWS - 1) / SM_ROWS);
  return 256 + sm_al(nb * (size_t)K * sizeof(float)) + sm_al(nb * sizeof(double));
}

int launch_softmax_rows(const float* l, int ldl, const float* bias, long long N, int K, float* p, int ldp, float* logZn,
                        float* NA, float* logZ, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (N < 0 || K < 1 || ldl < K || ldp < K || (size_t)K * 8 * sizeof(float) > 200 * 1024) {
    set_error("softmax_rows: bad shape N=%lld K=%d ldl=%d ldp=%d (K <= 6400)", N, K, ldl, ldp);
    return VBMP_ERR_SHAPE;
  }
  if (N == 0) {
    cudaMemsetAsync(NA, 0, sizeof(float) * K, st);
    cudaMemsetAsync(logZ, 0, sizeof(float), st);
    return VBMP_OK;
  }
  if (ws_bytes < softmax_rows_workspace_bytes(N, K)) { set_error("softmax_rows: workspace too small"); return VBMP_ERR_WORKSPACE; }
  const int nb = (int)((N + SM_ROWS - 1) / SM_ROWS);
  char* q = (char*)sm_al((size_t)ws);
  float* NA_part = (float*)q; q += sm_al((size_t)nb * K * sizeof(float));
  double* logZ_part = (double*)q;
  const size_t smem = (size_t)8 * K * sizeof(float);
  if (smem > 48 * 1024) cudaFuncSetAttribute(softmax_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  softmax_rows_kernel<<<nb, 256, smem, st>>>(l, ldl, bias, N, K, p, ldp, logZn, NA_part, logZ_part);
  int rc = check_launch("softmax_rows");
  if (rc) return rc;
  return launch_estep_reduce(NA_part, logZ_part, nb, 1, K, NA, logZ, st);
}

}  // namespace vbmp
