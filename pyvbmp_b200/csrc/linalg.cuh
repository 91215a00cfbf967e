// Block-cooperative dense linear algebra on one small SPD matrix (n <= 128) held in shared memory
// as packed lower triangles in fp64.  One CTA per matrix; every routine must be called by all
// threads of the block.  These replace the reference's batched LU calls (Tensor.inverse / logdet /
// linalg.solve: dists/Wishart.py:55-56, transforms/MatrixNormalWishart.py:108,134-135) with a
// Cholesky route, which is valid because every matrix on the path is SPD.
#pragma once
#include "common.cuh"

namespace vbmp {

// In-place Cholesky A = L L^T on a packed lower triangle.  Returns logdet(A) = 2 sum log L_ii.
// *info (shared int, pre-zeroed) is set to (column+1) of the first non-positive pivot.
__device__ inline double chol_packed(double* A, int n, int* info, double* red) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int k = 0; k < n; ++k) {
    __syncthreads();
    double akk = A[tri(k, k)];
    if (!(akk > 0.0)) {            // uniform across the block: everyone reads the same value
      if (tid == 0 && *info == 0) *info = k + 1;
      akk = nan("");
    }
    const double lkk = sqrt(akk), inv = 1.0 / lkk;
    __syncthreads();
    if (tid == 0) A[tri(k, k)] = lkk;
    for (int i = k + 1 + tid; i < n; i += nt) A[tri(i, k)] *= inv;
    __syncthreads();
    // trailing update: rows i > k, columns k < j <= i
    const int m = n - k - 1;
    for (int e = tid; e < m * m; e += nt) {
      const int i = k + 1 + e / m, j = k + 1 + e % m;
      if (j <= i) A[tri(i, j)] -= A[tri(i, k)] * A[tri(j, k)];
    }
  }
  __syncthreads();
  double v = 0.0;
  for (int i = tid; i < n; i += nt) v += log(A[tri(i, i)]);
  return 2.0 * block_sum(v, red);
}

// Li = L^{-1} (lower, packed) into a second packed buffer: one thread per column, forward substitution.
__device__ inline void tri_inverse_packed(const double* L, double* Li, int n) {
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    Li[tri(j, j)] = 1.0 / L[tri(j, j)];
    for (int i = j + 1; i < n; ++i) {
      double s = 0.0;
      for (int k = j; k < i; ++k) s += L[tri(i, k)] * Li[tri(k, j)];
      Li[tri(i, j)] = -s / L[tri(i, i)];
    }
  }
  __syncthreads();
}

// (A^{-1})_{ab} = sum_{k >= max(a,b)} Li[k][a] Li[k][b]
__device__ inline double inv_entry(const double* Li, int n, int a, int b) {
  const int k0 = a > b ? a : b;
  double s = 0.0;
  for (int k = k0; k < n; ++k) s += Li[tri(k, a)] * Li[tri(k, b)];
  return s;
}

}  // namespace vbmp
