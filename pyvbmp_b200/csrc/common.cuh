// Shared helpers for the pyvbmp_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define VBMP_MAX_D 128          // largest feature dimension (d, or n+p for MNW) the path accepts
#define VBMP_OK 0
#define VBMP_ERR_SHAPE 1
#define VBMP_ERR_CUDA 2
#define VBMP_ERR_WORKSPACE 3
#define VBMP_ERR_UNSUPPORTED 4

namespace vbmp {

// kernel argument blocks shared by the launchers and api.cu
struct EstepArgs {
  const float* z0; const float* z1; int d0, d1;      // z = [z0 | z1], each (N, GX, d_i) row-major
  long long N; int GX; const int* xg;                // xg[G]: data column of theta group g (NULL -> 0)
  const float* W; const float* m; const float* cst;  // (G,K,Dp,Dp), (G,K,Dp), (G,K)
  int G, K, Dp;
  float* out; float* logZn; float* NA_part; double* logZ_part;
  unsigned char* rpack = nullptr;                    // mode 1, tcgen05 fp16 path: also emit K3's pre-split weight images
  int ldo = 0;                                       // tcgen05 kernel, mode 0: row stride of `out` in floats (0 = K): a block of
                                                     // components writes its columns of a wider logits array (K > 512, api.cu)
};
struct GramArgs {
  const float* z0; const float* z1; int d0, d1;
  long long N; int GX; const int* xg;
  const float* p; int GP; const int* pg;              // p == nullptr -> unit weights (p=None branches)
  int G, K, Dp;
  long long S_per; int splits;
  float* part;                                        // [splits][G][K][(D+1)^2]
  const unsigned char* rpack = nullptr;               // pre-split weight images written by K2 (vbmp_estep_rpack)
  const unsigned char* zpack = nullptr;               // sample image written by vbmp_gram_zpack (column maxima + transposed chunks)
  int diag = 0;                                       // flags bit 1: only diag(SExx), SEx, N are needed (diagonal-precision nodes)
};

void set_error(const char* fmt, ...);
int check_launch(const char* what);
void count_launch(int n);

// SM count of the CURRENT device, cached per device (a process may drive several GPUs)
inline int num_sms() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 148; }
  if (dev >= 0 && dev < 64 && cache[dev] > 0) return cache[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 148; }
  if (dev >= 0 && dev < 64) cache[dev] = n;
  return n;
}

__host__ __device__ inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
__host__ __device__ inline int tri(int i, int j) { return i * (i + 1) / 2 + j; }   // packed lower index, j <= i

// digamma for x > 0 in fp64: upward recurrence to x >= 6, then the asymptotic series
// (the Cephes scheme torch's digamma uses on CPU; error ~1e-13 here).
__device__ inline double digamma_d(double x) {
  if (!(x > 0.0)) return nan("");
  double r = 0.0;
  while (x < 6.0) { r -= 1.0 / x; x += 1.0; }
  double f = 1.0 / (x * x);
  double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 + f * (-1.0 / 132.0
             + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
  return r + log(x) - 0.5 / x + t;
}

// sum_{i<d} psi(a - i/2) and sum_{i<d} lgamma(a - i/2), cooperatively over the block (dists/Wishart.py:37-41)
__device__ inline double block_sum(double v, double* red) {
  // red: shared scratch of >= 32 doubles; all threads of the block must call
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < nw; ++i) s += red[i];      // fixed order: deterministic
  return s;
}

__device__ inline double mv_digamma_block(double a, int d, double* red) {
  double v = 0.0;
  for (int i = threadIdx.x; i < d; i += blockDim.x) v += digamma_d(a - 0.5 * i);
  return block_sum(v, red);
}

__device__ inline double mv_lgamma_block(double a, int d, double* red) {
  double v = 0.0;
  for (int i = threadIdx.x; i < d; i += blockDim.x) v += lgamma(a - 0.5 * i);
  return block_sum(v, red);
}

}  // namespace vbmp
