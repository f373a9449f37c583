// FP64 SIMT GEMM building block (B200 keeps a full-rate FP64 pipe: 64 DFMA/clk/SM).
// C[M×N] = alpha · A[M×K] · op(B) + beta · C, with per-tile K-range clipping for triangular operands.
// Used by the blocked Cholesky (syrk trailing update), the triangular inverse and the FP64
// (checker-precision) variance contraction.
#pragma once
#include <cuda_runtime.h>

enum { KM_FULL = 0, KM_UPTO_N = 1, KM_FROM_N = 2, KM_UPTO_M = 3, KM_FROM_M = 4 };
enum { EPI_STORE = 0, EPI_ROWSUMSQ = 1 };
enum { TS_NONE = 0, TS_LOWER = 1 };

#define DG_BM 128
#define DG_BN 64
#define DG_BK 16

// TRANSB = true : B is N×K row-major (C = A·Bᵀ, "NT");  false: B is K×N row-major ("NN").
// 128×64 tile, 8×4 outputs per thread: 12 LDS per 32 DFMA, so the FP64 pipe — not shared memory — is the limiter
// (the first version's 4×4 tile ran at 16 TFLOP/s, shared-memory bound).  The next K-slab is prefetched into registers
// while the current one is consumed.
template <bool TRANSB, int EPI>
__global__ void __launch_bounds__(256, 2)
dgemm64_kernel(int M, int N, int K, const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
               double* __restrict__ C, int ldc, double alpha, double beta, int kmode, int kbegin, int tileskip,
               long long strideA, long long strideB, long long strideC) {
  A += (long long)blockIdx.z * strideA;
  B += (long long)blockIdx.z * strideB;
  C += (long long)blockIdx.z * strideC;
  __shared__ double As[DG_BK][DG_BM + 2];
  __shared__ double Bs[DG_BK][DG_BN + 2];
  const int m0 = blockIdx.y * DG_BM, n0 = blockIdx.x * DG_BN;
  if (tileskip == TS_LOWER && n0 > m0 + DG_BM - 1) return;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  int kb = kbegin, ke = K;
  if (kmode == KM_UPTO_N) ke = min(K, n0 + DG_BN);
  if (kmode == KM_FROM_N) kb = max(kbegin, n0);
  if (kmode == KM_UPTO_M) ke = min(K, m0 + DG_BM);
  if (kmode == KM_FROM_M) kb = max(kbegin, m0);
  kb = kb & ~(DG_BK - 1);

  double acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.0;

  double pa[8], pb[4];
  auto prefetch = [&](int k0) {
    {
      const int k = tid & 15, gk = k0 + k;
      const bool kok = gk < ke && gk >= kbegin;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int gm = m0 + (tid >> 4) + 16 * i;
        pa[i] = (kok && gm < M) ? A[(size_t)gm * lda + gk] : 0.0;
      }
    }
    if (TRANSB) {
      const int k = tid & 15, gk = k0 + k;
      const bool kok = gk < ke && gk >= kbegin;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int gn = n0 + (tid >> 4) + 16 * i;
        pb[i] = (kok && gn < N) ? B[(size_t)gn * ldb + gk] : 0.0;
      }
    } else {
      const int gn = n0 + (tid & 63);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int gk = k0 + (tid >> 6) + 4 * i;
        pb[i] = (gn < N && gk < ke && gk >= kbegin) ? B[(size_t)gk * ldb + gn] : 0.0;
      }
    }
  };
  auto commit = [&]() {
#pragma unroll
    for (int i = 0; i < 8; i++) As[tid & 15][(tid >> 4) + 16 * i] = pa[i];
    if (TRANSB) {
#pragma unroll
      for (int i = 0; i < 4; i++) Bs[tid & 15][(tid >> 4) + 16 * i] = pb[i];
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++) Bs[(tid >> 6) + 4 * i][tid & 63] = pb[i];
    }
  };
  if (kb < ke) {
    prefetch(kb);
    commit();
  }
  __syncthreads();
  for (int k0 = kb; k0 < ke; k0 += DG_BK) {
    const bool more = k0 + DG_BK < ke;
    if (more) prefetch(k0 + DG_BK);
#pragma unroll
    for (int k = 0; k < DG_BK; k++) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = As[k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
    if (more) commit();
    __syncthreads();
  }

  if (EPI == EPI_STORE) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int gm = m0 + ty + 16 * i;
      if (gm >= M) continue;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int gn = n0 + tx + 16 * j;
        if (gn >= N) continue;
        double* c = C + (size_t)gm * ldc + gn;
        *c = (beta == 0.0) ? alpha * acc[i][j] : alpha * acc[i][j] + beta * (*c);
      }
    }
  } else {  // EPI_ROWSUMSQ: C is part[M × ldc], column = this block's n-tile; fixed reduction order
#pragma unroll
    for (int i = 0; i < 8; i++) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int gn = n0 + tx + 16 * j;
        if (gn < N) s = fma(acc[i][j], acc[i][j], s);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const int gm = m0 + ty + 16 * i;
      if (tx == 0 && gm < M) C[(size_t)gm * ldc + blockIdx.x] = s;
    }
  }
}

template <bool TRANSB, int EPI>
static inline void dgemm64_launch(cudaStream_t s, int M, int N, int K, const double* A, int lda, const double* B, int ldb,
                                  double* C, int ldc, double alpha, double beta, int kmode, int kbegin, int tileskip,
                                  int batch = 1, long long strideA = 0, long long strideB = 0, long long strideC = 0) {
  dim3 grid((N + DG_BN - 1) / DG_BN, (M + DG_BM - 1) / DG_BM, batch);
  dgemm64_kernel<TRANSB, EPI><<<grid, 256, 0, s>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, kmode, kbegin, tileskip,
                                                   strideA, strideB, strideC);
}
