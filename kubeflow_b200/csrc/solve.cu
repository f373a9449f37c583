// Triangular solves with the Cholesky factor, FP64, by 256-row panels — what the product path uses INSTEAD of the explicit
// inverse W = L⁻¹ ($SK/_gpr.py:363 alpha_ = cho_solve(L, y); :460 V = solve_triangular(L, K*ᵀ)).
//
// The pruned sweep needs exact variances for a handful of candidates and alpha for one right-hand side; forming W costs N³/3
// FP64 flop (10 ms at N = 8192, a third of the fit) to serve them.  A solve needs the factor L and the inverses of its 256×256
// diagonal blocks, which the look-ahead factorisation computes anyway (fit.cu: W_PP, kept in the diagonal blocks of the W
// buffer).  Forward, right-looking:   v_P = W_PP·b_P ;  b_>P −= L_>P,P·v_P      (P = 0, 1, …)
// backward (alpha = L⁻ᵀ z):            a_P = W_PPᵀ·z_P ; z_<P −= (L_P,<P)ᵀ·a_P   (P = last, …, 0)
// Up to 8 right-hand sides ride together (row-major N × 8): one pass over the triangle of L (268 MB at N = 8192) serves all of
// them, 2 small launches per panel.  Every right-hand side goes through the same operations in the same order whatever
// rides beside it, so a candidate's value does not depend on which other survivors it was grouped with.
#include "kbo_internal.cuh"

#define SV_P 256   // panel width = the factorisation's outer panel
#define SV_R 8     // right-hand sides per pass

namespace {

// V_P = W_PP · B_P : one warp per row of the panel (lanes stride the columns k <= r), butterfly reduction in a fixed order
__global__ void __launch_bounds__(256)
sv_diag_fwd_kernel(const double* __restrict__ Wpp, int ldw, int Wd, const double* __restrict__ Bp, double* __restrict__ Vp) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= Wd) return;
  double acc[SV_R];
#pragma unroll
  for (int q = 0; q < SV_R; q++) acc[q] = 0.0;
  const double* w = Wpp + (size_t)r * ldw;
  for (int k = lane; k <= r; k += 32) {
    const double wv = w[k];
#pragma unroll
    for (int q = 0; q < SV_R; q++) acc[q] = fma(wv, Bp[(size_t)k * SV_R + q], acc[q]);
  }
#pragma unroll
  for (int q = 0; q < SV_R; q++) {
    double v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) Vp[(size_t)r * SV_R + q] = v;
  }
}
// B_>P −= L_>P,P · V_P : one warp per row below the panel, V_P (≤ 256 × 8) in shared memory
__global__ void __launch_bounds__(256)
sv_update_fwd_kernel(const double* __restrict__ Lp /* rows below, panel's columns */, int ldl, int rows, int Wd, const double* __restrict__ Vp,
                     double* __restrict__ Bb /* rows below */) {
  __shared__ double vs[SV_P * SV_R];
  for (int e = threadIdx.x; e < Wd * SV_R; e += 256) vs[e] = Vp[e];
  __syncthreads();
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  double acc[SV_R];
#pragma unroll
  for (int q = 0; q < SV_R; q++) acc[q] = 0.0;
  const double* l = Lp + (size_t)r * ldl;
  for (int k = lane; k < Wd; k += 32) {
    const double lv = l[k];
#pragma unroll
    for (int q = 0; q < SV_R; q++) acc[q] = fma(lv, vs[k * SV_R + q], acc[q]);
  }
#pragma unroll
  for (int q = 0; q < SV_R; q++) {
    double v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) Bb[(size_t)r * SV_R + q] -= v;
  }
}
// A_P = W_PPᵀ · Z_P : thread c owns output row c, walks column c of W_PP (k >= c) — coalesced across the threads
__global__ void __launch_bounds__(256)
sv_diag_bwd_kernel(const double* __restrict__ Wpp, int ldw, int Wd, const double* __restrict__ Zp, double* __restrict__ Ap) {
  __shared__ double zs[SV_P * SV_R];
  for (int e = threadIdx.x; e < Wd * SV_R; e += 256) zs[e] = Zp[e];
  __syncthreads();
  const int c = threadIdx.x;
  if (c >= Wd) return;
  double acc[SV_R];
#pragma unroll
  for (int q = 0; q < SV_R; q++) acc[q] = 0.0;
  for (int k = c; k < Wd; k++) {
    const double wv = Wpp[(size_t)k * ldw + c];
#pragma unroll
    for (int q = 0; q < SV_R; q++) acc[q] = fma(wv, zs[k * SV_R + q], acc[q]);
  }
#pragma unroll
  for (int q = 0; q < SV_R; q++) Ap[(size_t)c * SV_R + q] = acc[q];
}
// Z_<P −= (L_P,<P)ᵀ · A_P : thread k owns row k < K0 of Z, walks the panel's rows r — coalesced across the threads
__global__ void __launch_bounds__(256)
sv_update_bwd_kernel(const double* __restrict__ Lrow /* panel's rows, columns 0.. */, int ldl, int K0, int Wd, const double* __restrict__ Ap,
                     double* __restrict__ Z) {
  __shared__ double as[SV_P * SV_R];
  for (int e = threadIdx.x; e < Wd * SV_R; e += 256) as[e] = Ap[e];
  __syncthreads();
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= K0) return;
  double acc[SV_R];
#pragma unroll
  for (int q = 0; q < SV_R; q++) acc[q] = 0.0;
  for (int r = 0; r < Wd; r++) {
    const double lv = Lrow[(size_t)r * ldl + k];
#pragma unroll
    for (int q = 0; q < SV_R; q++) acc[q] = fma(lv, as[r * SV_R + q], acc[q]);
  }
#pragma unroll
  for (int q = 0; q < SV_R; q++) Z[(size_t)k * SV_R + q] -= acc[q];
}

// right-hand sides of one group: B[j][q] = Ks[(c0 + q)·ld + j] (q < nq; the rest zero)
__global__ void sv_pack_kernel(const double* __restrict__ Ks, int ld, int N, int c0, int nq, double* __restrict__ B) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
#pragma unroll
  for (int q = 0; q < SV_R; q++) B[(size_t)j * SV_R + q] = q < nq ? Ks[(size_t)(c0 + q) * ld + j] : 0.0;
}
// varn[c0 + q] = amp − Σ_j V[j][q]² : one CTA per right-hand side, fixed-order tree
__global__ void __launch_bounds__(256) sv_sumsq_kernel(const double* __restrict__ V, int N, double amp, double* __restrict__ varn) {
  __shared__ double red[256];
  const int q = blockIdx.x;
  double a = 0.0;
  for (int j = threadIdx.x; j < N; j += 256) {
    const double v = V[(size_t)j * SV_R + q];
    a = fma(v, v, a);
  }
  red[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) varn[q] = amp - red[0];
}
__global__ void sv_col_kernel(const double* __restrict__ src, int N, int stride_src, double* __restrict__ dst, int stride_dst) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < N) dst[(size_t)j * stride_dst] = src[(size_t)j * stride_src];
}

}  // namespace

// B (N × 8, overwritten: scratch) -> V = L⁻¹·B (N × 8).  Needs the diagonal-block inverses in W's diagonal 256-blocks.
int kbo_i_solve_fwd(kbo_handle* h, double* B, double* V, cudaStream_t s) {
  const int N = h->N, ld = h->ld;
  const double* L = (const double*)h->K.p;
  const double* W = (const double*)h->W.p;
  for (int K0 = 0; K0 < N; K0 += SV_P) {
    const int Wd = N - K0 < SV_P ? N - K0 : SV_P;
    sv_diag_fwd_kernel<<<(Wd + 7) / 8, 256, 0, s>>>(W + (size_t)K0 * ld + K0, ld, Wd, B + (size_t)K0 * SV_R, V + (size_t)K0 * SV_R);
    KBO_LAUNCH_CHECK(h);
    const int rows = N - (K0 + Wd);
    if (rows > 0) {
      sv_update_fwd_kernel<<<(rows + 7) / 8, 256, 0, s>>>(L + (size_t)(K0 + Wd) * ld + K0, ld, rows, Wd, V + (size_t)K0 * SV_R, B + (size_t)(K0 + Wd) * SV_R);
      KBO_LAUNCH_CHECK(h);
    }
  }
  return KBO_OK;
}
// Z (N × 8, overwritten) -> A = L⁻ᵀ·Z (N × 8)
int kbo_i_solve_bwd(kbo_handle* h, double* Z, double* A, cudaStream_t s) {
  const int N = h->N, ld = h->ld;
  const double* L = (const double*)h->K.p;
  const double* W = (const double*)h->W.p;
  const int last = (N - 1) / SV_P * SV_P;
  for (int K0 = last; K0 >= 0; K0 -= SV_P) {
    const int Wd = N - K0 < SV_P ? N - K0 : SV_P;
    sv_diag_bwd_kernel<<<1, 256, 0, s>>>(W + (size_t)K0 * ld + K0, ld, Wd, Z + (size_t)K0 * SV_R, A + (size_t)K0 * SV_R);
    KBO_LAUNCH_CHECK(h);
    if (K0 > 0) {
      sv_update_bwd_kernel<<<(K0 + 255) / 256, 256, 0, s>>>(L + (size_t)K0 * ld, ld, K0, Wd, A + (size_t)K0 * SV_R, Z);
      KBO_LAUNCH_CHECK(h);
    }
  }
  return KBO_OK;
}

// alpha = L⁻ᵀ(L⁻¹·yn) into h->alpha (what fit_finish computes as Wᵀ(W·yn) when W is formed)
int kbo_i_alpha_by_solves(kbo_handle* h, cudaStream_t s) {
  const int N = h->N;
  KBO_TRY(kbo_reserve(h, h->sv_B, sizeof(double) * (size_t)(N + SV_P) * SV_R));
  KBO_TRY(kbo_reserve(h, h->sv_V, sizeof(double) * (size_t)(N + SV_P) * SV_R));
  double* B = (double*)h->sv_B.p;
  double* V = (double*)h->sv_V.p;
  KBO_CUDA(h, cudaMemsetAsync(B, 0, sizeof(double) * (size_t)N * SV_R, s));
  sv_col_kernel<<<(N + 255) / 256, 256, 0, s>>>((const double*)h->yn.p, N, 1, B, SV_R);
  KBO_LAUNCH_CHECK(h);
  KBO_TRY(kbo_i_solve_fwd(h, B, V, s));      // V[:,0] = z
  KBO_TRY(kbo_i_solve_bwd(h, V, B, s));      // B[:,0] = alpha (V is consumed)
  sv_col_kernel<<<(N + 255) / 256, 256, 0, s>>>(B, N, SV_R, (double*)h->alpha.p, 1);
  KBO_LAUNCH_CHECK(h);
  return KBO_OK;
}

// varn64[c] = amp − ‖L⁻¹ k*_c‖² for the n rows of Ks (n × ld, FP64 K*), groups of 8
int kbo_i_variance_by_solves(kbo_handle* h, const double* Ks, int n, double* varn64, cudaStream_t s) {
  const int N = h->N, ld = h->ld;
  KBO_TRY(kbo_reserve(h, h->sv_B, sizeof(double) * (size_t)(N + SV_P) * SV_R));
  KBO_TRY(kbo_reserve(h, h->sv_V, sizeof(double) * (size_t)(N + SV_P) * SV_R));
  double* B = (double*)h->sv_B.p;
  double* V = (double*)h->sv_V.p;
  for (int c0 = 0; c0 < n; c0 += SV_R) {
    const int nq = n - c0 < SV_R ? n - c0 : SV_R;
    sv_pack_kernel<<<(N + 255) / 256, 256, 0, s>>>(Ks, ld, N, c0, nq, B);
    KBO_LAUNCH_CHECK(h);
    KBO_TRY(kbo_i_solve_fwd(h, B, V, s));
    sv_sumsq_kernel<<<nq, 256, 0, s>>>(V, N, h->prm.amplitude, varn64 + c0);
    KBO_LAUNCH_CHECK(h);
  }
  return KBO_OK;
}
