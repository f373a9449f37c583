// The diagonal-block kernel of the blocked Cholesky (fit.cu), in a header so that tests/studies/potf2_probe.cu can time its
// phases in isolation.  Needs KBO_NB (= 64) defined.
#pragma once
#ifdef KBO_POTF2_PROBE
#define POTF2_MARK(i) do { if (threadIdx.x == 0 && probe) probe[i] = clock64(); } while (0)
#else
#define POTF2_MARK(i) do { } while (0)
#endif

// potf2_inv: factor one 64×64 diagonal block in shared memory and invert it.  This kernel IS the Cholesky's dependent chain
// (128 of them at N = 8192), so its latency — not its flops — is what counts.  Blocked 4 × 16:
//   (A) ONE WARP factors the 16×16 diagonal sub-block in registers (lane i holds row i; the pivot and the column travel by
//       shuffles, no barrier inside the 16 column steps) and inverts it (lane c owns column c of the inverse — no cross-lane
//       dependency);
//   (B) all threads: rows below ← rows below · Dinvᵀ;   (C) all threads: rank-16 update of the trailing block.
// The 64×64 inverse is then assembled from the four 16×16 inverses by two levels of recursive doubling (the upper triangle of
// S is the scratch).  Arithmetic per element is the right-looking column-by-column recurrence in the same order as an
// unblocked factorisation (pivot reciprocal by rsqrt, diagonal by sqrt).  Optionally also writes the inverse into W's diagonal.
// 1/sqrt(d), d a positive normal double: the hardware seed (2⁻²⁰) and one third-order step — 4 dependent operations after the
// seed, where the library routine (special cases, two Newton steps) has more than twice that on the Cholesky's critical chain
__device__ __forceinline__ double potf2_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double t = d * y;
  const double e = fma(-t, y, 1.0);          // 1 − d·y²
  const double p = fma(0.375, e, 0.5);       // y·(1 + e/2 + 3e²/8)
  return fma(y * e, p, y);
}
// The 16 column steps of the one-warp 16×16 factorisation.  Lane (i, h) = (lane & 15, lane >> 4) holds row i's columns
// c ≡ h (mod 2) in a[c >> 1]: both half-warps work, each on half of the rank-1 update.  Column J: the pivot comes by shuffle
// from its owner, every lane forms 1/sqrt, the owning half scales its column into shared memory (double-buffered by J's
// parity), and after one __syncwarp every lane updates its columns > J from two broadcast reads per element.  The warp is
// ISSUE-bound here, so the step is written for few instructions: no predicated upper-triangle guards (the strict upper
// triangle of the register tile is scratch and never stored), the diagonal by d·r plus one correction instead of sqrt().
// Unrolled by template recursion: a plain `#pragma unroll` left the column loop partly rolled and the row in local memory.
template <int J>
struct potf2_cols {
  static __device__ __forceinline__ void run(double (&a)[8], int i, int h, int kb, double* colbuf, double& rsv, int& bad) {
    constexpr int hJ = J & 1, qJ = J >> 1;
    double d = __shfl_sync(0xffffffffu, a[qJ], J + 16 * hJ);
    if (!(d > 1e-300)) {   // uniform over the warp
      if (!bad) bad = kb + J + 1;
      d = 1.0;
    }
    const double rs = potf2_rsqrt(d);
    double* cb = colbuf + 16 * hJ;
    if (h == hJ) {
      const double lij = a[qJ] * rs;
      cb[i] = lij;
      if (i == J) {
        const double sq = d * rs;
        a[qJ] = fma(fma(-sq, sq, d), 0.5 * rs, sq);   // sqrt(d)
      } else {
        a[qJ] = lij;
      }
    }
    if (i == J) rsv = rs;
    __syncwarp();
    if (J < 15) {
      const double lij = cb[i];
#pragma unroll
      for (int q = 0; q < 8; q++) {
        if (2 * q + 1 > J) {                       // columns 2q, 2q+1: at least one is > J (static)
          const int c = 2 * q + h;
          if (2 * q > J) a[q] = fma(-lij, cb[c], a[q]);            // both halves' columns are > J
          else if (h == 1) a[q] = fma(-lij, cb[c], a[q]);          // 2q == J: only the odd column 2q+1 is > J
        }
      }
    }
    potf2_cols<J + 1>::run(a, i, h, kb, colbuf, rsv, bad);
  }
};
template <>
struct potf2_cols<16> {
  static __device__ __forceinline__ void run(double (&)[8], int, int, int, double*, double&, int&) {}
};
// one output of a 16-term product row·row out of shared memory, two accumulators (halves the dependent chain)
template <int NK>
__device__ __forceinline__ double potf2_dot(const double* __restrict__ x, const double* __restrict__ y) {
  double s0 = 0.0, s1 = 0.0;
#pragma unroll
  for (int k = 0; k < NK; k += 2) {
    s0 = fma(x[k], y[k], s0);
    s1 = fma(x[k + 1], y[k + 1], s1);
  }
  return s0 + s1;
}
// recursive-doubling level of the 64×64 inverse: T21 = −T22·(L21·T11) for every pair of B-blocks; X = L21·T11 sits in S's
// upper triangle.  Full-length sums (the zeros of the triangular factors make them exact) so the loops unroll.
template <int B>
__device__ __forceinline__ void potf2_double_level(double (*S)[KBO_NB + 1], double (*T)[KBO_NB + 1], int t, int nthreads) {
  constexpr int npair = KBO_NB / (2 * B), per = B * B;
  for (int e = t; e < npair * per; e += nthreads) {
    const int pr = e / per, r0 = pr * 2 * B, aa = (e % per) / B, bb = e % B;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k < B; k += 2) {
      s0 = fma(S[r0 + B + aa][r0 + k], T[r0 + k][r0 + bb], s0);
      s1 = fma(S[r0 + B + aa][r0 + k + 1], T[r0 + k + 1][r0 + bb], s1);
    }
    S[r0 + aa][r0 + B + bb] = s0 + s1;   // X[aa][bb]
  }
  __syncthreads();
  for (int e = t; e < npair * per; e += nthreads) {
    const int pr = e / per, r0 = pr * 2 * B, aa = (e % per) / B, bb = e % B;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k < B; k += 2) {
      s0 = fma(T[r0 + B + aa][r0 + B + k], S[r0 + k][r0 + B + bb], s0);
      s1 = fma(T[r0 + B + aa][r0 + B + k + 1], S[r0 + k + 1][r0 + B + bb], s1);
    }
    T[r0 + B + aa][r0 + bb] = -(s0 + s1);
  }
  __syncthreads();
}
#define POTF2_THREADS 512
__global__ void __launch_bounds__(POTF2_THREADS) potf2_inv_kernel(double* __restrict__ A, int lda, int jb, int k_global,
                                                                  double* __restrict__ Linv, int* __restrict__ info,
                                                                  double* __restrict__ Wd = nullptr, int ldw = 0
#ifdef KBO_POTF2_PROBE
                                                                  , long long* probe = nullptr
#endif
) {
  extern __shared__ double sm[];
  double(*S)[KBO_NB + 1] = reinterpret_cast<double(*)[KBO_NB + 1]>(sm);
  double(*T)[KBO_NB + 1] = reinterpret_cast<double(*)[KBO_NB + 1]>(sm + KBO_NB * (KBO_NB + 1));
  __shared__ int s_fail;
  __shared__ __align__(16) double colbuf[32];
  const int t = threadIdx.x;
  if (*info != 0) return;  // an earlier panel already failed
  for (int e = t; e < KBO_NB * KBO_NB; e += POTF2_THREADS) {
    const int r = e >> 6, c = e & 63;
    S[r][c] = (r < jb && c <= r) ? A[(size_t)r * lda + c] : (r == c ? 1.0 : 0.0);
    T[r][c] = 0.0;
  }
  if (t == 0) s_fail = 0;
  __syncthreads();
  POTF2_MARK(0);
  for (int kb = 0; kb < KBO_NB; kb += 16) {
    if (t < 32) {
      // ---- (A) 16×16 factor + inverse, one warp ----
      const int i = t & 15, h = t >> 4;
      double a[8];
#pragma unroll
      for (int q = 0; q < 8; q++) a[q] = S[kb + i][kb + 2 * q + h];   // the upper triangle rides along as scratch
      double rsv = 0.0;
      int bad = 0;
      potf2_cols<0>::run(a, i, h, kb, colbuf, rsv, bad);
#pragma unroll
      for (int q = 0; q < 8; q++)
        if (2 * q + h <= i) S[kb + i][kb + 2 * q + h] = a[q];
      __syncwarp();
      if (kb == 0) POTF2_MARK(1);
      // inverse: lane c owns column c:  x_c = 1/L_cc ;  x_r = −(Σ_{k<r} L_rk·x_k)/L_rr  (x_k = 0 for k < c); two
      // accumulators (k even / odd) halve the dependent chain
      double x[16];
#pragma unroll
      for (int r = 0; r < 16; r++) {
        const double rsr = __shfl_sync(0xffffffffu, rsv, r);
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < r; k++)
          if ((k & 1) == 0) sacc = fma(S[kb + r][kb + k], x[k], sacc);
        double sodd = 0.0;
#pragma unroll
        for (int k = 0; k < r; k++)
          if ((k & 1) == 1) sodd = fma(S[kb + r][kb + k], x[k], sodd);
        sacc += sodd;
        x[r] = r < i ? 0.0 : (r == i ? rsr : -sacc * rsr);
      }
      if (t < 16) {
#pragma unroll
        for (int r = 0; r < 16; r++) T[kb + r][kb + i] = x[r];
      }
      if (t == 0 && bad) s_fail = bad;
    }
    __syncthreads();
    if (kb == 0) POTF2_MARK(2);
    if (s_fail) {
      if (t == 0) *info = k_global + s_fail;
      return;
    }
    const int r1 = kb + 16, nr = KBO_NB - r1;
    if (nr > 0) {
      // ---- (B) rows below: P ← P · Dinvᵀ   (P[r][c] = Σ_k P[r][k]·Dinv[c][k]; Dinv[c][k] = 0 for k > c) ----
      double pv[2] = {0.0, 0.0};
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const int e = t + u * POTF2_THREADS;
        if (e < nr * 16) pv[u] = potf2_dot<16>(&S[r1 + (e >> 4)][kb], &T[kb + (e & 15)][kb]);
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const int e = t + u * POTF2_THREADS;
        if (e < nr * 16) S[r1 + (e >> 4)][kb + (e & 15)] = pv[u];
      }
      __syncthreads();
      if (kb == 0) POTF2_MARK(3);
      // ---- (C) trailing block: S[i][c] −= Σ_k L[i][k]·L[c][k], k ascending (the column-by-column order); 2×2 register tiles
      // over the lower triangle of tiles, one tile per thread ----
      {
        const int nt = nr >> 1;
        if (t < nt * (nt + 1) / 2) {
          int ti = (int)((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
          while (ti * (ti + 1) / 2 > t) ti--;
          while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
          const int tc = t - ti * (ti + 1) / 2;
          const int i0 = r1 + 2 * ti, c0 = r1 + 2 * tc;
          double a00 = S[i0][c0], a01 = S[i0][c0 + 1], a10 = S[i0 + 1][c0], a11 = S[i0 + 1][c0 + 1];
#pragma unroll
          for (int k = 0; k < 16; k++) {
            const double li0 = S[i0][kb + k], li1 = S[i0 + 1][kb + k], lc0 = S[c0][kb + k], lc1 = S[c0 + 1][kb + k];
            a00 = fma(-li0, lc0, a00);
            a01 = fma(-li0, lc1, a01);
            a10 = fma(-li1, lc0, a10);
            a11 = fma(-li1, lc1, a11);
          }
          S[i0][c0] = a00;
          if (tc != ti) S[i0][c0 + 1] = a01;   // on a diagonal tile that element is in the upper triangle (scratch for the doubling)
          S[i0 + 1][c0] = a10;
          S[i0 + 1][c0 + 1] = a11;
        }
      }
      __syncthreads();
      if (kb == 0) POTF2_MARK(4);
    }
  }
  POTF2_MARK(5);
  // ---- the 64×64 inverse from its 16×16 diagonal inverses: T21 = −T22·(L21·T11), block sizes 16 then 32; X in S's upper triangle ----
  potf2_double_level<16>(S, T, t, POTF2_THREADS);
  POTF2_MARK(6);
  potf2_double_level<32>(S, T, t, POTF2_THREADS);
  POTF2_MARK(7);
  for (int e = t; e < KBO_NB * KBO_NB; e += POTF2_THREADS) {
    const int r = e >> 6, c = e & 63;
    if (r < jb && c <= r) A[(size_t)r * lda + c] = S[r][c];
    Linv[e] = T[r][c];
    if (Wd && r < jb && c < jb) Wd[(size_t)r * ldw + c] = T[r][c];
  }
}

