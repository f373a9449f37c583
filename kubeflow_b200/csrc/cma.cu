// (μ/μ_w, λ)-CMA-ES generation on the device, FP64 — the sampler behind Katib's `cmaes` algorithm (goptuna).
// N. Hansen, "The CMA Evolution Strategy: A Tutorial" (arXiv:1604.00772), active-weights variant as in CyberAgent `cmaes`:
//   ask : z ~ N(0,I) (Philox4x32-10 + Box–Muller, or caller-supplied), y = B·(d∘z), x = m + σ·y
//   tell: rank by fitness (ties → lower sample index), y_w, m, p_σ, σ, h_σ, p_c, C ← a·C + c1·p_c p_cᵀ + cμ·Σ w°_i y_i y_iᵀ,
//         then C = B diag(d²) Bᵀ by one-sided Jacobi warm-started from the previous basis (G = C·B_prev is already nearly
//         column-orthogonal, so one or two sweeps suffice instead of ~8 from scratch).
// Everything is tiny (D ≤ 128, λ ≤ 8192: C is 128 KiB) — the generation is launch/latency bound, so each phase is one
// kernel and the whole state lives in HBM/L2 between them; nothing returns to the host until the caller asks.
#include "kbo_internal.cuh"
#include "dgemm.cuh"

#define CMA_EPS 1e-8
#define CMA_MAXD 128

struct kbo_cma {
  int D = 0, lambda = 0, mu = 0, lam_pow2 = 0, lam_pad = 0;  // lam_pad: λ rounded up to the 256-wide split-K slices
  uint64_t seed = 0;
  long long gen = 0;
  // strategy parameters (host copies)
  double mu_eff, c1, cmu, cm, c_sigma, d_sigma, cc, chi_n, wsum;
  DevBuf mean, C, B, Dv, ps, pc, weights, Z, Zs, Y, Ys, YwT, zn2, order, R, G, scal, yw;
  bool asked = false;
};

enum CmaScal { CS_SIGMA = 0, CS_ALPHA_C = 1, CS_BEST_F = 2, CS_SWEEPS = 3, CS_COUNT = 8 };

// ------------------------------------------------------------------------------------------------ RNG
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// two standard normals per call from 128 random bits (53-bit uniforms, Box–Muller in FP64)
__global__ void cma_normal_kernel(double* __restrict__ Z, int64_t n, uint64_t seed, uint64_t gen) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // pair index
  if (2 * p >= n) return;
  uint32_t r[4];
  philox4x32_10((uint32_t)p, (uint32_t)(p >> 32), (uint32_t)gen, (uint32_t)(gen >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), r);
  const double u1 = ((((uint64_t)r[0] << 21) ^ (uint64_t)(r[1] >> 11)) + 1.0) * (1.0 / 9007199254740993.0);  // (0,1]
  const double u2 = (((uint64_t)r[2] << 21) ^ (uint64_t)(r[3] >> 11)) * (1.0 / 9007199254740992.0);          // [0,1)
  const double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  Z[2 * p] = rad * c;
  if (2 * p + 1 < n) Z[2 * p + 1] = rad * s;
}

// Zs = Z∘d, zn2[k] = ‖z_k‖²  (one warp per sample)
__global__ void cma_scale_kernel(const double* __restrict__ Z, const double* __restrict__ Dv, int lambda, int D, double* __restrict__ Zs,
                                 double* __restrict__ zn2) {
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (k >= lambda) return;
  double s = 0.0;
  for (int d = lane; d < D; d += 32) {
    const double z = Z[(size_t)k * D + d];
    Zs[(size_t)k * D + d] = z * Dv[d];
    s = fma(z, z, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) zn2[k] = s;
}
__global__ void cma_x_kernel(const double* __restrict__ Y, const double* __restrict__ mean, const double* __restrict__ scal, int64_t n, int D,
                             double* __restrict__ X) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) X[e] = fma(scal[CS_SIGMA], Y[e], mean[e % D]);
}
// synthetic fitness for the benchmark of record (cfg4): 0 = sphere, 1 = Rastrigin (one warp per sample)
__global__ void cma_fitness_kernel(const double* __restrict__ X, int lambda, int D, int kind, double* __restrict__ f) {
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (k >= lambda) return;
  double s = 0.0;
  for (int d = lane; d < D; d += 32) {
    const double x = X[(size_t)k * D + d];
    s += kind == 0 ? x * x : (x * x - 10.0 * cospi(2.0 * x) + 10.0);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) f[k] = s;
}

// ------------------------------------------------------------------------------------------------ rank
// single-CTA bitonic sort of (fitness, index) in shared memory; NaN sorts last; ties → lower index first
__global__ void __launch_bounds__(1024) cma_sort_kernel(const double* __restrict__ f, int lambda, int n2, int* __restrict__ order,
                                                        double* __restrict__ scal) {
  extern __shared__ unsigned char smc[];
  double* key = reinterpret_cast<double*>(smc);
  int* idx = reinterpret_cast<int*>(key + n2);
  for (int i = threadIdx.x; i < n2; i += 1024) {
    double v = i < lambda ? f[i] : INFINITY;
    if (v != v) v = INFINITY;
    key[i] = v;
    idx[i] = i;
  }
  __syncthreads();
  for (int k = 2; k <= n2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n2; i += 1024) {
        const int l = i ^ j;
        if (l > i) {
          const bool up = (i & k) == 0;
          const double a = key[i], b = key[l];
          const int ia = idx[i], ib = idx[l];
          const bool gt = a > b || (a == b && ia > ib);
          if (gt == up) {
            key[i] = b; key[l] = a;
            idx[i] = ib; idx[l] = ia;
          }
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < lambda; i += 1024) order[i] = idx[i];
  if (threadIdx.x == 0) scal[CS_BEST_F] = key[0];
}

// Ys[r] = Y[order[r]];  YwT[d][r] = w°_r·Ys[r][d],  w°_r = w_r (w_r ≥ 0) or w_r·n/(‖z‖²+ε)
// rows r >= lambda (padding up to lam_pad) are written as zeros so the split-K GEMM can use uniform 256-wide slices
__global__ void cma_gather_kernel(const double* __restrict__ Y, const int* __restrict__ order, const double* __restrict__ w,
                                  const double* __restrict__ zn2, int lambda, int lam_pad, int D, double* __restrict__ Ys,
                                  double* __restrict__ YwT) {
  const int r = blockIdx.x;
  const bool live = r < lambda;
  const int src = live ? order[r] : 0;
  const double wr = live ? w[r] : 0.0;
  const double wio = wr >= 0.0 ? wr : wr * (double)D / (zn2[src] + CMA_EPS);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const double v = live ? Y[(size_t)src * D + d] : 0.0;
    Ys[(size_t)r * D + d] = v;
    YwT[(size_t)d * lam_pad + r] = wio * v;
  }
}

// ------------------------------------------------------------------------------------------------ update (one CTA)
struct CmaConst {
  double mu_eff, c1, cmu, cm, c_sigma, d_sigma, cc, chi_n, wsum;
  int mu;
  long long gen;
};
__global__ void __launch_bounds__(1024) cma_update_kernel(const double* __restrict__ Ys, const double* __restrict__ w, int D, CmaConst k,
                                                          double* __restrict__ mean, const double* __restrict__ B,
                                                          const double* __restrict__ Dv, double* __restrict__ ps, double* __restrict__ pc,
                                                          double* __restrict__ scal, double* __restrict__ yw_out) {
  __shared__ double yw[CMA_MAXD], u[CMA_MAXD], red[1024];
  __shared__ double s_norm;
  const int t = threadIdx.x;
  // y_w = Σ_{r<μ} w_r Ys[r]: 1024/D threads per coordinate, fixed-order combine
  {
    const int per = 1024 / CMA_MAXD;  // 8 partial sums per coordinate
    const int d = t % CMA_MAXD, part = t / CMA_MAXD;
    double s = 0.0;
    if (d < D)
      for (int r = part; r < k.mu; r += per) s = fma(w[r], Ys[(size_t)r * D + d], s);
    red[t] = s;
    __syncthreads();
    if (t < D) {
      double a = 0.0;
      for (int p = 0; p < per; p++) a += red[p * CMA_MAXD + t];
      yw[t] = a;
      yw_out[t] = a;
    }
    __syncthreads();
  }
  const double sigma = scal[CS_SIGMA];
  if (t < D) mean[t] = fma(k.cm * sigma, yw[t], mean[t]);
  // C^{-1/2} y_w = B·((Bᵀ y_w)/d)
  if (t < D) {
    double a = 0.0;
    for (int i = 0; i < D; i++) a = fma(B[(size_t)i * D + t], yw[i], a);
    u[t] = a / Dv[t];
  }
  __syncthreads();
  if (t < D) {
    double a = 0.0;
    for (int j = 0; j < D; j++) a = fma(B[(size_t)t * D + j], u[j], a);
    const double v = (1.0 - k.c_sigma) * ps[t] + sqrt(k.c_sigma * (2.0 - k.c_sigma) * k.mu_eff) * a;
    ps[t] = v;
    red[t] = v * v;
  }
  __syncthreads();
  if (t == 0) {
    double a = 0.0;
    for (int i = 0; i < D; i++) a += red[i];
    s_norm = sqrt(a);
  }
  __syncthreads();
  const double norm_ps = s_norm;
  const double h_left = norm_ps / sqrt(1.0 - pow(1.0 - k.c_sigma, 2.0 * (double)(k.gen + 1)));
  const double h_sigma = h_left < (1.4 + 2.0 / (D + 1.0)) * k.chi_n ? 1.0 : 0.0;
  if (t < D) pc[t] = (1.0 - k.cc) * pc[t] + h_sigma * sqrt(k.cc * (2.0 - k.cc) * k.mu_eff) * yw[t];
  if (t == 0) {
    const double delta_h = (1.0 - h_sigma) * k.cc * (2.0 - k.cc);
    scal[CS_ALPHA_C] = 1.0 + k.c1 * delta_h - k.c1 - k.cmu * k.wsum;
    scal[CS_SIGMA] = sigma * exp((k.c_sigma / k.d_sigma) * (norm_ps / k.chi_n - 1.0));
  }
}
// C ← a·C + c1·p_c p_cᵀ + cμ·R, symmetrised ((C+Cᵀ)/2 as the reference does before its eigendecomposition)
// R arrives as `nsl` split-K partial sums (D×D each), added here in slice order (deterministic)
__global__ void cma_c_finalize_kernel(double* __restrict__ C, const double* __restrict__ R, int nsl, const double* __restrict__ pc,
                                      const double* __restrict__ scal, int D, double c1, double cmu) {
  const int i = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= D || j > i) return;
  const double a = scal[CS_ALPHA_C];
  double rij = 0.0, rji = 0.0;
  for (int z = 0; z < nsl; z++) {
    rij += R[(size_t)z * D * D + (size_t)i * D + j];
    rji += R[(size_t)z * D * D + (size_t)j * D + i];
  }
  const double v = a * 0.5 * (C[(size_t)i * D + j] + C[(size_t)j * D + i]) + c1 * pc[i] * pc[j] + cmu * 0.5 * (rij + rji);
  C[(size_t)i * D + j] = v;
  C[(size_t)j * D + i] = v;
}

// ------------------------------------------------------------------------------------------------ eigendecomposition
// One-sided (Hestenes) Jacobi on the columns of G = C·B_prev in shared memory (column-major, one column contiguous).
// Round-robin ordering: D/2 disjoint column pairs per step, 16 threads per pair.  On exit column j of G is λ_j·b_j.
// Rotations stop at |g_a·g_b| ≤ 1e-13·‖g_a‖‖g_b‖ (orthogonality of B to ~1e-13; the last quadratic sweep to 1e-16 buys nothing).
// DT = compile-time D (128: fully unrolled 8-row loops, loads pipelined) or 0 = runtime D.
template <int DT>
__global__ void __launch_bounds__(1024) cma_jacobi_kernel(const double* __restrict__ G, int Drt, double* __restrict__ B, double* __restrict__ Dv,
                                                          double* __restrict__ scal, int max_sweeps) {
  const int D = DT > 0 ? DT : Drt;
  extern __shared__ double Gt[];  // [Dp][D+1], Dp = D rounded up to even
  __shared__ int s_rot;
  const int Dp = (D + 1) & ~1, ldg = D + 1;
  const int t = threadIdx.x, pair = t >> 4, q = t & 15, npairs = Dp >> 1;
  for (int e = t; e < Dp * D; e += 1024) {
    const int j = e / D, i = e % D;
    Gt[j * ldg + i] = j < D ? G[(size_t)i * D + j] : 0.0;
  }
  __syncthreads();
  int sweep = 0;
  for (; sweep < max_sweeps; sweep++) {
    if (t == 0) s_rot = 0;
    __syncthreads();
    for (int step = 0; step < Dp - 1; step++) {
      for (int pk0 = 0; pk0 < npairs; pk0 += 64) {   // warp-uniform trip count: every lane takes part in the shuffles
        const int pk = pk0 + pair;
        const bool live = pk < npairs;
        int a = 0, b = 0;
        if (live) {
          if (pk == 0) {
            a = Dp - 1;
            b = step;
          } else {
            a = (step + pk) % (Dp - 1);
            b = (step - pk + Dp - 1) % (Dp - 1);
          }
        }
        const bool real = live && a < D && b < D;
        double* ga = Gt + a * ldg;
        double* gb = Gt + b * ldg;
        double al = 0.0, be = 0.0, ga_ = 0.0;
        if (real) {
#pragma unroll
          for (int i = q; i < D; i += 16) {
            const double x = ga[i], y = gb[i];
            al = fma(x, x, al);
            be = fma(y, y, be);
            ga_ = fma(x, y, ga_);
          }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          al += __shfl_xor_sync(0xffffffffu, al, o);
          be += __shfl_xor_sync(0xffffffffu, be, o);
          ga_ += __shfl_xor_sync(0xffffffffu, ga_, o);
        }
        if (real && fabs(ga_) > 1e-13 * sqrt(al * be) && al > 0.0 && be > 0.0) {
          const double zeta = (be - al) / (2.0 * ga_);
          const double tt = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double c = rsqrt(1.0 + tt * tt), s = c * tt;
#pragma unroll
          for (int i = q; i < D; i += 16) {
            const double x = ga[i], y = gb[i];
            ga[i] = c * x - s * y;
            gb[i] = s * x + c * y;
          }
          if (q == 0) s_rot = 1;
        }
      }
      __syncthreads();
    }
    const int rotated = s_rot;
    __syncthreads();
    if (!rotated) break;
  }
  // column norms are the eigenvalues of C (SPD); normalised columns the eigenvectors
  for (int j0 = 0; j0 < D; j0 += 64) {
    const int j = j0 + pair;
    double n2 = 0.0;
    if (j < D)
      for (int i = q; i < D; i += 16) n2 = fma(Gt[j * ldg + i], Gt[j * ldg + i], n2);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    const double lam = sqrt(n2);
    if (j < D && lam > 1e-300) {
      for (int i = q; i < D; i += 16) B[(size_t)i * D + j] = Gt[j * ldg + i] / lam;
      if (q == 0) Dv[j] = sqrt(lam < CMA_EPS ? CMA_EPS : lam);
    }
  }
  if (t == 0) scal[CS_SWEEPS] = (double)(sweep + 1);
}

__global__ void cma_init_kernel(double* __restrict__ C, double* __restrict__ B, double* __restrict__ Dv, double* __restrict__ ps,
                                double* __restrict__ pc, int D) {
  const int i = blockIdx.x, j = threadIdx.x;
  if (j < D) {
    C[(size_t)i * D + j] = i == j ? 1.0 : 0.0;
    B[(size_t)i * D + j] = i == j ? 1.0 : 0.0;
  }
  if (j == 0) {
    Dv[i] = 1.0;
    ps[i] = 0.0;
    pc[i] = 0.0;
  }
}

// ------------------------------------------------------------------------------------------------ host side
static int cma_generation_tell(kbo_handle* h, kbo_cma* c, const double* fitness, cudaStream_t s) {
  const int D = c->D, lam = c->lambda;
  const int smem_sort = c->lam_pow2 * 12;
  cma_sort_kernel<<<1, 1024, smem_sort, s>>>(fitness, lam, c->lam_pow2, (int*)c->order.p, (double*)c->scal.p);
  KBO_LAUNCH_CHECK(h);
  cma_gather_kernel<<<c->lam_pad, 128, 0, s>>>((const double*)c->Y.p, (const int*)c->order.p, (const double*)c->weights.p, (const double*)c->zn2.p, lam,
                                               c->lam_pad, D, (double*)c->Ys.p, (double*)c->YwT.p);
  KBO_LAUNCH_CHECK(h);
  CmaConst k{c->mu_eff, c->c1, c->cmu, c->cm, c->c_sigma, c->d_sigma, c->cc, c->chi_n, c->wsum, c->mu, c->gen};
  cma_update_kernel<<<1, 1024, 0, s>>>((const double*)c->Ys.p, (const double*)c->weights.p, D, k, (double*)c->mean.p, (const double*)c->B.p,
                                       (const double*)c->Dv.p, (double*)c->ps.p, (double*)c->pc.p, (double*)c->scal.p, (double*)c->yw.p);
  KBO_LAUNCH_CHECK(h);
  // R = Σ w°_i y_i y_iᵀ  (D×D = YwT[D×λ]·Ys[λ×D]) as split-K over 256-sample slices: a 128×128 output is only two
  // tiles of the GEMM, so without the split a K = 4096 product ran on 2 CTAs (337 µs); the slices fill 2·λ/256 CTAs
  const int nsl = c->lam_pad / 256;
  dgemm64_launch<false, EPI_STORE>(s, D, D, 256, (const double*)c->YwT.p, c->lam_pad, (const double*)c->Ys.p, D, (double*)c->R.p, D, 1.0, 0.0, KM_FULL,
                                   0, TS_NONE, nsl, 256, 256LL * D, (long long)D * D);
  KBO_LAUNCH_CHECK(h);
  dim3 g((D + 127) / 128, D);
  cma_c_finalize_kernel<<<g, 128, 0, s>>>((double*)c->C.p, (const double*)c->R.p, nsl, (const double*)c->pc.p, (const double*)c->scal.p, D, c->c1,
                                          c->cmu);
  KBO_LAUNCH_CHECK(h);
  // G = C·B_prev, then Jacobi
  dgemm64_launch<false, EPI_STORE>(s, D, D, D, (const double*)c->C.p, D, (const double*)c->B.p, D, (double*)c->G.p, D, 1.0, 0.0, KM_FULL, 0, TS_NONE);
  KBO_LAUNCH_CHECK(h);
  const int Dp = (D + 1) & ~1;
  if (D == 128)
    cma_jacobi_kernel<128><<<1, 1024, sizeof(double) * Dp * (D + 1), s>>>((const double*)c->G.p, D, (double*)c->B.p, (double*)c->Dv.p, (double*)c->scal.p, 30);
  else
    cma_jacobi_kernel<0><<<1, 1024, sizeof(double) * Dp * (D + 1), s>>>((const double*)c->G.p, D, (double*)c->B.p, (double*)c->Dv.p, (double*)c->scal.p, 30);
  KBO_LAUNCH_CHECK(h);
  c->gen++;
  c->asked = false;
  return KBO_OK;
}

static int cma_generation_ask(kbo_handle* h, kbo_cma* c, double* X, const double* z_in, cudaStream_t s) {
  const int D = c->D, lam = c->lambda;
  const int64_t n = (int64_t)lam * D;
  const double* Z = z_in;
  if (!Z) {
    cma_normal_kernel<<<(unsigned)((n / 2 + 1 + 255) / 256), 256, 0, s>>>((double*)c->Z.p, n, c->seed, (uint64_t)c->gen);
    KBO_LAUNCH_CHECK(h);
    Z = (const double*)c->Z.p;
  }
  cma_scale_kernel<<<(lam + 7) / 8, 256, 0, s>>>(Z, (const double*)c->Dv.p, lam, D, (double*)c->Zs.p, (double*)c->zn2.p);
  KBO_LAUNCH_CHECK(h);
  // Y[k,i] = Σ_j Zs[k,j]·B[i,j]
  dgemm64_launch<true, EPI_STORE>(s, lam, D, D, (const double*)c->Zs.p, D, (const double*)c->B.p, D, (double*)c->Y.p, D, 1.0, 0.0, KM_FULL, 0, TS_NONE);
  KBO_LAUNCH_CHECK(h);
  cma_x_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const double*)c->Y.p, (const double*)c->mean.p, (const double*)c->scal.p, n, D, X);
  KBO_LAUNCH_CHECK(h);
  c->asked = true;
  return KBO_OK;
}

extern "C" {

int kbo_cma_create(kbo_handle* h, kbo_cma** out, int32_t D, int32_t lambda, const double* mean0, double sigma0, uint64_t seed) {
  if (!h || !out) return KBO_ERR_INVALID;
  *out = nullptr;
  if (D < 1 || D > CMA_MAXD) KBO_FAIL(h, KBO_ERR_INVALID, "kbo_cma_create: 1 <= D <= %d supported (got %d)", CMA_MAXD, D);
  if (lambda < 4 || lambda > 8192) KBO_FAIL(h, KBO_ERR_INVALID, "kbo_cma_create: 4 <= lambda <= 8192 supported (got %d)", lambda);
  if (!mean0 || !(sigma0 > 0.0)) KBO_FAIL(h, KBO_ERR_INVALID, "kbo_cma_create: mean0 must be given and sigma0 > 0");
  KBO_CUDA(h, cudaSetDevice(h->device));
  kbo_cma* c = new (std::nothrow) kbo_cma();
  if (!c) return KBO_ERR_NOMEM;
  c->D = D;
  c->lambda = lambda;
  c->mu = lambda / 2;
  c->seed = seed;
  int p2 = 1;
  while (p2 < lambda) p2 <<= 1;
  c->lam_pow2 = p2;
  c->lam_pad = round_up(lambda, 256);
  // strategy parameters (tutorial Table 1; active weights as in `cmaes`)
  std::vector<double> w(lambda);
  const int mu = c->mu, n = D;
  double s1 = 0, s2 = 0, m1 = 0, m2 = 0;
  for (int i = 0; i < lambda; i++) {
    w[i] = log((lambda + 1) / 2.0) - log((double)(i + 1));
    if (i < mu) { s1 += w[i]; s2 += w[i] * w[i]; } else { m1 += w[i]; m2 += w[i] * w[i]; }
  }
  const double mu_eff = s1 * s1 / s2, mu_eff_minus = m1 * m1 / m2, alpha_cov = 2.0;
  const double c1 = alpha_cov / ((n + 1.3) * (n + 1.3) + mu_eff);
  double cmu = alpha_cov * (mu_eff - 2 + 1 / mu_eff) / ((n + 2.0) * (n + 2.0) + alpha_cov * mu_eff / 2);
  if (cmu > 1 - c1 - 1e-8) cmu = 1 - c1 - 1e-8;
  double min_alpha = 1 + c1 / cmu;
  if (1 + (2 * mu_eff_minus) / (mu_eff + 2) < min_alpha) min_alpha = 1 + (2 * mu_eff_minus) / (mu_eff + 2);
  if ((1 - c1 - cmu) / (n * cmu) < min_alpha) min_alpha = (1 - c1 - cmu) / (n * cmu);
  double pos = 0, neg = 0;
  for (int i = 0; i < lambda; i++) (w[i] > 0 ? pos : neg) += fabs(w[i]);
  double wsum = 0;
  for (int i = 0; i < lambda; i++) {
    w[i] = w[i] >= 0 ? w[i] / pos : min_alpha / neg * w[i];
    wsum += w[i];
  }
  c->mu_eff = mu_eff; c->c1 = c1; c->cmu = cmu; c->cm = 1.0; c->wsum = wsum;
  c->c_sigma = (mu_eff + 2) / (n + mu_eff + 5);
  const double t = sqrt((mu_eff - 1) / (n + 1)) - 1;
  c->d_sigma = 1 + 2 * (t > 0 ? t : 0) + c->c_sigma;
  c->cc = (4 + mu_eff / n) / (n + 4 + 2 * mu_eff / n);
  c->chi_n = sqrt((double)n) * (1.0 - 1.0 / (4.0 * n) + 1.0 / (21.0 * n * n));
  const size_t LD = (size_t)lambda * D, LPD = (size_t)c->lam_pad * D;
  int r = KBO_OK;
  auto R_ = [&](DevBuf& b, size_t bytes) { if (r == KBO_OK) r = kbo_reserve(h, b, bytes); };
  R_(c->mean, 8 * D); R_(c->C, 8 * (size_t)D * D); R_(c->B, 8 * (size_t)D * D); R_(c->Dv, 8 * D); R_(c->ps, 8 * D); R_(c->pc, 8 * D);
  R_(c->weights, 8 * lambda); R_(c->Z, 8 * (LD + 2)); R_(c->Zs, 8 * LD); R_(c->Y, 8 * LD); R_(c->Ys, 8 * LPD); R_(c->YwT, 8 * LPD);
  R_(c->zn2, 8 * lambda); R_(c->order, 4 * lambda); R_(c->R, 8 * (size_t)D * D * (c->lam_pad / 256)); R_(c->G, 8 * (size_t)D * D); R_(c->scal, 8 * CS_COUNT);
  R_(c->yw, 8 * D);
  if (r != KBO_OK) { delete c; return r; }
  double sc[CS_COUNT] = {sigma0, 0, 0, 0, 0, 0, 0, 0};
  cudaMemcpy(c->mean.p, mean0, 8 * D, cudaMemcpyHostToDevice);
  cudaMemcpy(c->weights.p, w.data(), 8 * lambda, cudaMemcpyHostToDevice);
  cudaMemcpy(c->scal.p, sc, sizeof sc, cudaMemcpyHostToDevice);
  cma_init_kernel<<<D, 128>>>((double*)c->C.p, (double*)c->B.p, (double*)c->Dv.p, (double*)c->ps.p, (double*)c->pc.p, D);
  cudaFuncSetAttribute(cma_jacobi_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * CMA_MAXD * (CMA_MAXD + 1)));
  cudaFuncSetAttribute(cma_jacobi_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * CMA_MAXD * (CMA_MAXD + 1)));
  cudaFuncSetAttribute(cma_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 12);
  KBO_CUDA(h, cudaDeviceSynchronize());
  *out = c;
  return KBO_OK;
}

void kbo_cma_destroy(kbo_cma* c) {
  if (!c) return;
  DevBuf* bufs[] = {&c->mean, &c->C, &c->B, &c->Dv, &c->ps, &c->pc, &c->weights, &c->Z, &c->Zs, &c->Y, &c->Ys, &c->YwT, &c->zn2, &c->order,
                    &c->R, &c->G, &c->scal, &c->yw};
  for (DevBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  delete c;
}

int kbo_cma_ask(kbo_handle* h, kbo_cma* c, double* X_dev, const double* z_in_dev, void* stream) {
  if (!h || !c) return KBO_ERR_INVALID;
  if (!X_dev) KBO_FAIL(h, KBO_ERR_INVALID, "kbo_cma_ask: null output");
  return cma_generation_ask(h, c, X_dev, z_in_dev, (cudaStream_t)stream);
}

int kbo_cma_tell(kbo_handle* h, kbo_cma* c, const double* fitness_dev, void* stream) {
  if (!h || !c) return KBO_ERR_INVALID;
  if (!fitness_dev) KBO_FAIL(h, KBO_ERR_INVALID, "kbo_cma_tell: null fitness");
  if (!c->asked) KBO_FAIL(h, KBO_ERR_STATE, "kbo_cma_tell: call kbo_cma_ask first (tell ranks the samples of the last ask)");
  return cma_generation_tell(h, c, fitness_dev, (cudaStream_t)stream);
}

int kbo_cma_state(kbo_handle* h, kbo_cma* c, double* mean, double* sigma, double* C, double* p_sigma, double* p_c, double* B, double* Dv,
                  double* Y_last, int64_t* generation, void* stream) {
  if (!h || !c) return KBO_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t D = c->D;
  if (mean) KBO_CUDA(h, cudaMemcpyAsync(mean, c->mean.p, 8 * D, cudaMemcpyDeviceToHost, s));
  if (sigma) KBO_CUDA(h, cudaMemcpyAsync(sigma, c->scal.p, 8, cudaMemcpyDeviceToHost, s));
  if (C) KBO_CUDA(h, cudaMemcpyAsync(C, c->C.p, 8 * D * D, cudaMemcpyDeviceToHost, s));
  if (p_sigma) KBO_CUDA(h, cudaMemcpyAsync(p_sigma, c->ps.p, 8 * D, cudaMemcpyDeviceToHost, s));
  if (p_c) KBO_CUDA(h, cudaMemcpyAsync(p_c, c->pc.p, 8 * D, cudaMemcpyDeviceToHost, s));
  if (B) KBO_CUDA(h, cudaMemcpyAsync(B, c->B.p, 8 * D * D, cudaMemcpyDeviceToHost, s));
  if (Dv) KBO_CUDA(h, cudaMemcpyAsync(Dv, c->Dv.p, 8 * D, cudaMemcpyDeviceToHost, s));
  if (Y_last) KBO_CUDA(h, cudaMemcpyAsync(Y_last, c->Y.p, 8 * D * c->lambda, cudaMemcpyDeviceToHost, s));
  KBO_CUDA(h, cudaStreamSynchronize(s));
  if (generation) *generation = c->gen;
  return KBO_OK;
}

int kbo_cma_run_synthetic(kbo_handle* h, kbo_cma* c, int32_t fitness_kind, int32_t generations, double* best_f_host, float* elapsed_ms,
                          double* jacobi_sweeps_last) {
  if (!h || !c) return KBO_ERR_INVALID;
  if (fitness_kind < 0 || fitness_kind > 1 || generations < 1) KBO_FAIL(h, KBO_ERR_INVALID, "kbo_cma_run_synthetic: bad argument");
  cudaStream_t s = 0;
  DevBuf X{}, f{};
  KBO_TRY(kbo_reserve(h, X, 8 * (size_t)c->lambda * c->D));
  KBO_TRY(kbo_reserve(h, f, 8 * (size_t)c->lambda));
  cudaEventRecord(h->ev[5], s);
  int r = KBO_OK;
  for (int g = 0; g < generations && r == KBO_OK; g++) {
    r = cma_generation_ask(h, c, (double*)X.p, nullptr, s);
    if (r != KBO_OK) break;
    cma_fitness_kernel<<<(c->lambda + 7) / 8, 256, 0, s>>>((const double*)X.p, c->lambda, c->D, fitness_kind, (double*)f.p);
    h->launches++;
    r = cma_generation_tell(h, c, (const double*)f.p, s);
  }
  cudaEventRecord(h->ev[6], s);
  double sc[CS_COUNT];
  cudaMemcpyAsync(sc, c->scal.p, sizeof sc, cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  cudaFree(X.p);
  cudaFree(f.p);
  if (r != KBO_OK) return r;
  if (e != cudaSuccess) KBO_FAIL(h, KBO_ERR_CUDA, "kbo_cma_run_synthetic: %s", cudaGetErrorString(e));
  if (best_f_host) *best_f_host = sc[CS_BEST_F];
  if (jacobi_sweeps_last) *jacobi_sweeps_last = sc[CS_SWEEPS];
  if (elapsed_ms) cudaEventElapsedTime(elapsed_ms, h->ev[5], h->ev[6]);
  return KBO_OK;
}

}  // extern "C"
