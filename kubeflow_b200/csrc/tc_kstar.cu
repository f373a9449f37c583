// K* = amp·k(Xc, X) on the tensor cores, for the RANKING pass of the sweep ($SK/_gpr.py:446, $SK/kernels.py:1559-1570 RBF,
// :1713-1729 Matérn-5/2).  sm_100a only: tcgen05.mma + TMEM + TMA.
//
// The FP64 cross kernel (sweep.cu) spends ~75 FP64 operations per (candidate, trial) pair: 59 ms per suggestion at cfg3, on a
// pass whose output the ranking pass then rounds to fp16.  Here the pairwise distances are the dense product they are:
//     d²(c, x) = |c|² + |x|² − 2 c·x ,   c·x = Σ_d ĉ_d x̂_d   (ĉ, x̂: coordinates / ℓ_d, centred on the trials' mean, pre-scaled
//                                                              by √5 (Matérn) or 1/√2 (RBF) so the epilogue needs no constant)
// with c·x on the tensor cores.  Coordinates are split into fp16 hi + lo (|x̂ − hi − lo| ≤ 2⁻²²|x̂|) and all four partial
// products are accumulated in fp32 TMEM (small ones first), so c·x carries ~1e-7 relative error — the kernel value is then
// evaluated in fp32 (MUFU sqrt / ex2) per element by the epilogue warps, written ONCE as the fp16 hi plane the ranking kernel
// consumes (no lo plane, no FP64), and the normalised mean μ̃ = Σ_n K̃*[m,n]·alpha[n] is accumulated on the fly (four fp32 FMA
// chains of four trials, summed and added in FP64 once per 16 trials).  Both μ̃ and the variance built from this plane are only used to RANK: their error
// is measured per sweep on stratified calibration rows against the FP64 path and every candidate that could still be the
// maximum is re-evaluated in FP64 (sweep.cu).
//
// CTA = 128 candidate rows (TMEM lanes); it walks all 256-trial tiles.  Candidate planes (128 × Dp, hi and lo) stay in shared
// memory; trial planes stream through a 3-stage TMA ring one 32-feature slab at a time (SWIZZLE_64B);
//   warp 0      TMA producer
//   warp 1      MMA issuer (one thread): per (tile, slab) 2 k-steps × 4 products of M128 N256 K16 → one of two TMEM accumulators
//   warps 2-17  epilogue: tcgen05.ld 16 columns → d² → kernel → fp16 → one 32-byte store per thread per batch; μ̃ partials
// Per 128×256 tile the epilogue (~12 instructions and 2 MUFU per element: MUFU-bound, profiles/r11_tckstar_cfg3_summary.csv) is the
// limiter, the MMAs take a sixth of that.
#include "kbo_internal.cuh"
#include <type_traits>

#include "tc_common.cuh"

#define KS_BM 128
#define KS_BN 256
#define KS_BK 32
#define KS_STAGES 3
#define KS_EPI_WARPS 16
#define KS_THREADS (64 + 32 * KS_EPI_WARPS)
#define KS_A_PLANE_BYTES (KS_BM * KS_BK * 2)    // 8 KB: one plane of one slab of the candidate tile
#define KS_B_PLANE_BYTES (KS_BN * KS_BK * 2)    // 16 KB
#define KS_STAGE_BYTES (2 * KS_B_PLANE_BYTES)   // hi + lo
#define KS_MAX_SLABS 4                          // D <= 128

namespace {
using namespace tcx;

struct KsSmem {
  uint64_t full[KS_STAGES];
  uint64_t empty[KS_STAGES];
  uint64_t a_full;
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

static size_t ks_smem_bytes(int ns) {
  return 1024 + (size_t)ns * 2 * KS_A_PLANE_BYTES + (size_t)KS_STAGES * KS_STAGE_BYTES + (size_t)KS_EPI_WARPS * 2 * 512 + 4 * KS_BM * sizeof(double) +
         256;
}

__device__ __forceinline__ float ex2_ftz(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sqrt_abs_ftz(float x) {   // MUFU.SQRT |x|: one instruction where max + rsqrt + multiply were three
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fabsf(x)));
  return r;
}
// kernel value from the dot product: KIND 0 = RBF (coordinates pre-scaled by 1/√2: k = exp(−d2)), 1 = Matérn-5/2 (pre-scaled by
// √5: s = √d2, k = (1 + s + s²/3)·exp(−s)).  log2amp folds the amplitude into the exponent.
template <int KIND>
__device__ __forceinline__ float ks_kernel_value(float dot, float nsum, float log2amp) {
  float d2 = fmaf(-2.f, dot, nsum);
  if (KIND == 0) {
    d2 = fmaxf(d2, 0.f);
    return ex2_ftz(fmaf(d2, -1.4426950408889634f, log2amp));
  }
  const float s = sqrt_abs_ftz(d2);   // d2 < 0 only by rounding (|d2| tiny): its magnitude is as good an estimate as 0
  const float p = fmaf(fmaf(s, 0.33333334f, 1.f), s, 1.f);
  return p * ex2_ftz(fmaf(s, -1.4426950408889634f, log2amp));
}

template <int KIND>
__global__ void __launch_bounds__(KS_THREADS, 1)
tc_kstar_kernel(const __grid_constant__ CUtensorMap tmCh, const __grid_constant__ CUtensorMap tmCl, const __grid_constant__ CUtensorMap tmXh,
                const __grid_constant__ CUtensorMap tmXl, int ns, int ntiles, int N, const float* __restrict__ ncand,
                const float4* __restrict__ nxal4 /* (|x̂_n|², alpha_n) pairs, two trials per float4 */, float log2amp,
                __half* __restrict__ Ksh, int Npad, float* __restrict__ mun, int store_tiles /* trial tiles whose K* columns are written */) {
  extern __shared__ unsigned char ks_smem_raw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)ks_smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* smA = base;                                            // [ns][hi 8 KB | lo 8 KB]
  unsigned char* ring = smA + (size_t)ns * 2 * KS_A_PLANE_BYTES;        // [KS_STAGES][hi 16 KB | lo 16 KB]
  float4* nxs = (float4*)(ring + KS_STAGES * KS_STAGE_BYTES);           // [16 warps][2][32 lanes]
  double* musum = (double*)(nxs + KS_EPI_WARPS * 2 * 32);               // [4][128]
  KsSmem* S = (KsSmem*)(musum + 4 * KS_BM);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * KS_BM;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmCh);
    prefetch_tmap(&tmCl);
    prefetch_tmap(&tmXh);
    prefetch_tmap(&tmXl);
    for (int i = 0; i < KS_STAGES; i++) {
      mbar_init(smem_u32(&S->full[i]), 1);
      mbar_init(smem_u32(&S->empty[i]), 1);
    }
    mbar_init(smem_u32(&S->a_full), 1);
    for (int i = 0; i < 2; i++) {
      mbar_init(smem_u32(&S->tmem_full[i]), 1);
      mbar_init(smem_u32(&S->tmem_empty[i]), KS_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S->tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = S->tmem_base;

  if (warp == 0) {
    // ===================================== TMA producer ===========================================
    if (lane == 0) {
      const uint32_t abar = smem_u32(&S->a_full);
      mbar_expect_tx(abar, (uint32_t)ns * 2 * KS_A_PLANE_BYTES);
      for (int sl = 0; sl < ns; sl++) {
        tma_load_2d(smem_u32(smA + (size_t)sl * 2 * KS_A_PLANE_BYTES), &tmCh, abar, sl * KS_BK, m0);
        tma_load_2d(smem_u32(smA + (size_t)sl * 2 * KS_A_PLANE_BYTES + KS_A_PLANE_BYTES), &tmCl, abar, sl * KS_BK, m0);
      }
      uint32_t c = 0;
      for (int t = 0; t < ntiles; t++)
        for (int sl = 0; sl < ns; sl++, c++) {
          const uint32_t st = c % KS_STAGES, use = c / KS_STAGES;
          mbar_wait(smem_u32(&S->empty[st]), (use & 1) ^ 1, 21);
          const uint32_t bar = smem_u32(&S->full[st]);
          mbar_expect_tx(bar, KS_STAGE_BYTES);
          const uint32_t dst = smem_u32(ring + (size_t)st * KS_STAGE_BYTES);
          tma_load_2d(dst, &tmXh, bar, sl * KS_BK, t * KS_BN);
          tma_load_2d(dst + KS_B_PLANE_BYTES, &tmXl, bar, sl * KS_BK, t * KS_BN);
        }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================== MMA issuer =============================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(KS_BM, KS_BN);
      mbar_wait(smem_u32(&S->a_full), 0, 22);
      tc_fence_after();
      uint32_t c = 0;
      for (int t = 0; t < ntiles; t++) {
        const uint32_t buf = t & 1;
        mbar_wait(smem_u32(&S->tmem_empty[buf]), ((t >> 1) & 1) ^ 1, 23);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * KS_BN;
        for (int sl = 0; sl < ns; sl++, c++) {
          const uint32_t st = c % KS_STAGES, suse = c / KS_STAGES;
          mbar_wait(smem_u32(&S->full[st]), suse & 1, 24);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smA + (size_t)sl * 2 * KS_A_PLANE_BYTES), a_lo = a_hi + KS_A_PLANE_BYTES;
          const uint32_t b_hi = smem_u32(ring + (size_t)st * KS_STAGE_BYTES), b_lo = b_hi + KS_B_PLANE_BYTES;
          // small products first: the accumulator is still tiny when they are added, so their bits are not truncated away
#pragma unroll
          for (int prod = 0; prod < 4; prod++) {
            const uint32_t a = (prod < 2) ? a_lo : a_hi;
            const uint32_t b = (prod & 1) ? b_hi : b_lo;   // lo·lo, lo·hi, hi·lo, hi·hi
#pragma unroll
            for (int k = 0; k < KS_BK / 16; k++)
              umma_f16(d_tmem, umma_desc_sw64(a + k * 32), umma_desc_sw64(b + k * 32), idesc, (sl > 0 || prod > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(smem_u32(&S->empty[st]));
        }
        umma_commit(smem_u32(&S->tmem_full[buf]));
      }
    }
    __syncwarp();
  } else {
    // ===================================== epilogue warps =========================================
    const int ew = warp - 2;
    const int quarter = warp & 3;        // TMEM lane quarter this warp may access
    const int cg = ew >> 2;              // which 64 of the tile's 256 trial columns
    const int row = quarter * 32 + lane;
    const float ncr = ncand[m0 + row];
    float4* slot = nxs + (size_t)ew * 2 * 32;
    __half* out_row = Ksh + (size_t)(m0 + row) * Npad + cg * 64;
    double mu_d = 0.0;
    float4 pf = nxal4[(size_t)(cg * 64) / 2 + lane];
    for (int t = 0; t < ntiles; t++) {
      const uint32_t buf = t & 1;
      slot[(t & 1) * 32 + lane] = pf;
      __syncwarp();
      if (t + 1 < ntiles) pf = nxal4[(size_t)((t + 1) * KS_BN + cg * 64) / 2 + lane];
      const float4* sl4 = slot + (t & 1) * 32;
      mbar_wait(smem_u32(&S->tmem_full[buf]), (t >> 1) & 1, 25);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * KS_BN + cg * 64;
      const int n_base = t * KS_BN + cg * 64;
      const bool edge = n_base + 64 > N;   // warp-uniform: only the last tile(s) mask trials >= N
      // EDGE: only the last tile(s) mask trials >= N.  Two instantiations behind one warp-uniform branch: as a runtime flag inside the
      // body the mask compiled to ISETP + FSEL + IADD per element — 3 of 15 instructions on every tile.
      auto tile_body = [&](auto edge_tag) {
        constexpr bool EDGE = decltype(edge_tag)::value;
        uint32_t rr[2][16];
        tmem_ld16(taddr, rr[0]);
#pragma unroll
        for (int q = 0; q < 4; q++) {
          uint32_t(&r)[16] = rr[q & 1];
          tmem_ld_wait();
          if (q < 3) tmem_ld16(taddr + (q + 1) * 16, rr[(q + 1) & 1]);   // the next 16 columns travel while these are evaluated
          if (q == 3) {   // everything this warp needs from the accumulator is in registers: hand it back to the MMA thread
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&S->tmem_empty[buf]));
          }
          uint32_t o[8];
          float macc[4];   // four independent fp32 chains of four trials each, summed in FP64 once per 16 trials
#pragma unroll
          for (int g = 0; g < 4; g++) {
            macc[g] = 0.f;
#pragma unroll
            for (int e = 0; e < 4; e += 2) {
              const int i = g * 4 + e;
              const float4 na = sl4[(q * 16 + i) >> 1];   // broadcast: every lane reads the same word
              float k0 = ks_kernel_value<KIND>(__uint_as_float(r[i]), ncr + na.x, log2amp);
              float k1 = ks_kernel_value<KIND>(__uint_as_float(r[i + 1]), ncr + na.z, log2amp);
              if (EDGE) {
                k0 = (n_base + q * 16 + i < N) ? k0 : 0.f;
                k1 = (n_base + q * 16 + i + 1 < N) ? k1 : 0.f;
              }
              macc[g] = fmaf(k0, na.y, macc[g]);
              macc[g] = fmaf(k1, na.w, macc[g]);
              const __half2 h2 = __floats2half2_rn(k0, k1);
              o[i >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
            }
          }
          mu_d += (double)((macc[0] + macc[1]) + (macc[2] + macc[3]));
          if (t < store_tiles)   // a pruning pass contracts a prefix of the trials only: the mean needs every column, the plane does not
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(out_row + (size_t)t * KS_BN + q * 16), "r"(o[0]), "r"(o[1]),
                         "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                         : "memory");
        }
      };
      if (edge)
        tile_body(std::true_type{});
      else
        tile_body(std::false_type{});
    }
    musum[cg * KS_BM + row] = mu_d;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * KS_EPI_WARPS) : "memory");
    if (cg == 0) mun[m0 + row] = (float)(((musum[row] + musum[KS_BM + row]) + musum[2 * KS_BM + row]) + musum[3 * KS_BM + row]);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// Operand preparation.  x̂ = (x/ℓ − centre)·kscale, clamped to ±32768 (fp16 range; anything that far from the trials has
// k = 0 either way), hi = fp16(x̂), lo = fp16(x̂ − hi); the squared norm is taken of hi + lo, the point the MMAs actually see.
__global__ void __launch_bounds__(256) ks_center_kernel(const double* __restrict__ XsT, int ldx, int N, double* __restrict__ center) {
  __shared__ double red[256];
  const double* col = XsT + (size_t)blockIdx.x * ldx;
  double a = 0.0;
  for (int i = threadIdx.x; i < N; i += 256) a += col[i];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) center[blockIdx.x] = red[0] / N;
}

__device__ __forceinline__ void ks_split(double v, __half& hi, __half& lo, double& seen) {
  v = fmin(fmax(v, -32768.0), 32768.0);
  hi = __double2half(v);
  lo = __double2half(v - (double)__half2float(hi));
  seen = (double)__half2float(hi) + (double)__half2float(lo);
}

// one warp per row; rows >= n_valid and features >= D are written as zeros (planes are rows_pad × Dp)
template <typename XT, bool SCALED /* input already divided by ℓ (the trials' Xs) */>
__global__ void __launch_bounds__(256)
ks_split_rows_kernel(const XT* __restrict__ X, int64_t n_valid, int64_t rows_pad, int D, int Dp, const double* __restrict__ inv_ls, int n_ls,
                     const double* __restrict__ center, double kscale, __half* __restrict__ Ph, __half* __restrict__ Pl, float* __restrict__ nrm,
                     int nrm_stride, const double* __restrict__ alpha /* trials: second float of the (norm, alpha) pair */) {
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows_pad) return;
  double s = 0.0;
  for (int d = lane; d < Dp; d += 32) {
    __half hi = __float2half(0.f), lo = hi;
    if (r < n_valid && d < D) {
      double v = (double)X[r * D + d];
      if (!SCALED) v *= inv_ls[n_ls == 1 ? 0 : d];
      double seen;
      ks_split((v - center[d]) * kscale, hi, lo, seen);
      s = fma(seen, seen, s);
    }
    Ph[r * Dp + d] = hi;
    Pl[r * Dp + d] = lo;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    nrm[r * nrm_stride] = (float)s;
    if (alpha) nrm[r * nrm_stride + 1] = r < n_valid ? (float)alpha[r] : 0.f;
  }
}

}  // namespace

static double ks_kscale(int kind) { return kind == KBO_KERNEL_RBF ? 0.70710678118654752440 : 2.23606797749978969641; }

// Trial-side operands of the tensor-core K* kernel: planes of the centred, scaled trials and the (|x̂|², alpha) pairs.  Cheap
// (N·D); redone by every fit / append / rebase finish because alpha changes with each of them.
int kbo_i_tc_trials_prep(kbo_handle* h, bool new_center, cudaStream_t s) {
  const int N = h->N, D = h->D, Dp = round_up(D, KS_BK), Np = round_up(N, KS_BN);
  h->ks_ready = false;
  if (D > KS_MAX_SLABS * KS_BK) return KBO_OK;   // the FP64 K* kernel serves wider spaces
  KBO_TRY(kbo_reserve(h, h->ks_center, sizeof(double) * 512));
  KBO_TRY(kbo_reserve(h, h->ks_Xh, sizeof(__half) * (size_t)Np * Dp));
  KBO_TRY(kbo_reserve(h, h->ks_Xl, sizeof(__half) * (size_t)Np * Dp));
  KBO_TRY(kbo_reserve(h, h->ks_nxal, sizeof(float) * 2 * (size_t)Np));
  if (new_center) {
    ks_center_kernel<<<D, 256, 0, s>>>((const double*)h->XsT.p, h->ld, N, (double*)h->ks_center.p);
    KBO_LAUNCH_CHECK(h);
  }
  ks_split_rows_kernel<double, true><<<(unsigned)((Np + 7) / 8), 256, 0, s>>>((const double*)h->Xs.p, N, Np, D, Dp, nullptr, 1, (const double*)h->ks_center.p,
                                                                          ks_kscale(h->prm.kernel), (__half*)h->ks_Xh.p, (__half*)h->ks_Xl.p,
                                                                          (float*)h->ks_nxal.p, 2, (const double*)h->alpha.p);
  KBO_LAUNCH_CHECK(h);
  h->ks_ready = true;
  return KBO_OK;
}

// K̃* hi plane (rows_pad × Npad fp16, rows_pad = rows rounded up to 256) and μ̃ for `rows` candidates starting at Xc.
int kbo_i_tc_kstar(kbo_handle* h, const void* Xc, int xc_dtype, int64_t rows, __half* Ksh, float* mun, cudaStream_t s, int store_cols) {
  if (!h->ks_ready) KBO_FAIL(h, KBO_ERR_STATE, "tc_kstar: trial operands not prepared");
  const int N = h->N, D = h->D, Dp = round_up(D, KS_BK), Np = round_up(N, KS_BN), ns = Dp / KS_BK;
  const int64_t rows_pad = round_up64(rows, 256);
  if (h->Npad % KS_BN != 0 || h->Npad < Np) KBO_FAIL(h, KBO_ERR_STATE, "tc_kstar: Npad %d does not cover N %d", h->Npad, N);
  KBO_TRY(kbo_reserve(h, h->ks_Ch, sizeof(__half) * (size_t)rows_pad * Dp));
  KBO_TRY(kbo_reserve(h, h->ks_Cl, sizeof(__half) * (size_t)rows_pad * Dp));
  KBO_TRY(kbo_reserve(h, h->ks_nc, sizeof(float) * (size_t)rows_pad));
  const unsigned g = (unsigned)((rows_pad + 7) / 8);
  if (xc_dtype == KBO_F64)
    ks_split_rows_kernel<double, false><<<g, 256, 0, s>>>((const double*)Xc, rows, rows_pad, D, Dp, (const double*)h->d_inv_ls.p, (int)h->inv_ls.size(),
                                                         (const double*)h->ks_center.p, ks_kscale(h->prm.kernel), (__half*)h->ks_Ch.p, (__half*)h->ks_Cl.p,
                                                         (float*)h->ks_nc.p, 1, nullptr);
  else
    ks_split_rows_kernel<float, false><<<g, 256, 0, s>>>((const float*)Xc, rows, rows_pad, D, Dp, (const double*)h->d_inv_ls.p, (int)h->inv_ls.size(),
                                                        (const double*)h->ks_center.p, ks_kscale(h->prm.kernel), (__half*)h->ks_Ch.p, (__half*)h->ks_Cl.p,
                                                        (float*)h->ks_nc.p, 1, nullptr);
  KBO_LAUNCH_CHECK(h);
  CUtensorMap tmCh, tmCl, tmXh, tmXl;
  KBO_TRY(kbo_i_encode_map_f16(h, &tmCh, h->ks_Ch.p, (uint64_t)Dp, (uint64_t)rows_pad, KS_BK, KS_BM));
  KBO_TRY(kbo_i_encode_map_f16(h, &tmCl, h->ks_Cl.p, (uint64_t)Dp, (uint64_t)rows_pad, KS_BK, KS_BM));
  KBO_TRY(kbo_i_encode_map_f16(h, &tmXh, h->ks_Xh.p, (uint64_t)Dp, (uint64_t)Np, KS_BK, KS_BN));
  KBO_TRY(kbo_i_encode_map_f16(h, &tmXl, h->ks_Xl.p, (uint64_t)Dp, (uint64_t)Np, KS_BK, KS_BN));
  const size_t smem = ks_smem_bytes(ns);
  const float log2amp = (float)log2(h->prm.amplitude);
  const int ntiles = h->Npad / KS_BN;   // tiles past Np read out-of-bounds trial rows (TMA zero fill) and are masked to 0
  const int store_tiles = (store_cols < 0 || store_cols >= h->Npad) ? ntiles : (store_cols + KS_BN - 1) / KS_BN;
  if (h->prm.kernel == KBO_KERNEL_RBF) {
    KBO_CUDA(h, cudaFuncSetAttribute(tc_kstar_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_kstar_kernel<0><<<(unsigned)(rows_pad / KS_BM), KS_THREADS, smem, s>>>(tmCh, tmCl, tmXh, tmXl, ns, ntiles, N, (const float*)h->ks_nc.p,
                                                                            (const float4*)h->ks_nxal.p, log2amp, Ksh, h->Npad, mun, store_tiles);
  } else {
    KBO_CUDA(h, cudaFuncSetAttribute(tc_kstar_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_kstar_kernel<1><<<(unsigned)(rows_pad / KS_BM), KS_THREADS, smem, s>>>(tmCh, tmCl, tmXh, tmXl, ns, ntiles, N, (const float*)h->ks_nc.p,
                                                                            (const float4*)h->ks_nxal.p, log2amp, Ksh, h->Npad, mun, store_tiles);
  }
  KBO_LAUNCH_CHECK(h);
  return KBO_OK;
}
