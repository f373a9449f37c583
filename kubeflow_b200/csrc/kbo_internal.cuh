// Internal declarations shared by the libkbo translation units (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/kbo.h"

#define KBO_NB 64  // Cholesky / trtri block size (one diagonal block = one CTA in shared memory)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

struct kbo_handle {
  int device = 0;
  int sm_count = 148;
  std::string err;
  uint64_t scratch_limit = 4ull << 30;

  // ---- fit state -------------------------------------------------------------------------
  bool fitted = false;
  bool have_planes = false;  // fp16 hi/lo planes of W built by the last fit
  int N = 0, D = 0, ld = 0;  // ld = leading dimension of L / W (multiple of 64)
  int Npad = 0;              // multiple of 256: extent of the fp16 W planes and of K* scratch rows
  kbo_params prm{};
  std::vector<double> inv_ls;  // 1/ℓ_d, host copy
  DevBuf d_inv_ls;             // D doubles
  DevBuf XsT;                  // Xs transposed: D × ld (coalesced trial-tile loads in the K* kernel)
  DevBuf Xs, nx, yn, K, W, Linv, T, alpha, z;
  DevBuf yraw, lrow;   // raw y (re-normalised on append) and the append's work vectors
  DevBuf Wh, Wl;        // fp16 planes Npad×Npad (TC mode)
  DevBuf scal;          // device scalars, see ScalIdx
  DevBuf info;          // int32: potrf info
  DevBuf stage_X, stage_y, stage_Xc;  // H2D staging for host-pointer entry points
  // ---- sweep workspace ---------------------------------------------------------------------
  DevBuf Ks64;          // chunk × ldks fp64 (F64 mode)
  DevBuf Ksh, Ksl;      // chunk × Npad fp16 planes (TC mode)
  DevBuf mun;           // M (fp64 in F64 mode, fp32 in TC mode)
  DevBuf part;          // chunk × n_jtiles partial Σv² (fp64) — F64 mode
  DevBuf varn;          // M (same dtype as mun)
  DevBuf blockbest;     // per-block argmax partials
  DevBuf best;          // one kbo_best for suggest_host
  DevBuf refine, refine_x;  // contender list / gathered rows of the FP64 refinement (tensor-core mode)
  int last_contenders = 0;
  float last_rank_err = 0.f;   // largest |σ²(1 product) − σ²(3 products)| over the calibration rows of the last fast sweep
  bool tc_refine = true;
  bool tc_fast = true;      // array-free tensor-core sweeps rank with ONE fp16 product and let the FP64 refinement decide
  DevBuf var_cal;           // three-product variance of the calibration rows
  kbo_timings tim{};
  cudaEvent_t ev[8] = {};
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_var, ev_cross, ev_acq, ev_cal;
  size_t ev_var_used = 0, ev_cross_used = 0, ev_acq_used = 0, ev_cal_used = 0;
  int launches = 0;
  bool time_kernels = false;
  bool attr_fit = false, attr_tc = false;
  int tc_pair = 1;  // variance kernel variant (kbo_set_tc_pair)  // cudaFuncSetAttribute done for this handle's device
  // ---- tensor-core K* generation (tc_kstar.cu) and cta_group::2 ranking kernel (tc_rank.cu) ---------------------------
  bool rank_tc = true;        // kbo_set_rank_tc: the ranking pass builds K̃* on the tensor cores and contracts with cta_group::2 MMAs
  bool ks_ready = false;      // trial-side operands below match the current fit
  DevBuf ks_center;           // D doubles: centre subtracted from the scaled coordinates (any fixed vector is valid)
  DevBuf ks_Xh, ks_Xl;        // trials, fp16 hi/lo planes, round_up(N,256) × round_up(D,32)
  DevBuf ks_nxal;             // (|x̂_n|², alpha_n) float pairs
  DevBuf ks_Ch, ks_Cl, ks_nc; // candidate planes / squared norms of the current chunk
  DevBuf rk_part;             // [tile pairs][rows] partial Σ v² of the ranking kernel
  struct RkSched {
    int key[5] = {-1, -1, -1, -1, -1};  // (row groups, j-tiles, clusters, first tile pair, end tile pair)
    DevBuf dev;
    std::vector<int> host;
  } rk_sched[8];
  int rk_sched_next = 0;
  bool attr_rank = false;
  int rank_prefix = -1;            // kbo_set_rank_prefix: tile pairs of the pruning pass (-1: an eighth of them, 0: no pruning pass)
  int last_prefix_survivors = -1;  // candidates whose prefix upper bound reached the calibration rows' best value (-1: pass not run)
  DevBuf pr_list, pr_x, pr_mu, pr_var;   // pruning pass: survivor indices, their gathered rows, their ranking-pass mean / variance
  DevBuf cal_mu_rk, cal_var_rk;          // the ranking arithmetic on the calibration rows
  DevBuf mu_part;                  // FP64 K* kernel with the trial tiles split over blockIdx.y: partial means [split][row]
  DevBuf cal_idx, cal_x, cal_mu;   // stratified calibration rows of the ranking pass: indices, gathered rows, FP64-path mean
  float last_rank_mu_err = 0.f;    // largest |μ̃ − μ| (normalised units) on the calibration rows of the last ranking sweep
  int last_unrefined = 0;          // 1: the last tensor-core sweep could not decide in FP64 (more near-ties than the cap)
  // ---- fit: Cholesky chain on a high-priority stream, row-panel inverse on a second one (fit.cu) ----------------------
  cudaStream_t s_hi = nullptr, s_lo = nullptr, s_upd = nullptr, s_copy = nullptr;
  DevBuf T2;                        // N × 256 panel-solve scratch of the look-ahead factorisation (v2)
  // v3: the chain runs on its own SM partition (green contexts); shadow panel solve / trailing update / inverse share the rest
  bool part_tried = false, part_ok = false;
  void *gctx_chain = nullptr, *gctx_rest = nullptr;   // CUgreenCtx
  // stream roles: 0 chain, 1 near shadow (both on the chain's partition), 2 far shadow, 3..8 column-block updates at distance 1..6,
  // 9 bulk trailing update, 10 inverse, 11 W_PP and 12 MID rows (both on the chain's partition)
  cudaStream_t s3g[13] = {};       // green-context set
  cudaStream_t s3p[13] = {};       // the same roles as plain priority streams (small N, profilers, no green contexts)
  DevBuf Linv4;                     // the four 64×64 block inverses of the current panel
  std::vector<cudaEvent_t> ev_panel;
  cudaEvent_t ev_gram = nullptr;   // the first column block of the Gram matrix is complete (the factorisation starts on it)
  // ---- lazy inverse (fit.cu, solve.cu): the product path never needs all of W = L⁻¹ -----------------------------------------
  bool lazy_w = true;              // kbo_set_lazy_inverse: tensor-core fits form only the leading rows of W the pruning pass reads
  bool w_full = true;              // all rows of W (and the full fp16 planes) exist
  int w_lead = 0;                  // rows of W formed so far (the diagonal 256-blocks exist for every panel)
  DevBuf sv_B, sv_V;               // N × 8 right-hand sides / solutions of the panel solves
  DevBuf sv_bar;                   // grid-barrier counter of the cooperative solve kernels
  DevBuf zf;                       // b (consumed) and z = L⁻¹·yn of the forward substitution carried along by the factorisation
  bool z_ready = false;            // zf holds z of the current factorisation
  // ---- kbo_lml_batch: concurrent factorisations for several θ (fit.cu) --------------------------------------------------
  void* lml_lanes = nullptr;       // std::vector<LmlLane>*
  DevBuf lml_yn, lml_scal;
  cudaEvent_t lml_ev = nullptr;
  // ---- multi-GPU exchange (comm.cu): NCCL communicator bound at run time ------------------------------------------------
  void* comm = nullptr;            // ncclComm_t
  int comm_ranks = 1, comm_rank = 0;
  DevBuf comm_buf;                 // n_ranks kbo_best gathered
};

enum ScalIdx { S_YMEAN = 0, S_YSTD = 1, S_YOPT = 2, S_LML = 3, S_LOGDET = 4, S_QUAD = 5, S_COUNT = 8 };

#define KBO_FAIL(h, code, ...)                         \
  do {                                                 \
    char _b[512];                                      \
    snprintf(_b, sizeof _b, __VA_ARGS__);              \
    (h)->err = _b;                                     \
    return (code);                                     \
  } while (0)

#define KBO_CUDA(h, expr)                                                                          \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      KBO_FAIL(h, KBO_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define KBO_LAUNCH_CHECK(h)                                                                        \
  do {                                                                                             \
    (h)->launches++;                                                                               \
    cudaError_t _e = cudaGetLastError();                                                           \
    if (_e != cudaSuccess)                                                                         \
      KBO_FAIL(h, KBO_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

static inline int kbo_reserve(kbo_handle* h, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return KBO_OK;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
  cudaError_t e = cudaMalloc(&b.p, bytes);
  if (e != cudaSuccess) KBO_FAIL(h, KBO_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  b.cap = bytes;
  return KBO_OK;
}
#define KBO_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != KBO_OK) return _r; \
  } while (0)

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline int64_t round_up64(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ---- fit.cu ------------------------------------------------------------------------------------
int kbo_i_gram(kbo_handle* h, const double* Xs, int N, int D, int kernel, double amp, double noise, double* K, int ldk,
               cudaStream_t s);
int kbo_i_potrf(kbo_handle* h, double* A, int N, int lda, int* info_dev, cudaStream_t s);
int kbo_i_fit_append(kbo_handle* h, const double* x_dev, double y, cudaStream_t s);
int kbo_i_fit_rebase(kbo_handle* h, int n_keep, const double* y_dev, cudaStream_t s);
int kbo_i_trtri(kbo_handle* h, const double* L, int N, int ldl, double* W, int ldw, cudaStream_t s);
int kbo_i_zero_upper(kbo_handle* h, double* A, int N, int lda, cudaStream_t s);
int kbo_i_fit(kbo_handle* h, const double* X_dev, const double* y_dev, int N, int D, const kbo_params* p, cudaStream_t s);
int kbo_i_lml_batch(kbo_handle* h, const double* X_dev, const double* y_dev, int N, int D, int G, const kbo_params* params, double* lml_host,
                    int32_t* info_host, cudaStream_t s);
void kbo_i_lml_batch_free(kbo_handle* h);
void kbo_i_fit_partition_free(kbo_handle* h);
int kbo_i_ensure_w(kbo_handle* h, cudaStream_t s);   // form the rest of W and the full planes if the fit left them out
// ---- solve.cu ----------------------------------------------------------------------------------
int kbo_i_alpha_by_solves(kbo_handle* h, cudaStream_t s);
int kbo_i_variance_by_solves(kbo_handle* h, const double* Ks, int n, double* varn64, cudaStream_t s);
int kbo_i_zsolve_begin(kbo_handle* h, cudaStream_t s);
int kbo_i_zsolve_diag(kbo_handle* h, int K0, int Wd, cudaStream_t s);
int kbo_i_zsolve_update(kbo_handle* h, int K0, int Wd, cudaStream_t s);
// ---- sweep.cu ----------------------------------------------------------------------------------
int kbo_i_sweep(kbo_handle* h, const void* Xc_dev, int xc_dtype, int64_t M, int64_t goff, double* mu_out, double* std_out,
                double* acq_out, kbo_best* best_dev, cudaStream_t s);
int kbo_i_debug_cross_planes(kbo_handle* h, const void* Xc, int xc_dtype, int64_t M, cudaStream_t s);   // FP64 K* → Ksh/Ksl, mun (test hook)
int kbo_i_acq_argmax_f32(kbo_handle* h, const float* mu_n, const float* var_n, int64_t M, int64_t goff, int acq, double y_mean,
                         double y_std, double y_opt, double xi, double kappa, double amp, float* acq_out, kbo_best* best_dev,
                         cudaStream_t s);
int kbo_i_encode_map_f16(kbo_handle* h, CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint32_t box_inner,
                         uint32_t box_outer);
// ---- tc_kstar.cu / tc_rank.cu ------------------------------------------------------------------
int kbo_i_tc_trials_prep(kbo_handle* h, bool new_center, cudaStream_t s);
// store_cols: leading columns of the plane actually written (-1: all); the mean always runs over every trial
int kbo_i_tc_kstar(kbo_handle* h, const void* Xc, int xc_dtype, int64_t rows, __half* Ksh, float* mun, cudaStream_t s, int store_cols = -1);
int kbo_i_tc_rank(kbo_handle* h, const __half* Ksh, int64_t rows, const __half* Wh, int Npad, double amp, float* var_n_out, cudaStream_t s,
                  int p_begin = 0, int p_end = -1);
// ---- tc_var.cu ---------------------------------------------------------------------------------
// var_n[m] = amp − Σ_j (Σ_k K*[m,k] W[j,k])²  for the rows of one chunk, on tcgen05 tensor cores.
int kbo_i_tc_variance(kbo_handle* h, const __half* Ksh, const __half* Ksl, int64_t rows, const __half* Wh, const __half* Wl,
                      int Npad, double w_scale_inv, double amp, float* var_n_out, int k_span, cudaStream_t s, int nprod = 3, int jtiles = -1);
