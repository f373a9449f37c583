/*
 * kbo.h — C ABI of libkbo.so, the B200 (sm_100a) GP-suggestion engine.
 *
 * This is the drop-in boundary (SURVEY.md §8(b) "B-inner").  The reference platform repo
 * (/root/reference = kubeflow/kubeflow @ 6d6b78bc) reaches the suggestion service only through
 * Kubernetes (testing/katib_studyjob_test.py:155-161, testing/kfctl/kf_is_ready_test.py:65-70) and
 * holds no FFI for it; the arithmetic these entry points replace is the one Katib's
 * `bayesianoptimization` service executes (upstream kubeflow/katib
 * pkg/suggestion/v1beta1/skopt/base_service.py -> skopt.Optimizer.tell/ask ->
 * sklearn.gaussian_process.GaussianProcessRegressor).  Each entry point cites the
 * scikit-learn 1.9.0 lines ($SK = sklearn/gaussian_process) or the skopt function it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++/torch types.
 *   - every call returns KBO_OK (0) or a negative kbo_status; kbo_last_error(h) gives the text.
 *   - "dev" pointers are device memory owned by the caller (e.g. torch tensors), row-major,
 *     contiguous, 16-byte aligned.  The library never frees caller memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *     asynchronous on that stream unless the name ends in _host or the doc says "synchronises".
 *   - one handle per device per thread; a handle owns its workspace (grow-only device buffers).
 *   - there is NO CPU fallback: without a CUDA device kbo_create fails with KBO_ERR_CUDA.
 */
#ifndef KBO_H_
#define KBO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KBO_VERSION 100 /* 0.1.0 */

typedef struct kbo_handle kbo_handle;

typedef enum kbo_status {
  KBO_OK = 0,
  KBO_ERR_INVALID = -1, /* bad argument (maps to gRPC INVALID_ARGUMENT)            */
  KBO_ERR_CUDA = -2,    /* CUDA runtime/driver error (maps to gRPC INTERNAL)       */
  KBO_ERR_NOT_PD = -3,  /* Cholesky met a non-positive pivot ($SK/_gpr.py:353-362) */
  KBO_ERR_NOMEM = -4,
  KBO_ERR_STATE = -5    /* sweep before fit, etc.                                   */
} kbo_status;

typedef enum kbo_kernel_kind {
  KBO_KERNEL_RBF = 0,     /* $SK/kernels.py:1559-1570 */
  KBO_KERNEL_MATERN52 = 1 /* $SK/kernels.py:1713-1729 (nu = 2.5; skopt's default GP kernel) */
} kbo_kernel_kind;

typedef enum kbo_acq_kind {
  KBO_ACQ_EI = 0,  /* skopt.acquisition.gaussian_ei  */
  KBO_ACQ_LCB = 1, /* skopt.acquisition.gaussian_lcb (value reported is -(mu - kappa*sigma)) */
  KBO_ACQ_PI = 2   /* skopt.acquisition.gaussian_pi  */
} kbo_acq_kind;

typedef enum kbo_var_mode {
  KBO_VAR_F64 = 0,     /* variance contraction V = K* W^T on the FP64 SIMT pipe (checker precision) */
  KBO_VAR_TC_F16X3 = 1, /* tcgen05 tensor cores, fp16 hi/lo split (3 MMAs per product), fp32 accum */
  KBO_VAR_AUTO = 2      /* FP64 while M·N² <= 2e11 (a few ms of FP64 pipe), tensor cores above that    */
} kbo_var_mode;

typedef enum kbo_dtype { KBO_F64 = 0, KBO_F32 = 1 } kbo_dtype;

/* Fixed hyper-parameters θ and acquisition settings of one tell/ask (SURVEY.md §7 "parity at fixed θ"). */
typedef struct kbo_params {
  int32_t kernel;          /* kbo_kernel_kind */
  int32_t acq;             /* kbo_acq_kind */
  int32_t normalize_y;     /* 1 = $SK/_gpr.py:275-280 */
  int32_t var_mode;        /* kbo_var_mode */
  double amplitude;        /* ConstantKernel value ($SK/kernels.py:1278) */
  double noise;            /* GaussianProcessRegressor.alpha, added to diag(K) ($SK/_gpr.py:350) */
  double xi;               /* EI / PI offset (skopt default 0.01) */
  double kappa;            /* LCB weight (skopt default 1.96) */
  const double* length_scale; /* HOST pointer, n_length_scale entries */
  int32_t n_length_scale;  /* 1 (isotropic) or D (anisotropic) */
  int32_t tc_k_span;       /* TC mode: trials accumulated in TMEM before a drain to fp32 registers; 0 = default */
} kbo_params;

/* Result of one sweep: best acquisition value and its GLOBAL candidate index (lowest index wins ties,
 * = np.argmin(-values) in skopt Optimizer._tell), with the posterior at that point (raw y scale). */
typedef struct kbo_best {
  double value;
  int64_t index;
  double mu;
  double std;
} kbo_best;

/* Timings of the last kbo_suggest_host call, milliseconds (CUDA events on the call's stream). */
typedef struct kbo_timings {
  float h2d_ms, fit_ms, sweep_ms, d2h_ms, total_ms;
  float var_kernel_ms;   /* sum over chunks of the variance-contraction kernel */
  float cross_kernel_ms; /* sum over chunks of the K* / mean kernel           */
  float acq_kernel_ms;   /* acquisition + argmax kernels                      */
  int32_t launches;      /* kernels launched by the call                      */
  int32_t chunks;
  float calib_ms;        /* ranking pass: the stratified calibration rows (FP64 K* + three-product contraction) */
} kbo_timings;

int kbo_version(void);
int kbo_create(kbo_handle** out, int device);
void kbo_destroy(kbo_handle* h);
const char* kbo_last_error(const kbo_handle* h);
/* cap (bytes) on the per-sweep K* scratch; default 4 GiB.  Determines the candidate chunk size (rounded down to a
 * multiple of sm_count*128 rows so every launch is a whole number of waves). */
int kbo_set_scratch_limit(kbo_handle* h, uint64_t bytes);
/* tensor-core variance kernel variant: 1 (default) = 2-CTA cluster per 128-row panel with TMA multicast of the K* chunks,
 * 0 = one CTA per panel.  Same arithmetic, same results to the last bit per j-tile; the default halves the DRAM re-reads.
 * The environment variable KBO_TC_PAIR sets the initial value for new handles. */
int kbo_set_tc_pair(kbo_handle* h, int enabled);
/* Tensor-core mode only: re-evaluate on the FP64 path every candidate whose fp32 acquisition value is within 2e-4 of the
 * maximum (at most 4096 of them) and take the first-index argmax over those FP64 values, so the returned suggestion (index,
 * value, mu, std) has FP64 accuracy.  Default on; costs one extra pass over (mu, sigma²) and one stream synchronisation.
 * kbo_last_contenders returns how many candidates the last sweep refined (> 4096: refinement skipped). */
int kbo_set_tc_refine(kbo_handle* h, int enabled);
int kbo_last_contenders(kbo_handle* h);
/* Tensor-core sweeps that return only the suggestion (no mu/std/acq arrays) rank the grid with ONE fp16 product per term —
 * a third of the MMAs — and keep every candidate that could still be the maximum given that pass's error on sigma²
 * (calibrated per sweep against the three-product kernel on the first wave of rows, x8 + 1e-6); the survivors are decided
 * by the FP64 refinement above, so the returned suggestion is unchanged.  More than 4096 survivors: the sweep is redone
 * with three products.  Default on; needs kbo_set_tc_refine and kbo_set_tc_pair on.  kbo_last_rank_error returns the
 * largest |sigma²(1 product) - sigma²(3 products)| seen on the calibration rows of the last such sweep. */
int kbo_set_tc_fast(kbo_handle* h, int enabled);
double kbo_last_rank_error(kbo_handle* h);
/* The ranking pass builds its K* plane on the tensor cores (c·x as tcgen05 MMAs of fp16 hi/lo splits, kernel values in fp32,
 * only the hi plane written, the mean accumulated on the fly) and contracts it with cta_group::2 MMAs; 0 selects the FP64 K*
 * kernel + single-CTA-MMA cluster kernel of round 1 for that pass.  Default 1 (needs D <= 128; wider spaces use the FP64
 * kernel).  Either way the calibration rows are STRATIFIED over the grid (row i*M/n) and go through the FP64 K* kernel and
 * the three-product contraction; kbo_last_rank_mu_error is the largest |mean(ranking) - mean(FP64)| seen on them, in
 * normalised-y units (0 with the FP64 K* kernel).  The environment variable KBO_RANK_TC sets the initial value. */
int kbo_set_rank_tc(kbo_handle* h, int enabled);
double kbo_last_rank_mu_error(kbo_handle* h);
/* Pruning pass in front of the ranking pass (needs kbo_set_rank_tc).  sigma² = amp - sum_j v_j² with every term >= 0, so the sum
 * over a PREFIX of the trial tiles bounds sigma² from above, and with it EI / LCB (they grow with sigma; PI is bounded by 1 where
 * the improvement is positive) — at the prefix's share of the triangular contraction: the first eighth of the trials costs 1/64
 * of the MMAs.  The calibration rows' (near-exact) values give a lower bound on the maximum for free.  Candidates whose upper
 * bound stays below it are out; the rest (at most 16384, else the full ranking pass runs as before) get the full contraction,
 * the interval test and the FP64 decision.  tile_pairs: 512-trial tile pairs in the prefix; -1 (default) = an eighth of them,
 * 0 = no pruning pass.  kbo_last_prefix_survivors: how many candidates the last pruning pass kept (-1: it did not run).
 * The environment variable KBO_RANK_PREFIX sets the initial value. */
int kbo_set_rank_prefix(kbo_handle* h, int tile_pairs);
/* Lazy inverse (default on).  A tensor-core fit whose sweeps will prune forms only the rows of W = L^-1 the pruning pass reads
 * (the first 512·tile_pairs; plus the 256×256 diagonal blocks of every panel, which the look-ahead factorisation needs anyway):
 * alpha comes from two panel solves with L, the survivors' exact variances from panel solves with L (solve.cu), the lower bound
 * on the maximum from the sigma -> 0 limit of the acquisition function.  N³/3 FP64 flop — a third of the fit — are not spent.
 * The rest of W (and the full fp16 planes) are formed on demand by anything that needs them: array-returning sweeps, FP64-mode
 * sweeps, kbo_fit_append / kbo_fit_rebase, kbo_lml_grad, kbo_fit_state, and the sweep itself when more than 64 candidates
 * pass the prefix bound.  0 = always form W in kbo_fit.  Results do not depend on the setting beyond FP64 rounding. */
int kbo_set_lazy_inverse(kbo_handle* h, int enabled);
int kbo_last_prefix_survivors(kbo_handle* h);
/* How the last tensor-core sweep decided its suggestion: 0 = FP64 evaluation of every candidate that could still be the
 * maximum; 1 = more such candidates than the cap (4096), FP64 decision among the best 4096 by fp32 value; 2 = not refined
 * (exact fp32 ties beyond the cap, or refinement switched off): the suggestion carries the mode's own accuracy. */
int kbo_last_unrefined(kbo_handle* h);

/* ---- tell: GaussianProcessRegressor.fit at fixed θ ($SK/_gpr.py:275-280, 349-368) ---------------
 * X: N×D fp64, y: N fp64, device pointers (x_on_host = 0) or host pointers (x_on_host = 1).
 * Computes X/ℓ, normalised y, K = amp·k(X,X)+noise·I, L = chol(K), W = L^-1, alpha = W^T W y, LML, and
 * (TC mode) the fp16 hi/lo split of W.  All on the device; asynchronous on `stream` except that a
 * non-positive pivot is reported by the next synchronising call (kbo_fit_info / kbo_best_to_host). */
int kbo_fit(kbo_handle* h, const double* X, const double* y, int32_t N, int32_t D, const kbo_params* p,
            int x_on_host, void* stream);
/* Append ONE trial (x: D doubles, y) to the fitted history at the same theta without refactorising: the bordered Cholesky
 * row l = W k, d = sqrt(amp + noise - l.l), the matching row of W = L^-1, then the y statistics, alpha, LML and the fp16
 * planes are brought up to date — O(N^2) instead of O(N^3) (skopt's constant-liar `ask(n_points=k)` tells
 * k-1 lies one after another, `Optimizer.tell` refits each time).  Works in place while the history fits its 64-row pitch:
 * kbo_fit_room() says how many more trials can be appended (0: call kbo_fit with the whole history).  Not positive definite
 * is reported by kbo_fit_info like after a fit; the handle then needs a kbo_fit.  Results equal a refit's to ~1e-12. */
int kbo_fit_append(kbo_handle* h, const double* x, double y, int x_on_host, void* stream);
int kbo_fit_room(kbo_handle* h);
/* Keep the first n_keep trials of the fitted history, optionally with new targets y[0..n_keep) (NULL: unchanged).  The
 * leading blocks of L and L^-1 are the factors of the shorter history and y enters only through yn / alpha / LML, so this is
 * O(N^2): it drops constant-liar rows before the real trials are appended and replaces a lie by the observed value. */
int kbo_fit_rebase(kbo_handle* h, int32_t n_keep, const double* y, int y_on_host, void* stream);

/* synchronises; any out pointer may be NULL.  info = 0 or 1-based index of the failed pivot. */
int kbo_fit_info(kbo_handle* h, double* lml, double* y_mean, double* y_std, double* y_opt, int32_t* info,
                 void* stream);
/* Gradient of the log-marginal likelihood of the last kbo_fit w.r.t. θ = (log amplitude, log noise, log ℓ_1..ℓ_P), P =
 * n_length_scale of that fit ($SK/_gpr.py:621-653).  grad_host: n_out = 2 + P doubles.  Synchronises. */
int kbo_lml_grad(kbo_handle* h, double* grad_host, int32_t n_out, void* stream);
/* Log-marginal likelihood of G hyper-parameter settings over ONE history (X, y as for kbo_fit), for theta search: what sklearn's
 * optimiser evaluates one theta after another ($SK/_gpr.py:299-339, :604-618, :658-667).  The G factorisations run on G streams
 * and hide each other's latency (a Cholesky of a few thousand trials is a chain of single-CTA blocks): 8 thetas cost about two
 * fits.  Per theta: Gram, Cholesky, z = L^-1 yn by forward substitution, LML = -z.z/2 - sum log L_ii - N/2 log 2pi.  Does not
 * touch the handle's fitted state.  params: G entries (kernel, amplitude, noise, length_scale, n_length_scale, normalize_y of
 * entry 0 used); lml_host: G doubles (-inf where the Gram matrix is not positive definite); info_host (may be NULL): G failed
 * pivots (0 = ok).  Synchronises.  1 <= G <= 64. */
int kbo_lml_batch(kbo_handle* h, const double* X, const double* y, int32_t N, int32_t D, int32_t G, const kbo_params* params,
                  int x_on_host, double* lml_host, int32_t* info_host, void* stream);
/* copies the fit state into caller-owned DEVICE buffers (any may be NULL): L_out, W_out are N×N row-major
 * (L: lower Cholesky factor, strict upper part zeroed; W = L^-1), alpha_out has N entries. For parity tests. */
int kbo_fit_state(kbo_handle* h, double* L_out, double* W_out, double* alpha_out, void* stream);

/* ---- ask: predict(return_std) + acquisition + first-index argmax over M candidates ---------------
 * ($SK/_gpr.py:445-500; skopt gaussian_ei/lcb/pi; Optimizer._tell `X_cand[np.argmin(values)]`).
 * Xc: M×D, dtype xc_dtype, device (xc_on_host=0) or host.  global_offset is added to reported indices
 * (the rank's row offset when the grid is sharded).
 * Optional device outputs (NULL to skip): mu_out, std_out, acq_out — fp64, M entries each.
 * best_dev: device kbo_best written at the end of the stream work. */
int kbo_sweep(kbo_handle* h, const void* Xc, int32_t xc_dtype, int64_t M, int64_t global_offset, int xc_on_host,
              double* mu_out, double* std_out, double* acq_out, kbo_best* best_dev, void* stream);
/* synchronises `stream`, copies *best_dev to host, and returns KBO_ERR_NOT_PD if the fit failed. */
int kbo_best_to_host(kbo_handle* h, const kbo_best* best_dev, kbo_best* best_host, void* stream);

/* ---- multi-GPU: the grid shards by rows, one exchange step ---------------------------------------------
 * One process (or thread) per GPU, one handle each.  Rank r sweeps rows [lo_r, hi_r) of the grid with global_offset = lo_r;
 * kbo_allreduce_argmax then replaces *best_dev on every rank by the global first-index argmax (maximum value, lowest global
 * index among equals = np.argmin(-values) over the concatenated grid): ONE ncclAllGather of 32 bytes per rank on `stream`
 * plus a one-thread reduce.  NCCL is bound at run time (dlopen libnccl.so.2): kbo_comm_unique_id on one rank gives the
 * 128-byte ncclUniqueId the caller distributes by any means (the gRPC SPMD server broadcasts it); kbo_comm_init is
 * collective over the n_ranks handles.  Without a communicator (or with one rank) kbo_allreduce_argmax is a no-op. */
#define KBO_COMM_ID_BYTES 128
int kbo_comm_unique_id(void* id_out /* KBO_COMM_ID_BYTES */);
int kbo_comm_init(kbo_handle* h, int32_t n_ranks, int32_t rank, const void* id /* KBO_COMM_ID_BYTES */);
int kbo_comm_destroy(kbo_handle* h);
int kbo_comm_size(kbo_handle* h);
int kbo_allreduce_argmax(kbo_handle* h, kbo_best* best_dev, void* stream);

/* ---- one call, HOST buffers in, host result out (tell + ask); synchronises -----------------------
 * The end-to-end entry: H2D of X, y, Xc and D2H of the result are inside the call.  With a communicator (kbo_comm_init) the
 * call is collective: every rank passes its row block and global_offset, and best_host is the global argmax on every rank. */
int kbo_suggest_host(kbo_handle* h, const double* X, const double* y, int32_t N, int32_t D, const void* Xc,
                     int32_t xc_dtype, int64_t M, int64_t global_offset, const kbo_params* p,
                     kbo_best* best_host, kbo_timings* timings /* may be NULL */);
int kbo_last_timings(kbo_handle* h, kbo_timings* out);

/* ---- building blocks, caller-owned device memory (used by the parity tests one kernel at a time) --
 * kbo_gram:  K (N×ldk, fp64) = amplitude·k(Xs,Xs) + noise·I, Xs already divided by ℓ. ($SK/kernels.py:1561,1716)
 * kbo_potrf: in-place lower Cholesky of A (N×lda); the strict upper triangle is scratch afterwards; info_dev = 0 or failed pivot. ($SK/_gpr.py:352)
 * kbo_trtri: W (N×ldw, zero upper) = L^-1 for lower-triangular L.  Needs the potrf of the same handle just before
 *            (reuses its inverted diagonal blocks).
 * kbo_acq_argmax: standalone acquisition + argmax pass over (mu, var) in the NORMALISED y scale:
 *            mu_n, var_n fp32 (8 B/candidate read) -> optional acq fp32 (4 B write) + best. The HBM-bound kernel
 *            whose GB/s the headline metric asks for. */
int kbo_gram(kbo_handle* h, const double* Xs, int32_t N, int32_t D, int32_t kernel, double amplitude, double noise,
             double* K, int32_t ldk, void* stream);
int kbo_potrf(kbo_handle* h, double* A, int32_t N, int32_t lda, int32_t* info_dev, void* stream);
int kbo_trtri(kbo_handle* h, const double* L, int32_t N, int32_t ldl, double* W, int32_t ldw, void* stream);
int kbo_acq_argmax(kbo_handle* h, const float* mu_n, const float* var_n, int64_t M, int64_t global_offset,
                   int32_t acq, double y_mean, double y_std, double y_opt, double xi, double kappa,
                   float* acq_out, kbo_best* best_dev, void* stream);

/* ---- request ingestion (host only) -------------------------------------------------------------------
 * A serialized api.v1.beta1.GetSuggestionsRequest (kubeflow/katib pkg/apis/manager/v1beta1/api.proto; what grpc hands the
 * suggestion service before `GetSuggestionsRequest.FromString`) scanned into flat arrays, replacing the walk over the parsed
 * message that katib's pkg/suggestion/v1beta1/internal/trial.py `Trial.convert` + skopt/base_service.py `getSuggestions`
 * do per request.  `bytes` must stay alive while the kbo_req is open; offsets are relative to it.
 * kbo_req_trials fills one row per trial (any output pointer may be NULL), n_params columns in the order of param_names:
 *   usable          condition SUCCEEDED/EARLYSTOPPED and the objective metric present (Trial.convert's filter)
 *   objective       strtod of the objective metric value; objective_flags bit0 = plain decimal literal (else NaN)
 *   values / flags  strtod of the assignment; bit0 = plain decimal literal, bit1 = integer literal (<= 15 digits);
 *                   value_len 0xFFFFFFFF = the trial has no assignment of that name
 * Values that are not plain literals are left to the caller (Python float()/int() semantics, categorical strings).
 * select (NULL = all): per-trial mask; assignments of unselected trials are not looked at (their rows read as missing) —
 * a steady request parses numbers only for the trials the service has not seen (strtod dominates the scan). */
typedef struct kbo_req kbo_req;
int kbo_req_open(const void* bytes, uint64_t len, kbo_req** out);
void kbo_req_close(kbo_req* r);
int kbo_req_header(const kbo_req* r, uint64_t* experiment_off, uint64_t* experiment_len, int32_t* current_request_number,
                   int32_t* total_request_number, int32_t* n_trials);
int kbo_req_trials(const kbo_req* r, int32_t n_params, const char* const* param_names, uint64_t* name_off, uint32_t* name_len,
                   uint64_t* name_hash, int32_t* condition, uint8_t* usable, double* objective, uint8_t* objective_flags,
                   uint64_t* objective_off, uint32_t* objective_len, double* values, uint8_t* value_flags, uint64_t* value_off,
                   uint32_t* value_len, const uint8_t* select);
uint64_t kbo_hash64(const void* bytes, uint64_t len);   /* the FNV-1a hash kbo_req_trials puts in name_hash */

/* ---- CMA-ES (Katib algorithm `cmaes`, goptuna; N. Hansen's tutorial arXiv:1604.00772, active weights) ---------------
 * State (mean, sigma, C, evolution paths, eigenbasis) lives on the device.  One generation = ask + tell:
 *   kbo_cma_ask : X (lambda×D, device, fp64) = m + sigma·B·(d∘z); z from the built-in Philox stream (seed, generation) or,
 *                 for parity tests, from z_in_dev (lambda×D device, may be NULL).
 *   kbo_cma_tell: fitness (lambda, device, minimised, ties → lower sample index) for the X of the last ask → rank-μ/rank-1
 *                 covariance update, step-size control and a warm-started Jacobi eigendecomposition.
 * Limits: 1 <= D <= 128, 4 <= lambda <= 8192.  kbo_cma_state copies to HOST buffers (any may be NULL) and synchronises.
 * kbo_cma_run_synthetic runs `generations` ask/fitness/tell rounds entirely on the device with a built-in fitness
 * (0 = sphere, 1 = Rastrigin) — the BASELINE.json config-4 benchmark. */
typedef struct kbo_cma kbo_cma;
int kbo_cma_create(kbo_handle* h, kbo_cma** out, int32_t D, int32_t lambda, const double* mean0_host, double sigma0, uint64_t seed);
void kbo_cma_destroy(kbo_cma* c);
int kbo_cma_ask(kbo_handle* h, kbo_cma* c, double* X_dev, const double* z_in_dev, void* stream);
int kbo_cma_tell(kbo_handle* h, kbo_cma* c, const double* fitness_dev, void* stream);
int kbo_cma_state(kbo_handle* h, kbo_cma* c, double* mean, double* sigma, double* C, double* p_sigma, double* p_c, double* B,
                  double* d, double* Y_last, int64_t* generation, void* stream);
int kbo_cma_run_synthetic(kbo_handle* h, kbo_cma* c, int32_t fitness_kind, int32_t generations, double* best_f_host,
                          float* elapsed_ms, double* jacobi_sweeps_last);

#ifdef __cplusplus
}
#endif
#endif /* KBO_H_ */
