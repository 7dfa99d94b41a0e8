/* gpbo.h -- C ABI of the B200-native step2_fitgps hot path (libgpbo.so).
 *
 * Drop-in boundary for GP-BayesOpInf's GP fitting / posterior-moment stage.  The reference has no
 * FFI: the seam is the Python module `gpkernels` imported by each `step2_fitgps.py`.
 * Each entry point below names the reference call it replaces; the ctypes binding a maintainer adds
 * is shown in INTEGRATION.md and lives in gp-bayesopinf_b200/_lib.py.
 *
 * Conventions
 *  - all matrices/vectors are IEEE float64, row-major, batch-strided; sizes are plain ints;
 *  - `theta` is the log-space hyper-parameter vector (log sigma^2, log ell, log chi), the order of
 *    sklearn's `kernel_.theta` for (ConstantKernel * RBF) + WhiteKernel;
 *  - a "pair" is one (GP, theta) combination; `gp_of[b]` selects which GP's (t, y) pair b uses
 *    (NULL: pair b uses GP b); all GPs of a call share the training size m;
 *  - functions with the suffix `_host` take HOST pointers and perform the H2D/D2H copies themselves; they reject
 *    non-finite t / y / evaluation points with GPBO_EINVAL (scikit-learn validates X and y the same way);
 *    the others take DEVICE pointers and enqueue on `stream` (a cudaStream_t, NULL = default stream)
 *    and return after the work has been enqueued AND completed (they synchronise the stream);
 *  - every function returns 0 on success or a negative GPBO_E* code; `gpbo_last_error()` describes
 *    the last failure of the calling thread.  No exceptions cross the ABI.  There is no CPU fallback:
 *    without a CUDA device every compute entry point fails with GPBO_ECUDA.
 *  - per-pair numerical status: 0 ok, 1 = kernel matrix not positive definite (the condition for which
 *    sklearn returns (-inf, 0) inside the optimiser, _gpr.py:589-593, and raises LinAlgError in the
 *    final factorisation, _gpr.py:351-361).
 */
#ifndef GPBO_H
#define GPBO_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPBO_OK 0
#define GPBO_EINVAL (-1)   /* bad argument */
#define GPBO_ECUDA (-2)    /* CUDA runtime error / no device */
#define GPBO_ENOMEM (-3)   /* workspace does not fit the memory limit */

typedef struct gpbo_ctx gpbo_ctx;

/* Library version (major*10000 + minor*100 + patch). */
int gpbo_version(void);
/* Message of the last error on this thread ("" if none). */
const char* gpbo_last_error(void);

/* Create / destroy a workspace handle on CUDA device `device`.  `max_workspace_bytes` bounds the
 * scratch the library may allocate (0 = 80% of the device's free memory at first use). */
int gpbo_create(gpbo_ctx** out, int device, size_t max_workspace_bytes);
int gpbo_destroy(gpbo_ctx* ctx);

/* Kernel family used by every fit / LML / prediction entry point of this handle: 0 = RBF, the reference's
 * (ConstantKernel * RBF) + WhiteKernel (default); 3 / 5 = Matern nu = 3/2, 5/2 with scikit-learn's Matern semantics
 * (kernels.py:1601-1790) -- an extension, the reference has no Matern kernel.  The derivative quantities of
 * gpbo_lstsq_moments* are then the analytic d/dt' and d^2/dt'dt of the Matern kernel. */
int gpbo_set_kernel_family(gpbo_ctx* ctx, int twice_nu);

/* Training sizes m <= max_m (at most 224, the default) take the in-shared small-matrix path: one CTA per pair
 * evaluates LML + gradient in ONE launch and gpbo_fit_host runs the whole multi-start fit as ONE persistent kernel
 * with the L-BFGS-B state machines on the device (csrc/kernels_small.cuh) -- the regime of the reference's own
 * experiments (m = 10 ... 200, ODEs/experiments.sh:11-18, PDEs/experiments.sh:13-26).  0 forces the blocked
 * 128-tile path for every size. */
int gpbo_set_small_path(gpbo_ctx* ctx, int max_m);

/* The cudaStream_t on which the *_host entry points of this handle enqueue their copies and kernels (so that a
 * caller can record CUDA events on it). */
int gpbo_get_stream(gpbo_ctx* ctx, void** stream);

/* Number of CUDA kernels this handle has launched so far (for bench accounting). */
long long gpbo_launch_count(const gpbo_ctx* ctx);
/* Number of pairs a wave can hold at training size m (derived from the memory limit). */
int gpbo_wave_capacity(gpbo_ctx* ctx, int m);

/* Kernel-matrix assembly to HBM.  Replaces sklearn `kernel_(X[, Y])` (kernels.py:838-871, 936-971,
 * 1244-1296, 1374-1419, 1530-1587) and the NumPy broadcasting of gpkernels.py:591-609, 630-641.
 * kind: 0 K(theta)=sigma^2 R+chi I (sklearn order)   1 K_yy (rbf_eval order)   2 K(t1,t2) sklearn cross
 *       3 kappa(t1,t2)   4 K_zy=d/dt1 kappa   5 K_zz=d2/dt1dt2 kappa   6 dK/dlog(ell)
 * t1: [B][n1] with stride t1_stride (0 = shared), t2 likewise; theta [B][3]; out [B][n1][n2].  The train kinds 0 / 1
 * (white noise on the diagonal, sklearn `kernel_(X)`) need n1 == n2 (GPBO_EINVAL otherwise). */
int gpbo_assemble(gpbo_ctx* ctx, int kind, const double* t1, long t1_stride, int n1, const double* t2,
                  long t2_stride, int n2, const double* theta, int B, double* out, void* stream);

/* Matern (nu = twice_nu / 2, twice_nu = 3 or 5) counterpart of gpbo_assemble -- an extension: the reference has no
 * Matern kernel (BASELINE.json's north star names RBF/Matern assembly).  K follows scikit-learn's Matern
 * (kernels.py:1601-1790) in (ConstantKernel * Matern) + WhiteKernel; same arguments and kinds as gpbo_assemble
 * (kinds 1 / 3 alias 0 / 2; 4 = d k / d t1; 5 = d^2 k / d t1 d t2; 6 = dK / dlog(ell)). */
int gpbo_assemble_matern(gpbo_ctx* ctx, int twice_nu, int kind, const double* t1, long t1_stride, int n1,
                         const double* t2, long t2_stride, int n2, const double* theta, int B, double* out,
                         void* stream);

/* Batched log-marginal likelihood and gradient at fixed theta.
 * Replaces GaussianProcessRegressor.log_marginal_likelihood(theta, eval_gradient=True)
 * (sklearn _gpr.py:541-656) for B pairs at once.
 * t, y: [G][m]; theta: [B][3]; gp_of: [B] or NULL; outputs lml [B], grad [B][3] (NULL: LML only),
 * status [B] (may be NULL). */
int gpbo_lml_grad(gpbo_ctx* ctx, const double* t, const double* y, int G, int m, const double* theta,
                  const int* gp_of, int B, double* lml, double* grad, int* status, void* stream);
int gpbo_lml_grad_host(gpbo_ctx* ctx, const double* t, const double* y, int G, int m, const double* theta,
                       const int* gp_of, int B, double* lml, double* grad, int* status);

/* Multi-start hyper-parameter optimisation, all (GP, start) pairs in lock-step.
 * Replaces the loop of scipy L-BFGS-B runs in GaussianProcessRegressor.fit (sklearn _gpr.py:298-340,
 * 658-668).  bounds_log: [3][2] log-space box; starts: [B][3] (clipped into the box like scipy);
 * outputs per pair: theta_opt [B][3], fun [B] (= -LML at theta_opt), nfev [B], nit [B],
 * opt_status [B] (0 projected-gradient test, 1 relative-reduction test, 2 abnormal line search,
 * 3 maxiter, 4 maxfun, 5 objective not finite at the start).  opts may be NULL for scipy's defaults:
 * opts = {factr, pgtol, maxiter, maxfun, maxls} as doubles.  total_evals (may be NULL) receives the
 * number of LML+gradient evaluations performed, rounds the number of lock-step launches. */
int gpbo_fit_host(gpbo_ctx* ctx, const double* t, const double* y, int G, int m, const double* bounds_log,
                  const double* starts, const int* gp_of, int B, const double* opts, double* theta_opt,
                  double* fun, int* nfev, int* nit, int* opt_status, long long* total_evals, int* rounds);

/* The pieces of gpbo_fit_host, for drivers that interleave a collective with every lock-step round (multi-GPU:
 * gp-bayesopinf_b200/sharding.py re-balances the live pairs over the ranks each round and all-gathers the
 * evaluations).  An optimiser pool holds the B L-BFGS-B state machines (host memory only, no CUDA device needed):
 *   gpbo_optpool_live   -> the running pairs' indices idx[n] and current trial points theta[n][3] (arrays of B entries);
 *   gpbo_optpool_feed   <- lml[n], grad[n][3] for pairs idx[n] (the pool minimises -lml like sklearn _gpr.py:300-307);
 *   gpbo_optpool_result -> per-pair optimum as in gpbo_fit_host, total evaluations fed and number of feed rounds;
 *                          a pair that is still running (driver stopped early) reports its last accepted iterate
 *                          with opt_status -1.
 * Replaces, together with gpbo_lml_grad_resident_host, the same reference loop as gpbo_fit_host. */
typedef struct gpbo_optpool gpbo_optpool;
int gpbo_optpool_create(gpbo_optpool** out, int B, const double* bounds_log, const double* starts, const double* opts);
int gpbo_optpool_destroy(gpbo_optpool* pool);
int gpbo_optpool_live(gpbo_optpool* pool, int* idx, double* theta, int* n_live);
int gpbo_optpool_feed(gpbo_optpool* pool, int n, const int* idx, const double* lml, const double* grad);
int gpbo_optpool_result(gpbo_optpool* pool, double* theta_opt, double* fun, int* nfev, int* nit, int* opt_status,
                        long long* total_evals, int* rounds);

/* Keep one problem (t, y: [G][m], HOST pointers) resident in HBM, then evaluate LML + gradient for batches of
 * pairs against it (theta [B][3], gp_of [B], outputs as gpbo_lml_grad_host; all HOST pointers).  Any other *_host
 * entry point of the handle replaces the resident problem. */
int gpbo_problem_upload_host(gpbo_ctx* ctx, const double* t, const double* y, int G, int m);
int gpbo_lml_grad_resident_host(gpbo_ctx* ctx, const double* theta, const int* gp_of, int B, double* lml,
                                double* grad, int* status);

/* Host-only driver of the SAME optimiser state machine that gpbo_fit_host advances in lock-step, for one
 * 3-parameter objective given as a callback (fn returns f and writes g[3]).  Needs no CUDA device.
 * Replaces one scipy.optimize.minimize(method="L-BFGS-B", jac=True, bounds=...) call
 * (sklearn _gpr.py:658-668).  Outputs: x[3], *f, *nfev, *nit, *status (codes as in gpbo_fit_host). */
typedef double (*gpbo_objective_fn)(const double* x, double* g, void* user);
int gpbo_lbfgsb_minimize(gpbo_objective_fn fn, void* user, const double* x0, const double* bounds_log,
                         const double* opts, double* x, double* f, int* nfev, int* nit, int* status);

/* Posterior mean and standard deviation at n_star points for G GPs with fixed theta [G][3].
 * Replaces _BaseGP.predict (gpkernels.py:350-365) -> GaussianProcessRegressor.predict(return_std=True)
 * (sklearn _gpr.py:444-500).  t_star: [G][n_star] with stride tstar_stride (0 = shared).
 * Outputs mean, std: [G][n_star]; alpha [G][m] (may be NULL) = K^-1 y; status [G]. */
int gpbo_predict_host(gpbo_ctx* ctx, const double* t, const double* y, int G, int m, const double* theta,
                      const double* t_star, long tstar_stride, int n_star, double* mean, double* std,
                      double* alpha, int* status);

/* State / time-derivative estimates and derivative covariance at the estimation points.
 * Replaces GP_RBFW.compute_lstsq_matrices + _BaseGP._compute_estimates_and_weights up to and
 * including ddt_covariance (gpkernels.py:612-649, 445-493).  t_est: [G][n_est] with stride
 * test_stride (0 = shared).  Outputs state, ddt: [G][n_est]; cov: [G][n_est][n_est] or NULL;
 * status [G] (1 = K_yy not positive definite, the condition for scipy cho_factor's LinAlgError). */
int gpbo_lstsq_moments_host(gpbo_ctx* ctx, const double* t, const double* y, int G, int m, const double* theta,
                            const double* t_est, long test_stride, int n_est, double* state, double* ddt,
                            double* cov, int* status);
/* Device-pointer form of the same (used for HBM-resident timing). */
int gpbo_lstsq_moments(gpbo_ctx* ctx, const double* t, const double* y, int G, int m, const double* theta,
                       const double* t_est, long test_stride, int n_est, double* state, double* ddt, double* cov,
                       int* status, void* stream);

/* Weight matrix of the weighted least squares: sqrtW = (C + eta I)^(-1/2), symmetric.
 * Replaces la.eigh(C + eta I) -> V diag(lambda^-1/2) V^T of _BaseGP._compute_estimates_and_weights
 * (gpkernels.py:496-504).  Computed with the coupled Newton-Schulz iteration on the FP64 tensor tiles
 * (csrc/kernels_sqrtw.cuh).  cov, sqrtw: [G][n][n] DEVICE pointers (the lower triangle of cov is read);
 * status [G] and iters [G] are HOST arrays (may be NULL): status 1 = C + eta I has an eigenvalue <= 0 to working
 * precision (the condition for the reference's ValueError "inverse covariance not positive definite,
 * increase eta", gpkernels.py:500-503); iters = Newton-Schulz iterations used. */
int gpbo_sqrtw(gpbo_ctx* ctx, const double* cov, int G, int n, double eta, double* sqrtw, int* status, int* iters,
               void* stream);
/* Same with HOST cov / sqrtw buffers. */
int gpbo_sqrtw_host(gpbo_ctx* ctx, const double* cov, int G, int n, double eta, double* sqrtw, int* status,
                    int* iters);
/* gpbo_lstsq_moments_host followed by gpbo_sqrtw on the covariance while it is still in HBM: everything
 * GP_RBFW.compute_lstsq_matrices produces (gpkernels.py:612-649, 445-504) in one call.  w_status / w_iters as in
 * gpbo_sqrtw (HOST arrays, may be NULL). */
int gpbo_lstsq_weights_host(gpbo_ctx* ctx, const double* t, const double* y, int G, int m, const double* theta,
                            const double* t_est, long test_stride, int n_est, double eta, double* state, double* ddt,
                            double* cov, double* sqrtw, int* status, int* w_status, int* w_iters);

/* Products of the weighted least squares of step 3: for every GP g, out_lhs[g] = sqrtW[g] @ lhs and
 * out_rhs[g] = sqrtW[g] @ rhs[g].  Replaces `self.weights[i] @ lhs, self.weights[i] @ rhs[i]` of
 * WeightedLSTSQSolver.fit (codebase/wlstsq.py:183-188).  All pointers HOST.  sqrtw: [G][n][n], or NULL to use the
 * weight matrices still resident in HBM from the last gpbo_lstsq_weights_host / gpbo_sqrtw_host call of this handle
 * (same G and n) -- they then never travel to the host for this step.  lhs: [n][d] (shared data matrix D),
 * rhs: [G][n]; outputs out_lhs [G][n][d], out_rhs [G][n]. */
int gpbo_weighted_products_host(gpbo_ctx* ctx, const double* sqrtw, int G, int n, const double* lhs, int d,
                                const double* rhs, double* out_lhs, double* out_rhs);

/* Posterior assembly of step 3 for a grid of regularizers (the linear algebra of
 * PDEs/step3_estimate.py:75-95 = get_bayesian_model(reg), for all candidates of the grid search :131-148 at once).
 * For every GP / mode g and every regularizer regs[k]:
 *   A_g = sqrtW[g] @ lhs, b_g = sqrtW[g] @ rhs[g]          (codebase/wlstsq.py:183-188)
 *   gram[g] = A_g^T A_g, proj[g] = A_g^T b_g;  precision P = gram[g] + regs[k]^2 I      (step3_estimate.py:86-90)
 *   means[k][g] = argmin |A_g o - b_g|^2 + regs[k]^2 |o|^2  (lstsq_solver.solve(), :78-79; opinf L2Solver)
 *   chol[k][g]  = lower Cholesky factor of P -- what scipy.stats.Covariance.from_precision computes inside
 *                 bayes.BayesianROM (codebase/bayes.py:283-287); status[k][g] = 1 when P is not positive definite
 *                 (the reference's LinAlgError "Matrix is not positive definite" -> candidate skipped, :92-95;
 *                 means[k][g] is then NaN), else 0.
 * All pointers HOST.  sqrtw [G][n][n] or NULL (use the stack resident from gpbo_lstsq_weights_host / gpbo_sqrtw_host,
 * same G and n); lhs [n][d] with d <= 128; rhs [G][n]; regs [nreg]; means [nreg][G][d]; chol [nreg][G][d][d] or NULL;
 * gram [G][d][d] or NULL; proj [G][d] or NULL; status [nreg][G].  Integrating the ROM for the posterior draws of each
 * candidate (the rest of the grid search) needs `opinf` and stays in the reference. */
int gpbo_posterior_grid_host(gpbo_ctx* ctx, const double* sqrtw, int G, int n, const double* lhs, int d, const double* rhs,
                             const double* regs, int nreg, double* means, double* chol, double* gram, double* proj,
                             int* status);

/* Per-kernel-class device timing (CUDA events on the launching stream), for bench.py's roofline.
 * Classes: 0 prep 1 chol_diag 2 chol_panel 3 trsv 4 trtri 5 lauum_grad 6 finalize 7 cross_panel
 *          8 schur 9 mean_std (fused kernel-row x alpha means) 10 assemble 11 sqrtw 12 small (in-shared path)
 *          13 std (row norms of V for the predictive std).  `ms` and `launches` are arrays of
 * GPBO_NCLASS entries, accumulated since the last gpbo_profile_enable(ctx, 1). */
#define GPBO_NCLASS 14
int gpbo_profile_enable(gpbo_ctx* ctx, int on);
int gpbo_profile_get(gpbo_ctx* ctx, double* ms, long long* launches);

/* Micro-benchmarks used by bench.py for the roofline denominators (not part of the reference API):
 * FP64 tensor-pipe (DMMA) issue-rate peak in TFLOP/s over `ms` milliseconds of back-to-back DMMA. */
int gpbo_bench_dmma_peak(gpbo_ctx* ctx, int iters, double* tflops, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* GPBO_H */
