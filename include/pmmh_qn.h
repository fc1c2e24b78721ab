/*
 * pmmh_qn.h -- C ABI of libpmmh_qn_b200.so, the B200 (sm_100a) implementation of the
 * likelihood-estimation hot path of compops/pmmh-qn.
 *
 * Every entry point replaces one native call the reference's Python estimator classes make
 * into its Cython extensions (paths relative to /root/reference/python):
 *
 *   pmmh_flps_sv_corr            flps_sv_corr(obs, params, rvr, rvp, compute_hessian)
 *                                state/particle_methods/stochastic_volatility.pyx:205,
 *                                called at state/particle_methods/cython.py:97
 *   pmmh_bpf_sv_corr             bpf_sv_corr(obs, params, rvr, rvp)
 *                                ...stochastic_volatility.pyx:61, called at cython.py:57,63
 *   pmmh_sv_workspace_bytes      (new) the reference malloc()s scratch inside each call
 *                                (:208-238); here the caller owns one reusable workspace
 *   pmmh_flps_sv_corr_streamed   the same call fed straight from the sampler's host array
 *                                (copies overlap the kernel)
 *   pmmh_split_rvs               rvs.flatten() / rvs[NOBS:] of cython.py:54-56,89-91 plus
 *                                the change to the device's time-major layout
 *   pmmh_norm_cdf                scipy.stats.norm.cdf of cython.py:55,90 / standard.py:52,75
 *   pmmh_importance_discrete     importance_discrete(obs, params, rvr, rvp)
 *                                state/importance_sampling/random_effects.pyx:21,
 *                                called at state/importance_sampling/cython.py:59,89
 *   pmmh_crank_nicolson          _propose_rvs, parameter/mcmc/base_class.py:221-241
 *   pmmh_subsample_indices       np.sort(norm.cdf(u)) + stratified(rnd)
 *                                state/direct/standard.py:52-53,75-76, subsampling.pyx:34-51
 *   pmmh_logistic_loglike        LogisticRegressionModel.get_loglike_gradient
 *                                models/logistic_regression.py:108-176
 *   pmmh_*_host                  the same calls with HOST buffers in the reference's own
 *                                layouts (what a cgo/ctypes/Cython stub would bind 1:1)
 *
 * Conventions: plain pointers and sizes; `d_` = device memory, no prefix = host memory;
 * all reals are IEEE fp64; `stream` is a cudaStream_t passed as void* (NULL = default
 * stream); functions are asynchronous on `stream` unless named *_host; return value 0 on
 * success, non-zero otherwise (pmmh_last_error() gives the text).  The reference's estimator
 * contract maps non-zero / non-finite results to `return False`
 * (state/particle_methods/cython.py:71-75,133-137).  No function allocates device memory
 * except the *_host convenience wrappers.  There is no CPU fallback.
 */
#ifndef PMMH_QN_H
#define PMMH_QN_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMMH_OK 0
#define PMMH_ERR_INVALID 1      /* bad sizes / null pointers */
#define PMMH_ERR_WORKSPACE 2    /* workspace too small */
#define PMMH_ERR_CUDA 3         /* CUDA runtime error */
#define PMMH_ERR_NO_DEVICE 4    /* no sm_100 device */

/* pmmh_bpf_sv_corr leverage-term read */
#define PMMH_BPF_PARITY 0       /* reference behaviour: reads the time-i column (serial chain) */
#define PMMH_BPF_INTENDED 1     /* the evidently intended time i-1 read; NOT reference parity */

/* diag[] layout (int64 per problem) written by the SV kernels */
#define PMMH_DIAG_NEAR_TIES 0   /* ancestor decisions within 64 ulp of a cumulative-weight tie */
#define PMMH_DIAG_MAX_BIN 1     /* largest sort-bin occupancy */
#define PMMH_DIAG_STATUS 2      /* 0 ok, 1 degenerate particle cloud (evaluation abandoned) */
#define PMMH_DIAG_KEY_TIES 3    /* equal keys met while sorting */
#define PMMH_DIAG_WAVEFRONT 4   /* bpf parity mode: deepest dependency chain */
#define PMMH_DIAG_TRAJ_IDX 5    /* bpf: sampled trajectory index */
#define PMMH_DIAG_KERNEL 6      /* kernel that produced the outputs: 1 general, 2 exchange, 3 chain */
#define PMMH_DIAG_FAST_INFO 7   /* exchange kernel: abandon reason (1 run / 2 chunk overflow) |
                                   step << 8 | longest mailbox run << 32 */
#define PMMH_DIAG_COUNT 8

int pmmh_version(void);
const char* pmmh_last_error(void);
/* sm count / compute capability of the current device */
int pmmh_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------- SV particle methods -- */

/* Kernel selection for pmmh_flps_sv_corr (process-wide; default 0).
 *   0  automatic, for log-likelihood + gradient (compute_hessian == 0): the "chain" kernel (one
 *      CTA per problem, everything in shared memory) when n_particles <= 4096 and lag <= 10, the
 *      two-exchange "exchange" kernel for teams of CTAs; the general kernel for everything else
 *      and as the fallback for problems those abandon (degenerate particle clouds)
 *   1  general kernel only
 *   2  diagnostics: exchange kernel where eligible, without the fallback pass
 *   3  diagnostics: chain kernel where eligible, without the fallback pass
 * Call before pmmh_sv_workspace_bytes: the workspace size depends on it. */
int pmmh_sv_set_algorithm(int algorithm);

/* Development hook: d_clocks = device buffer of [n_ctas][16] int64 (or NULL to switch off);
 * the exchange kernel adds the SM clock cycles each CTA spent per phase. */
int pmmh_sv_debug_profile(long long* d_clocks);

/* Bytes of device workspace pmmh_flps_sv_corr / pmmh_bpf_sv_corr need for these sizes.
 * mode: 0 = flps, 1 = bpf.  have_history: caller passes d_x_hist / d_a_hist.
 * ctas_per_problem: 0 = choose automatically. */
int pmmh_sv_workspace_bytes(int n_obs, int n_particles, int lag, int batch, int compute_hessian,
                            int mode, int have_history, int ctas_per_problem, size_t* bytes);

/* Fixed-lag particle smoother for `batch` independent (params, u) problems.
 *   d_obs      [n_obs] (obs_stride = 0) or [batch][obs_stride]
 *   d_params   [batch][4]  = mu, phi, sigma_v, rho
 *   d_rvr      [batch][n_obs]  uniforms for the resampler (Phi already applied)
 *   d_u        [batch][n_obs][n_particles]  time-major: d_u[t][j] == rvp[t + j*n_obs]
 * outputs (per problem): d_filt[n_obs], d_smo[n_obs], d_log_like[1], d_gradient[4][n_obs],
 *   d_traj[n_obs], d_hess1[4][4], d_hess2[4][4], d_diag[PMMH_DIAG_COUNT];
 *   optional d_x_hist / d_a_hist [n_obs][n_particles]: sorted particles and composed
 *   one-step ancestors of every time step (both NULL: only a ring of depth lag+1 is kept). */
int pmmh_flps_sv_corr(const double* d_obs, long long obs_stride, const double* d_params,
                      const double* d_rvr, const double* d_u, int n_obs, int n_particles, int lag,
                      int batch, int compute_hessian, double* d_filt, double* d_smo,
                      double* d_log_like, double* d_gradient, double* d_traj, double* d_hess1,
                      double* d_hess2, long long* d_diag, double* d_x_hist, int* d_a_hist,
                      void* d_workspace, size_t workspace_bytes, int ctas_per_problem, void* stream);

/* The same evaluation (log-likelihood + gradient, one problem) with the auxiliary variables still
 * in HOST memory -- what the samplers hand to estimator.smoother(model, rvs={'rvs': ndarray})
 * (mh_quasi_newton.py:333, state/particle_methods/cython.py:89-91).  h_rvs is the reference's
 * (n_obs, N+1) row-major array (pinned memory makes the copies asynchronous).  The copy engine
 * moves it in chunks of 64 time steps on an internal stream into d_stage (particle-major chunks,
 * no layout kernel) while the persistent kernel is already running; the kernel waits for a
 * chunk only when it reaches it.  d_rvr = Phi of the first n_obs flat entries (computed by the
 * caller as in cython.py:90).  Only sizes the exchange kernel takes (pmmh_sv_streamed_eligible);
 * there is no fallback inside: if d_diag[PMMH_DIAG_STATUS] == 1 afterwards, upload the array and
 * call pmmh_flps_sv_corr.  d_stage needs pmmh_sv_stage_bytes() bytes and must stay untouched
 * until the kernel has finished; the workspace size is that of pmmh_sv_workspace_bytes(batch 1). */
int pmmh_sv_stage_bytes(int n_obs, int n_particles, size_t* bytes);
int pmmh_sv_streamed_eligible(int n_obs, int n_particles, int lag, int ctas_per_problem);
int pmmh_flps_sv_corr_streamed(const double* h_rvs, const double* d_obs, const double* d_params,
                               const double* d_rvr, int n_obs, int n_particles, int lag, void* d_stage,
                               size_t stage_bytes, double* d_filt, double* d_smo, double* d_log_like,
                               double* d_gradient, double* d_traj, double* d_hess1, double* d_hess2,
                               long long* d_diag, void* d_workspace, size_t workspace_bytes,
                               int ctas_per_problem, void* stream);

/* Bootstrap particle filter (filter only).  read_mode: PMMH_BPF_PARITY / PMMH_BPF_INTENDED. */
int pmmh_bpf_sv_corr(const double* d_obs, long long obs_stride, const double* d_params,
                     const double* d_rvr, const double* d_u, int n_obs, int n_particles, int batch,
                     int read_mode, double* d_filt, double* d_log_like, double* d_traj,
                     long long* d_diag, double* d_x_hist, int* d_a_hist, void* d_workspace,
                     size_t workspace_bytes, int ctas_per_problem, void* stream);

/* rvs [batch][n_obs][n_particles + 1] row-major (the reference's dim_rvs) ->
 *   d_r_raw [batch][n_obs]   the first n_obs FLAT entries (not yet Phi-transformed)
 *   d_u     [batch][n_obs][n_particles] time-major view of the flat remainder */
int pmmh_split_rvs(const double* d_rvs, int n_obs, int n_particles, int batch, double* d_r_raw,
                   double* d_u, void* stream);

/* out[k] = Phi(in[k]) (standard normal cdf), in place allowed */
int pmmh_norm_cdf(const double* d_in, double* d_out, long long n, void* stream);

/* ------------------------------------------------------ random-effects importance sampler -- */

/*   d_obs [n_obs] (obs_stride 0) or [batch][obs_stride]; d_params [batch][2] = mu, sigma;
 *   d_rvr [batch] uniforms; d_rvp [batch][n_obs * n_particles] in the reference's flat layout
 *   rvp[i + j*n_obs].  outputs per problem: d_filt[n_obs], d_log_like[1], d_traj[n_obs],
 *   d_gradient[2], d_traj_idx[1]. */
int pmmh_importance_discrete(const double* d_obs, long long obs_stride, const double* d_params,
                             const double* d_rvr, const double* d_rvp, int n_obs, int n_particles,
                             int batch, double* d_filt, double* d_log_like, double* d_traj,
                             double* d_gradient, int* d_traj_idx, void* stream);

/* ------------------------------------------------------------------ Crank-Nicolson on u -- */

/* d_out[k] = sqrt(1 - sigma_u^2) * d_u[k] + sigma_u * xi[k].  If d_xi is NULL, xi is drawn on
 * the device (Philox4x32-10 + Box-Muller, counter = philox_offset + k/2, key = seed). */
int pmmh_crank_nicolson(const double* d_u, const double* d_xi, double* d_out, long long n,
                        double sigma_u, unsigned long long seed, unsigned long long philox_offset,
                        void* stream);

/* ------------------------------------------------------------- data-subsampling estimator -- */

/* Workspace bytes for pmmh_subsample_indices. */
int pmmh_subsample_workspace_bytes(int m, size_t* bytes);

/* d_u [m] standard normals -> d_idx [m] data indices in [0, n_data):
 * sort(Phi(u)) then the stratified merge walk in closed form.  If apply_cdf == 0, d_u is
 * taken as already-transformed uniforms.  d_sorted (optional, [m]) receives the sorted
 * uniforms. */
int pmmh_subsample_indices(const double* d_u, int m, int n_data, int apply_cdf, int* d_idx,
                           double* d_sorted, void* d_workspace, size_t workspace_bytes,
                           void* stream);

/* Subsampled logistic log-likelihood, gradient and (optionally) Hessian over the rows
 * d_idx[0..m) that fall inside [row_begin, row_end) (the shard this device owns; pass 0 and
 * n_data for all).  d_x is [*][d] row-major and holds rows row_begin..row_end-1.
 * d_out [1 + d + d*d]: log_like, gradient[d], hessian[d][d] (Hessian = -sum s_i x_i x_i^T as
 * in logistic_regression.py:155-165; zero-filled when compute_hessian == 0).
 * The sums are reductions (atomic-free, fixed order); d_out is overwritten. */
int pmmh_logistic_workspace_bytes(int m, int d, int compute_hessian, size_t* bytes);
int pmmh_logistic_loglike(const double* d_x, const double* d_y, const int* d_idx, int m, int d,
                          long long row_begin, long long row_end, const double* d_beta,
                          int compute_hessian, double* d_out, void* d_workspace,
                          size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------- host-buffer wrappers -- */

/* Same signature shape as the reference's Cython entry point: host arrays in the reference's
 * own layouts (rvp[i + j*n_obs]); copies in, runs on the current device, copies out,
 * synchronises.  gradient is [4][n_obs]. */
int pmmh_flps_sv_corr_host(const double* obs, const double* params, const double* rvr,
                           const double* rvp, int n_obs, int n_particles, int lag,
                           int compute_hessian, double* filt, double* smo, double* log_like,
                           double* gradient, double* traj, double* hess1, double* hess2,
                           long long* diag);
int pmmh_bpf_sv_corr_host(const double* obs, const double* params, const double* rvr,
                          const double* rvp, int n_obs, int n_particles, int read_mode,
                          double* filt, double* log_like, double* traj, long long* diag);
int pmmh_importance_discrete_host(const double* obs, const double* params, double rvr,
                                  const double* rvp, int n_obs, int n_particles, double* filt,
                                  double* log_like, double* traj, double* gradient);
int pmmh_stratified_host(const double* rnd_sorted, int m, int n_data, int* indices);

#ifdef __cplusplus
}
#endif
#endif /* PMMH_QN_H */
