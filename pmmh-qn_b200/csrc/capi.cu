// capi.cu -- extern "C" boundary of libpmmh_qn_b200.so (declared in include/pmmh_qn.h).
// Plain pointers and sizes in, status codes out; no torch types, no CPU fallback.
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/pmmh_qn.h"
#include "aux_kernels.cuh"
#include "sv_filter.cuh"
#include "sv_grid.cuh"
#include "sv_split.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* what) {
    g_last_error = what;
    return code;
}
int fail_cuda(cudaError_t err, const char* where) {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s: %s", where, cudaGetErrorString(err));
    g_last_error = buf;
    return PMMH_ERR_CUDA;
}
#define PMMH_CUDA(call)                                       \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return fail_cuda(e__, #call); \
    } while (0)

}  // namespace

namespace pmmh {
// error plumbing for the other translation units (sv_split.cu)
int set_error(int code, const char* what) { return fail(code, what); }
int set_cuda_error(cudaError_t err, const char* where) { return fail_cuda(err, where); }
}  // namespace pmmh

namespace {

struct DeviceInfo {
    int sm = 0, major = 0, minor = 0, coop = 0;
    bool ok = false;
};

int get_device_info(DeviceInfo* out) {
    static thread_local DeviceInfo cache[64];
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return fail_cuda(err, "cudaGetDevice");
    if (dev < 0 || dev >= 64) return fail(PMMH_ERR_NO_DEVICE, "device ordinal out of range");
    DeviceInfo& d = cache[dev];
    if (!d.ok) {
        PMMH_CUDA(cudaDeviceGetAttribute(&d.sm, cudaDevAttrMultiProcessorCount, dev));
        PMMH_CUDA(cudaDeviceGetAttribute(&d.major, cudaDevAttrComputeCapabilityMajor, dev));
        PMMH_CUDA(cudaDeviceGetAttribute(&d.minor, cudaDevAttrComputeCapabilityMinor, dev));
        PMMH_CUDA(cudaDeviceGetAttribute(&d.coop, cudaDevAttrCooperativeLaunch, dev));
        d.ok = true;
    }
    *out = d;
    return PMMH_OK;
}

struct SvPlan {
    int G, n_teams, grid, NB, RING, SQ, SQW;
    size_t stamp_bytes, sync_bytes, team_stride, total;
    // exchange kernel (sv_fast.cu); use_fast = 0 when the problem is not eligible
    int use_fast, NSUB, CP;
    int use_chain;
    size_t chain_stride, chain_total;
    int skip_fast;           // the exchange kernel is eligible but the streaming kernels are preferred
    int use_split;           // streaming kernels (sv_split.cu), one problem, log-likelihood + gradient
    int split_path;          // ... with path storage (1) or with records (0)
    size_t split_total;
    size_t fast_sync_bytes, fast_team_stride, fast_total, general_total;
    int use_grid, grid_G;    // grid kernel (sv_grid.cu): one problem, log-likelihood + gradient, one tile per CTA
    size_t grid_total;
};

// 0 = automatic (chain kernel for problems that fit one CTA, exchange kernel for teams of CTAs,
//     general kernel otherwise and as their fallback),
// 1 = general kernel only, 2 = exchange kernel where eligible WITHOUT the fallback pass (diagnostics),
// 3 = chain kernel where eligible WITHOUT the fallback pass (diagnostics)
// The selection is per host thread (thread_local): a thread that calls pmmh_sv_set_algorithm changes the
// sizing and the kernel choice of its own later calls only, so a toggle on one thread cannot land between
// another thread's workspace query and its run call.
thread_local int g_sv_algorithm = 0;
int g_split_min_particles = 1 << 20;   // automatic selection of the streaming kernels from this N on
int g_split_path_max_particles = 1 << 23;   // ... with path storage below this N, with records from it on
int g_grid_min_particles = 1 << 14;         // automatic selection of the grid kernel from this N on (while a tile fits one CTA)
int g_grid_hess_min_particles = 1 << 14;    // ... with the Hessian branch (T = 1000: 51.8 ms against 101.8 ms on the general kernel at 2^14)
long long* g_sv_prof = nullptr;   // development: per-CTA phase clocks of the exchange kernel
constexpr int kMaxDynSmem = 227 * 1024;

int sv_make_plan(int nobs, int n, int lag, int batch, int hess, int mode, int have_hist, int ctas,
                 SvPlan* p, bool for_host_streamed = false) {
    if (nobs < 2 || n < 1 || batch < 1) return fail(PMMH_ERR_INVALID, "sizes must be positive");
    if (mode == pmmh::kSvFlps) {
        if (lag < 2 || lag >= 64) return fail(PMMH_ERR_INVALID, "lag must be in [2, 63]");
        if (nobs < lag + 1) return fail(PMMH_ERR_INVALID, "n_obs must be at least lag + 1");
    } else {
        lag = 2;
    }
    if ((long long)n + nobs >= (1ll << 31)) return fail(PMMH_ERR_INVALID, "n_particles too large");
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc != PMMH_OK) return rc;
    if (di.major < 10) return fail(PMMH_ERR_NO_DEVICE, "an sm_100 (B200) device is required");
    if (!di.coop) return fail(PMMH_ERR_NO_DEVICE, "device lacks cooperative launch");
    // log-likelihood + gradient of a problem that fits one CTA: the chain kernel (everything in
    // shared memory, no exchanges)
    const bool chain_ok = (mode == pmmh::kSvFlps) && pmmh::sv_chain_eligible(n, lag) &&
                          (g_sv_algorithm == 0 || g_sv_algorithm == 3) && (ctas <= 1);
    int G;
    if (ctas > 0) G = ctas;
    else if (chain_ok) G = 1;
    else if (batch > 1) G = (n + 4095) / 4096;
    else G = (n + 1023) / 1024;
    if (G < 1) G = 1;
    if (G > di.sm) G = di.sm;
    int n_teams = di.sm / G;
    if (n_teams > batch) n_teams = batch;
    if (n_teams < 1) n_teams = 1;
    p->G = G;
    p->n_teams = n_teams;
    p->grid = G * n_teams;
    p->NB = ((n + 31) / 32) * 32;
    if (p->NB < 64) p->NB = 64;
    p->RING = lag + 1;
    p->SQ = hess ? (nobs + n - 2) / nobs + 1 : 1;
    p->SQW = (mode != pmmh::kSvFlps) ? (n + nobs - 1) / nobs : 0;
    p->stamp_bytes = pmmh::sv_align((size_t)p->grid * sizeof(unsigned));
    p->sync_bytes = p->stamp_bytes +
                    pmmh::sv_align((size_t)n_teams * 2 * G * pmmh::kMaxAllgatherHost * sizeof(double));
    p->team_stride = pmmh::sv_ws_layout(n, nobs, lag, p->NB, p->RING, hess, mode, p->SQ, p->SQW,
                                        have_hist, nullptr, nullptr);
    p->general_total = p->sync_bytes + (size_t)n_teams * p->team_stride;
    p->total = p->general_total;
    p->use_chain = 0;
    p->chain_stride = p->chain_total = 0;
    if (chain_ok && G == 1 && pmmh::sv_chain_smem_bytes(n) <= kMaxDynSmem) {
        p->use_chain = 1;
        p->chain_stride = pmmh::sv_chain_ws_bytes(n, lag, nobs, hess);
        p->chain_total = (size_t)p->grid * p->chain_stride;
        if (p->chain_total > p->total) p->total = p->chain_total;
    }
    p->use_split = 0;
    p->split_path = 0;
    p->split_total = 0;
    p->use_fast = 0;
    p->NSUB = 0;
    p->fast_sync_bytes = p->fast_team_stride = p->fast_total = 0;
    p->CP = 0;
    // one CTA per problem (batches of small problems) has no exchange to save: the general kernel
    // is the faster one there (measured: 8.8e9 vs 7.0e9 particle-steps/s at 1024 x N=4096)
    const bool want_fast = !p->use_chain && g_sv_algorithm != 4 && g_sv_algorithm != 5 && g_sv_algorithm != 6 && ((g_sv_algorithm == 2) || (g_sv_algorithm == 0 && G > 1));
    if (mode == pmmh::kSvFlps && !hess && want_fast && pmmh::sv_fast_eligible(n, G)) {
        const int S = pmmh::sv_fast_nsub(n, G);
        const int CP = pmmh::sv_fast_pair_cap(n, G);
        const long long nv = (long long)G * G * (pmmh::kFastThreads / 32) * CP;
        if (nv < (1ll << 31) && pmmh::sv_fast_smem_bytes(n, G, S) <= kMaxDynSmem &&
            2 * S + 10 <= pmmh::kMaxAllgatherHost) {
            p->use_fast = 1;
            p->NSUB = S;
            p->CP = CP;
            p->fast_sync_bytes = pmmh::sv_fast_sync_bytes(G, n_teams);
            p->fast_team_stride = pmmh::sv_fast_ws_bytes(n, G, S, CP, lag, have_hist);
            p->fast_total = p->fast_sync_bytes + (size_t)n_teams * p->fast_team_stride;
            if (p->fast_total > p->total) p->total = p->fast_total;
        }
    }
    // one problem, log-likelihood + gradient: the grid kernel (one persistent launch, one tile of the
    // sorted generation per CTA) while N / #SMs fits the shared memory of one CTA
    p->use_grid = 0;
    p->grid_G = 0;
    p->grid_total = 0;
    // (with compute_hessian: the kernel's second instantiation, automatically from g_grid_hess_min_particles on)
    if (mode == pmmh::kSvFlps && batch == 1 && !p->use_chain &&
        (g_sv_algorithm == 6 || (g_sv_algorithm == 0 && ctas == 0 && !for_host_streamed &&
                                 n >= (hess ? g_grid_hess_min_particles : g_grid_min_particles)))) {
        const int GG = pmmh::sv_grid_ctas(n, di.sm, ctas);
        if (pmmh::sv_grid_eligible(nobs, n, lag, GG)) {
            p->use_grid = 1;
            p->grid_G = GG;
            p->grid_total = pmmh::sv_grid_ws_bytes(nobs, n, lag, GG, have_hist, hess);
            // (the general kernel's fallback pass uses the head of the same workspace)
            if (p->grid_total > p->total) p->total = p->grid_total;
        }
    }
    // one large problem, log-likelihood + gradient, no history dump: the streaming kernels -- on
    // request (algorithm 4 / 5), or automatically from N = 2^20 on: there they are as fast as the
    // exchange kernel or faster (174.7 vs 181.0 ms at N = 2^20, T = 1000) and they have no size
    // limit (the exchange kernel stops at ~1.16 M particles; the general kernel is ~4x slower).
    // The host-streamed entry point (pmmh_flps_sv_corr_streamed) stays on the exchange kernel.
    p->skip_fast = 0;
    if (mode == pmmh::kSvFlps && !hess && batch == 1 && !have_hist && ctas == 0 && !p->use_chain && !p->use_grid &&
        (g_sv_algorithm == 4 || g_sv_algorithm == 5 ||
         (g_sv_algorithm == 0 && !for_host_streamed && n >= g_split_min_particles)) &&
        pmmh::sv_split_single_eligible(nobs, n, lag)) {
        p->use_split = 1;
        p->skip_fast = 1;
        // automatic selection: path storage while its jump tables stay (mostly) L2 resident
        // (7.2e9 vs 6.4e9 particle-steps/s at N = 2^22), the record variant from N = 2^23 on
        // (6.45e9 vs 6.24e9 at 2^23, 5.8e9 vs 4.9e9 at 2^24: random 4-byte reads from HBM)
        p->split_path = (g_sv_algorithm == 5 || (g_sv_algorithm == 0 && n < g_split_path_max_particles)) ? 1 : 0;
        p->split_total = p->split_path ? pmmh::sv_split_path_ws_bytes(nobs, n, lag)
                                               : pmmh::sv_split_single_ws_bytes(nobs, n, lag);
        // the general kernel (fallback pass) reuses the head of the same workspace
        if (p->split_total > p->total) p->total = p->split_total;
    }
    return PMMH_OK;
}

int sv_run(int mode, const double* d_obs, long long obs_stride, const double* d_params,
           const double* d_rvr, const double* d_u, int nobs, int n, int lag, int batch, int hess,
           double* d_filt, double* d_smo, double* d_ll, double* d_grad, double* d_traj, double* d_h1,
           double* d_h2, long long* d_diag, double* d_xh, int* d_ah, void* d_ws, size_t ws_bytes,
           int ctas, void* stream) {
    if (!d_obs || !d_params || !d_rvr || !d_u || !d_filt || !d_ll || !d_traj || !d_diag || !d_ws)
        return fail(PMMH_ERR_INVALID, "null pointer argument");
    if (mode == pmmh::kSvFlps && (!d_smo || !d_grad || !d_h1 || !d_h2))
        return fail(PMMH_ERR_INVALID, "null output pointer");
    if ((d_xh == nullptr) != (d_ah == nullptr))
        return fail(PMMH_ERR_INVALID, "d_x_hist and d_a_hist must be given together");
    SvPlan p;
    int rc = sv_make_plan(nobs, n, lag, batch, hess, mode, d_xh != nullptr, ctas, &p);
    if (rc != PMMH_OK) return rc;
    if (ws_bytes < p.total) return fail(PMMH_ERR_WORKSPACE, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    pmmh::SvArgs a;
    memset(&a, 0, sizeof(a));
    a.N = n;
    a.NOBS = nobs;
    a.LAG = (mode == pmmh::kSvFlps) ? lag : 2;
    a.B = batch;
    a.G = p.G;
    a.n_teams = p.n_teams;
    a.NB = p.NB;
    a.RING = p.RING;
    a.mode = mode;
    a.hess = hess;
    a.SQ = p.SQ;
    a.SQW = p.SQW;
    a.obs = d_obs;
    a.obs_stride = obs_stride;
    a.params = d_params;
    a.rvr = d_rvr;
    a.U = d_u;
    a.filt = d_filt;
    a.smo = d_smo;
    a.loglike = d_ll;
    a.grad = d_grad;
    a.traj = d_traj;
    a.hess1 = d_h1;
    a.hess2 = d_h2;
    a.diag = d_diag;
    a.Xhist = d_xh;
    a.Ahist = d_ah;
    a.prof = g_sv_prof;
    a.ws = (char*)d_ws;
    if (p.use_chain) {
        // chain kernel first; problems it abandons (diag status 1) are re-run by the general kernel
        a.ws_sync_bytes = 0;
        a.ws_team_stride = p.chain_stride;
        PMMH_CUDA(pmmh::sv_chain_launch(a, p.grid, st));
        if (g_sv_algorithm == 3) return PMMH_OK;   // diagnostics: no fallback pass
        a.only_failed = 1;
    }
    if (p.use_grid) {
        // grid kernel first; an abandoned evaluation (diag status 1) is re-run by the general kernel
        rc = pmmh::sv_grid_run(d_obs, d_params, d_rvr, d_u, nobs, n, lag, p.grid_G, d_filt, d_smo, d_ll, d_grad,
                               d_traj, d_diag, d_xh, d_ah, d_ws, ws_bytes, g_sv_prof, st, 0, nullptr,
                               hess ? d_h1 : nullptr, hess ? d_h2 : nullptr);
        if (rc != PMMH_OK) return rc;
        if (!hess) {
            PMMH_CUDA(cudaMemsetAsync(d_h1, 0, 16 * sizeof(double), st));
            PMMH_CUDA(cudaMemsetAsync(d_h2, 0, 16 * sizeof(double), st));
        }
        if (g_sv_algorithm == 6) return PMMH_OK;   // diagnostics: no fallback pass
        a.only_failed = 1;
    }
    if (p.use_split) {
        // streaming kernels first; an abandoned evaluation (diag status 1) is re-run by the general
        // kernel in the same stream
        rc = p.split_path
                 ? pmmh::sv_split_path_run(d_obs, d_params, d_rvr, d_u, nobs, n, lag, d_filt, d_smo, d_ll,
                                           d_grad, d_traj, d_diag, d_ws, ws_bytes, st)
                 : pmmh::sv_split_single_run(d_obs, d_params, d_rvr, d_u, nobs, n, lag, d_filt, d_smo, d_ll,
                                             d_grad, d_traj, d_diag, d_ws, ws_bytes, st);
        if (rc != PMMH_OK) return rc;
        PMMH_CUDA(cudaMemsetAsync(d_h1, 0, 16 * sizeof(double), st));
        PMMH_CUDA(cudaMemsetAsync(d_h2, 0, 16 * sizeof(double), st));
        if (g_sv_algorithm == 4 || g_sv_algorithm == 5) return PMMH_OK;   // no fallback pass
        a.only_failed = 1;
    }
    if (p.use_fast && !p.skip_fast && !p.use_grid) {
        // exchange kernel first; problems it abandons (diag status 1) are re-run by the general
        // kernel in the same stream, reusing the workspace
        a.NSUB = p.NSUB;
        a.CP = p.CP;
        a.ws_sync_bytes = p.fast_sync_bytes;
        a.ws_team_stride = p.fast_team_stride;
        PMMH_CUDA(cudaMemsetAsync(d_ws, 0, p.fast_sync_bytes, st));
        PMMH_CUDA(pmmh::sv_fast_launch(a, p.grid, st));
        if (g_sv_algorithm == 2) return PMMH_OK;   // diagnostics: no fallback pass
        a.only_failed = 1;
    }
    a.ws_sync_bytes = p.sync_bytes;
    a.ws_team_stride = p.team_stride;
    PMMH_CUDA(cudaMemsetAsync(d_ws, 0, p.stamp_bytes, st));
    PMMH_CUDA(pmmh::sv_launch(a, p.grid, st));
    return PMMH_OK;
}

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <typename T>
    T* as() {
        return (T*)p;
    }
};

}  // namespace

extern "C" {

int pmmh_version(void) { return 100; }

const char* pmmh_last_error(void) { return g_last_error.c_str(); }

int pmmh_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc != PMMH_OK) return rc;
    if (sm_count) *sm_count = di.sm;
    if (cc_major) *cc_major = di.major;
    if (cc_minor) *cc_minor = di.minor;
    return PMMH_OK;
}

int pmmh_sv_set_algorithm(int algorithm) {
    if (algorithm < 0 || algorithm > 6) return fail(PMMH_ERR_INVALID, "algorithm must be 0 .. 6");
    g_sv_algorithm = algorithm;
    return PMMH_OK;
}

int pmmh_sv_debug_profile(long long* d_clocks) {
    g_sv_prof = d_clocks;
    return PMMH_OK;
}

int pmmh_sv_workspace_bytes(int n_obs, int n_particles, int lag, int batch, int compute_hessian,
                            int mode, int have_history, int ctas_per_problem, size_t* bytes) {
    if (!bytes) return fail(PMMH_ERR_INVALID, "bytes is null");
    SvPlan p;
    int rc = sv_make_plan(n_obs, n_particles, lag, batch, compute_hessian,
                          mode == 0 ? pmmh::kSvFlps : pmmh::kSvBpfParity, have_history,
                          ctas_per_problem, &p);
    if (rc != PMMH_OK) return rc;
    *bytes = p.total;
    return PMMH_OK;
}

int pmmh_flps_sv_corr(const double* d_obs, long long obs_stride, const double* d_params,
                      const double* d_rvr, const double* d_u, int n_obs, int n_particles, int lag,
                      int batch, int compute_hessian, double* d_filt, double* d_smo,
                      double* d_log_like, double* d_gradient, double* d_traj, double* d_hess1,
                      double* d_hess2, long long* d_diag, double* d_x_hist, int* d_a_hist,
                      void* d_workspace, size_t workspace_bytes, int ctas_per_problem, void* stream) {
    return sv_run(pmmh::kSvFlps, d_obs, obs_stride, d_params, d_rvr, d_u, n_obs, n_particles, lag, batch,
                  compute_hessian ? 1 : 0, d_filt, d_smo, d_log_like, d_gradient, d_traj, d_hess1,
                  d_hess2, d_diag, d_x_hist, d_a_hist, d_workspace, workspace_bytes, ctas_per_problem,
                  stream);
}

// ---- auxiliary variables defined by a Philox stream (never stored) --------------------------
int pmmh_flps_sv_corr_philox(const double* d_obs, const double* d_params, const double* d_rvr,
                             unsigned long long seed, unsigned long long philox_offset, int n_obs,
                             int n_particles, int lag, double* d_filt, double* d_smo, double* d_log_like,
                             double* d_gradient, double* d_traj, long long* d_diag, void* d_workspace,
                             size_t workspace_bytes, void* stream) {
    if (!d_obs || !d_params || !d_rvr || !d_filt || !d_smo || !d_log_like || !d_gradient || !d_traj || !d_diag ||
        !d_workspace)
        return fail(PMMH_ERR_INVALID, "pmmh_flps_sv_corr_philox: null pointer argument");
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc != PMMH_OK) return rc;
    if (di.major < 10) return fail(PMMH_ERR_NO_DEVICE, "an sm_100 (B200) device is required");
    if (!pmmh::sv_split_single_eligible(n_obs, n_particles, lag))
        return fail(PMMH_ERR_INVALID, "pmmh_flps_sv_corr_philox: needs lag in [2, 63] and n_obs >= 2 * lag");
    if (workspace_bytes < pmmh::sv_split_path_ws_bytes(n_obs, n_particles, lag))
        return fail(PMMH_ERR_WORKSPACE, "workspace too small");
    return pmmh::sv_split_path_run(d_obs, d_params, d_rvr, nullptr, n_obs, n_particles, lag, d_filt, d_smo,
                                   d_log_like, d_gradient, d_traj, d_diag, d_workspace, workspace_bytes,
                                   (cudaStream_t)stream, 0, nullptr, seed, philox_offset);
}

int pmmh_flps_sv_corr_philox_workspace_bytes(int n_obs, int n_particles, int lag, size_t* bytes) {
    if (!bytes || !pmmh::sv_split_single_eligible(n_obs, n_particles, lag))
        return fail(PMMH_ERR_INVALID, "pmmh_flps_sv_corr_philox_workspace_bytes: bad arguments");
    *bytes = pmmh::sv_split_path_ws_bytes(n_obs, n_particles, lag);
    return PMMH_OK;
}

// ---- host-resident auxiliary variables, copied while the kernel runs ----------------------
namespace {
constexpr int kUChunk = 64;   // time steps per copy chunk (512-byte rows for the copy engine)
struct StreamedState {
    std::vector<cudaEvent_t> ev_chunk;   // streaming kernels: one event per landed chunk
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_reset = nullptr, ev_done = nullptr;
};
thread_local StreamedState g_streamed[64];

int streamed_chunks(int n_obs) { return (n_obs + kUChunk - 1) / kUChunk; }
size_t streamed_data_bytes(int n_obs, int n) {
    return (size_t)streamed_chunks(n_obs) * (size_t)n * kUChunk * sizeof(double);
}
}  // namespace

namespace {
// grid kernel with host-resident u: chunks of g_grid_u_chunk time steps (2 KB rows for the copy engine)
int grid_u_chunk() {
    static int v = 0;
    if (!v) {
        const char* e = getenv("PMMH_GRID_U_CHUNK");
        v = e ? atoi(e) : 256;
        if (v < 16 || v > 4096 || (v & 3)) v = 256;   // a multiple of 4 (32-byte sectors of the staged rows)
    }
    return v;
}
// smallest copy of the tapered schedule of the last chunk slot, in time steps (0 = no taper)
int grid_u_taper_min() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PMMH_GRID_U_TAPER");
        v = e ? atoi(e) : 48;
        if (v < 0 || v > 4096) v = 48;
    }
    return v;
}
// next piece of the copy schedule of the host-streamed grid path (see pmmh_flps_sv_corr_streamed): time steps
// [t0, t0 + return value); never crosses a slot of `ch` steps, ends on a multiple of 4 steps unless it ends the series
int grid_u_next_piece(int t0, int n_obs, int ch, int taper_min) {
    const int c = t0 / ch, in = t0 - c * ch;
    const int rem = n_obs - t0;
    int wsteps = (rem < ch - in) ? rem : ch - in;            // what is left of this slot
    if (taper_min > 0 && rem >= 2 * taper_min) {
        int w = (3 * rem) / 10 + (7 * taper_min) / 10;
        if (w < taper_min) w = taper_min;
        if (w < wsteps) wsteps = w;
    }
    if (t0 + wsteps < n_obs) wsteps = (wsteps + 3) & ~3;   // pieces end on 32-byte sectors of the staged rows (ch % 4 == 0)
    const int whole = (rem < ch - in) ? rem : ch - in;
    // no sliver at the end of a slot (what is left over at the end of the series may be half as short: it is the last piece)
    if (wsteps > whole || (taper_min > 0 && whole - wsteps < (whole == rem ? taper_min / 2 : taper_min))) wsteps = whole;
    return wsteps;
}
size_t streamed_grid_data_bytes(int n_obs, int n) {
    const int ch = grid_u_chunk();
    return (size_t)((n_obs + ch - 1) / ch) * (size_t)n * ch * sizeof(double);
}
// immutable pinned table tab[i] = i: the copy engine writes &tab[rows landed] to the device flag, so
// no call ever rewrites a host value that a queued copy may still read
const int* step_table(int need) {
    static std::mutex mu;
    static int* tab = nullptr;
    static int cap = 0;
    std::lock_guard<std::mutex> lk(mu);
    if (cap < need + 1) {
        int ncap = need + 1 > 65536 ? need + 1 : 65536;
        int* t = nullptr;
        if (cudaHostAlloc((void**)&t, (size_t)ncap * sizeof(int), cudaHostAllocPortable) != cudaSuccess) return nullptr;
        for (int i = 0; i < ncap; ++i) t[i] = i;
        tab = t;   // an older, smaller table stays allocated: copies queued earlier may still read it
        cap = ncap;
    }
    return tab;
}
bool streamed_grid_ok(int n_obs, int n, int lag, int ctas, int* G_out) {
    if (!(g_sv_algorithm == 0 || g_sv_algorithm == 6) || ctas != 0) return false;
    if (g_sv_algorithm == 0 && n < g_grid_min_particles) return false;
    DeviceInfo di;
    if (get_device_info(&di) != PMMH_OK || di.major < 10 || !di.coop) return false;
    const int GG = pmmh::sv_grid_ctas(n, di.sm, 0);
    if (!pmmh::sv_grid_eligible(n_obs, n, lag, GG)) return false;
    if (G_out) *G_out = GG;
    return true;
}
}  // namespace

int pmmh_sv_stream_schedule(int n_obs, int* pieces, int max_pieces) {
    if (n_obs < 1 || (max_pieces > 0 && !pieces)) return -1;
    const int ch = grid_u_chunk(), taper = grid_u_taper_min();
    int t0 = 0, k = 0;
    while (t0 < n_obs) {
        const int w = grid_u_next_piece(t0, n_obs, ch, taper);
        if (k < max_pieces) pieces[k] = w;
        ++k;
        t0 += w;
    }
    return k;
}

int pmmh_sv_stage_bytes(int n_obs, int n_particles, size_t* bytes) {
    if (!bytes || n_obs < 2 || n_particles < 1) return fail(PMMH_ERR_INVALID, "pmmh_sv_stage_bytes: bad arguments");
    size_t b = streamed_data_bytes(n_obs, n_particles);
    const size_t bg = streamed_grid_data_bytes(n_obs, n_particles);
    if (bg > b) b = bg;
    *bytes = b + 256;
    return PMMH_OK;
}

int pmmh_sv_streamed_workspace_bytes(int n_obs, int n_particles, int lag, int ctas_per_problem, size_t* bytes) {
    if (!bytes) return fail(PMMH_ERR_INVALID, "bytes is null");
    SvPlan p;
    int rc = sv_make_plan(n_obs, n_particles, lag, 1, 0, pmmh::kSvFlps, 0, ctas_per_problem, &p, true);
    if (rc != PMMH_OK) return rc;
    size_t b = p.total;
    int GG = 0;
    if (streamed_grid_ok(n_obs, n_particles, lag, ctas_per_problem, &GG)) {
        const size_t g = pmmh::sv_grid_ws_bytes(n_obs, n_particles, lag, GG, 0);
        if (g > b) b = g;
    }
    if (pmmh::sv_split_single_eligible(n_obs, n_particles, lag)) {
        const size_t sp = pmmh::sv_split_path_ws_bytes(n_obs, n_particles, lag);
        if (sp > b) b = sp;
    }
    *bytes = b;
    return PMMH_OK;
}

namespace {
// the streaming kernels take host-resident u as well (and are the preferred kernels from N = 2^20)
bool streamed_prefers_split(int n_obs, int n, int lag, int ctas) {
    return (g_sv_algorithm == 0 || g_sv_algorithm == 5) && ctas == 0 && n >= g_split_min_particles &&
           pmmh::sv_split_single_eligible(n_obs, n, lag);
}
}  // namespace

int pmmh_sv_streamed_eligible(int n_obs, int n_particles, int lag, int ctas_per_problem) {
    SvPlan p;
    if (sv_make_plan(n_obs, n_particles, lag, 1, 0, pmmh::kSvFlps, 0, ctas_per_problem, &p, true) != PMMH_OK) return 0;
    return (p.use_fast || streamed_prefers_split(n_obs, n_particles, lag, ctas_per_problem) ||
            streamed_grid_ok(n_obs, n_particles, lag, ctas_per_problem, nullptr)) ? 1 : 0;
}

int pmmh_flps_sv_corr_streamed(const double* h_rvs, const double* d_obs, const double* d_params,
                               const double* d_rvr, int n_obs, int n_particles, int lag, void* d_stage,
                               size_t stage_bytes, double* d_filt, double* d_smo, double* d_log_like,
                               double* d_gradient, double* d_traj, double* d_hess1, double* d_hess2,
                               long long* d_diag, void* d_workspace, size_t workspace_bytes,
                               int ctas_per_problem, void* stream) {
    if (!h_rvs || !d_obs || !d_params || !d_rvr || !d_stage || !d_filt || !d_smo || !d_log_like || !d_gradient ||
        !d_traj || !d_hess1 || !d_hess2 || !d_diag || !d_workspace)
        return fail(PMMH_ERR_INVALID, "pmmh_flps_sv_corr_streamed: null pointer argument");
    // The grid kernel where it takes the size: the copy engine lays the reference's particle-major host
    // array down in chunks of 256 time steps (2 KB rows), the kernel reads its column of the chunk with a
    // stride and polls one flag per time step; the call is bound by the host link, not by the kernel.
    int GGs = 0;
    if (streamed_grid_ok(n_obs, n_particles, lag, ctas_per_problem, &GGs)) {
        const int ch = grid_u_chunk();
        const size_t gws = pmmh::sv_grid_ws_bytes(n_obs, n_particles, lag, GGs, 0);
        if (workspace_bytes < gws) return fail(PMMH_ERR_WORKSPACE, "workspace too small (pmmh_sv_streamed_workspace_bytes)");
        const size_t gdata = streamed_grid_data_bytes(n_obs, n_particles);
        if (stage_bytes < gdata + 256) return fail(PMMH_ERR_WORKSPACE, "staging buffer too small");
        if (((uintptr_t)d_stage & 31) != 0) return fail(PMMH_ERR_INVALID, "d_stage must be 32-byte aligned (the kernel reads 32-byte sectors of the staged rows)");
        int devg = 0;
        PMMH_CUDA(cudaGetDevice(&devg));
        if (devg < 0 || devg >= 64) return fail(PMMH_ERR_NO_DEVICE, "device ordinal out of range");
        StreamedState& sg = g_streamed[devg];
        if (!sg.copy_stream) {
            PMMH_CUDA(cudaStreamCreateWithFlags(&sg.copy_stream, cudaStreamNonBlocking));
            PMMH_CUDA(cudaEventCreateWithFlags(&sg.ev_start, cudaEventDisableTiming));
            PMMH_CUDA(cudaEventCreateWithFlags(&sg.ev_reset, cudaEventDisableTiming));
        }
        if (!sg.ev_done) PMMH_CUDA(cudaEventCreateWithFlags(&sg.ev_done, cudaEventDisableTiming));
        const int* tab = step_table(n_obs + ch);
        if (!tab) return fail(PMMH_ERR_CUDA, "pinned step table allocation failed");
        cudaStream_t stg = (cudaStream_t)stream;
        char* stageg = (char*)d_stage;
        int* d_flagg = (int*)(stageg + gdata);
        // development: PMMH_STREAM_TIMING=1 prints when the copies and the kernel of this call ended (blocks the host)
        static const bool timing = getenv("PMMH_STREAM_TIMING") != nullptr;
        cudaEvent_t tv[3] = {nullptr, nullptr, nullptr};
        if (timing)
            for (int i = 0; i < 3; ++i) PMMH_CUDA(cudaEventCreate(&tv[i]));
        PMMH_CUDA(cudaEventRecord(sg.ev_start, stg));
        PMMH_CUDA(cudaStreamWaitEvent(sg.copy_stream, sg.ev_start, 0));
        if (timing) PMMH_CUDA(cudaEventRecord(tv[0], sg.copy_stream));
        PMMH_CUDA(cudaMemsetAsync(d_flagg, 0, sizeof(int), sg.copy_stream));
        PMMH_CUDA(cudaEventRecord(sg.ev_reset, sg.copy_stream));
        // Copy schedule.  The kernel needs ~0.11 ms per time step, the host link ~0.157 ms (2 KB rows: 53.5 GB/s; 1 KB:
        // 51.7; 512 B: 48.4; and ~8 ns per row whatever its width, i.e. 8.4 ms for any piece of fewer than ~52 steps at
        // N = 2^20).  The kernel can only start a piece when all of it has landed, so after the last copy it still has
        // to work off what it was behind: a piece of p steps that starts with `rem` steps to go delays the end unless
        // 0.157 p <= (0.157 - 0.11) rem + 0.11 p_last, i.e. p <= 0.3 rem + 0.7 p_last.  Hence: whole slots (widest rows) while
        // that holds, then pieces of ~0.3 x what remains, the last one ~52 steps.  Measured at T = 1000, N = 2^20
        // (PMMH_STREAM_TIMING=1): one copy per slot 157.7 ms of copies + 26.3 ms of kernel after the last one; this
        // schedule: see DESIGN 3.  PMMH_GRID_U_TAPER=0: one copy per slot.  The staging layout does not change.
        int t0 = 0;
        while (t0 < n_obs) {
            const int c = t0 / ch, in = t0 - c * ch;
            const int wsteps = grid_u_next_piece(t0, n_obs, ch, grid_u_taper_min());
            const double* src = h_rvs + (size_t)n_obs + (size_t)t0;   // rvp[i + j * n_obs], cython.py:89-91
            char* dst = stageg + ((size_t)c * (size_t)n_particles * ch + (size_t)in) * sizeof(double);
            PMMH_CUDA(cudaMemcpy2DAsync(dst, (size_t)ch * sizeof(double), src, (size_t)n_obs * sizeof(double),
                                        (size_t)wsteps * sizeof(double), (size_t)n_particles, cudaMemcpyHostToDevice,
                                        sg.copy_stream));
            PMMH_CUDA(cudaMemcpyAsync(d_flagg, &tab[t0 + wsteps], sizeof(int), cudaMemcpyHostToDevice, sg.copy_stream));
            t0 += wsteps;
        }
        if (timing) PMMH_CUDA(cudaEventRecord(tv[1], sg.copy_stream));
        PMMH_CUDA(cudaEventRecord(sg.ev_done, sg.copy_stream));
        PMMH_CUDA(cudaStreamWaitEvent(stg, sg.ev_reset, 0));   // the kernel must not start before the flag is reset
        int rcg = pmmh::sv_grid_run(d_obs, d_params, d_rvr, (const double*)d_stage, n_obs, n_particles, lag, GGs, d_filt,
                                    d_smo, d_log_like, d_gradient, d_traj, d_diag, nullptr, nullptr, d_workspace,
                                    workspace_bytes, g_sv_prof, stg, ch, d_flagg);
        if (rcg != PMMH_OK) return rcg;
        PMMH_CUDA(cudaMemsetAsync(d_hess1, 0, 16 * sizeof(double), stg));
        PMMH_CUDA(cudaMemsetAsync(d_hess2, 0, 16 * sizeof(double), stg));
        // an evaluation that is abandoned early leaves copies behind: the caller's stream ends after them
        PMMH_CUDA(cudaStreamWaitEvent(stg, sg.ev_done, 0));
        if (timing) {
            PMMH_CUDA(cudaEventRecord(tv[2], stg));
            PMMH_CUDA(cudaEventSynchronize(tv[2]));
            float copies_ms = 0.f, all_ms = 0.f;
            cudaEventElapsedTime(&copies_ms, tv[0], tv[1]);
            cudaEventElapsedTime(&all_ms, tv[0], tv[2]);
            fprintf(stderr, "[pmmh stream timing] copies %.2f ms (%.1f GB/s), kernel ends %.2f ms after the last copy\n",
                    copies_ms, (double)n_obs * n_particles * 8.0 / (copies_ms * 1e6), all_ms - copies_ms);
            for (int i = 0; i < 3; ++i) cudaEventDestroy(tv[i]);
        }
        return PMMH_OK;
    }
    // Host-resident u runs on the exchange kernel where it takes the size (measured at N = 2^20:
    // 206 ms per call against 242 ms on the streaming kernels, whose strided reads of the
    // particle-major chunks cost more than they save) and on the streaming kernels beyond it.
    bool exchange_ok = false;
    {
        SvPlan p0;
        if (sv_make_plan(n_obs, n_particles, lag, 1, 0, pmmh::kSvFlps, 0, ctas_per_problem, &p0, true) == PMMH_OK)
            exchange_ok = p0.use_fast != 0;
    }
    if (!exchange_ok && streamed_prefers_split(n_obs, n_particles, lag, ctas_per_problem)) {
        // streaming kernels with path storage: the copy engine fills particle-major chunks of 64
        // time steps; the (host-driven) step loop waits for a chunk's event when it enters it
        if (workspace_bytes < pmmh::sv_split_path_ws_bytes(n_obs, n_particles, lag))
            return fail(PMMH_ERR_WORKSPACE, "workspace too small");
        const size_t data_bytes2 = streamed_data_bytes(n_obs, n_particles);
        if (stage_bytes < data_bytes2 + 256) return fail(PMMH_ERR_WORKSPACE, "staging buffer too small");
        int dev2 = 0;
        PMMH_CUDA(cudaGetDevice(&dev2));
        if (dev2 < 0 || dev2 >= 64) return fail(PMMH_ERR_NO_DEVICE, "device ordinal out of range");
        StreamedState& s2 = g_streamed[dev2];
        const int chunks2 = streamed_chunks(n_obs);
        if (!s2.copy_stream) {
            PMMH_CUDA(cudaStreamCreateWithFlags(&s2.copy_stream, cudaStreamNonBlocking));
            PMMH_CUDA(cudaEventCreateWithFlags(&s2.ev_start, cudaEventDisableTiming));
            PMMH_CUDA(cudaEventCreateWithFlags(&s2.ev_reset, cudaEventDisableTiming));
        }
        while ((int)s2.ev_chunk.size() < chunks2) {
            cudaEvent_t e;
            PMMH_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            s2.ev_chunk.push_back(e);
        }
        cudaStream_t st2 = (cudaStream_t)stream;
        PMMH_CUDA(cudaEventRecord(s2.ev_start, st2));
        PMMH_CUDA(cudaStreamWaitEvent(s2.copy_stream, s2.ev_start, 0));
        for (int c = 0; c < chunks2; ++c) {
            const int t0 = c * kUChunk;
            const int wsteps = (n_obs - t0 < kUChunk) ? (n_obs - t0) : kUChunk;
            const double* src = h_rvs + (size_t)n_obs + (size_t)t0;   // rvp[i + j * n_obs], cython.py:89-91
            char* dst = (char*)d_stage + (size_t)c * (size_t)n_particles * kUChunk * sizeof(double);
            PMMH_CUDA(cudaMemcpy2DAsync(dst, (size_t)kUChunk * sizeof(double), src, (size_t)n_obs * sizeof(double),
                                        (size_t)wsteps * sizeof(double), (size_t)n_particles,
                                        cudaMemcpyHostToDevice, s2.copy_stream));
            PMMH_CUDA(cudaEventRecord(s2.ev_chunk[c], s2.copy_stream));
        }
        PMMH_CUDA(cudaMemsetAsync(d_hess1, 0, 16 * sizeof(double), st2));
        PMMH_CUDA(cudaMemsetAsync(d_hess2, 0, 16 * sizeof(double), st2));
        return pmmh::sv_split_path_run(d_obs, d_params, d_rvr, (const double*)d_stage, n_obs, n_particles, lag,
                                       d_filt, d_smo, d_log_like, d_gradient, d_traj, d_diag, d_workspace,
                                       workspace_bytes, st2, kUChunk, s2.ev_chunk.data());
    }
    SvPlan p;
    int rc = sv_make_plan(n_obs, n_particles, lag, 1, 0, pmmh::kSvFlps, 0, ctas_per_problem, &p, true);
    if (rc != PMMH_OK) return rc;
    if (!p.use_fast) return fail(PMMH_ERR_INVALID, "pmmh_flps_sv_corr_streamed: exchange kernel not eligible for these sizes");
    if (workspace_bytes < p.total) return fail(PMMH_ERR_WORKSPACE, "workspace too small");
    const size_t data_bytes = streamed_data_bytes(n_obs, n_particles);
    if (stage_bytes < data_bytes + 256) return fail(PMMH_ERR_WORKSPACE, "staging buffer too small");
    int dev = 0;
    PMMH_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(PMMH_ERR_NO_DEVICE, "device ordinal out of range");
    StreamedState& ss = g_streamed[dev];
    const int chunks = streamed_chunks(n_obs);
    if (!ss.copy_stream) {
        PMMH_CUDA(cudaStreamCreateWithFlags(&ss.copy_stream, cudaStreamNonBlocking));
        PMMH_CUDA(cudaEventCreateWithFlags(&ss.ev_start, cudaEventDisableTiming));
        PMMH_CUDA(cudaEventCreateWithFlags(&ss.ev_reset, cudaEventDisableTiming));
    }
    if (!ss.ev_done) PMMH_CUDA(cudaEventCreateWithFlags(&ss.ev_done, cudaEventDisableTiming));
    const int* tabx = step_table(n_obs + kUChunk);   // immutable pinned table: no host value is ever rewritten
    if (!tabx) return fail(PMMH_ERR_CUDA, "pinned step table allocation failed");
    cudaStream_t st = (cudaStream_t)stream;
    char* stage = (char*)d_stage;
    int* d_flag = (int*)(stage + data_bytes);
    // the copies may start once everything queued so far on the caller's stream has finished
    // (the previous evaluation may still be reading the staging buffer)
    PMMH_CUDA(cudaEventRecord(ss.ev_start, st));
    PMMH_CUDA(cudaStreamWaitEvent(ss.copy_stream, ss.ev_start, 0));
    PMMH_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), ss.copy_stream));
    PMMH_CUDA(cudaEventRecord(ss.ev_reset, ss.copy_stream));
    for (int c = 0; c < chunks; ++c) {
        const int t0 = c * kUChunk;
        const int wsteps = (n_obs - t0 < kUChunk) ? (n_obs - t0) : kUChunk;
        // rows = particles: rvp[i + j * n_obs] with rvp = rvs_flat + n_obs (cython.py:89-91)
        const double* src = h_rvs + (size_t)n_obs + (size_t)t0;
        char* dst = stage + (size_t)c * (size_t)n_particles * kUChunk * sizeof(double);
        PMMH_CUDA(cudaMemcpy2DAsync(dst, (size_t)kUChunk * sizeof(double), src, (size_t)n_obs * sizeof(double),
                                    (size_t)wsteps * sizeof(double), (size_t)n_particles, cudaMemcpyHostToDevice,
                                    ss.copy_stream));
        PMMH_CUDA(cudaMemcpyAsync(d_flag, &tabx[t0 + wsteps], sizeof(int), cudaMemcpyHostToDevice, ss.copy_stream));
    }
    PMMH_CUDA(cudaEventRecord(ss.ev_done, ss.copy_stream));
    // the kernel must not start before the flag has been reset
    PMMH_CUDA(cudaStreamWaitEvent(st, ss.ev_reset, 0));
    pmmh::SvArgs a;
    memset(&a, 0, sizeof(a));
    a.N = n_particles;
    a.NOBS = n_obs;
    a.LAG = lag;
    a.B = 1;
    a.G = p.G;
    a.n_teams = p.n_teams;
    a.NB = p.NB;
    a.RING = p.RING;
    a.mode = pmmh::kSvFlps;
    a.hess = 0;
    a.SQ = p.SQ;
    a.SQW = p.SQW;
    a.obs = d_obs;
    a.obs_stride = 0;
    a.params = d_params;
    a.rvr = d_rvr;
    a.U = (const double*)d_stage;
    a.u_chunk = kUChunk;
    a.u_ready = d_flag;
    a.filt = d_filt;
    a.smo = d_smo;
    a.loglike = d_log_like;
    a.grad = d_gradient;
    a.traj = d_traj;
    a.hess1 = d_hess1;
    a.hess2 = d_hess2;
    a.diag = d_diag;
    a.prof = g_sv_prof;
    a.ws = (char*)d_workspace;
    a.NSUB = p.NSUB;
    a.CP = p.CP;
    a.ws_sync_bytes = p.fast_sync_bytes;
    a.ws_team_stride = p.fast_team_stride;
    PMMH_CUDA(cudaMemsetAsync(d_workspace, 0, p.fast_sync_bytes, st));
    PMMH_CUDA(pmmh::sv_fast_launch(a, p.grid, st));
    // an evaluation that is abandoned early leaves copies behind: the caller's stream ends after them
    PMMH_CUDA(cudaStreamWaitEvent(st, ss.ev_done, 0));
    return PMMH_OK;
}

// ---- model-generic entry point: the chain kernel instantiated for a model of pf_model.cuh ----------
int pmmh_flps_model_workspace_bytes(int n_obs, int n_particles, int lag, int batch, size_t* bytes) {
    if (!bytes || n_obs < 2 || batch < 1) return fail(PMMH_ERR_INVALID, "pmmh_flps_model_workspace_bytes: bad arguments");
    if (!pmmh::sv_chain_eligible(n_particles, lag) || pmmh::sv_chain_smem_bytes(n_particles) > kMaxDynSmem)
        return fail(PMMH_ERR_INVALID, "pmmh_flps_model_corr: 2 <= n_particles <= 4096 and 2 <= lag <= 10");
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc != PMMH_OK) return rc;
    const int grid = batch < di.sm ? batch : di.sm;
    *bytes = (size_t)grid * pmmh::sv_chain_ws_bytes(n_particles, lag, n_obs, 0);
    return PMMH_OK;
}

int pmmh_flps_model_corr(int model_id, const double* d_obs, long long obs_stride, const double* d_params,
                         const double* d_rvr, const double* d_u, int n_obs, int n_particles, int lag, int batch,
                         double* d_filt, double* d_smo, double* d_log_like, double* d_gradient, double* d_traj,
                         long long* d_diag, double* d_x_hist, int* d_a_hist, void* d_workspace,
                         size_t workspace_bytes, void* stream) {
    if (!d_obs || !d_params || !d_rvr || !d_u || !d_filt || !d_smo || !d_log_like || !d_gradient || !d_traj ||
        !d_diag || !d_workspace)
        return fail(PMMH_ERR_INVALID, "pmmh_flps_model_corr: null pointer argument");
    if (model_id != PMMH_MODEL_SV_LEVERAGE && model_id != PMMH_MODEL_LINEAR_GAUSSIAN &&
        model_id != PMMH_MODEL_LINEAR_GAUSSIAN_FA)
        return fail(PMMH_ERR_INVALID, "pmmh_flps_model_corr: unknown model id");
    if ((d_x_hist == nullptr) != (d_a_hist == nullptr))
        return fail(PMMH_ERR_INVALID, "d_x_hist and d_a_hist must be given together");
    if (n_obs < lag + 1) return fail(PMMH_ERR_INVALID, "n_obs must be at least lag + 1");
    size_t need = 0;
    int rc = pmmh_flps_model_workspace_bytes(n_obs, n_particles, lag, batch, &need);
    if (rc != PMMH_OK) return rc;
    if (workspace_bytes < need) return fail(PMMH_ERR_WORKSPACE, "workspace too small");
    DeviceInfo di;
    rc = get_device_info(&di);
    if (rc != PMMH_OK) return rc;
    if (di.major < 10) return fail(PMMH_ERR_NO_DEVICE, "an sm_100 (B200) device is required");
    pmmh::SvArgs a;
    memset(&a, 0, sizeof(a));
    a.N = n_particles;
    a.NOBS = n_obs;
    a.LAG = lag;
    a.B = batch;
    a.G = 1;
    a.n_teams = batch < di.sm ? batch : di.sm;
    a.mode = pmmh::kSvFlps;
    a.model_id = model_id;
    a.obs = d_obs;
    a.obs_stride = obs_stride;
    a.params = d_params;
    a.rvr = d_rvr;
    a.U = d_u;
    a.filt = d_filt;
    a.smo = d_smo;
    a.loglike = d_log_like;
    a.grad = d_gradient;
    a.traj = d_traj;
    a.hess1 = nullptr;
    a.hess2 = nullptr;
    a.diag = d_diag;
    a.Xhist = d_x_hist;
    a.Ahist = d_a_hist;
    a.prof = nullptr;
    a.ws = (char*)d_workspace;
    a.ws_sync_bytes = 0;
    a.ws_team_stride = pmmh::sv_chain_ws_bytes(n_particles, lag, n_obs, 0);
    PMMH_CUDA(pmmh::sv_chain_launch(a, a.n_teams, (cudaStream_t)stream));
    return PMMH_OK;
}

int pmmh_bpf_sv_corr(const double* d_obs, long long obs_stride, const double* d_params,
                     const double* d_rvr, const double* d_u, int n_obs, int n_particles, int batch,
                     int read_mode, double* d_filt, double* d_log_like, double* d_traj,
                     long long* d_diag, double* d_x_hist, int* d_a_hist, void* d_workspace,
                     size_t workspace_bytes, int ctas_per_problem, void* stream) {
    const int mode = (read_mode == PMMH_BPF_INTENDED) ? pmmh::kSvBpfIntended : pmmh::kSvBpfParity;
    return sv_run(mode, d_obs, obs_stride, d_params, d_rvr, d_u, n_obs, n_particles, 2, batch, 0, d_filt,
                  nullptr, d_log_like, nullptr, d_traj, nullptr, nullptr, d_diag, d_x_hist, d_a_hist,
                  d_workspace, workspace_bytes, ctas_per_problem, stream);
}

int pmmh_split_rvs(const double* d_rvs, int n_obs, int n_particles, int batch, double* d_r_raw,
                   double* d_u, void* stream) {
    if (!d_rvs || !d_r_raw || !d_u || n_obs < 1 || n_particles < 1 || batch < 1)
        return fail(PMMH_ERR_INVALID, "pmmh_split_rvs: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const long long per = (long long)n_obs * ((long long)n_particles + 1);
    PMMH_CUDA(pmmh::launch_copy_head(d_rvs, d_r_raw, n_obs, batch, per, n_obs, st));
    // flat remainder viewed as [n_particles][n_obs] -> [n_obs][n_particles]
    PMMH_CUDA(pmmh::launch_transpose(d_rvs + n_obs, d_u, n_particles, n_obs, batch, per,
                                     (long long)n_obs * n_particles, st));
    return PMMH_OK;
}

int pmmh_norm_cdf(const double* d_in, double* d_out, long long n, void* stream) {
    if (!d_in || !d_out || n < 0) return fail(PMMH_ERR_INVALID, "pmmh_norm_cdf: bad arguments");
    PMMH_CUDA(pmmh::launch_norm_cdf(d_in, d_out, n, (cudaStream_t)stream));
    return PMMH_OK;
}

int pmmh_importance_discrete(const double* d_obs, long long obs_stride, const double* d_params,
                             const double* d_rvr, const double* d_rvp, int n_obs, int n_particles,
                             int batch, double* d_filt, double* d_log_like, double* d_traj,
                             double* d_gradient, int* d_traj_idx, void* stream) {
    if (!d_obs || !d_params || !d_rvr || !d_rvp || !d_filt || !d_log_like || !d_traj || !d_gradient ||
        !d_traj_idx || n_obs < 1 || n_particles < 1 || batch < 1)
        return fail(PMMH_ERR_INVALID, "pmmh_importance_discrete: bad arguments");
    if ((size_t)3 * n_particles * sizeof(double) > 200 * 1024)
        return fail(PMMH_ERR_INVALID, "pmmh_importance_discrete: n_particles above 8533 unsupported");
    PMMH_CUDA(pmmh::launch_importance_discrete(d_obs, obs_stride, d_params, d_rvr, d_rvp, n_obs,
                                               n_particles, batch, d_filt, d_log_like, d_traj,
                                               d_gradient, d_traj_idx, (cudaStream_t)stream));
    return PMMH_OK;
}

int pmmh_crank_nicolson(const double* d_u, const double* d_xi, double* d_out, long long n,
                        double sigma_u, unsigned long long seed, unsigned long long philox_offset,
                        void* stream) {
    if (!d_u || !d_out || n < 0) return fail(PMMH_ERR_INVALID, "pmmh_crank_nicolson: bad arguments");
    const double a = sqrt(1.0 - sigma_u * sigma_u);
    PMMH_CUDA(pmmh::launch_crank_nicolson(d_u, d_xi, d_out, n, a, sigma_u, seed, philox_offset,
                                          (cudaStream_t)stream));
    return PMMH_OK;
}

int pmmh_subsample_workspace_bytes(int m, size_t* bytes) {
    if (!bytes || m < 1) return fail(PMMH_ERR_INVALID, "pmmh_subsample_workspace_bytes: bad arguments");
    *bytes = pmmh::subsample_ws_bytes(m);
    return PMMH_OK;
}

int pmmh_subsample_indices(const double* d_u, int m, int n_data, int apply_cdf, int* d_idx,
                           double* d_sorted, void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!d_u || !d_idx || !d_workspace || m < 1 || n_data < 1)
        return fail(PMMH_ERR_INVALID, "pmmh_subsample_indices: bad arguments");
    if (workspace_bytes < pmmh::subsample_ws_bytes(m)) return fail(PMMH_ERR_WORKSPACE, "workspace too small");
    PMMH_CUDA(pmmh::launch_subsample_indices(d_u, m, n_data, apply_cdf, d_idx, d_sorted, d_workspace,
                                             (cudaStream_t)stream));
    return PMMH_OK;
}

int pmmh_logistic_workspace_bytes(int m, int d, int compute_hessian, size_t* bytes) {
    if (!bytes || m < 1 || d < 1) return fail(PMMH_ERR_INVALID, "pmmh_logistic_workspace_bytes: bad arguments");
    *bytes = pmmh::logistic_ws_bytes(m, d, compute_hessian);
    return PMMH_OK;
}

int pmmh_logistic_loglike(const double* d_x, const double* d_y, const int* d_idx, int m, int d,
                          long long row_begin, long long row_end, const double* d_beta,
                          int compute_hessian, double* d_out, void* d_workspace,
                          size_t workspace_bytes, void* stream) {
    if (!d_x || !d_y || !d_idx || !d_beta || !d_out || !d_workspace || m < 1 || d < 1)
        return fail(PMMH_ERR_INVALID, "pmmh_logistic_loglike: bad arguments");
    if (d > 32) return fail(PMMH_ERR_INVALID, "pmmh_logistic_loglike: d > 32 not supported yet");
    if (workspace_bytes < pmmh::logistic_ws_bytes(m, d, compute_hessian))
        return fail(PMMH_ERR_WORKSPACE, "workspace too small");
    PMMH_CUDA(pmmh::launch_logistic(d_x, d_y, d_idx, m, d, row_begin, row_end, d_beta,
                                    compute_hessian ? 1 : 0, d_out, d_workspace, (cudaStream_t)stream));
    return PMMH_OK;
}

// ------------------------------------------------------------------ host-buffer wrappers

static int sv_host(int mode, const double* obs, const double* params, const double* rvr,
                   const double* rvp, int n_obs, int n, int lag, int hess, int read_mode, double* filt,
                   double* smo, double* log_like, double* gradient, double* traj, double* hess1,
                   double* hess2, long long* diag) {
    if (!obs || !params || !rvr || !rvp || !filt || !log_like || !traj)
        return fail(PMMH_ERR_INVALID, "null host pointer");
    const size_t nt = (size_t)n_obs * n;
    size_t ws_bytes = 0;
    int rc = pmmh_sv_workspace_bytes(n_obs, n, lag, 1, hess, mode, 0, 0, &ws_bytes);
    if (rc != PMMH_OK) return rc;
    DevBuf b_obs, b_par, b_rvr, b_rvp, b_u, b_out, b_diag, b_ws;
    const size_t n_out = (size_t)n_obs * 7 + 1 + 32;   // filt smo traj grad[4] ll h1 h2
    PMMH_CUDA(b_obs.alloc(n_obs * sizeof(double)));
    PMMH_CUDA(b_par.alloc(4 * sizeof(double)));
    PMMH_CUDA(b_rvr.alloc(n_obs * sizeof(double)));
    PMMH_CUDA(b_rvp.alloc(nt * sizeof(double)));
    PMMH_CUDA(b_u.alloc(nt * sizeof(double)));
    PMMH_CUDA(b_out.alloc(n_out * sizeof(double)));
    PMMH_CUDA(b_diag.alloc(PMMH_DIAG_COUNT * sizeof(long long)));
    PMMH_CUDA(b_ws.alloc(ws_bytes));
    PMMH_CUDA(cudaMemcpy(b_obs.p, obs, n_obs * sizeof(double), cudaMemcpyHostToDevice));
    PMMH_CUDA(cudaMemcpy(b_par.p, params, 4 * sizeof(double), cudaMemcpyHostToDevice));
    PMMH_CUDA(cudaMemcpy(b_rvr.p, rvr, n_obs * sizeof(double), cudaMemcpyHostToDevice));
    PMMH_CUDA(cudaMemcpy(b_rvp.p, rvp, nt * sizeof(double), cudaMemcpyHostToDevice));
    PMMH_CUDA(cudaMemset(b_out.p, 0, n_out * sizeof(double)));
    PMMH_CUDA(cudaMemset(b_diag.p, 0, PMMH_DIAG_COUNT * sizeof(long long)));
    PMMH_CUDA(pmmh::launch_transpose(b_rvp.as<double>(), b_u.as<double>(), n, n_obs, 1, 0, 0, 0));
    double* o = b_out.as<double>();
    double* d_filt = o;
    double* d_smo = o + n_obs;
    double* d_traj = o + 2 * (size_t)n_obs;
    double* d_grad = o + 3 * (size_t)n_obs;
    double* d_ll = o + 7 * (size_t)n_obs;
    double* d_h1 = d_ll + 1;
    double* d_h2 = d_h1 + 16;
    if (mode == 0)
        rc = pmmh_flps_sv_corr(b_obs.as<double>(), 0, b_par.as<double>(), b_rvr.as<double>(),
                               b_u.as<double>(), n_obs, n, lag, 1, hess, d_filt, d_smo, d_ll, d_grad,
                               d_traj, d_h1, d_h2, b_diag.as<long long>(), nullptr, nullptr, b_ws.p,
                               ws_bytes, 0, nullptr);
    else
        rc = pmmh_bpf_sv_corr(b_obs.as<double>(), 0, b_par.as<double>(), b_rvr.as<double>(),
                              b_u.as<double>(), n_obs, n, 1, read_mode, d_filt, d_ll, d_traj,
                              b_diag.as<long long>(), nullptr, nullptr, b_ws.p, ws_bytes, 0, nullptr);
    if (rc != PMMH_OK) return rc;
    PMMH_CUDA(cudaDeviceSynchronize());
    std::vector<double> h(n_out);
    PMMH_CUDA(cudaMemcpy(h.data(), o, n_out * sizeof(double), cudaMemcpyDeviceToHost));
    memcpy(filt, h.data(), n_obs * sizeof(double));
    if (smo) memcpy(smo, h.data() + n_obs, n_obs * sizeof(double));
    memcpy(traj, h.data() + 2 * (size_t)n_obs, n_obs * sizeof(double));
    if (gradient) memcpy(gradient, h.data() + 3 * (size_t)n_obs, 4 * (size_t)n_obs * sizeof(double));
    *log_like = h[7 * (size_t)n_obs];
    if (hess1) memcpy(hess1, h.data() + 7 * (size_t)n_obs + 1, 16 * sizeof(double));
    if (hess2) memcpy(hess2, h.data() + 7 * (size_t)n_obs + 17, 16 * sizeof(double));
    if (diag)
        PMMH_CUDA(cudaMemcpy(diag, b_diag.p, PMMH_DIAG_COUNT * sizeof(long long), cudaMemcpyDeviceToHost));
    return PMMH_OK;
}

int pmmh_flps_sv_corr_host(const double* obs, const double* params, const double* rvr,
                           const double* rvp, int n_obs, int n_particles, int lag,
                           int compute_hessian, double* filt, double* smo, double* log_like,
                           double* gradient, double* traj, double* hess1, double* hess2,
                           long long* diag) {
    if (!smo || !gradient || !hess1 || !hess2) return fail(PMMH_ERR_INVALID, "null host pointer");
    return sv_host(0, obs, params, rvr, rvp, n_obs, n_particles, lag, compute_hessian ? 1 : 0, 0, filt,
                   smo, log_like, gradient, traj, hess1, hess2, diag);
}

int pmmh_bpf_sv_corr_host(const double* obs, const double* params, const double* rvr,
                          const double* rvp, int n_obs, int n_particles, int read_mode,
                          double* filt, double* log_like, double* traj, long long* diag) {
    return sv_host(1, obs, params, rvr, rvp, n_obs, n_particles, 2, 0, read_mode, filt, nullptr,
                   log_like, nullptr, traj, nullptr, nullptr, diag);
}

int pmmh_importance_discrete_host(const double* obs, const double* params, double rvr,
                                  const double* rvp, int n_obs, int n_particles, double* filt,
                                  double* log_like, double* traj, double* gradient) {
    if (!obs || !params || !rvp || !filt || !log_like || !traj || !gradient)
        return fail(PMMH_ERR_INVALID, "null host pointer");
    const size_t nt = (size_t)n_obs * n_particles;
    DevBuf b_obs, b_par, b_rvr, b_rvp, b_out, b_idx;
    const size_t n_out = (size_t)2 * n_obs + 3;
    PMMH_CUDA(b_obs.alloc(n_obs * sizeof(double)));
    PMMH_CUDA(b_par.alloc(2 * sizeof(double)));
    PMMH_CUDA(b_rvr.alloc(sizeof(double)));
    PMMH_CUDA(b_rvp.alloc(nt * sizeof(double)));
    PMMH_CUDA(b_out.alloc(n_out * sizeof(double)));
    PMMH_CUDA(b_idx.alloc(sizeof(int)));
    PMMH_CUDA(cudaMemcpy(b_obs.p, obs, n_obs * sizeof(double), cudaMemcpyHostToDevice));
    PMMH_CUDA(cudaMemcpy(b_par.p, params, 2 * sizeof(double), cudaMemcpyHostToDevice));
    PMMH_CUDA(cudaMemcpy(b_rvr.p, &rvr, sizeof(double), cudaMemcpyHostToDevice));
    PMMH_CUDA(cudaMemcpy(b_rvp.p, rvp, nt * sizeof(double), cudaMemcpyHostToDevice));
    double* o = b_out.as<double>();
    int rc = pmmh_importance_discrete(b_obs.as<double>(), 0, b_par.as<double>(), b_rvr.as<double>(),
                                      b_rvp.as<double>(), n_obs, n_particles, 1, o, o + 2 * (size_t)n_obs,
                                      o + n_obs, o + 2 * (size_t)n_obs + 1, b_idx.as<int>(), nullptr);
    if (rc != PMMH_OK) return rc;
    PMMH_CUDA(cudaDeviceSynchronize());
    std::vector<double> h(n_out);
    PMMH_CUDA(cudaMemcpy(h.data(), o, n_out * sizeof(double), cudaMemcpyDeviceToHost));
    memcpy(filt, h.data(), n_obs * sizeof(double));
    memcpy(traj, h.data() + n_obs, n_obs * sizeof(double));
    *log_like = h[2 * (size_t)n_obs];
    gradient[0] = h[2 * (size_t)n_obs + 1];
    gradient[1] = h[2 * (size_t)n_obs + 2];
    return PMMH_OK;
}

int pmmh_stratified_host(const double* rnd_sorted, int m, int n_data, int* indices) {
    if (!rnd_sorted || !indices || m < 1 || n_data < 1) return fail(PMMH_ERR_INVALID, "bad arguments");
    size_t ws_bytes = pmmh::subsample_ws_bytes(m);
    DevBuf b_u, b_idx, b_ws;
    PMMH_CUDA(b_u.alloc(m * sizeof(double)));
    PMMH_CUDA(b_idx.alloc(m * sizeof(int)));
    PMMH_CUDA(b_ws.alloc(ws_bytes));
    PMMH_CUDA(cudaMemcpy(b_u.p, rnd_sorted, m * sizeof(double), cudaMemcpyHostToDevice));
    int rc = pmmh_subsample_indices(b_u.as<double>(), m, n_data, 0, b_idx.as<int>(), nullptr, b_ws.p,
                                    ws_bytes, nullptr);
    if (rc != PMMH_OK) return rc;
    PMMH_CUDA(cudaDeviceSynchronize());
    PMMH_CUDA(cudaMemcpy(indices, b_idx.p, m * sizeof(int), cudaMemcpyDeviceToHost));
    return PMMH_OK;
}

}  // extern "C"
