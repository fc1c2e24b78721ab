// aux_kernels.cu -- the smaller device kernels of the pmmh-qn hot path:
//   * layout change of the auxiliary variables u (reference flat layout -> time-major)
//   * Phi (standard normal cdf), Crank-Nicolson update of u (optionally with Philox noise)
//   * correlated importance sampler of the random-effects model
//   * data-subsampling estimator: sort of Phi(u), stratified indices, logistic gather-reduce
//
// Reference paths relative to /root/reference/python; see include/pmmh_qn.h for the mapping.
// All of it is HBM/L2-bound integer and fp64 streaming work: coalesced loads, shared-memory
// tiles where a transpose or a reduction needs them, no tensor cores.  -fmad=false.
#include "aux_kernels.cuh"
#include "philox.cuh"

#include <math.h>

#include "common.cuh"

namespace pmmh {

// ------------------------------------------------------------------------------------------
// transpose: in [rows][cols] row-major -> out [cols][rows]   (rvp[j*NOBS + i] -> u[i][j])
// ------------------------------------------------------------------------------------------
__global__ void transpose_f64_kernel(const double* __restrict__ in, double* __restrict__ out,
                                     long long rows, int cols, long long in_batch_stride,
                                     long long out_batch_stride) {
    __shared__ double tile[32][33];
    const double* src = in + (size_t)blockIdx.z * in_batch_stride;
    double* dst = out + (size_t)blockIdx.z * out_batch_stride;
    const long long r0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const long long r = r0 + k;
        const int cc = c0 + threadIdx.x;
        if (r < rows && cc < cols) tile[k][threadIdx.x] = src[(size_t)r * cols + cc];
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int cc = c0 + k;
        const long long r = r0 + threadIdx.x;
        if (r < rows && cc < cols) dst[(size_t)cc * rows + r] = tile[threadIdx.x][k];
    }
}

cudaError_t launch_transpose(const double* in, double* out, long long rows, int cols, int batch,
                             long long in_stride, long long out_stride, cudaStream_t st) {
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32), (unsigned)batch);
    dim3 block(32, 8);
    transpose_f64_kernel<<<grid, block, 0, st>>>(in, out, rows, cols, in_stride, out_stride);
    return cudaGetLastError();
}

__global__ void copy_head_kernel(const double* __restrict__ in, double* __restrict__ out, int n,
                                 long long in_stride, long long out_stride) {
    const double* src = in + (size_t)blockIdx.y * in_stride;
    double* dst = out + (size_t)blockIdx.y * out_stride;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
        dst[k] = src[k];
}

cudaError_t launch_copy_head(const double* in, double* out, int n, int batch, long long in_stride,
                             long long out_stride, cudaStream_t st) {
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)batch);
    copy_head_kernel<<<grid, 256, 0, st>>>(in, out, n, in_stride, out_stride);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Phi and Crank-Nicolson
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double norm_cdf_dev(double x) {
    return 0.5 * erfc(-x * 0.70710678118654752440);
}

__global__ void norm_cdf_kernel(const double* __restrict__ in, double* __restrict__ out, long long n) {
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += (long long)gridDim.x * blockDim.x)
        out[k] = norm_cdf_dev(in[k]);
}

cudaError_t launch_norm_cdf(const double* in, double* out, long long n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int grid = (int)min((long long)148 * 8, (n + 255) / 256);
    norm_cdf_kernel<<<grid, 256, 0, st>>>(in, out, n);
    return cudaGetLastError();
}

// parameter/mcmc/base_class.py:231-233:  out = sqrt(1 - s^2) * u + s * xi   (two roundings)
__global__ void crank_nicolson_kernel(const double* __restrict__ u, const double* __restrict__ xi,
                                      double* __restrict__ out, long long n, double a, double b,
                                      unsigned long long seed, unsigned long long offset) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (xi) {
        for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
            const double mean = a * u[k];
            out[k] = mean + b * xi[k];
        }
        return;
    }
    const long long npair = (n + 1) / 2;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npair; p += stride) {
        const unsigned long long ctr = offset + (unsigned long long)p;
        uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        const double u1 = u53(c[0], c[1]), u2 = u53(c[2], c[3]);
        const double r = sqrt(-2.0 * log(u1));
        double sn, cs;
        sincospi(2.0 * u2, &sn, &cs);
        const long long k0 = 2 * p, k1 = 2 * p + 1;
        out[k0] = a * u[k0] + b * (r * cs);
        if (k1 < n) out[k1] = a * u[k1] + b * (r * sn);
    }
}

cudaError_t launch_crank_nicolson(const double* u, const double* xi, double* out, long long n,
                                  double a, double b, unsigned long long seed,
                                  unsigned long long offset, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const long long work = xi ? n : (n + 1) / 2;
    const int grid = (int)min((long long)148 * 16, (work + 255) / 256);
    crank_nicolson_kernel<<<grid, 256, 0, st>>>(u, xi, out, n, a, b, seed, offset);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// importance_discrete, state/importance_sampling/random_effects.pyx:21-104
// one CTA per evaluation; dynamic smem: W[N] (joint log-weights, then normalised weights),
// A1[N] = sum_i (x_ij - mu), A2[N] = sum_i (x_ij - mu)^2
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) importance_discrete_kernel(
    const double* __restrict__ obs, long long obs_stride, const double* __restrict__ params,
    const double* __restrict__ rvr, const double* __restrict__ rvp, int NOBS, int N,
    double* __restrict__ filt, double* __restrict__ loglike, double* __restrict__ traj,
    double* __restrict__ grad, int* __restrict__ traj_idx) {
    extern __shared__ double sm[];
    double* W = sm;
    double* A1 = sm + N;
    double* A2 = sm + 2 * (size_t)N;
    __shared__ double s_red[3 * 32];
    __shared__ double s_scal[4];
    __shared__ int s_idx;
    const int prob = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const double* y = obs + (size_t)prob * obs_stride;
    const double mu = params[(size_t)prob * 2], sigma = params[(size_t)prob * 2 + 1];
    const double* rp = rvp + (size_t)prob * NOBS * N;

    // :54-67 particles and joint log-weights (warp per particle, lanes over observations)
    for (int j = warp; j < N; j += nwarp) {
        double wsum = 0.0, a1 = 0.0, a2 = 0.0;
        for (int i = lane; i < NOBS; i += 32) {
            const double p = mu + sigma * rp[(size_t)i + (size_t)j * NOBS];
            const double part3 = -0.5 * (y[i] - p) * (y[i] - p) / (1.0 * 1.0);
            const double g = -0.91893853320467267 + (-log(1.0)) + part3;
            if (isfinite(g)) wsum += g;
            const double d = p - mu;
            a1 += d;
            a2 += d * d;
        }
        wsum = warp_sum(wsum);
        a1 = warp_sum(a1);
        a2 = warp_sum(a2);
        if (lane == 0) {
            W[j] = wsum;
            A1[j] = a1;
            A2[j] = a2;
        }
    }
    __syncthreads();
    // :69 my_max quirk (Q4): last j with W[j] > W[0] and finite
    {
        int best = 0;
        const double w0 = W[0];
        for (int j = tid; j < N; j += blockDim.x)
            if (j >= 1 && W[j] > w0 && isfinite(W[j])) best = max(best, j);
        best = warp_max(best);
        if (lane == 0) s_red[warp] = (double)best;
        __syncthreads();
        if (warp == 0) {
            int b = (lane < nwarp) ? (int)s_red[lane] : 0;
            b = warp_max(b);
            if (lane == 0) s_scal[0] = W[b];
        }
        __syncthreads();
    }
    const double shift = s_scal[0];
    // :70-76
    double v[1] = {0.0};
    for (int j = tid; j < N; j += blockDim.x) {
        double sh = exp(W[j] - shift);
        if (!isfinite(sh)) sh = 0.0;
        W[j] = sh;
        v[0] += sh;
    }
    block_sum<1>(v, s_red);
    const double norm = v[0];
    // :79
    if (tid == 0) loglike[prob] = shift + log(norm) - NOBS * log((double)N);
    // :82-85 weights, filtered means (thread per observation, coalesced over i)
    for (int j = tid; j < N; j += blockDim.x) W[j] = W[j] / norm;
    __syncthreads();
    for (int i = tid; i < NOBS; i += blockDim.x) {
        double f = 0.0;
        for (int j = 0; j < N; ++j) f += W[j] * (mu + sigma * rp[(size_t)i + (size_t)j * NOBS]);
        filt[(size_t)prob * NOBS + i] = f;
    }
    // :88-90 trajectory: sampleParticle_corr, sequential as in the reference (:128-149)
    if (tid == 0) {
        const double rnd = rvr[prob];
        double sum = W[0];
        for (int j = 1; j < N; ++j) sum += W[j];
        // the reference accumulates cum and sum with identical additions, so sum == cum[N-1]
        double cum = W[0];
        int cur = 0;
        for (int j = 0; j < N; ++j) {
            const double cn = (cur == 0) ? W[0] : cum / sum;
            if (cn < rnd) {
                cur++;
                if (cur < N) cum = cum + W[cur];
            } else break;
        }
        if (cur >= N) cur = N - 1;   // reference reads out of bounds here; out of contract
        s_idx = cur;
        traj_idx[prob] = cur;
    }
    __syncthreads();
    {
        const int idx = s_idx;
        for (int i = tid; i < NOBS; i += blockDim.x)
            traj[(size_t)prob * NOBS + i] = mu + sigma * rp[(size_t)i + (size_t)idx * NOBS];
    }
    // :93-99 gradient wrt (mu, log sigma)
    {
        const double is2 = 1.0 / (sigma * sigma);
        double gacc[2] = {0.0, 0.0};
        for (int j = tid; j < N; j += blockDim.x) {
            gacc[0] += W[j] * (is2 * A1[j]);
            gacc[1] += W[j] * (is2 * A2[j] - (double)NOBS);
        }
        block_sum<2>(gacc, s_red);
        if (tid == 0) {
            grad[(size_t)prob * 2] = gacc[0];
            grad[(size_t)prob * 2 + 1] = gacc[1];
        }
    }
}

cudaError_t launch_importance_discrete(const double* obs, long long obs_stride, const double* params,
                                       const double* rvr, const double* rvp, int nobs, int n, int batch,
                                       double* filt, double* ll, double* traj, double* grad,
                                       int* traj_idx, cudaStream_t st) {
    const size_t smem = (size_t)3 * n * sizeof(double);
    cudaError_t err = cudaFuncSetAttribute(importance_discrete_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    importance_discrete_kernel<<<batch, 256, smem, st>>>(obs, obs_stride, params, rvr, rvp, nobs, n,
                                                         filt, ll, traj, grad, traj_idx);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// data subsampling: sort(Phi(u)) + stratified (standard.py:75-76, subsampling.pyx:34-51)
// keys are uniform on [0,1] so m equal-width bins hold ~Poisson(1) keys each.
// ------------------------------------------------------------------------------------------
__global__ void ss_hist_kernel(const double* __restrict__ u, int m, int apply_cdf,
                               double* __restrict__ key, int* __restrict__ hist,
                               int* __restrict__ rnk) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) {
        const double v = apply_cdf ? norm_cdf_dev(u[k]) : u[k];
        double t = v * (double)m;
        int b = (!(t >= 0.0)) ? 0 : (t >= (double)m ? m - 1 : (int)t);
        key[k] = v;
        rnk[k] = atomicAdd(&hist[b], 1);
    }
}

// single-CTA exclusive scan of hist[0..m) -> start[0..m], start[m] = total
__global__ void __launch_bounds__(1024) ss_scan_kernel(const int* __restrict__ hist,
                                                       int* __restrict__ start, int m) {
    __shared__ int s_wtot[32], s_wbase[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Tile t;
    t.init(m, 1, 0);
    const int sb = t.seg_begin(warp), se = t.seg_end(warp);
    int carry = 0;
    for (int base = sb; base < se; base += 32) {
        const int b = base + lane;
        const int v = (b < se) ? hist[b] : 0;
        const int incl = warp_incl_scan(v, lane);
        carry += __shfl_sync(kFullMask, incl, 31);
    }
    if (lane == 0) s_wtot[warp] = carry;
    __syncthreads();
    if (warp == 0) {
        const int v = s_wtot[lane];
        const int incl = warp_incl_scan(v, lane);
        s_wbase[lane] = incl - v;
        if (lane == 31) start[m] = incl;
    }
    __syncthreads();
    const int wb = s_wbase[warp];
    carry = 0;
    for (int base = sb; base < se; base += 32) {
        const int b = base + lane;
        const int v = (b < se) ? hist[b] : 0;
        const int incl = warp_incl_scan(v, lane);
        if (b < se) start[b] = wb + carry + (incl - v);
        carry += __shfl_sync(kFullMask, incl, 31);
    }
}

__global__ void ss_scatter_kernel(const double* __restrict__ key, const int* __restrict__ rnk,
                                  const int* __restrict__ start, int m, double* __restrict__ tkey,
                                  int* __restrict__ tidx) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) {
        const double v = key[k];
        double t = v * (double)m;
        int b = (!(t >= 0.0)) ? 0 : (t >= (double)m ? m - 1 : (int)t);
        const int slot = start[b] + rnk[k];
        tkey[slot] = v;
        tidx[slot] = k;
    }
}

// orders each bin by (key, original index), then maps position p to its data index
// (stratified: first k with (k + 1.0) / n >= (r_p + p) / m, clamped to n - 1)
__global__ void ss_rank_kernel(const double* __restrict__ tkey, const int* __restrict__ tidx,
                               const int* __restrict__ start, int m, int n,
                               double* __restrict__ sorted, int* __restrict__ idx) {
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < m; s += gridDim.x * blockDim.x) {
        const double v = tkey[s];
        const int oi = tidx[s];
        double t = v * (double)m;
        int b = (!(t >= 0.0)) ? 0 : (t >= (double)m ? m - 1 : (int)t);
        const int st = start[b], en = start[b + 1];
        int rank = 0;
        for (int q = st; q < en; ++q) {
            if (q == s) continue;
            const double k2 = tkey[q];
            if (k2 < v || (k2 == v && tidx[q] < oi)) rank++;
        }
        const int p = st + rank;
        if (sorted) sorted[p] = v;
        const double cp = (v + (double)p) / (double)m;
        long long k0 = (long long)ceil(cp * (double)n) - 1;
        if (k0 < 0) k0 = 0;
        if (k0 > n - 1) k0 = n - 1;
        while (k0 > 0 && ((double)(k0 - 1) + 1.0) / (double)n >= cp) k0--;
        while (k0 < n - 1 && ((double)k0 + 1.0) / (double)n < cp) k0++;
        idx[p] = (int)k0;
    }
}

size_t subsample_ws_bytes(int m) {
    size_t off = 0;
    off += sv_align_aux((size_t)m * sizeof(double));       // key
    off += sv_align_aux((size_t)m * sizeof(double));       // tkey
    off += sv_align_aux((size_t)m * sizeof(int));          // rnk
    off += sv_align_aux((size_t)m * sizeof(int));          // tidx
    off += sv_align_aux((size_t)m * sizeof(int));          // hist
    off += sv_align_aux((size_t)(m + 1) * sizeof(int));    // start
    return off;
}

cudaError_t launch_subsample_indices(const double* u, int m, int n, int apply_cdf, int* idx,
                                     double* sorted, void* ws, cudaStream_t st) {
    char* base = (char*)ws;
    size_t off = 0;
    double* key = (double*)(base + off);
    off += sv_align_aux((size_t)m * sizeof(double));
    double* tkey = (double*)(base + off);
    off += sv_align_aux((size_t)m * sizeof(double));
    int* rnk = (int*)(base + off);
    off += sv_align_aux((size_t)m * sizeof(int));
    int* tidx = (int*)(base + off);
    off += sv_align_aux((size_t)m * sizeof(int));
    int* hist = (int*)(base + off);
    off += sv_align_aux((size_t)m * sizeof(int));
    int* start = (int*)(base + off);
    cudaError_t err = cudaMemsetAsync(hist, 0, (size_t)m * sizeof(int), st);
    if (err != cudaSuccess) return err;
    const int grid = min(148 * 8, (m + 255) / 256);
    ss_hist_kernel<<<grid, 256, 0, st>>>(u, m, apply_cdf, key, hist, rnk);
    ss_scan_kernel<<<1, 1024, 0, st>>>(hist, start, m);
    ss_scatter_kernel<<<grid, 256, 0, st>>>(key, rnk, start, m, tkey, tidx);
    ss_rank_kernel<<<grid, 256, 0, st>>>(tkey, tidx, start, m, n, sorted, idx);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// subsampled logistic log-likelihood / gradient / Hessian
// models/logistic_regression.py:108-176.  Warp per row (224-byte rows at d = 28: 7 aligned
// sectors), lane k owns feature k; per-warp partials go to the workspace and a second kernel
// sums them in a fixed order (deterministic, atomic-free).
// ------------------------------------------------------------------------------------------
template <bool HESS>
__global__ void __launch_bounds__(256) logistic_partial_kernel(
    const double* __restrict__ x, const double* __restrict__ y, const int* __restrict__ idx, int m,
    int d, long long row_begin, long long row_end, const double* __restrict__ beta,
    double* __restrict__ partial, int stride) {
    extern __shared__ double s_acc[];   // [stride] block accumulator
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const double bk = (lane < d) ? beta[lane] : 0.0;
    double ll = 0.0, gk = 0.0;
    double h[HESS ? 32 : 1];
#pragma unroll
    for (int l = 0; l < (HESS ? 32 : 1); ++l) h[l] = 0.0;
    constexpr int U = HESS ? 2 : 4;   // rows in flight per warp (independent gathers)
    for (int r0 = gwarp; r0 < m; r0 += U * nwarps) {
        double xk[U], yy[U];
        bool ok[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
            const int r = r0 + q * nwarps;
            long long row = -1;
            if (r < m) row = idx[r];
            ok[q] = (row >= row_begin && row < row_end);   // warp-uniform
            const size_t lrow = ok[q] ? (size_t)(row - row_begin) : 0;
            xk[q] = (ok[q] && lane < d) ? x[lrow * d + lane] : 0.0;
            yy[q] = ok[q] ? y[lrow] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < U; ++q) {
            if (!ok[q]) continue;
            const double xb = warp_sum(bk * xk[q]);
            const double en = exp(-1.0 * xb), ep = exp(xb);
            const double eta = 1.0 / (1.0 + en);
            double e1 = log(eta), e0 = log(1.0 - eta);
            if (isinf(e1)) e1 = 0.0;
            if (isinf(e0)) e0 = 0.0;
            ll += yy[q] * e1 + (1.0 - yy[q]) * e0;
            const double g1 = xk[q] / (1.0 + ep);
            const double g0 = -xk[q] / (1.0 + en);
            gk += yy[q] * g1 + (1.0 - yy[q]) * g0;
            if (HESS) {
                double s0 = -1.0 / ((1.0 + en) * (1.0 + en));
                s0 *= en;
                double s1 = -1.0 / ((1.0 + ep) * (1.0 + ep));
                s1 *= ep;
                const double sv = yy[q] * s1 + (1.0 - yy[q]) * s0;
#pragma unroll
                for (int l = 0; l < 32; ++l) {
                    const double xl = __shfl_sync(kFullMask, xk[q], l);
                    h[l] += sv * (xk[q] * xl);
                }
            }
        }
    }
    // block partial: the warps add their sums into shared memory one after the other (fixed order)
    for (int k = threadIdx.x; k < stride; k += blockDim.x) s_acc[k] = 0.0;
    __syncthreads();
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) {
        if (warp == wv) {
            if (lane == 0) s_acc[0] += ll;
            if (lane < d) s_acc[1 + lane] += gk;
            if (HESS && lane < d) {
#pragma unroll
                for (int l = 0; l < 32; ++l)
                    if (l < d) s_acc[1 + d + lane * d + l] += -h[l];
            }
        }
        __syncthreads();
    }
    double* out = partial + (size_t)blockIdx.x * stride;
    for (int k = threadIdx.x; k < stride; k += blockDim.x) out[k] = s_acc[k];
}

// Log-likelihood + gradient only (the call the QN sampler makes): a warp takes 32 sampled rows at
// a time.  Lane k owns feature k for the loads and the dot products (32 independent row gathers
// in flight), then every lane does the scalar exp / log work of ONE row, so the transcendental
// work is shared out over the lanes instead of being repeated by all 32.
__global__ void __launch_bounds__(256) logistic_grad_kernel(
    const double* __restrict__ x, const double* __restrict__ y, const int* __restrict__ idx, int m,
    int d, long long row_begin, long long row_end, const double* __restrict__ beta,
    double* __restrict__ partial, int stride) {
    extern __shared__ double s_acc[];   // [stride] block accumulator
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const double bk = (lane < d) ? beta[lane] : 0.0;
    double ll = 0.0, gk = 0.0;
    for (int r0 = gwarp * 32; r0 < m; r0 += nwarps * 32) {
        const int r = r0 + lane;
        long long row = -1;
        if (r < m) row = idx[r];
        const bool ok = (row >= row_begin && row < row_end);
        const long long lrow = ok ? (row - row_begin) : -1;
        const double yy = ok ? y[lrow] : 0.0;
        double xk[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            const long long rq = __shfl_sync(kFullMask, lrow, q);
            xk[q] = (rq >= 0 && lane < d) ? x[(size_t)rq * d + lane] : 0.0;
        }
        double xb = 0.0;
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            const double pq = warp_sum(bk * xk[q]);
            if (lane == q) xb = pq;
        }
        double coef = 0.0;
        if (ok) {
            const double en = exp(-1.0 * xb), ep = exp(xb);
            const double eta = 1.0 / (1.0 + en);
            double e1 = log(eta), e0 = log(1.0 - eta);
            if (isinf(e1)) e1 = 0.0;
            if (isinf(e0)) e0 = 0.0;
            ll += yy * e1 + (1.0 - yy) * e0;
            coef = yy / (1.0 + ep) - (1.0 - yy) / (1.0 + en);
        }
#pragma unroll
        for (int q = 0; q < 32; ++q) gk += __shfl_sync(kFullMask, coef, q) * xk[q];
    }
    ll = warp_sum(ll);   // one row per lane: the warp's log-likelihood is the sum over its lanes
    for (int k = threadIdx.x; k < stride; k += blockDim.x) s_acc[k] = 0.0;
    __syncthreads();
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) {
        if (warp == wv) {
            if (lane == 0) s_acc[0] += ll;
            if (lane < d) s_acc[1 + lane] += gk;
        }
        __syncthreads();
    }
    double* out = partial + (size_t)blockIdx.x * stride;
    for (int k = threadIdx.x; k < stride; k += blockDim.x) out[k] = s_acc[k];
}

// one warp per output: lanes stride over the block partials, then a fixed tree
__global__ void logistic_reduce_kernel(const double* __restrict__ partial, int nparts, int stride,
                                       int nout_valid, int nout, double* __restrict__ out) {
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= nout) return;
    double s = 0.0;
    if (k < nout_valid)
        for (int w = lane; w < nparts; w += 32) s += partial[(size_t)w * stride + k];
    s = warp_sum(s);
    if (lane == 0) out[k] = s;
}

static int logistic_grid(int hess) { return 148 * (hess ? 2 : 4); }

size_t logistic_ws_bytes(int m, int d, int hess) {
    (void)m;
    const int stride = 1 + d + (hess ? d * d : 0);
    return (size_t)logistic_grid(hess) * stride * sizeof(double);
}

cudaError_t launch_logistic(const double* x, const double* y, const int* idx, int m, int d,
                            long long row_begin, long long row_end, const double* beta, int hess,
                            double* out, void* ws, cudaStream_t st) {
    const int grid = logistic_grid(hess);
    const int stride = 1 + d + (hess ? d * d : 0);
    const size_t smem = (size_t)stride * sizeof(double);
    double* partial = (double*)ws;
    if (hess)
        logistic_partial_kernel<true><<<grid, 256, smem, st>>>(x, y, idx, m, d, row_begin, row_end, beta,
                                                               partial, stride);
    else
        logistic_grad_kernel<<<grid, 256, smem, st>>>(x, y, idx, m, d, row_begin, row_end, beta, partial,
                                                      stride);
    const int nout = 1 + d + d * d;
    logistic_reduce_kernel<<<(nout * 32 + 127) / 128, 128, 0, st>>>(partial, grid, stride, stride, nout, out);
    return cudaGetLastError();
}

}  // namespace pmmh
