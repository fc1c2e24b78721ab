// sv_filter.cu -- persistent cooperative kernel for the stochastic-volatility particle filter
// with sorted, correlated systematic resampling and the fixed-lag smoother that feeds the
// quasi-Newton proposal.
//
// Replaces (reference paths relative to /root/reference/python):
//   flps_sv_corr      state/particle_methods/stochastic_volatility.pyx:205-655
//   bpf_sv_corr       state/particle_methods/stochastic_volatility.pyx:61-201
//   systematic_corr   ...:694-715     my_max ...:738-746     norm_logpdf ...:659-664
//   argsort/qsort     ...:23-52
//
// Design (B200-first, not a port):
//   * one launch = one whole evaluation (all T steps) for every problem of the batch; a TEAM
//     of G co-resident CTAs owns one problem, teams loop over the batch.  Cross-CTA steps use
//     team_allgather() (common.cuh) -- 7 per time step -- instead of kernel boundaries.
//   * particles, weights, ancestors are fp64/int32 structure-of-arrays in global memory
//     (L2-resident at N <= 2^20); u is streamed time-major [t][j], read exactly once.
//   * the per-step sort is a one-pass monotone bucket split: bins are uniform over a range
//     predicted from the sorted parents (exact range of the propagation mean over
//     [x_min, x_max] +- 6.5 sd, outliers clamp into the end bins), atomics give each key a
//     slot inside its bin, a scan gives bin offsets, and a tiny all-pairs rank orders each bin.
//     Any correct sort reproduces the reference's qsort order for distinct keys.
//   * resampling = team-wide scan of the shifted weights in POSITION order (deterministic) +
//     per-child binary search in a shared-memory window of the cumulative weights, which are
//     normalised by their total at lookup.
//   * the genealogy is O(T N): composed one-step ancestors A_t in a ring of depth LAG+1; the
//     fixed-lag pair (x_{t-L+1}, x_{t-L+2}) of each particle is found by chasing A.
//   * quirks Q1, Q3-Q8, Q11 of the reference are reproduced (see oracle/pmmh_oracle.c).
//
// fp64 arithmetic follows the reference's operation order; sums over particles are tree sums
// in a fixed order (the reference sums sequentially), hence tolerance-level, deterministic
// agreement.  Compiled with -fmad=false.
#include "sv_filter.cuh"

#include <math.h>

#include "common.cuh"
#include "sv_math.cuh"

namespace pmmh {

namespace {

constexpr int kMaxLag = 64;
static_assert(kMaxAllgatherHost == kMaxAllgather, "all-gather width mismatch");

// lower_bound of cp in the normalised cumulative weights cum[m] / sum_w restricted to [l, h];
// returns h if every entry is below cp (the reference's `cur < N-1` clamp when h = N-1).
__device__ __forceinline__ int search_cum_global(const double* cum, double sum_w, double cp, int l,
                                                 int h) {
    while (l < h) {
        int m = (l + h) >> 1;
        const double v = cum[m] / sum_w;
        if (v < cp) l = m + 1;
        else h = m;
    }
    return l;
}

// Block-wide inclusive scan step used only by the (rare) bpf trajectory draw.
__device__ __forceinline__ double block_incl_scan(double v, double* s_w /*[32]*/, double* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const double incl = warp_incl_scan(v, lane);
    __syncthreads();
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    double base = 0.0;
    for (int k = 0; k < warp; ++k) base = base + s_w[k];
    double tot = 0.0;
    for (int k = 0; k < nwarp; ++k) tot = tot + s_w[k];
    *total = tot;
    return base + incl;
}

template <bool HESS>
__global__ void __launch_bounds__(kSvThreads, 1) sv_pf_kernel(SvArgs a) {
    extern __shared__ double dsm[];
    double* s_gather = dsm;                                      // [G * kMaxAllgather]
    double* s_stage = dsm + (size_t)a.G * kMaxAllgather;         // [kStageDoubles]
    __shared__ double s_vals[kMaxAllgather];
    __shared__ double s_tot[kMaxAllgather];
    __shared__ double s_red[kMaxAllgather * 32];
    __shared__ double s_wtot[32], s_wbase[32];
    __shared__ int s_iwtot[32], s_iwbase[32];
    __shared__ int s_lohi[2];
    __shared__ double s_bin[2];
    __shared__ double s_S[kMaxLag];      // sum of shifted weights of the last kMaxLag steps
    __shared__ double s_hacc[20];        // running hessian1 / hessian2 (upper triangles)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarp = blockDim.x >> 5;
    const int N = a.N, NOBS = a.NOBS, LAG = a.LAG, NB = a.NB, G = a.G;
    const int SQ = a.SQ, SQW = a.SQW;
    const bool flps = (a.mode == kSvFlps);

    Team tm;
    tm.G = G;
    tm.rank = blockIdx.x % G;
    tm.epoch = 0;
    const int team_id = blockIdx.x / G;
    {
        unsigned* all_stamps = (unsigned*)a.ws;
        double* all_slots = (double*)(a.ws + sv_align((size_t)gridDim.x * sizeof(unsigned)));
        tm.stamps = all_stamps + (size_t)team_id * G;
        tm.slots = all_slots + (size_t)team_id * 2 * G * kMaxAllgather;
    }
    char* wsbase = a.ws + a.ws_sync_bytes + (size_t)team_id * a.ws_team_stride;
    const bool lead = (tm.rank == 0);

    Tile tp, tb;
    tp.init(N, G, tm.rank);
    tb.init(NB, G, tm.rank);

    constexpr int NACC = HESS ? 26 : 6;   // filt, smo, grad[4], (hess1[10], hess2[10])

    for (int prob = team_id; prob < a.B; prob += a.n_teams) {
        // fallback pass: only the problems the exchange kernel (sv_fast.cu) abandoned
        if (a.only_failed && a.diag[(size_t)prob * kDiagCount + kDiagStatus] != 1) continue;
        const double* obs = a.obs + (size_t)prob * a.obs_stride;
        const double* par = a.params + (size_t)prob * 4;
        const double* rvr = a.rvr + (size_t)prob * NOBS;
        const double* U = a.U + (size_t)prob * NOBS * N;
        double* o_filt = a.filt + (size_t)prob * NOBS;
        double* o_smo = flps ? a.smo + (size_t)prob * NOBS : nullptr;
        double* o_grad = flps ? a.grad + (size_t)prob * 4 * NOBS : nullptr;
        double* o_traj = a.traj + (size_t)prob * NOBS;
        long long* o_diag = a.diag + (size_t)prob * kDiagCount;

        SvWs w;
        sv_ws_layout(N, NOBS, LAG, NB, a.RING, a.hess, a.mode, SQ, SQW, a.Xhist != nullptr, wsbase, &w);
        double* Xh;
        int* Ah;
        int RING;
        if (a.Xhist) {
            Xh = a.Xhist + (size_t)prob * NOBS * N;
            Ah = a.Ahist + (size_t)prob * NOBS * N;
            RING = NOBS;
        } else {
            Xh = w.Xring;
            Ah = w.Aring;
            RING = a.RING;
        }
        const int RR = a.RING;   // ring depth of R (never the full history)
#define XT(t) (Xh + (size_t)((t) % RING) * N)
#define AT(t) (Ah + (size_t)((t) % RING) * N)
#define RT(t, c) (w.Rring + ((size_t)((t) % RR) * 4 + (c)) * N)

        SvConst c;
        c.mu = par[0];
        c.phi = par[1];
        c.sigmav = par[2];
        c.rho = par[3];
        c.rho_term = 1.0 - c.rho * c.rho;
        c.q = 1.0 / (c.sigmav * c.sigmav * (1.0 - c.rho * c.rho));
        c.sd = sqrt(c.rho_term) * c.sigmav;
        c.one_m_phi = 1.0 - c.phi;
        c.one_m_phi2 = 1.0 - c.phi * c.phi;
        c.inv_sv = 1.0 / c.sigmav;
        c.inv_sv2 = 1.0 / (c.sigmav * c.sigmav);
        c.sr = c.sigmav * c.rho;
        const double logN = log((double)N);

        // ---------------- time 0
        // flps (stochastic_volatility.pyx:306-323, Q1): every particle equals mu + stDev*0.0
        // bpf  (:110-122): x0_j = mu + stDev * rvp[0 + j*NOBS], then sorted (handled as step i = 0)
        const double stdev0 = c.sigmav / sqrt(1.0 - (c.phi * c.phi));
        const double x0 = c.mu + stdev0 * 0.0;
        for (int b = tb.p0 + tid; b < tb.p1; b += blockDim.x) w.hist[b] = 0;
        if (flps) {
            for (int j = tp.p0 + tid; j < tp.p1; j += blockDim.x) {
                XT(0)[j] = x0;
                AT(0)[j] = j;
                w.sh0[j] = 1.0;
                if (j < SQ) w.Xlow[j] = x0;
                if (HESS) {
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) RT(0, cc)[j] = 0.0;
                }
            }
            if (lead) {
                for (int t = tid; t < NOBS; t += blockDim.x) {
                    o_smo[t] = 0.0;
                    o_grad[t] = 0.0;
                    o_grad[NOBS + t] = 0.0;
                    o_grad[2 * NOBS + t] = 0.0;
                    o_grad[3 * NOBS + t] = 0.0;
                }
                if (tid == 0) o_traj[0] = x0;
            }
        }
        if (tid < 20) s_hacc[tid] = 0.0;
        double loglike = 0.0;
        double shift_prev = 0.0;
        long long near_ties = 0, key_ties2 = 0;   // key_ties2 counts each tie pair twice
        int max_occ_seen = 0;
        int wavefront_max = 0;
        int status = 0;
        team_barrier(tm, s_gather);

        for (int i = flps ? 1 : 0; i <= NOBS; ++i) {
            const bool init_step = (i == 0);
            const int t = i - 1;
            const double* shp = (t & 1) ? w.sh1 : w.sh0;
            double* shn = (i & 1) ? w.sh1 : w.sh0;
            const double* Xp = init_step ? nullptr : XT(t);
            double S_t = 1.0;

            if (!init_step) {
                // =========== phase A: weights of time t (position order => deterministic sums):
                //   S_t = sum sh, filter mean, fixed-lag smoother terms, then the cumulative sums.
                //   All sums are accumulated un-normalised and divided by S_t afterwards.
                double acc[NACC];
#pragma unroll
                for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
                const int sb = tp.seg_begin(warp), se = tp.seg_end(warp);
                {
                    double carry = 0.0;
                    for (int base = sb; base < se; base += 32) {
                        const int j = base + lane;
                        double sj = 0.0;
                        if (j < se) {
                            sj = shp[j];
                            if (!isfinite(sj)) sj = 0.0;
                            const double wx = sj * Xp[j];
                            if (isfinite(wx)) acc[0] += wx;
                        }
                        const double incl = warp_incl_scan(sj, lane);
                        carry = carry + __shfl_sync(kFullMask, incl, 31);
                    }
                    if (lane == 0) s_wtot[warp] = carry;
                }
                if (flps && t >= LAG) {
                    // fixed-lag smoother, stochastic_volatility.pyx:445-534 (Q5: obs[t - LAG])
                    const double yl = obs[t - LAG];
                    for (int j = tp.p0 + tid; j < tp.p1; j += blockDim.x) {
                        double sj = shp[j];
                        if (!isfinite(sj)) sj = 0.0;
                        int b = j;
                        for (int k = 0; k < LAG - 2; ++k) b = AT(t - k)[b];
                        const double next = XT(t - LAG + 2)[b];
                        const int bc = AT(t - LAG + 2)[b];
                        const double curr = XT(t - LAG + 1)[bc];
                        double sq, g[4];
                        sv_score_main(c, curr, next, yl, sq, g);
                        acc[1] += sj * curr;
                        acc[2] += g[0] * sj;
                        acc[3] += g[1] * sj;
                        acc[4] += g[2] * sj;
                        acc[5] += g[3] * sj;
                        if (HESS) {
                            double al[4];
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) al[cc] = RT(t - LAG + 2, cc)[b];
                            sv_hessian_terms(c, curr, sq, yl, g, al, sj, &acc[6]);
                        }
                    }
                }
                block_sum<NACC>(acc, s_red);
                if (warp == 0) {
                    const double v = (lane < nwarp) ? s_wtot[lane] : 0.0;
                    const double incl = warp_incl_scan(v, lane);
                    const double excl = __shfl_up_sync(kFullMask, incl, 1);
                    s_wbase[lane] = (lane == 0) ? 0.0 : excl;
                    if (lane == 31) s_vals[0] = incl;
                }
                if (tid < NACC) s_vals[1 + tid] = acc[tid];
                team_allgather(tm, s_vals, 1 + NACC, s_gather);
                for (int k = warp; k < 1 + NACC; k += nwarp) {
                    const double s = gathered_sum(s_gather, 1 + NACC, k, G, lane);
                    if (lane == 0) s_tot[k] = s;
                }
                if (warp == nwarp - 1) {
                    const double s = gathered_sum(s_gather, 1 + NACC, 0, tm.rank, lane);
                    if (lane == 0) s_tot[kMaxAllgather - 1] = s;
                }
                __syncthreads();
                S_t = s_tot[0];
                const double P = s_tot[kMaxAllgather - 1];
                if (t >= 1) loglike += shift_prev + log(S_t) - logN;   // :537 / :186
                if (lead && tid == 0) {
                    o_filt[t] = s_tot[1] / S_t;
                    if (flps && t >= LAG) {
                        const int tt = t - LAG + 1;
                        o_smo[tt] = s_tot[2] / S_t;
                        o_grad[tt] = s_tot[3] / S_t;
                        o_grad[NOBS + tt] = s_tot[4] / S_t;
                        o_grad[2 * NOBS + tt] = s_tot[5] / S_t;
                        o_grad[3 * NOBS + tt] = s_tot[6] / S_t;
                    }
                }
                if (tid == 0) s_S[t % kMaxLag] = S_t;
                if (HESS && tid < 20 && flps && t >= LAG) s_hacc[tid] += s_tot[7 + tid] / S_t;
                if (!flps) {
                    // Q10: keep the normalised weights of the low slots of every time step
                    for (int j = tp.p0 + tid; j < tp.p1 && j < SQW; j += blockDim.x) {
                        double sj = shp[j];
                        if (!isfinite(sj)) sj = 0.0;
                        w.Wlow[(size_t)t * SQW + j] = sj / S_t;
                    }
                }
                if (i == NOBS) break;

                // pass 2: cumulative (un-normalised) weights; normalised by S_t at lookup
                {
                    const double wb = s_wbase[warp];
                    double carry = 0.0;
                    for (int base = sb; base < se; base += 32) {
                        const int j = base + lane;
                        double sj = (j < se) ? shp[j] : 0.0;
                        if (!isfinite(sj)) sj = 0.0;
                        const double incl = warp_incl_scan(sj, lane);
                        if (j < se) w.cum[j] = P + ((wb + carry) + incl);
                        carry = carry + __shfl_sync(kFullMask, incl, 31);
                    }
                }
                team_barrier(tm, s_gather);
            }

            // =========== phase B: resample (systematic_corr :694-715), propagate (:354-358 /
            //             :142-146), bin histogram
            const double y1 = init_step ? 0.0 : obs[i - 1];
            const double yi = obs[i];
            double tmin = INFINITY, tmax = -INFINITY;
            if (flps && tid == 0) {
                // bin range predicted from the sorted parents: exact range of the propagation
                // mean over [x_min, x_max], widened by 6.5 sd (outliers clamp into end bins)
                const double xmin = Xp[0], xmax = Xp[N - 1];
                const double cc = c.sr * y1;
                auto f = [&](double x) { return (c.mu + c.phi * (x - c.mu)) + cc * exp(-0.5 * x); };
                const double fa = f(xmin), fb = f(xmax);
                double fmn = fmin(fa, fb), fmx = fmax(fa, fb);
                if (c.phi * cc > 0.0) {
                    const double xs = -2.0 * log(2.0 * c.phi / cc);
                    if (xs > xmin && xs < xmax) {
                        const double fs = f(xs);
                        fmn = fmin(fmn, fs);
                        fmx = fmax(fmx, fs);
                    }
                }
                const double lo = fmn - 6.5 * c.sd, hi = fmx + 6.5 * c.sd;
                const double width = hi - lo;
                double scale = (double)NB / width;
                if (!(width > 0.0) || !isfinite(scale) || !isfinite(lo)) scale = 0.0;
                s_bin[0] = isfinite(lo) ? lo : 0.0;
                s_bin[1] = scale;
            }
            if (init_step) {
                // bpf time 0: unsorted initial cloud
                for (int j = tp.p0 + tid; j < tp.p1; j += blockDim.x) {
                    const double xn = c.mu + stdev0 * ld_stream_f64(&U[j]);
                    w.xnew[j] = xn;
                    w.aun[j] = j;
                    tmin = fmin(tmin, xn);
                    tmax = fmax(tmax, xn);
                }
            } else {
                const double u = rvr[i];
                if (tp.p1 > tp.p0) {
                    if (tid == 0)
                        s_lohi[0] = search_cum_global(w.cum, S_t, (u + (double)tp.p0) / (double)N, 0, N - 1);
                    if (tid == 32)
                        s_lohi[1] =
                            search_cum_global(w.cum, S_t, (u + (double)(tp.p1 - 1)) / (double)N, 0, N - 1);
                }
                __syncthreads();
                if (tp.p1 > tp.p0) {
                    const int wlo = s_lohi[0], whi = s_lohi[1];
                    const int wlen = whi - wlo + 1;
                    const bool staged = (wlen <= kStageDoubles);
                    if (staged) {
                        for (int k = tid; k < wlen; k += blockDim.x) s_stage[k] = w.cum[wlo + k] / S_t;
                        __syncthreads();
                    }
                    const double bin_lo = s_bin[0], bin_scale = s_bin[1];
                    const double* Ui = U + (size_t)i * N;
                    for (int j = tp.p0 + tid; j < tp.p1; j += blockDim.x) {
                        const double cp = (u + (double)j) / (double)N;
                        int aj;
                        double cv_hi, cv_lo;
                        if (staged) {
                            int l = 0, h = wlen - 1;
                            while (l < h) {
                                const int m = (l + h) >> 1;
                                if (s_stage[m] < cp) l = m + 1;
                                else h = m;
                            }
                            aj = wlo + l;
                            cv_hi = s_stage[l];
                            cv_lo = (l > 0) ? s_stage[l - 1] : -1.0;
                        } else {
                            aj = search_cum_global(w.cum, S_t, cp, wlo, whi);
                            cv_hi = w.cum[aj] / S_t;
                            cv_lo = (aj > 0) ? w.cum[aj - 1] / S_t : -1.0;
                        }
                        {   // diagnostics: decisions within 64 ulp of a cumulative-weight tie
                            const double tol = 64.0 * 2.220446049250313e-16 * cp;
                            if (fabs(cv_hi - cp) <= tol || (cv_lo >= 0.0 && fabs(cp - cv_lo) <= tol))
                                near_ties++;
                        }
                        const double xp = Xp[aj];
                        double mean = c.mu + c.phi * (xp - c.mu);
                        w.aun[j] = aj;
                        if (flps || a.mode == kSvBpfIntended) {
                            mean += c.sr * exp(-0.5 * xp) * y1;
                            const double xn = mean + c.sd * ld_stream_f64(&Ui[j]);
                            w.xnew[j] = xn;
                            tmin = fmin(tmin, xn);
                            tmax = fmax(tmax, xn);
                            if (flps) {
                                const int b = sv_bin(xn, bin_lo, bin_scale, NB);
                                w.rnk[j] = atomicAdd(&w.hist[b], 1);
                            }
                        } else {
                            // bpf parity mode (Q2): the leverage term reads x_new[a_j] of THIS step
                            // if a_j < j (else 0.0).  Resolve the dependency chains in rounds.
                            w.tkey[j] = mean;              // mean without the leverage term
                            if (aj >= j) {
                                mean += c.sr * exp(-0.5 * 0.0) * y1;
                                const double xn = mean + c.sd * ld_stream_f64(&Ui[j]);
                                w.xnew[j] = xn;
                                w.rnk[j] = 0;              // resolved in round 0
                                tmin = fmin(tmin, xn);
                                tmax = fmax(tmax, xn);
                            } else {
                                w.rnk[j] = 0x7fffffff;     // unresolved
                            }
                        }
                    }
                }
                if (a.mode == kSvBpfParity) {
                    const double* Ui = U + (size_t)i * N;
                    for (int round = 1;; ++round) {
                        team_barrier(tm, s_gather);
                        double rem[1] = {0.0};
                        for (int j = tp.p0 + tid; j < tp.p1; j += blockDim.x) {
                            if (w.rnk[j] != 0x7fffffff) continue;
                            const int aj = w.aun[j];
                            if (w.rnk[aj] < round) {
                                double mean = w.tkey[j];
                                mean += c.sr * exp(-0.5 * w.xnew[aj]) * y1;
                                const double xn = mean + c.sd * ld_stream_f64(&Ui[j]);
                                w.xnew[j] = xn;
                                w.rnk[j] = round;
                                tmin = fmin(tmin, xn);
                                tmax = fmax(tmax, xn);
                            } else {
                                rem[0] += 1.0;
                            }
                        }
                        block_sum<1>(rem, s_red);
                        if (tid == 0) s_vals[0] = rem[0];
                        team_allgather(tm, s_vals, 1, s_gather);
                        double left = 0.0;
                        for (int cc = 0; cc < G; ++cc) left += s_gather[cc];
                        __syncthreads();
                        wavefront_max = max(wavefront_max, round);
                        if (left == 0.0) break;
                    }
                }
            }
            {   // tile min / max of the new keys
                tmin = warp_min(tmin);
                tmax = warp_max(tmax);
                if (lane == 0) {
                    s_red[warp] = tmin;
                    s_red[32 + warp] = tmax;
                }
                __syncthreads();
                if (warp == 0) {
                    double v = (lane < nwarp) ? s_red[lane] : INFINITY;
                    double v2 = (lane < nwarp) ? s_red[32 + lane] : -INFINITY;
                    v = warp_min(v);
                    v2 = warp_max(v2);
                    if (lane == 0) {
                        s_vals[0] = v;
                        s_vals[1] = v2;
                    }
                }
            }
            team_allgather(tm, s_vals, 2, s_gather);
            if (warp == 0) {
                const double v = gathered_min(s_gather, 2, 0, G, lane);
                const double v2 = gathered_max(s_gather, 2, 1, G, lane);
                if (lane == 0) {
                    s_tot[0] = v;
                    s_tot[3] = v2;
                }
            }
            __syncthreads();
            const double kmin = s_tot[0];
            if (!flps) {
                // bootstrap filter: data-driven bin range [min, max], then the histogram pass
                const double kmax = s_tot[3];
                const double width = kmax - kmin;
                double scale = (double)NB / width;
                if (!(width > 0.0) || !isfinite(scale) || !isfinite(kmin)) scale = 0.0;
                if (tid == 0) {
                    s_bin[0] = isfinite(kmin) ? kmin : 0.0;
                    s_bin[1] = scale;
                }
                __syncthreads();
                for (int j = tp.p0 + tid; j < tp.p1; j += blockDim.x) {
                    const int b = sv_bin(w.xnew[j], s_bin[0], s_bin[1], NB);
                    w.rnk[j] = atomicAdd(&w.hist[b], 1);
                }
                team_barrier(tm, s_gather);
            }
            const double bin_lo = s_bin[0], bin_scale = s_bin[1];

            // =========== phase C: scan of the bin histogram -> bin offsets
            const int bsb = tb.seg_begin(warp), bse = tb.seg_end(warp);
            {
                int occ = 0;
                int carry = 0;
                for (int base = bsb; base < bse; base += 32) {
                    const int b = base + lane;
                    const int v = (b < bse) ? w.hist[b] : 0;
                    occ = max(occ, v);
                    const int incl = warp_incl_scan(v, lane);
                    carry += __shfl_sync(kFullMask, incl, 31);
                }
                occ = warp_max(occ);
                if (lane == 0) {
                    s_iwtot[warp] = carry;
                    s_red[32 + warp] = (double)occ;
                }
            }
            __syncthreads();
            if (warp == 0) {
                const int v = (lane < nwarp) ? s_iwtot[lane] : 0;
                const int incl = warp_incl_scan(v, lane);
                const int excl = __shfl_up_sync(kFullMask, incl, 1);
                s_iwbase[lane] = (lane == 0) ? 0 : excl;
                double o = (lane < nwarp) ? s_red[32 + lane] : 0.0;
                o = warp_max(o);
                if (lane == 31) s_vals[0] = (double)incl;
                if (lane == 0) s_vals[1] = o;
            }
            team_allgather(tm, s_vals, 2, s_gather);
            if (warp == 0) {
                const double s = gathered_sum(s_gather, 2, 0, tm.rank, lane);
                const double m = gathered_max(s_gather, 2, 1, G, lane);
                if (lane == 0) {
                    s_tot[1] = s;
                    s_tot[2] = m;
                }
            }
            __syncthreads();
            {
                const int boff = (int)s_tot[1];
                const int mo = (int)s_tot[2];
                max_occ_seen = max(max_occ_seen, mo);
                if (mo > kBinCap) {
                    status = 1;   // degenerate particle cloud: give up on this problem (uniform)
                    break;
                }
                const int wb = s_iwbase[warp];
                int carry = 0;
                for (int base = bsb; base < bse; base += 32) {
                    const int b = base + lane;
                    const int v = (b < bse) ? w.hist[b] : 0;
                    const int incl = warp_incl_scan(v, lane);
                    if (b < bse) {
                        w.binstart[b] = boff + wb + carry + (incl - v);
                        w.hist[b] = 0;
                    }
                    carry += __shfl_sync(kFullMask, incl, 31);
                }
                if (tm.rank == G - 1 && tid == 0) w.binstart[NB] = N;
            }
            team_barrier(tm, s_gather);

            // =========== phase D: scatter into bins; Q4 shift (last log-weight above lw[0])
            const double lw0 = init_step ? 0.0 : sv_logw(kmin, yi);
            double pmax = -INFINITY;
            for (int j = tp.p0 + tid; j < tp.p1; j += blockDim.x) {
                const double key = w.xnew[j];
                const int b = sv_bin(key, bin_lo, bin_scale, NB);
                const int slot = w.binstart[b] + w.rnk[j];
                w.tkey[slot] = key;
                w.tpay[slot] = w.aun[j];
                if (HESS) w.tidx[slot] = j;
                double lw = 0.0;
                if (!init_step) {
                    lw = sv_logw(key, yi);
                    if (lw > lw0 && isfinite(lw)) pmax = fmax(pmax, key);
                }
                w.cum[slot] = lw;   // cum is free until the next phase A: log-weight scratch
            }
            {
                pmax = warp_max(pmax);
                if (lane == 0) s_red[warp] = pmax;
                __syncthreads();
                if (warp == 0) {
                    double v = (lane < nwarp) ? s_red[lane] : -INFINITY;
                    v = warp_max(v);
                    if (lane == 0) s_vals[0] = v;
                }
            }
            team_allgather(tm, s_vals, 1, s_gather);
            if (warp == 0) {
                const double v = gathered_max(s_gather, 1, 0, G, lane);
                if (lane == 0) s_tot[0] = v;
            }
            __syncthreads();
            const double kq = s_tot[0];
            const double shift = init_step ? 0.0 : ((kq > -INFINITY) ? sv_logw(kq, yi) : lw0);

            // =========== phase E: order each bin (all-pairs rank), write the sorted generation,
            //             weights (:427-437), alpha recursion (:361-390)
            const int* rootp = (t & 1) ? w.root1 : w.root0;
            int* rootn = (i & 1) ? w.root1 : w.root0;
            for (int s = tp.p0 + tid; s < tp.p1; s += blockDim.x) {
                const double key = w.tkey[s];
                const int pay = w.tpay[s];
                int oj = 0;
                if (HESS) oj = w.tidx[s];
                const int b = sv_bin(key, bin_lo, bin_scale, NB);
                const int start = w.binstart[b], end = w.binstart[b + 1];
                int rank = 0;
                for (int q = start; q < end; ++q) {
                    if (q == s) continue;
                    const double k2 = w.tkey[q];
                    if (k2 < key) rank++;
                    else if (k2 == key) {
                        key_ties2++;
                        bool less;
                        if (HESS) less = w.tidx[q] < oj;
                        else {
                            const int p2 = w.tpay[q];
                            less = (p2 < pay) || (p2 == pay && q < s);
                        }
                        if (less) rank++;
                    }
                }
                const int p = start + rank;
                XT(i)[p] = key;
                AT(i)[p] = pay;
                if (p < SQ) w.Xlow[(size_t)i * SQ + p] = key;
                if (p == 0 && i >= 1) o_traj[i] = key;   // Q11: traj[i] = X_i[0]
                double sh = 1.0;
                if (!init_step) {
                    sh = exp(w.cum[s] - shift);
                    if (!flps && !isfinite(sh)) sh = 0.0;   // :173-176
                }
                shn[p] = sh;
                if (flps && i >= NOBS - LAG) w.shtail[(size_t)(i - (NOBS - LAG)) * N + p] = sh;
                if (!flps) {
                    if (init_step) {
                        w.X0[p] = key;
                        rootn[p] = p;
                    } else {
                        rootn[p] = rootp[pay];
                    }
                }
                if (HESS) {
                    // Q7: particles[i - 1 + ancestors[j]] read through the flat layout
                    const long long q = (long long)i - 1 + pay;
                    const int tq = (int)(q % NOBS), sl = (int)(q / NOBS);
                    double curr;
                    if (tq < i) curr = w.Xlow[(size_t)tq * SQ + sl];
                    else if (tq == i) curr = (sl <= oj) ? w.xnew[sl] : 0.0;
                    else curr = 0.0;
                    const double ylag = obs_wrap(obs, i - LAG, NOBS);   // Q8
                    double sq = key - c.mu - c.phi * (curr - c.mu);
                    const double e = exp(-0.5 * curr);
                    sq -= c.sr * e * ylag;
                    const double a0 = c.q * sq * c.one_m_phi;
                    const double a1 = c.q * sq * (curr - c.mu) * c.one_m_phi2;
                    double a2 = sq;
                    a2 += c.sr * e * yi;
                    a2 *= c.q * sq;
                    a2 -= 1.0;
                    double a3 = c.rho - c.q * c.rho * sq * sq;
                    a3 += c.inv_sv * sq * e * yi;
                    RT(i, 0)[p] = a0 + RT(t, 0)[pay];
                    RT(i, 1)[p] = a1 + RT(t, 1)[pay];
                    RT(i, 2)[p] = a2 + RT(t, 2)[pay];
                    RT(i, 3)[p] = a3 + RT(t, 3)[pay];
                }
            }
            shift_prev = shift;
            team_barrier(tm, s_gather);
        }   // time loop
        team_barrier(tm, s_gather);

        // =========== tail (stochastic_volatility.pyx:540-626, Q6)
        if (flps && status == 0) {
            const int T = NOBS - 1;
            const double* shT = (T & 1) ? w.sh1 : w.sh0;
            const double S_T = s_S[T % kMaxLag];
            for (int k = 0; k < LAG; ++k) {
                const int ip = T - k;   // reference loop variable i
                constexpr int NT = HESS ? 25 : 5;
                double tacc[NT];
#pragma unroll
                for (int q = 0; q < NT; ++q) tacc[q] = 0.0;
                const double S_ip = s_S[ip % kMaxLag];
                const double* sh_ip = w.shtail + (size_t)(ip - (NOBS - LAG)) * N;
                const double y1 = obs_wrap(obs, ip - 1, NOBS);
                const double ylag = obs_wrap(obs, ip - LAG, NOBS);
                for (int j = tp.p0 + tid; j < tp.p1; j += blockDim.x) {
                    int b = j;
                    int bprev = j;     // index at lag k-1
                    for (int h = 0; h < k; ++h) {
                        bprev = b;
                        b = AT(T - h)[b];
                    }
                    const double curr = XT(ip)[b];
                    double sT = shT[j];
                    if (!isfinite(sT)) sT = 0.0;
                    tacc[0] += (sT / S_T) * curr;
                    if (k >= 1) {
                        const double next = XT(ip + 1)[bprev];
                        double sq, g[4];
                        sv_score_tail(c, curr, next, y1, sq, g);
                        double si = sh_ip[j];
                        if (!isfinite(si)) si = 0.0;
                        const double wi = si / S_ip;
                        tacc[1] += g[0] * wi;
                        tacc[2] += g[1] * wi;
                        tacc[3] += g[2] * wi;
                        tacc[4] += g[3] * wi;
                        if (HESS) {
                            int b_l2 = j;   // index at lag LAG-2 (for alpha)
                            for (int h = 0; h < LAG - 2; ++h) b_l2 = AT(T - h)[b_l2];
                            double al[4];
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) al[cc] = RT(T - LAG + 2, cc)[b_l2];
                            sv_hessian_terms(c, curr, sq, ylag, g, al, wi, &tacc[5]);
                        }
                    }
                }
                block_sum<NT>(tacc, s_red);
                if (tid < NT) s_vals[tid] = tacc[tid];
                team_allgather(tm, s_vals, NT, s_gather);
                for (int q = warp; q < NT; q += nwarp) {
                    const double s = gathered_sum(s_gather, NT, q, G, lane);
                    if (lane == 0) s_tot[q] = s;
                }
                __syncthreads();
                if (lead && tid == 0) {
                    o_smo[ip] += s_tot[0];
                    if (k >= 1) {
                        const int tt = ip - LAG + 1;
                        if (tt >= 0) {
                            o_grad[tt] += s_tot[1];
                            o_grad[NOBS + tt] += s_tot[2];
                            o_grad[2 * NOBS + tt] += s_tot[3];
                            o_grad[3 * NOBS + tt] += s_tot[4];
                        }
                    }
                }
                if (HESS && tid < 20 && k >= 1) s_hacc[tid] += s_tot[5 + tid];
                __syncthreads();
            }
        }

        // =========== bpf trajectory (:189-192): Q10 flat weights + Q11
        int traj_idx = 0;
        if (!flps && status == 0 && lead) {
            // cumulative sum over W_flat[k] = w_{k % NOBS}[k / NOBS], k < N (sampleParticle_corr)
            const double rnd = rvr[0];
            double total = 0.0;
            {
                double part[1] = {0.0};
                for (int k = tid; k < N; k += blockDim.x)
                    part[0] += w.Wlow[(size_t)(k % NOBS) * SQW + (k / NOBS)];
                block_sum<1>(part, s_red);
                total = part[0];
            }
            double carry = 0.0;
            int found = N;
            for (int base = 0; base < N; base += blockDim.x) {
                const int k = base + tid;
                const double v = (k < N) ? w.Wlow[(size_t)(k % NOBS) * SQW + (k / NOBS)] : 0.0;
                double chunk_total;
                const double incl = carry + block_incl_scan(v, s_wtot, &chunk_total);
                carry = carry + chunk_total;
                if (k < N) {
                    const double cn = (k == 0) ? incl : incl / total;   // Q3
                    if (!(cn < rnd)) found = min(found, k);
                }
                __syncthreads();
            }
            found = -warp_max(-found);
            if (lane == 0) s_iwtot[warp] = found;
            __syncthreads();
            if (tid == 0) {
                int f = N;
                for (int k = 0; k < nwarp; ++k) f = min(f, s_iwtot[k]);
                if (f >= N) f = N - 1;   // the reference would read out of bounds; out of contract
                const int* rootT = ((NOBS - 1) & 1) ? w.root1 : w.root0;
                o_traj[0] = w.X0[rootT[f]];
                s_iwbase[0] = f;
            }
            __syncthreads();
            traj_idx = s_iwbase[0];
        }

        // =========== outputs
        if (lead) {
            if (tid == 0) {
                a.loglike[prob] = (status == 0) ? loglike : NAN;
                o_diag[kDiagMaxBin] = max_occ_seen;
                o_diag[kDiagStatus] = status;
                o_diag[kDiagWavefront] = wavefront_max;
                o_diag[kDiagTrajIdx] = traj_idx;
                o_diag[kDiagKernel] = 1;
            }
            if (flps && tid < 16) {
                // expand the upper triangles into the symmetric 4x4 outputs
                const int r = tid >> 2, cidx = tid & 3;
                const int k = min(r, cidx), l = max(r, cidx);
                const int tri = k * 4 - (k * (k - 1)) / 2 + (l - k);
                a.hess1[(size_t)prob * 16 + tid] = HESS ? s_hacc[tri] : 0.0;
                a.hess2[(size_t)prob * 16 + tid] = HESS ? s_hacc[10 + tri] : 0.0;
            }
        }
        {   // tie counters: every CTA contributes
            double nt[2] = {(double)near_ties, (double)key_ties2};
            block_sum<2>(nt, s_red);
            if (tid < 2) s_vals[tid] = nt[tid];
            team_allgather(tm, s_vals, 2, s_gather);
            if (warp < 2) {
                const double s = gathered_sum(s_gather, 2, warp, G, lane);
                if (lane == 0 && lead) o_diag[warp == 0 ? kDiagNearTies : kDiagKeyTies] =
                    (long long)(warp == 0 ? s : s * 0.5);
            }
            __syncthreads();
        }
#undef XT
#undef AT
#undef RT
    }   // problem loop
}

}  // namespace

int sv_dynamic_smem_bytes(int G) {
    return (int)(((size_t)G * kMaxAllgather + kStageDoubles) * sizeof(double));
}

cudaError_t sv_launch(const SvArgs& a, int grid, cudaStream_t stream) {
    const int smem = sv_dynamic_smem_bytes(a.G);
    void* kargs[] = {(void*)&a};
    cudaError_t err;
    if (a.hess) {
        err = cudaFuncSetAttribute(sv_pf_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (err != cudaSuccess) return err;
        return cudaLaunchCooperativeKernel((void*)sv_pf_kernel<true>, dim3(grid), dim3(kSvThreads), kargs,
                                           smem, stream);
    }
    err = cudaFuncSetAttribute(sv_pf_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return err;
    return cudaLaunchCooperativeKernel((void*)sv_pf_kernel<false>, dim3(grid), dim3(kSvThreads), kargs,
                                       smem, stream);
}

}  // namespace pmmh
