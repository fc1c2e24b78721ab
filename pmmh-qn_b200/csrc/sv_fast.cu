// sv_fast.cu -- fast path of the SV fixed-lag particle smoother (log-likelihood + gradient).
//
// Same estimator as sv_filter.cu (flps_sv_corr, stochastic_volatility.pyx:205-655, hess = 0)
// re-organised around what the ncu profile of the first kernel showed (profiles/): the step was
// bound by barrier waits (7 team syncs + single-thread sections) and by dependent L2 gathers
// with one load in flight per thread.  Here one time step is TWO phases separated by TWO team
// all-gathers:
//
//   phase BG  (after the weights of time t are complete)
//        * resample: each CTA locates the window of cumulative weights its children need from
//          the all-gathered tile totals plus one parallel probe, stages it in shared memory and
//          every child does a branch-free binary search there              (:694-715)
//        * propagate (:354-358) using the parent's stored exp(-x/2); new exp(-x'/2) and
//          log-weight (:428) are computed once and travel with the particle
//        * coarse bucket split: one global atomic + one 32-byte record write per child
//        * fixed-lag smoother terms of time t (:445-470): ancestor chase with 4 independent
//          chains per thread, all reads are 32-byte records (one sector each)
//   all-gather #2: max log-weight (the shift), smoother partial sums
//   phase E   * every CTA sorts the buckets that cover ITS tile of output positions entirely in
//               shared memory (fine bins + all-pairs rank inside a bin), then writes the sorted
//               generation as coalesced 32-byte records {x, exp(-x/2), shifted weight, ancestor},
//               the tile-local cumulative weights, and its totals (position order =>
//               deterministic sums)
//   all-gather #1: tile totals of the weights (=> S_t, tile offsets), filter-mean partial sums
//
// Differences to the reference that stay inside the stated tolerances: sums over particles are
// fixed-order tree sums; the weight shift is the true maximum log-weight (the reference's
// my_max, Q4, returns another element; the shift cancels analytically); log N(y; 0, e^{x/2}) is
// evaluated as -0.9189.. - x/2 - y^2 e^{-x} / 2 with e^{-x} = (e^{-x/2})^2.
// A bucket that receives more than kChunk particles (a degenerate cloud) abandons the
// evaluation with status 1; callers fall back to the general kernel (sv_filter.cu).
#include <math.h>

#include "common.cuh"
#include "sv_filter.cuh"
#include "sv_math.cuh"

namespace pmmh {

namespace {

constexpr int kThreads = kSvThreads;
constexpr int kChunk = kFastChunk;       // records sorted per shared-memory pass == bucket capacity
constexpr int kFineBins = 4096;
constexpr int kGatherK = 8;              // doubles per CTA per all-gather in this kernel
constexpr int kMaxLagF = 64;
constexpr int kIlp = 4;

struct __align__(32) PRec {   // one sorted particle of one time step
    double x, e, sh;
    int a, pad;
};
struct __align__(32) BRec {   // one propagated child waiting in its bucket
    double x, e, lw;
    int a, j;
};
static_assert(sizeof(PRec) == 32 && sizeof(BRec) == 32, "records must be one 32-byte sector");

struct FastWs {
    int* bcount;      // [2][NBK]
    double* cumloc;   // [N] tile-local inclusive cumulative shifted weights
    PRec* P;          // [RING][N]
    BRec* BK;         // [NBK][kChunk]
};

__device__ __forceinline__ size_t fast_ws_carve(int N, int NBK, int RING, char* base, FastWs* w) {
    size_t off = 0;
    w->bcount = (int*)(base + off);
    off += sv_align((size_t)2 * NBK * sizeof(int));
    w->cumloc = (double*)(base + off);
    off += sv_align((size_t)N * sizeof(double));
    w->P = (PRec*)(base + off);
    off += sv_align((size_t)RING * N * sizeof(PRec));
    w->BK = (BRec*)(base + off);
    off += sv_align((size_t)NBK * kChunk * sizeof(BRec));
    return off;
}

__device__ __forceinline__ int bucket_of(double x, double lo, double scale, int NBK) {
    const double t = (x - lo) * scale;
    if (!(t >= 0.0)) return 0;
    if (t >= (double)NBK) return NBK - 1;
    return (int)t;
}

// strict weak order of the sort: key, then birth index (unique)
__device__ __forceinline__ bool rec_less(double xa, int ja, double xb, int jb) {
    return (xa < xb) || (xa == xb && ja < jb);
}

__global__ void __launch_bounds__(kThreads, 1) sv_fast_kernel(SvArgs a) {
    extern __shared__ __align__(32) unsigned char dsm_raw[];
    const int N = a.N, NOBS = a.NOBS, LAG = a.LAG, G = a.G, NBK = a.NBK, RING = a.RING;
    // dynamic shared memory carve-up
    double* s_gather = (double*)dsm_raw;                                   // [G * kGatherK]
    int* s_off = (int*)(s_gather + (size_t)G * kGatherK);                  // [NBK + 1]
    unsigned char* s_union = (unsigned char*)(s_off + ((NBK + 1 + 7) & ~7));
    //   view 1 (phase BG): staged cumulative-weight window
    double* s_stage = (double*)s_union;                                    // [kStageDoubles]
    //   view 2 (phase E): chunk sort
    BRec* s_rec = (BRec*)s_union;                                          // [kChunk]
    int* s_fh = (int*)(s_rec + kChunk);                                    // [kFineBins + 1]
    int* s_slot = s_fh + kFineBins + 8;                                    // [kChunk]
    int* s_inv = s_slot + kChunk;                                          // [kChunk]

    __shared__ double s_vals[kGatherK];
    __shared__ double s_tot[kGatherK + 2];
    __shared__ double s_red[8 * 32];
    __shared__ double s_tileP[160];          // prefix of the tile totals (G + 1 entries)
    __shared__ double s_w[32], s_wx[32];
    __shared__ int s_iw[32], s_iwx[32];
    __shared__ int s_misc[8];
    __shared__ double s_S[kMaxLagF];
    __shared__ double s_dmisc[4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarp = kThreads / 32;

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
    const int per_tile = (N + G - 1) / G;
    const int p0 = min(N, tm.rank * per_tile), p1 = min(N, p0 + per_tile);

    for (int prob = team_id; prob < a.B; prob += a.n_teams) {
        const double* obs = a.obs + (size_t)prob * a.obs_stride;
        const double* rvr = a.rvr + (size_t)prob * NOBS;
        const double* U = a.U + (size_t)prob * NOBS * N;
        double* o_filt = a.filt + (size_t)prob * NOBS;
        double* o_smo = a.smo + (size_t)prob * NOBS;
        double* o_grad = a.grad + (size_t)prob * 4 * NOBS;
        double* o_traj = a.traj + (size_t)prob * NOBS;
        long long* o_diag = a.diag + (size_t)prob * kDiagCount;
        double* Xh = a.Xhist ? a.Xhist + (size_t)prob * NOBS * N : nullptr;
        int* Ah = a.Ahist ? a.Ahist + (size_t)prob * NOBS * N : nullptr;

        FastWs w;
        fast_ws_carve(N, NBK, RING, wsbase, &w);
#define PT(t) (w.P + (size_t)((t) % RING) * N)

        SvConst c;
        sv_const_init(c, a.params + (size_t)prob * 4);
        const double logN = log((double)N);
        const double invN_exact = 1.0 / (double)N;
        const bool n_pow2 = (N & (N - 1)) == 0;   // then (u + j) / N == (u + j) * (1 / N) exactly

        // ---------------- time 0 (:306-323, Q1): every particle equals mu + stDev * 0.0
        const double stdev0 = c.sigmav / sqrt(1.0 - (c.phi * c.phi));
        const double x0 = c.mu + stdev0 * 0.0;
        const double e0 = exp(-0.5 * x0);
        for (int b = tm.rank * kThreads + tid; b < 2 * NBK; b += G * kThreads) w.bcount[b] = 0;
        for (int j = p0 + tid; j < p1; j += kThreads) {
            PRec r;
            r.x = x0;
            r.e = e0;
            r.sh = 1.0;
            r.a = j;
            r.pad = 0;
            PT(0)[j] = r;
            w.cumloc[j] = (double)(j - p0 + 1);
            if (Xh) {
                Xh[j] = x0;
                Ah[j] = j;
            }
        }
        if (lead) {
            for (int t = tid; t < NOBS; t += kThreads) {
                o_smo[t] = 0.0;
                o_grad[t] = 0.0;
                o_grad[NOBS + t] = 0.0;
                o_grad[2 * NOBS + t] = 0.0;
                o_grad[3 * NOBS + t] = 0.0;
            }
            if (tid == 0) o_traj[0] = x0;
        }
        double loglike = 0.0, shift = 0.0;
        long long near_ties = 0, key_ties2 = 0;
        int max_bucket = 0, status = 0;
        if (tid == 0) {
            s_vals[0] = (double)(p1 - p0);
            s_vals[1] = (double)(p1 - p0) * x0;
        }
        team_allgather(tm, s_vals, 2, s_gather);   // all-gather #1 of time 0

        for (int i = 1; i <= NOBS; ++i) {
            const int t = i - 1;
            const PRec* Pt = PT(t);
            // ------------- after all-gather #1: S_t, tile offsets, filter mean of time t
            for (int cc = warp; cc <= G; cc += nwarp) {
                const double s = gathered_sum(s_gather, 2, 0, cc, lane);
                if (lane == 0) s_tileP[cc] = s;
            }
            if (warp == nwarp - 1) {
                const double s = gathered_sum(s_gather, 2, 1, G, lane);
                if (lane == 0) s_tot[1] = s;
            }
            __syncthreads();
            const double S_t = s_tileP[G];
            if (t >= 1) loglike += shift + log(S_t) - logN;   // :537
            if (tid == 0) s_S[t % kMaxLagF] = S_t;
            if (lead && tid == 0) o_filt[t] = s_tot[1] / S_t;

            double acc[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) acc[k] = 0.0;
            acc[0] = -INFINITY;   // max log-weight of the children

            // ------------- phase BG, part 1: fixed-lag smoother terms of time t (:445-470)
            if (t >= LAG) {
                const double yl = obs[t - LAG];   // Q5
                const PRec* Pn = PT(t - LAG + 2);
                const PRec* Pc = PT(t - LAG + 1);
                for (int base = p0 + tid; base < p1; base += kIlp * kThreads) {
                    int b[kIlp];
                    double sj[kIlp];
#pragma unroll
                    for (int m = 0; m < kIlp; ++m) {
                        const int j = base + m * kThreads;
                        const bool ok = j < p1;
                        b[m] = ok ? j : p0;
                        sj[m] = ok ? Pt[b[m]].sh : 0.0;
                    }
                    for (int k = 0; k < LAG - 2; ++k) {
                        const PRec* Pk = PT(t - k);
#pragma unroll
                        for (int m = 0; m < kIlp; ++m) b[m] = Pk[b[m]].a;
                    }
                    double nx[kIlp];
                    int bc[kIlp];
#pragma unroll
                    for (int m = 0; m < kIlp; ++m) {
                        nx[m] = Pn[b[m]].x;
                        bc[m] = Pn[b[m]].a;
                    }
#pragma unroll
                    for (int m = 0; m < kIlp; ++m) {
                        const double2 xe = *reinterpret_cast<const double2*>(&Pc[bc[m]]);
                        double sq, g[4];
                        sv_score_main_e(c, xe.x, xe.y, nx[m], yl, sq, g);
                        acc[1] += sj[m] * xe.x;
                        acc[2] += g[0] * sj[m];
                        acc[3] += g[1] * sj[m];
                        acc[4] += g[2] * sj[m];
                        acc[5] += g[3] * sj[m];
                    }
                }
            }

            // ------------- phase BG, part 2: resample + propagate + bucket split (step i)
            double bk_lo = 0.0, bk_scale = 0.0;
            if (i < NOBS) {
                const double y1 = obs[i - 1], yi = obs[i];
                const double u = rvr[i];
                int* cnt = w.bcount + (size_t)(i & 1) * NBK;
                {
                    double lo, hi;
                    sv_child_range(c, Pt[0].x, Pt[N - 1].x, y1, 5.0, lo, hi);
                    const double width = hi - lo;
                    double scale = (double)NBK / width;
                    if (!(width > 0.0) || !isfinite(scale) || !isfinite(lo)) scale = 0.0;
                    bk_lo = isfinite(lo) ? lo : 0.0;
                    bk_scale = scale;
                }
                if (p1 > p0) {
                    const double cp_first = n_pow2 ? (u + (double)p0) * invN_exact : (u + (double)p0) / (double)N;
                    const double cp_last = n_pow2 ? (u + (double)(p1 - 1)) * invN_exact
                                                  : (u + (double)(p1 - 1)) / (double)N;
                    // tile-level bracket from the all-gathered totals
                    if (warp == 0) {
                        int c_lo = G - 1, c_hi = G - 1;
                        for (int cc = lane; cc < G; cc += 32) {
                            const double endc = s_tileP[cc + 1] / S_t;
                            if (endc >= cp_first) c_lo = min(c_lo, cc);
                            if (endc >= cp_last) c_hi = min(c_hi, cc);
                        }
                        c_lo = -warp_max(-c_lo);
                        c_hi = -warp_max(-c_hi);
                        if (lane == 0) {
                            s_misc[0] = min(N - 1, c_lo * per_tile);
                            s_misc[1] = min(N - 1, (c_hi + 1) * per_tile - 1);
                        }
                    }
                    __syncthreads();
                    int wlo = s_misc[0], whi = s_misc[1];
                    // one parallel probe tightens the window to ~1/1024 of the bracket
                    {
                        const int len = whi - wlo + 1;
                        const int stride = (len + kThreads - 1) / kThreads;
                        const int q = wlo + tid * stride;
                        int cand_lo = wlo, cand_hi = whi;
                        if (q <= whi) {
                            const double v = (s_tileP[q / per_tile] + w.cumloc[q]) / S_t;
                            if (v < cp_first) cand_lo = q;
                            if (v >= cp_last) cand_hi = q;
                        }
                        cand_lo = warp_max(cand_lo);
                        cand_hi = -warp_max(-cand_hi);
                        if (lane == 0) {
                            s_iw[warp] = cand_lo;
                            s_iwx[warp] = cand_hi;
                        }
                        __syncthreads();
                        if (warp == 0) {
                            int l2 = s_iw[lane], h2 = s_iwx[lane];
                            l2 = warp_max(l2);
                            h2 = -warp_max(-h2);
                            if (lane == 0) {
                                s_misc[2] = l2;
                                s_misc[3] = h2;
                            }
                        }
                        __syncthreads();
                        wlo = s_misc[2];
                        whi = s_misc[3];
                    }
                    const int wlen = whi - wlo + 1;
                    const bool staged = (wlen <= kStageDoubles);
                    if (staged) {
                        for (int k = tid; k < wlen; k += kThreads) {
                            const int m = wlo + k;
                            s_stage[k] = (s_tileP[m / per_tile] + w.cumloc[m]) / S_t;
                        }
                    }
                    __syncthreads();
                    int nsteps = 0;
                    while ((1 << nsteps) < wlen) ++nsteps;
                    const double* Ui = U + (size_t)i * N;
                    const double half_y2 = 0.5 * (yi * yi);
                    for (int base = p0 + tid; base < p1; base += kIlp * kThreads) {
                        int aj[kIlp];
                        double un[kIlp], cp[kIlp];
                        bool ok[kIlp];
#pragma unroll
                        for (int m = 0; m < kIlp; ++m) {
                            const int j = base + m * kThreads;
                            ok[m] = j < p1;
                            const int jj = ok[m] ? j : p0;
                            un[m] = ld_stream_f64(&Ui[jj]);
                            cp[m] = n_pow2 ? (u + (double)jj) * invN_exact : (u + (double)jj) / (double)N;
                        }
                        if (staged) {
                            // branch-free lower bound: first l with s_stage[l] >= cp (clamped)
                            int l[kIlp];
#pragma unroll
                            for (int m = 0; m < kIlp; ++m) l[m] = 0;
                            for (int s = nsteps - 1; s >= 0; --s) {
#pragma unroll
                                for (int m = 0; m < kIlp; ++m) {
                                    const int mid = l[m] + (1 << s);
                                    if (mid <= wlen - 1 && s_stage[mid - 1] < cp[m]) l[m] = mid;
                                }
                            }
#pragma unroll
                            for (int m = 0; m < kIlp; ++m) {
                                aj[m] = wlo + l[m];
                                const double cv_hi = s_stage[l[m]];
                                const double cv_lo = (l[m] > 0) ? s_stage[l[m] - 1] : -1.0;
                                const double tol = 64.0 * 2.220446049250313e-16 * cp[m];
                                if (ok[m] && (fabs(cv_hi - cp[m]) <= tol ||
                                              (cv_lo >= 0.0 && fabs(cp[m] - cv_lo) <= tol)))
                                    near_ties++;
                            }
                        } else {
#pragma unroll
                            for (int m = 0; m < kIlp; ++m) {
                                int lo2 = wlo, hi2 = whi;
                                while (lo2 < hi2) {
                                    const int mid = (lo2 + hi2) >> 1;
                                    const double v = (s_tileP[mid / per_tile] + w.cumloc[mid]) / S_t;
                                    if (v < cp[m]) lo2 = mid + 1;
                                    else hi2 = mid;
                                }
                                aj[m] = lo2;
                            }
                        }
                        double2 pxe[kIlp];
#pragma unroll
                        for (int m = 0; m < kIlp; ++m)
                            pxe[m] = *reinterpret_cast<const double2*>(&Pt[aj[m]]);
                        BRec r[kIlp];
                        int cb[kIlp], slot[kIlp];
#pragma unroll
                        for (int m = 0; m < kIlp; ++m) {
                            const double xp = pxe[m].x;
                            double mean = c.mu + c.phi * (xp - c.mu);
                            mean += c.sr * pxe[m].y * y1;            // pxe.y == exp(-0.5 * xp)
                            const double xn = mean + c.sd * un[m];
                            const double en = exp(-0.5 * xn);
                            r[m].x = xn;
                            r[m].e = en;
                            r[m].lw = -0.91893853320467267 - 0.5 * xn - half_y2 * (en * en);
                            r[m].a = aj[m];
                            r[m].j = base + m * kThreads;
                            cb[m] = bucket_of(xn, bk_lo, bk_scale, NBK);
                        }
#pragma unroll
                        for (int m = 0; m < kIlp; ++m) slot[m] = ok[m] ? atomicAdd(&cnt[cb[m]], 1) : kChunk;
#pragma unroll
                        for (int m = 0; m < kIlp; ++m) {
                            if (ok[m]) {
                                acc[0] = fmax(acc[0], r[m].lw);
                                if (slot[m] < kChunk) w.BK[(size_t)cb[m] * kChunk + slot[m]] = r[m];
                            }
                        }
                    }
                }
            }

            // ------------- all-gather #2: shift (max log-weight) and smoother sums
            {
                double mx = warp_max(acc[0]);
                if (lane == 0) s_red[7 * 32 + warp] = mx;
                double sums[5] = {acc[1], acc[2], acc[3], acc[4], acc[5]};
                block_sum<5>(sums, s_red);
                if (warp == 0) {
                    double v = s_red[7 * 32 + lane];
                    v = warp_max(v);
                    if (lane == 0) s_vals[0] = v;
                }
                if (tid < 5) s_vals[1 + tid] = sums[tid];
            }
            team_allgather(tm, s_vals, 6, s_gather);
            if (warp == 0) {
                const double v = gathered_max(s_gather, 6, 0, G, lane);
                if (lane == 0) s_tot[0] = v;
            } else if (warp <= 5) {
                const double s = gathered_sum(s_gather, 6, warp, G, lane);
                if (lane == 0) s_tot[warp] = s;
            }
            __syncthreads();
            if (lead && tid == 0 && t >= LAG) {
                const int tt = t - LAG + 1;
                o_smo[tt] = s_tot[1] / S_t;
                o_grad[tt] = s_tot[2] / S_t;
                o_grad[NOBS + tt] = s_tot[3] / S_t;
                o_grad[2 * NOBS + tt] = s_tot[4] / S_t;
                o_grad[3 * NOBS + tt] = s_tot[5] / S_t;
            }
            if (i == NOBS) break;
            shift = s_tot[0];

            // ------------- phase E: sort the buckets covering my output tile in shared memory
            const int* cnt = w.bcount + (size_t)(i & 1) * NBK;
            int* cnt_next = w.bcount + (size_t)((i + 1) & 1) * NBK;
            for (int b = tm.rank * kThreads + tid; b < NBK; b += G * kThreads) cnt_next[b] = 0;
            {   // bucket offsets: exclusive scan of the counts (4 consecutive buckets per thread)
                int v[4], tsum = 0, mx = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int b = tid * 4 + q;
                    v[q] = (b < NBK) ? __ldcg(&cnt[b]) : 0;
                    mx = max(mx, v[q]);
                    tsum += v[q];
                }
                const int incl = warp_incl_scan(tsum, lane);
                mx = warp_max(mx);
                if (lane == 31) s_iw[warp] = incl;
                if (lane == 0) s_iwx[warp] = mx;
                __syncthreads();
                if (warp == 0) {
                    const int tv = s_iw[lane];
                    const int ti = warp_incl_scan(tv, lane);
                    int m2 = s_iwx[lane];
                    m2 = warp_max(m2);
                    __syncwarp();
                    s_iw[lane] = ti - tv;
                    if (lane == 0) s_misc[4] = m2;
                }
                __syncthreads();
                int run = s_iw[warp] + (incl - tsum);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int b = tid * 4 + q;
                    if (b < NBK) s_off[b] = run;
                    run += v[q];
                }
                if (tid == 0) s_off[NBK] = N;
                __syncthreads();
            }
            max_bucket = max(max_bucket, s_misc[4]);
            if (s_misc[4] > kChunk) {
                status = 1;   // degenerate cloud: a bucket overflowed (uniform decision)
                break;
            }
            PRec* Pi = PT(i);
            double carry = 0.0;      // running tile-local cumulative weight
            double fx = 0.0;         // thread-local part of sum sh * x
            if (p1 > p0) {
                // first / last bucket overlapping [p0, p1)
                int cbA, cb_last;
                {
                    int lo2 = 0, hi2 = NBK - 1;
                    while (lo2 < hi2) {   // largest b with s_off[b] <= p0
                        const int mid = (lo2 + hi2 + 1) >> 1;
                        if (s_off[mid] <= p0) lo2 = mid;
                        else hi2 = mid - 1;
                    }
                    cbA = lo2;
                    lo2 = 0;
                    hi2 = NBK - 1;
                    while (lo2 < hi2) {
                        const int mid = (lo2 + hi2 + 1) >> 1;
                        if (s_off[mid] <= p1 - 1) lo2 = mid;
                        else hi2 = mid - 1;
                    }
                    cb_last = lo2;
                }
                while (cbA <= cb_last) {
                    int cbB = cbA;
                    while (cbB + 1 <= cb_last && s_off[cbB + 2] - s_off[cbA] <= kChunk) ++cbB;
                    const int cbase = s_off[cbA];
                    const int Lc = s_off[cbB + 1] - cbase;
                    // (a) load the chunk's records, find its key range
                    double kmn = INFINITY, kmx = -INFINITY;
                    for (int k = tid; k < Lc; k += kThreads) {
                        const int gp = cbase + k;
                        int lo2 = cbA, hi2 = cbB;
                        while (lo2 < hi2) {   // bucket of position gp
                            const int mid = (lo2 + hi2 + 1) >> 1;
                            if (s_off[mid] <= gp) lo2 = mid;
                            else hi2 = mid - 1;
                        }
                        const BRec r = w.BK[(size_t)lo2 * kChunk + (gp - s_off[lo2])];
                        s_rec[k] = r;
                        kmn = fmin(kmn, r.x);
                        kmx = fmax(kmx, r.x);
                    }
                    for (int k = tid; k <= kFineBins; k += kThreads) s_fh[k] = 0;
                    kmn = warp_min(kmn);
                    kmx = warp_max(kmx);
                    if (lane == 0) {
                        s_w[warp] = kmn;
                        s_wx[warp] = kmx;
                    }
                    __syncthreads();
                    if (warp == 0) {
                        double v1 = s_w[lane], v2 = s_wx[lane];
                        v1 = warp_min(v1);
                        v2 = warp_max(v2);
                        if (lane == 0) {
                            s_dmisc[0] = v1;
                            s_dmisc[1] = v2;
                        }
                    }
                    __syncthreads();
                    const double fmin_k = s_dmisc[0];
                    double fscale = (double)kFineBins / (s_dmisc[1] - fmin_k);
                    if (!(s_dmisc[1] > fmin_k) || !isfinite(fscale)) fscale = 0.0;
                    // (c) fine histogram
                    int fb[3], rf[3];
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const int k = tid + q * kThreads;
                        fb[q] = 0;
                        rf[q] = 0;
                        if (k < Lc) {
                            fb[q] = bucket_of(s_rec[k].x, fmin_k, fscale, kFineBins);
                            rf[q] = atomicAdd(&s_fh[fb[q]], 1);
                        }
                    }
                    __syncthreads();
                    {   // (d) exclusive scan of the fine histogram, in place (4 bins per thread)
                        int v[4], tsum = 0;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            v[q] = s_fh[tid * 4 + q];
                            tsum += v[q];
                        }
                        const int incl = warp_incl_scan(tsum, lane);
                        if (lane == 31) s_iw[warp] = incl;
                        __syncthreads();
                        if (warp == 0) {
                            const int tv = s_iw[lane];
                            const int ti = warp_incl_scan(tv, lane);
                            __syncwarp();
                            s_iw[lane] = ti - tv;
                        }
                        __syncthreads();
                        int run = s_iw[warp] + (incl - tsum);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            s_fh[tid * 4 + q] = run;
                            run += v[q];
                        }
                        if (tid == 0) s_fh[kFineBins] = Lc;
                    }
                    __syncthreads();
                    // (e) group the records by fine bin
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const int k = tid + q * kThreads;
                        if (k < Lc) s_slot[s_fh[fb[q]] + rf[q]] = k;
                    }
                    __syncthreads();
                    // (f) order inside each fine bin (all pairs; ~1 element per bin on average)
                    for (int s = tid; s < Lc; s += kThreads) {
                        const int k = s_slot[s];
                        const double key = s_rec[k].x;
                        const int kj = s_rec[k].j;
                        const int f = bucket_of(key, fmin_k, fscale, kFineBins);
                        const int st = s_fh[f], en = s_fh[f + 1];
                        int rank = 0;
                        for (int q = st; q < en; ++q) {
                            if (q == s) continue;
                            const int k2 = s_slot[q];
                            const double x2 = s_rec[k2].x;
                            if (x2 == key) key_ties2++;
                            if (rec_less(x2, s_rec[k2].j, key, kj)) rank++;
                        }
                        s_inv[st + rank] = k;
                    }
                    __syncthreads();
                    // (g) write my part of the sorted generation: records, cumulative weights
                    for (int rowb = 0; rowb < Lc; rowb += kThreads) {
                        const int pos = rowb + tid;
                        const int gp = cbase + pos;
                        double shv = 0.0;
                        const bool mine = (pos < Lc) && gp >= p0 && gp < p1;
                        if (mine) {
                            const BRec r = s_rec[s_inv[pos]];
                            shv = exp(r.lw - shift);
                            if (!isfinite(shv)) shv = 0.0;
                            PRec o;
                            o.x = r.x;
                            o.e = r.e;
                            o.sh = shv;
                            o.a = r.a;
                            o.pad = 0;
                            Pi[gp] = o;
                            const double sx = shv * r.x;
                            if (isfinite(sx)) fx += sx;
                            if (gp == 0) o_traj[i] = r.x;   // Q11: traj[i] = X_i[0]
                            if (Xh) {
                                Xh[(size_t)i * N + gp] = r.x;
                                Ah[(size_t)i * N + gp] = r.a;
                            }
                        }
                        // block-wide inclusive scan of this row of weights (fixed order)
                        const double incl = warp_incl_scan(shv, lane);
                        if (lane == 31) s_w[warp] = incl;
                        __syncthreads();
                        if (warp == 0) {
                            const double tv = s_w[lane];
                            const double ti = warp_incl_scan(tv, lane);
                            const double te = __shfl_up_sync(kFullMask, ti, 1);
                            s_wx[lane] = (lane == 0) ? 0.0 : te;
                            if (lane == 31) s_dmisc[2] = ti;
                        }
                        __syncthreads();
                        if (mine) w.cumloc[gp] = carry + (s_wx[warp] + incl);
                        carry = carry + s_dmisc[2];
                        __syncthreads();
                    }
                    cbA = cbB + 1;
                }
            }
            {
                double v[1] = {fx};
                block_sum<1>(v, s_red);
                if (tid == 0) {
                    s_vals[0] = carry;
                    s_vals[1] = v[0];
                }
            }
            team_allgather(tm, s_vals, 2, s_gather);   // all-gather #1 of time i
        }   // time loop

        // ---------------- tail (:540-562, Q6)
        if (status == 0) {
            const int T = NOBS - 1;
            const PRec* PTT = PT(T);
            const double S_T = s_S[T % kMaxLagF];
            for (int k = 0; k < LAG; ++k) {
                const int ip = T - k;
                double tacc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
                const double S_ip = s_S[ip % kMaxLagF];
                const PRec* Pip = PT(ip);
                const PRec* Pip1 = PT(ip + 1);
                const double y1 = obs_wrap(obs, ip - 1, NOBS);
                for (int j = p0 + tid; j < p1; j += kThreads) {
                    int b = j, bprev = j;
                    for (int h = 0; h < k; ++h) {
                        bprev = b;
                        b = PT(T - h)[b].a;
                    }
                    const double2 xe = *reinterpret_cast<const double2*>(&Pip[b]);
                    double sT = PTT[j].sh;
                    if (!isfinite(sT)) sT = 0.0;
                    tacc[0] += (sT / S_T) * xe.x;
                    if (k >= 1) {
                        const double next = Pip1[bprev].x;
                        double sq, g[4];
                        sv_score_tail_e(c, xe.x, xe.y, next, y1, sq, g);
                        double si = Pip[j].sh;
                        if (!isfinite(si)) si = 0.0;
                        const double wi = si / S_ip;
                        tacc[1] += g[0] * wi;
                        tacc[2] += g[1] * wi;
                        tacc[3] += g[2] * wi;
                        tacc[4] += g[3] * wi;
                    }
                }
                block_sum<5>(tacc, s_red);
                if (tid < 5) s_vals[tid] = tacc[tid];
                team_allgather(tm, s_vals, 5, s_gather);
                if (warp < 5) {
                    const double s = gathered_sum(s_gather, 5, warp, G, lane);
                    if (lane == 0) s_tot[warp] = s;
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
                __syncthreads();
            }
        }

        // ---------------- outputs
        if (lead) {
            if (tid == 0) {
                a.loglike[prob] = (status == 0) ? loglike : NAN;
                o_diag[kDiagMaxBin] = max_bucket;
                o_diag[kDiagStatus] = status;
                o_diag[kDiagWavefront] = 0;
                o_diag[kDiagTrajIdx] = 0;
            }
            if (tid < 16) {
                a.hess1[(size_t)prob * 16 + tid] = 0.0;
                a.hess2[(size_t)prob * 16 + tid] = 0.0;
            }
        }
        {
            double nt[2] = {(double)near_ties, (double)key_ties2};
            block_sum<2>(nt, s_red);
            if (tid < 2) s_vals[tid] = nt[tid];
            team_allgather(tm, s_vals, 2, s_gather);
            if (warp < 2) {
                const double s = gathered_sum(s_gather, 2, warp, G, lane);
                if (lane == 0 && lead)
                    o_diag[warp == 0 ? kDiagNearTies : kDiagKeyTies] = (long long)(warp == 0 ? s : s * 0.5);
            }
            __syncthreads();
        }
#undef PT
    }   // problem loop
}

}  // namespace

size_t sv_fast_ws_bytes(int N, int NBK, int RING) {
    size_t off = 0;
    off += sv_align((size_t)2 * NBK * sizeof(int));
    off += sv_align((size_t)N * sizeof(double));
    off += sv_align((size_t)RING * N * sizeof(PRec));
    off += sv_align((size_t)NBK * kChunk * sizeof(BRec));
    return off;
}

int sv_fast_smem_bytes(int G, int NBK) {
    size_t b = (size_t)G * kGatherK * sizeof(double);
    b += (size_t)((NBK + 1 + 7) & ~7) * sizeof(int);
    const size_t view1 = (size_t)kStageDoubles * sizeof(double);
    const size_t view2 = (size_t)kChunk * sizeof(BRec) + (size_t)(kFineBins + 8) * sizeof(int) +
                         (size_t)2 * kChunk * sizeof(int);
    b += (view1 > view2 ? view1 : view2);
    return (int)b + 32;
}

cudaError_t sv_fast_launch(const SvArgs& a, int grid, cudaStream_t stream) {
    const int smem = sv_fast_smem_bytes(a.G, a.NBK);
    cudaError_t err = cudaFuncSetAttribute(sv_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return err;
    void* kargs[] = {(void*)&a};
    return cudaLaunchCooperativeKernel((void*)sv_fast_kernel, dim3(grid), dim3(kThreads), kargs, smem, stream);
}

}  // namespace pmmh
