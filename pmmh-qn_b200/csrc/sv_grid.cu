// sv_grid.cu -- the "grid kernel": ONE stochastic-volatility fixed-lag smoother evaluation
// (log-likelihood + gradient) as ONE persistent cooperative launch over all SMs.
//
// Restates flps_sv_corr (/root/reference/python/state/particle_methods/stochastic_volatility.pyx
// :205-655: correlated systematic resampling :694-715, propagation :354-358, argsort :392-424 /
// :23-52, weights :427-442, fixed-lag score terms :445-470, tail :540-562, log-likelihood :537,
// trajectory :630-633 with quirks Q10/Q11).  B200 design, 148 CTAs x 1024 threads, one CTA per SM:
//
//   * the sorted generation is cut into G TILES of ~N/G particles (7 085 at N = 2^20, G = 148);
//     CTA c owns tile c while it is sorted and weighted (everything in shared memory) and owns
//     the CHILDREN [c*Wc, (c+1)*Wc) while they are generated (equal work for every CTA whatever
//     the weights look like);
//   * four grid barriers per time step (one atomic counter, arrive / wait split so that work that
//     only feeds outputs sits between the two):
//       C  owner of a tile: cumulative weights -> PARENT-side child ranges in closed form
//          ub(p) = #{j : (u + j)/N <= cum(p)} (exact predicate re-checked), head markers H[first
//          child] = parent                                                         | barrier 4
//       A  owner of a child range: max-scan of the head markers = ancestor of every child,
//          propagation, 8192-bin value histogram (shared-memory atomics, merged into a global
//          one)                                                                    | barrier 1
//          scan of the global histogram -> tile boundaries on bin edges (every tile gets N/G
//          particles +- one bin), slot reservation per (CTA, tile), scatter of 16-byte entries
//          (value, birth row, lagged ancestor row) into the tile's mailbox        | barrier 2
//       B  owner of a tile: counting sort over 8192 sub-bins + exact in-bin ranking by (value,
//          birth row) = the reference's argsort; weights, block scan, moments     | barrier 3
//          (after the arrive: fixed-lag score terms, copy-out)
//   * genealogy: a generation is stored ONCE in birth order as P[t][j] = (value, parent value,
//     exp(-parent value / 2)) and R[t][j] = birth rows of the ancestors 1..8 steps back (one
//     32-byte sector each).  The fixed-lag terms of step t need one random sector of P[t-lag+2];
//     a child copies its parent's R with one random sector read.  No history is ever moved.
//   * one exp per weight, one exp per particle shared by the weight, the propagation mean of the
//     children and the score terms (exp(-x/2) is stored next to x).
//
// Deviations from the reference's operation order: parallel sums / scans, log(exp(x/2)) = x/2 and
// 1/exp(x/2)^2 = exp(-x/2)^2 in the log-weight, cumulative weights multiplied by 1/S; any
// log-weight shift cancels (Q4).  Resampling decisions within 64 ulp of a cumulative-weight tie
// are counted in diag[0]; decisions closer than the sequential-vs-tree summation bound are
// counted separately (ctrl->soft_ties, reported by pmmh_sv_grid_last_info).
// fp64, -fmad=false.  Bound: L2 / HBM streaming and gathers; no tensor cores (no contraction).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/pmmh_qn.h"
#include "common.cuh"
#include "sv_grid.cuh"
#include "sv_math.cuh"

namespace pmmh {

int set_error(int code, const char* what);           // capi.cu
int set_cuda_error(cudaError_t err, const char* where);

namespace {

constexpr int kGT = 1024;          // threads per CTA
constexpr int kCap = 8192;         // entries of one tile (shared-memory capacity)
constexpr int kKpt = kCap / kGT;   // entries per thread (strided assignment)
constexpr int kNF = 8192;          // bins of the global value histogram
constexpr int kNSB = 8192;         // sub-bins of the in-tile counting sort
constexpr int kMaxSub = 1024;      // a sub-bin larger than this abandons the evaluation
constexpr int kMaxTiles = 160;     // >= SM count
constexpr double kZ = 6.5;         // histogram range: predicted mean +- 6.5 predicted sd
constexpr int kDynSmem = 192 * 1024;
constexpr int kProf = 16;

struct __align__(16) MailEntry {
    double x;
    int j, a;
};
struct __align__(32) PEntry {
    double n, c, e, pad;   // value, parent value, exp(-parent value / 2)
};
struct __align__(32) REntry {
    int a[8];              // birth rows of the ancestors 1 .. 8 steps back
};

struct GridCtrl {
    unsigned bar;                  // arrival counter of the grid barrier (monotone)
    int status;                    // 0, or (reason << 24) | first barrier index at which everybody stops
    int max_bin;
    int pad0;
    unsigned long long near_ties, soft_ties, key_ties;
    unsigned long long mn[2], mx[2];   // ordered encodings of min / max child value, by step parity
};

struct GridArgs {
    int N, NOBS, LAG, G, Wc, RP, hist;
    const double *obs, *params, *rvr, *U;
    GridCtrl* ctrl;
    int* ghist;        // [2][kNF]
    int* tilecnt;      // [2][kMaxTiles]
    double* tinfo;     // [kMaxTiles][4]  tot, n, sum sh m, sum sh m^2
    int* H;            // [N] head markers (-1 = none)
    int* Hcarry;       // [kMaxTiles]
    double *xs, *es;   // [N] sorted generation, exp(-x/2)
    int* perm;         // [N] sorted position -> birth row
    MailEntry* mail;   // [N]
    REntry* R;         // [2][N]
    PEntry* P;         // [RP][N]
    double* psum;      // [NOBS][G][8]
    double *shiftv, *xminv;   // [NOBS]
    double* shring;    // [LAG][N] sh of the last LAG generations (sorted order)
    int* parentpos;    // [N] (history dump only)
    double* Xhist;
    int* Ahist;
    long long* prof;
};

struct StepScalars {
    double S, invS, mhat, inv_shat, xmin, xmax, shift, lo, hi, tot;
    int carry, hc, total, abort_now;
};

// ---- small device helpers ---------------------------------------------------------------------
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long enc_f64(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec_f64(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ void prefetch_l2(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ int odd_chunk(int n) {
    const int c = (n + kGT - 1) / kGT;
    return c < 1 ? 1 : (c | 1);
}
__device__ __forceinline__ int warp_incl_max(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(kFullMask, v, d);
        if (lane >= d) v = max(v, o);
    }
    return v;
}
__device__ __forceinline__ int pick8(const int4 lo, const int4 hi, int idx) {
    switch (idx) {
        case 0: return lo.x;
        case 1: return lo.y;
        case 2: return lo.z;
        case 3: return lo.w;
        case 4: return hi.x;
        case 5: return hi.y;
        case 6: return hi.z;
        default: return hi.w;
    }
}

// #{ j in [0, N) : (u + j) / N <= c }: closed form, then the exact predicate of :703-711.
// frac_out = distance of c*N - u to the nearest integer (how close the decision is to a tie).
__device__ __forceinline__ int count_le(double c, double u, int N, double dn, double inv_n, bool pow2,
                                        double& frac_out) {
    const double e = c * dn - u;
    int est;
    if (!(e >= 0.0)) est = 0;
    else if (e >= dn) est = N;
    else est = (int)e + 1;
    const double fl = floor(e);
    const double fr = e - fl;
    frac_out = fmin(fr, 1.0 - fr);
    if (pow2) {
        while (est > 0 && (u + (double)(est - 1)) * inv_n > c) --est;
        while (est < N && (u + (double)est) * inv_n <= c) ++est;
    } else {
        while (est > 0 && (u + (double)(est - 1)) / dn > c) --est;
        while (est < N && (u + (double)est) / dn <= c) ++est;
    }
    return est;
}

__device__ __forceinline__ int fine_bin(double x, double mhat, double inv_shat) {
    const double t = ((x - mhat) * inv_shat + kZ) * ((double)kNF / (2.0 * kZ));
    if (!(t >= 0.0)) return 0;
    if (t >= (double)kNF) return kNF - 1;
    return (int)t;
}
__device__ __forceinline__ int sub_bin(double x, double lo, double scale) {
    const double t = (x - lo) * scale;
    if (!(t >= 0.0)) return 0;
    if (t >= (double)kNSB) return kNSB - 1;
    return (int)t;
}

// Block-wide exclusive scans over one value per thread (1024 threads).  s_w: shared [32].
// Two barriers each; s_w may be reused right after the call.
__device__ __forceinline__ int block_excl_scan_int(int v, int* s_w, int& total, int lane, int warp) {
    const int incl = warp_incl_scan(v, lane);
    __syncthreads();
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    int off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kGT / 32; ++w) {
        const int s = s_w[w];
        if (w < warp) off += s;
        tot += s;
    }
    total = tot;
    return off + incl - v;
}
__device__ __forceinline__ int block_excl_max_int(int v, int init, int* s_w, int lane, int warp) {
    const int incl = warp_incl_max(v, lane);
    __syncthreads();
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    int off = init;
#pragma unroll
    for (int w = 0; w < kGT / 32; ++w)
        if (w < warp) off = max(off, s_w[w]);
    int ex = __shfl_up_sync(kFullMask, incl, 1);
    if (lane == 0) ex = init;
    return max(off, ex);
}
__device__ __forceinline__ double block_excl_scan_f64(double v, double* s_w, int lane, int warp) {
    const double incl = warp_incl_scan(v, lane);
    __syncthreads();
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    double off = 0.0;
    for (int w = 0; w < warp; ++w) off = off + s_w[w];
    double ex = __shfl_up_sync(kFullMask, incl, 1);
    if (lane == 0) ex = 0.0;
    return off + ex;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGT, 1) sv_grid_kernel(const GridArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    // phase B / C view
    double* s_x = (double*)smem;                       // [kCap] sorted values, later exp(-x/2)
    double* s_sh = (double*)(smem + 65536);            // [kCap] unnormalised weights
    int* s_j = (int*)(smem + 131072);                  // [kCap] birth rows
    int* s_a = (int*)(smem + 163840);                  // [kCap] lagged ancestor rows
    int* s_sub = (int*)s_sh;                           // [kNSB] sub-bin counters (during the sort only)
    // phase A view
    int* s_fhist = (int*)smem;                         // [kNF] histogram of this CTA's children
    int* s_par = (int*)smem + kNF;                     // [kCap] ancestor (sorted position) of every child
    unsigned short* s_tileof = (unsigned short*)s_j;   // [kNF] tile of a histogram bin

    __shared__ SvConst s_k;
    __shared__ StepScalars s_sc;
    __shared__ double s_tot[kMaxTiles], s_m1[kMaxTiles], s_m2[kMaxTiles], s_off[kMaxTiles + 1];
    __shared__ int s_tstart[kMaxTiles + 1], s_tcnt[kMaxTiles], s_tbase[kMaxTiles];
    __shared__ double s_red[8 * 32];
    __shared__ double s_wd[32];
    __shared__ int s_wi[32];
    __shared__ long long s_prof[kProf];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x, G = a.G, N = a.N, NOBS = a.NOBS, L = a.LAG, Wc = a.Wc, RP = a.RP;
    GridCtrl* ctrl = a.ctrl;
    const double dn = (double)N, inv_n = 1.0 / dn;
    const bool pow2 = (N & (N - 1)) == 0;
    unsigned epoch = 0;                 // arrives done so far
    unsigned cnt_near = 0, cnt_soft = 0, cnt_key = 0;
    int my_max_bin = 0;
    long long pclk = 0;
    const bool prof = a.prof != nullptr;

#define GRID_FLAG(reason) atomicCAS(&ctrl->status, 0, (int)((epoch + 1u) | ((unsigned)(reason) << 24)))
#define GRID_ARRIVE()                                \
    do {                                             \
        __syncthreads();                             \
        if (tid == 0) {                              \
            __threadfence();                         \
            red_release_add(&ctrl->bar, 1u);         \
        }                                            \
        ++epoch;                                     \
    } while (0)
#define GRID_WAIT()                                                                   \
    do {                                                                              \
        if (tid == 0) {                                                               \
            const unsigned tgt = epoch * (unsigned)G;                                 \
            while (ld_acquire_u32(&ctrl->bar) < tgt) {                                \
            }                                                                         \
            __threadfence();                                                          \
            const int stv = *(volatile int*)&ctrl->status;                            \
            s_sc.abort_now = (stv != 0 && (unsigned)(stv & 0xffffff) <= epoch) ? 1 : 0; \
        }                                                                             \
        __syncthreads();                                                              \
    } while (0)
#define PROF_MARK(slot)                              \
    do {                                             \
        if (prof && tid == 0) {                      \
            const long long now__ = clock64();       \
            s_prof[slot] += now__ - pclk;            \
            pclk = now__;                            \
        }                                            \
    } while (0)

    if (tid == 0) {
        sv_const_init(s_k, a.params);
        s_sc.abort_now = 0;
        for (int i = 0; i < kProf; ++i) s_prof[i] = 0;
    }
    for (int i = tid; i < kMaxTiles; i += kGT) s_tcnt[i] = 0;
    __syncthreads();

    // ------------------------------------------------------------------------------------------
    // generation 0 (:306-323, Q1): every particle = mu, uniform weights, identity order
    // ------------------------------------------------------------------------------------------
    const int jb = min(N, c * Wc), je = min(N, jb + Wc), nc = je - jb;   // this CTA's children
    int pstart = jb, n = nc;                                             // this CTA's tile
    double toff;                                                          // cumulative weight in front of this thread's chunk
    {
        const double mu = s_k.mu;
        const double e0 = exp(-0.5 * mu);
        double m0 = s_k.mu + s_k.phi * (mu - s_k.mu);
        m0 += (s_k.sr * e0) * a.obs[0];
        for (int q = tid; q < n; q += kGT) {
            __stcg(&a.xs[pstart + q], mu);
            __stcg(&a.es[pstart + q], e0);
            __stcg(&a.perm[pstart + q], pstart + q);
            s_sh[q] = 1.0;
            int4 z = make_int4(0, 0, 0, 0);
            __stcg((int4*)&a.R[pstart + q].a[0], z);
            __stcg((int4*)&a.R[pstart + q].a[4], z);
            if (a.hist) {
                a.Xhist[pstart + q] = mu;
                a.Ahist[pstart + q] = pstart + q;
            }
        }
        const int Lc = odd_chunk(n);
        toff = (double)min(tid * Lc, n);
        if (tid == 0) {
            const double dnk = (double)n;
            __stcg((double2*)&a.tinfo[c * 4], make_double2(dnk, dnk));
            __stcg((double2*)&a.tinfo[c * 4 + 2], make_double2(dnk * m0, dnk * (m0 * m0)));
            double* ps = a.psum + ((size_t)0 * G + c) * 8;
            ps[0] = dnk;
            ps[1] = dnk * mu;
            for (int i = 2; i < 8; ++i) ps[i] = 0.0;
            if (c == 0) {
                a.shiftv[0] = 0.0;
                a.xminv[0] = mu;
                ctrl->mn[0] = ctrl->mn[1] = ~0ull;
                ctrl->mx[0] = ctrl->mx[1] = 0ull;
            }
        }
    }
    GRID_ARRIVE();
    GRID_WAIT();
    if (prof && tid == 0) pclk = clock64();

    for (int t = 1; t < NOBS; ++t) {
        const int par = t & 1;
        // --------------------------------------------------------------------------------------
        // phase C: totals of all tiles -> offsets; child ranges of this tile's parents (:694-715)
        // --------------------------------------------------------------------------------------
        if (tid < G) {
            const double2 t0 = __ldcg((const double2*)&a.tinfo[tid * 4]);
            const double2 t1 = __ldcg((const double2*)&a.tinfo[tid * 4 + 2]);
            s_tot[tid] = t0.x;
            s_m1[tid] = t1.x;
            s_m2[tid] = t1.y;
        }
        __syncthreads();
        if (warp == 0) {
            constexpr int kPer = kMaxTiles / 32;
            double loc = 0.0, l1 = 0.0, l2 = 0.0;
#pragma unroll
            for (int i = 0; i < kPer; ++i) {
                const int k = lane * kPer + i;
                if (k < G) {
                    loc = loc + s_tot[k];
                    l1 = l1 + s_m1[k];
                    l2 = l2 + s_m2[k];
                }
            }
            const double incl = warp_incl_scan(loc, lane);
            double run = __shfl_up_sync(kFullMask, incl, 1);
            if (lane == 0) run = 0.0;
#pragma unroll
            for (int i = 0; i < kPer; ++i) {
                const int k = lane * kPer + i;
                if (k < G) {
                    s_off[k] = run;
                    run = run + s_tot[k];
                }
            }
            const double S = __shfl_sync(kFullMask, incl, 31);
            const double M1 = warp_sum(l1), M2 = warp_sum(l2);
            if (lane == 0) {
                s_off[G] = S;
                const double invS = 1.0 / S;
                const double mhat = M1 / S;
                double var = M2 / S - mhat * mhat;
                if (!(var > 0.0)) var = 0.0;
                var = var + s_k.sd * s_k.sd;
                const double shat = sqrt(var);
                s_sc.S = S;
                s_sc.invS = invS;
                s_sc.mhat = mhat;
                s_sc.inv_shat = 1.0 / shat;
                if (!(S > 0.0) || !isfinite(S) || !isfinite(mhat) || !(shat > 0.0) || !isfinite(shat)) GRID_FLAG(2);
            }
        }
        __syncthreads();
        const double ur = a.rvr[t];
        if (warp == 0 && lane == 0) {
            // child range end of the tiles in front of this one (running maximum, see below)
            int carry = 0;
            double fr;
            for (int k = max(0, c - 2); k < c; ++k)
                carry = max(carry, count_le((s_off[k] + s_tot[k]) * s_sc.invS, ur, N, dn, inv_n, pow2, fr));
            s_sc.carry = carry;
        }
        {
            const double invS = s_sc.invS, offk = s_off[c];
            const int Lc = odd_chunk(n), q0 = tid * Lc;
            const double tol_near = 64.0 * 2.220446049250313e-16 * dn;
            const double tol_soft = 2.220446049250313e-16 * dn * (4.0 + 2.0 * sqrt(dn));
            int ubv[9];
            int rmax = 0;
            double run = 0.0;
#pragma unroll
            for (int kk = 0; kk < 9; ++kk) {
                const int q = q0 + kk;
                ubv[kk] = 0;
                if (kk < Lc && q < n) {
                    run = run + s_sh[q];
                    const double cN = (offk + (toff + run)) * invS;
                    int ub;
                    if (pstart + q == N - 1) {
                        ub = N;
                    } else {
                        double fr;
                        ub = count_le(cN, ur, N, dn, inv_n, pow2, fr);
                        if (fr < tol_soft) {
                            ++cnt_soft;
                            if (fr < tol_near * fmax(cN, inv_n)) ++cnt_near;
                        }
                    }
                    rmax = max(rmax, ub);
                    ubv[kk] = rmax;
                }
            }
            // parallel scans are monotone only up to an ulp: a running maximum over all parents
            // (and over the tiles in front) keeps the child ranges disjoint
            __syncthreads();   // s_sc.carry
            int prev = block_excl_max_int(rmax, s_sc.carry, s_wi, lane, warp);
#pragma unroll
            for (int kk = 0; kk < 9; ++kk) {
                const int q = q0 + kk;
                if (kk < Lc && q < n) {
                    const int ub = max(ubv[kk], prev);
                    if (ub > prev) {
                        const int P = pstart + q;
                        __stcg(&a.H[prev], P);
                        int m = prev / Wc + 1;              // first child-tile boundary behind prev
                        while (m < G && m * Wc < ub) {
                            __stcg(&a.Hcarry[m], P);
                            ++m;
                        }
                    }
                    prev = ub;
                }
            }
        }
        PROF_MARK(0);
        GRID_ARRIVE();   // ---- barrier 4: head markers complete
        for (int b = tid; b < kNF; b += kGT) s_fhist[b] = 0;
        PROF_MARK(1);
        GRID_WAIT();
        PROF_MARK(2);
        if (s_sc.abort_now) break;

        // --------------------------------------------------------------------------------------
        // phase A: ancestors of this CTA's children, propagation (:354-358), value histogram
        // --------------------------------------------------------------------------------------
        for (int i = tid; i < nc; i += kGT) {
            s_par[i] = __ldcg(&a.H[jb + i]);
            __stcg(&a.H[jb + i], -1);
        }
        if (tid == 0) {
            s_sc.hc = __ldcg(&a.Hcarry[c]);
            __stcg(&a.Hcarry[c], -1);
        }
        __syncthreads();
        {
            const int Lc2 = odd_chunk(nc), i0 = tid * Lc2;
            int mx = -1;
#pragma unroll
            for (int kk = 0; kk < 9; ++kk) {
                const int i = i0 + kk;
                if (kk < Lc2 && i < nc) mx = max(mx, s_par[i]);
            }
            int run = block_excl_max_int(mx, s_sc.hc, s_wi, lane, warp);
            bool orphan = false;
#pragma unroll
            for (int kk = 0; kk < 9; ++kk) {
                const int i = i0 + kk;
                if (kk < Lc2 && i < nc) {
                    run = max(run, s_par[i]);
                    if (run < 0 || run >= N) {
                        orphan = true;
                        run = 0;
                    }
                    s_par[i] = run;
                }
            }
            if (orphan) GRID_FLAG(4);
        }
        __syncthreads();
        double xn[kKpt];
        int bp[kKpt];
        {
            const double y1 = a.obs[t - 1];
            const double mhat = s_sc.mhat, inv_shat = s_sc.inv_shat;
            const double* Ut = a.U + (size_t)t * N;
            PEntry* Pt = a.P + (size_t)(t % RP) * N;
            double vmin = INFINITY, vmax = -INFINITY;
            bool bad = false;
#pragma unroll
            for (int kk = 0; kk < kKpt; ++kk) {
                const int i = kk * kGT + tid;
                xn[kk] = 0.0;
                bp[kk] = 0;
                if (i < nc) {
                    const int j = jb + i;
                    const int p = s_par[i];
                    const double xp = __ldcg(&a.xs[p]);
                    const double ep = __ldcg(&a.es[p]);
                    const int b = __ldcg(&a.perm[p]);
                    const double uu = ld_stream_f64(Ut + j);
                    double mean = s_k.mu + s_k.phi * (xp - s_k.mu);     // :355
                    mean += (s_k.sr * ep) * y1;                         // :356
                    const double x = mean + s_k.sd * uu;                // :357-358
                    if (!isfinite(x)) bad = true;
                    xn[kk] = x;
                    bp[kk] = b;
                    atomicAdd(&s_fhist[fine_bin(x, mhat, inv_shat)], 1);
                    vmin = fmin(vmin, x);
                    vmax = fmax(vmax, x);
                    __stcs((double2*)&Pt[j], make_double2(x, xp));
                    __stcs((double2*)&Pt[j] + 1, make_double2(ep, 0.0));
                    if (a.hist) a.parentpos[j] = p;
                }
            }
            if (bad) GRID_FLAG(2);
            vmin = warp_min(vmin);
            vmax = warp_max(vmax);
            if (lane == 0) {
                s_red[warp] = vmin;
                s_red[32 + warp] = vmax;
            }
        }
        __syncthreads();
        {
            int* gh = a.ghist + par * kNF;
            for (int b = tid; b < kNF; b += kGT) {
                const int cnt = s_fhist[b];
                if (cnt) atomicAdd(&gh[b], cnt);
            }
            if (warp == 0) {
                const double vmin = warp_min(s_red[lane]), vmax = warp_max(s_red[32 + lane]);
                if (lane == 0 && nc > 0) {
                    atomicMin(&ctrl->mn[par], enc_f64(vmin));
                    atomicMax(&ctrl->mx[par], enc_f64(vmax));
                }
            }
        }
        PROF_MARK(3);
        GRID_ARRIVE();   // ---- barrier 1: global histogram complete
        {
            // genealogy records (only feed outputs): child = (parent row, parent's ancestors 1..7)
            const REntry* Rp = a.R + (size_t)((t - 1) & 1) * N;
            REntry* Rc = a.R + (size_t)(t & 1) * N;
#pragma unroll
            for (int kk = 0; kk < kKpt; ++kk) {
                const int i = kk * kGT + tid;
                if (i < nc) {
                    const int j = jb + i;
                    const int b = bp[kk];
                    const int4 r0 = __ldcg((const int4*)&Rp[b].a[0]);
                    const int4 r1 = __ldcg((const int4*)&Rp[b].a[4]);
                    const int4 n0 = make_int4(b, r0.x, r0.y, r0.z);
                    const int4 n1 = make_int4(r0.w, r1.x, r1.y, r1.z);
                    __stcg((int4*)&Rc[j].a[0], n0);
                    __stcg((int4*)&Rc[j].a[4], n1);
                    bp[kk] = (L == 2) ? j : pick8(n0, n1, L - 3);   // row of the ancestor L-2 steps back
                }
            }
            // housekeeping for the next step
            const int zper = (kNF + G - 1) / G;
            int* ghn = a.ghist + (par ^ 1) * kNF;
            for (int b = c * zper + tid; b < min(kNF, (c + 1) * zper); b += kGT) __stcg(&ghn[b], 0);
            if (tid == 0) {
                __stcg(&a.tilecnt[(par ^ 1) * kMaxTiles + c], 0);
                if (c == 0) {
                    ctrl->mn[par ^ 1] = ~0ull;
                    ctrl->mx[par ^ 1] = 0ull;
                }
            }
        }
        PROF_MARK(4);
        GRID_WAIT();
        PROF_MARK(5);
        if (s_sc.abort_now) break;

        // scan of the global histogram: tile boundaries on bin edges, tile of every bin
        {
            const int* gh = a.ghist + par * kNF;
            const int4 v0 = __ldcg((const int4*)(gh + 8 * tid));
            const int4 v1 = __ldcg((const int4*)(gh + 8 * tid + 4));
            const int prevcnt = tid > 0 ? __ldcg(gh + 8 * tid - 1) : 0;
            const int cnt[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            int loc = 0, mxb = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                loc += cnt[i];
                mxb = max(mxb, cnt[i]);
            }
            my_max_bin = max(my_max_bin, mxb);
            int total;
            int start = block_excl_scan_int(loc, s_wi, total, lane, warp);
            int tprev = (tid == 0) ? -1 : min(G - 1, (start - prevcnt) / Wc);
            int tl = min(G - 1, start / Wc);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                while (tl < G - 1 && start >= (tl + 1) * Wc) ++tl;
                s_tileof[8 * tid + i] = (unsigned short)tl;
                for (int k = tprev + 1; k <= tl; ++k) s_tstart[k] = start;
                tprev = tl;
                start += cnt[i];
            }
            if (tid == kGT - 1) {
                for (int k = tprev + 1; k <= G; ++k) s_tstart[k] = N;
                if (total != N) GRID_FLAG(5);
            }
            if (tid == 0) {
                s_sc.xmin = dec_f64(*(volatile unsigned long long*)&ctrl->mn[par]);
                s_sc.xmax = dec_f64(*(volatile unsigned long long*)&ctrl->mx[par]);
            }
        }
        __syncthreads();
        int kr[kKpt];
        {
            const double mhat = s_sc.mhat, inv_shat = s_sc.inv_shat;
#pragma unroll
            for (int kk = 0; kk < kKpt; ++kk) {
                const int i = kk * kGT + tid;
                kr[kk] = 0;
                if (i < nc) {
                    const int tl = s_tileof[fine_bin(xn[kk], mhat, inv_shat)];
                    const int r = atomicAdd(&s_tcnt[tl], 1);
                    kr[kk] = (tl << 16) | r;
                }
            }
        }
        __syncthreads();
        if (tid < G) {
            const int cnt = s_tcnt[tid];
            int base = s_tstart[tid];
            if (cnt) base += atomicAdd(&a.tilecnt[par * kMaxTiles + tid], cnt);
            s_tbase[tid] = base;
            s_tcnt[tid] = 0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kKpt; ++kk) {
            const int i = kk * kGT + tid;
            if (i < nc) {
                const int pos = s_tbase[kr[kk] >> 16] + (kr[kk] & 0xffff);
                const long long xb = __double_as_longlong(xn[kk]);
                const int4 ent = make_int4((int)(xb & 0xffffffffll), (int)(xb >> 32), jb + i, bp[kk]);
                if (pos >= 0 && pos < N) __stcg((int4*)&a.mail[pos], ent);
            }
        }
        pstart = s_tstart[c];
        n = s_tstart[c + 1] - pstart;
        if (n > kCap || n < 0) {
            if (tid == 0) GRID_FLAG(1);
            n = n < 0 ? 0 : kCap;
        }
        PROF_MARK(6);
        GRID_ARRIVE();   // ---- barrier 2: mailboxes complete
        for (int b = tid; b < kNSB; b += kGT) s_sub[b] = 0;
        if (tid == 0) {
            // shift = largest log-weight over [xmin, xmax] (any shift cancels, Q4)
            const double y = a.obs[t];
            const double xmin = s_sc.xmin, xmax = s_sc.xmax;
            double xs_ = (y != 0.0) ? 2.0 * log(fabs(y)) : xmin;
            xs_ = fmin(fmax(xs_, xmin), xmax);
            const double es_ = exp(-0.5 * xs_);
            s_sc.shift = (-0.91893853320467267 - 0.5 * xs_) - (0.5 * y * y) * (es_ * es_);
            if (c == 0) {
                a.shiftv[t] = s_sc.shift;
                a.xminv[t] = xmin;
            }
            if (t + 1 < NOBS && nc > 0) {
                // next step's slice of u -> L2 (TMA-class bulk prefetch, no SM cycles)
                const char* p0 = (const char*)(a.U + (size_t)(t + 1) * N + jb);
                const char* p1 = (const char*)(a.U + (size_t)(t + 1) * N + je);
                const char* q0 = (const char*)(((uintptr_t)p0 + 15) & ~(uintptr_t)15);
                const char* q1 = (const char*)((uintptr_t)p1 & ~(uintptr_t)15);
                for (const char* q = q0; q < q1; q += 16384)
                    prefetch_l2(q, (unsigned)min((long long)16384, (long long)(q1 - q)));
            }
        }
        PROF_MARK(7);
        GRID_WAIT();
        PROF_MARK(8);
        if (s_sc.abort_now) break;

        // --------------------------------------------------------------------------------------
        // phase B: sort this tile (:392-424 / :23-52), weights (:427-442), block scan
        // --------------------------------------------------------------------------------------
        {
            const MailEntry* mb = a.mail + pstart;
            double ex[kKpt];
            int ej[kKpt], ea[kKpt], er[kKpt];
            double lmin = INFINITY, lmax = -INFINITY;
#pragma unroll
            for (int kk = 0; kk < kKpt; ++kk) {
                const int e = kk * kGT + tid;
                ex[kk] = 0.0;
                ej[kk] = ea[kk] = 0;
                if (e < n) {
                    const int4 raw = __ldcg((const int4*)(mb + e));
                    ex[kk] = __longlong_as_double(((long long)raw.y << 32) | (long long)(unsigned)raw.x);
                    ej[kk] = raw.z;
                    ea[kk] = raw.w;
                    lmin = fmin(lmin, ex[kk]);
                    lmax = fmax(lmax, ex[kk]);
                }
            }
            lmin = warp_min(lmin);
            lmax = warp_max(lmax);
            if (lane == 0) {
                s_red[warp] = lmin;
                s_red[32 + warp] = lmax;
            }
            __syncthreads();
            if (warp == 0) {
                const double lo = warp_min(s_red[lane]), hi = warp_max(s_red[32 + lane]);
                if (lane == 0) {
                    s_sc.lo = lo;
                    s_sc.hi = hi;
                }
            }
            __syncthreads();
            const double lo = s_sc.lo;
            const double scale = (s_sc.hi > lo) ? (double)kNSB / (s_sc.hi - lo) : 0.0;
#pragma unroll
            for (int kk = 0; kk < kKpt; ++kk) {
                const int e = kk * kGT + tid;
                er[kk] = 0;
                if (e < n) er[kk] = atomicAdd(&s_sub[sub_bin(ex[kk], lo, scale)], 1);
            }
            __syncthreads();
            {
                // exclusive scan of the sub-bin counters in place (8 consecutive bins per thread)
                int4 v0 = *(int4*)(s_sub + 8 * tid), v1 = *(int4*)(s_sub + 8 * tid + 4);
                int cnt[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                int loc = 0, mxb = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    loc += cnt[i];
                    mxb = max(mxb, cnt[i]);
                }
                if (mxb > kMaxSub) GRID_FLAG(3);
                int total;
                int start = block_excl_scan_int(loc, s_wi, total, lane, warp);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int cn = cnt[i];
                    cnt[i] = start;
                    start += cn;
                }
                *(int4*)(s_sub + 8 * tid) = make_int4(cnt[0], cnt[1], cnt[2], cnt[3]);
                *(int4*)(s_sub + 8 * tid + 4) = make_int4(cnt[4], cnt[5], cnt[6], cnt[7]);
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < kKpt; ++kk) {
                const int e = kk * kGT + tid;
                if (e < n) {
                    const int pos = s_sub[sub_bin(ex[kk], lo, scale)] + er[kk];
                    s_x[pos] = ex[kk];
                    s_j[pos] = ej[kk];
                    s_a[pos] = ea[kk];
                }
            }
            __syncthreads();
            // exact order inside a sub-bin: by value, then by birth row (:32-35 never returns 0)
            int np[kKpt];
#pragma unroll
            for (int kk = 0; kk < kKpt; ++kk) {
                const int q = kk * kGT + tid;
                np[kk] = q;
                if (q < n) {
                    const double x = s_x[q];
                    const int kb = sub_bin(x, lo, scale);
                    const int b0 = s_sub[kb];
                    const int b1 = (kb + 1 < kNSB) ? s_sub[kb + 1] : n;
                    if (b1 - b0 > 1 && b1 - b0 <= kMaxSub) {
                        const int j = s_j[q];
                        int rank = 0;
                        for (int m = b0; m < b1; ++m) {
                            const double xm = s_x[m];
                            if (xm < x) ++rank;
                            else if (xm == x && m != q) {
                                ++cnt_key;
                                if (s_j[m] < j) ++rank;
                            }
                        }
                        np[kk] = b0 + rank;
                    }
                }
            }
            // in-place permutation, one array at a time (keeps the register footprint small)
#pragma unroll
            for (int kk = 0; kk < kKpt; ++kk) {
                const int q = kk * kGT + tid;
                if (q < n) ex[kk] = s_x[q];
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < kKpt; ++kk) {
                const int q = kk * kGT + tid;
                if (q < n) {
                    s_x[np[kk]] = ex[kk];
                    ej[kk] = s_j[q];
                }
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < kKpt; ++kk) {
                const int q = kk * kGT + tid;
                if (q < n) {
                    s_j[np[kk]] = ej[kk];
                    ea[kk] = s_a[q];
                }
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < kKpt; ++kk) {
                const int q = kk * kGT + tid;
                if (q < n) s_a[np[kk]] = ea[kk];
            }
            __syncthreads();
        }
        PROF_MARK(9);
        // the new sorted generation (values, birth rows)
#pragma unroll
        for (int kk = 0; kk < kKpt; ++kk) {
            const int q = kk * kGT + tid;
            if (q < n) {
                const double x = s_x[q];
                const int j = s_j[q];
                __stcg(&a.xs[pstart + q], x);
                __stcg(&a.perm[pstart + q], j);
                if (a.hist) {
                    a.Xhist[(size_t)t * N + pstart + q] = x;
                    a.Ahist[(size_t)t * N + pstart + q] = __ldcg(&a.parentpos[j]);
                }
            }
        }
        __syncthreads();
        {
            // weights (:427-437): lw = -0.9189 - x/2 - y^2 exp(-x) / 2, sh = exp(lw - shift);
            // thread = Lc consecutive sorted particles, sequential running sum
            const double y = a.obs[t], hy2 = 0.5 * y * y, shift = s_sc.shift;
            const int Lc = odd_chunk(n), q0 = tid * Lc;
            double run = 0.0;
            double acc[3] = {0.0, 0.0, 0.0};
            bool bad = false;
#pragma unroll
            for (int kk = 0; kk < 9; ++kk) {
                const int q = q0 + kk;
                if (kk < Lc && q < n) {
                    const double x = s_x[q];
                    const double e = exp(-0.5 * x);
                    const double lw = (-0.91893853320467267 - 0.5 * x) - hy2 * (e * e);
                    double sh = exp(lw - shift);
                    if (!isfinite(sh)) {
                        bad = true;
                        sh = 0.0;
                    }
                    s_x[q] = e;
                    s_sh[q] = sh;
                    run = run + sh;
                    acc[0] += sh * x;
                    double m = s_k.mu + s_k.phi * (x - s_k.mu);
                    m += (s_k.sr * e) * y;
                    acc[1] += sh * m;
                    acc[2] += sh * (m * m);
                }
            }
            if (bad) GRID_FLAG(2);
            toff = block_excl_scan_f64(run, s_wd, lane, warp);
            if (n > 0 && tid == (n - 1) / Lc) s_sc.tot = toff + run;   // cumulative weight of the tile's last particle
            if (n == 0 && tid == 0) s_sc.tot = 0.0;
            block_sum<3>(acc, s_red);
            if (tid == 0) {
                const double tot = s_sc.tot;
                __stcg((double2*)&a.tinfo[c * 4], make_double2(tot, (double)n));
                __stcg((double2*)&a.tinfo[c * 4 + 2], make_double2(acc[1], acc[2]));
                double* ps = a.psum + ((size_t)t * G + c) * 8;
                ps[0] = tot;
                ps[1] = acc[0];
            }
        }
        PROF_MARK(10);
        GRID_ARRIVE();   // ---- barrier 3: tile totals published
        {
            // fixed-lag score terms (:445-470): ancestor pair (time t-L+1, t-L+2) from one sector
            const int Lc = odd_chunk(n), q0 = tid * Lc;
            double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
            if (t >= L) {
                const double ylag = a.obs[t - L];   // Q5: obs[i - LAG]
                const PEntry* Pg = a.P + (size_t)((t - (L - 2)) % RP) * N;
#pragma unroll
                for (int kk = 0; kk < 9; ++kk) {
                    const int q = q0 + kk;
                    if (kk < Lc && q < n) {
                        int an = s_a[q];
                        an = min(max(an, 0), N - 1);
                        const double sh = s_sh[q];
                        const double2 p0 = __ldcg((const double2*)&Pg[an]);
                        const double ec = __ldcg(&Pg[an].e);
                        double sq, g[4];
                        sv_score_main_e(s_k, p0.y, ec, p0.x, ylag, sq, g);
                        acc[0] += sh * p0.y;
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[1 + i] += g[i] * sh;
                    }
                }
            }
            block_sum<5>(acc, s_red);
            if (tid == 0) {
                double* ps = a.psum + ((size_t)t * G + c) * 8;
#pragma unroll
                for (int i = 0; i < 5; ++i) ps[2 + i] = acc[i];
                ps[7] = 0.0;
            }
            const bool keep = t >= NOBS - L;
#pragma unroll
            for (int kk = 0; kk < kKpt; ++kk) {
                const int q = kk * kGT + tid;
                if (q < n) {
                    __stcg(&a.es[pstart + q], s_x[q]);
                    if (keep) a.shring[(size_t)(t % L) * N + pstart + q] = s_sh[q];
                }
            }
        }
        PROF_MARK(11);
        GRID_WAIT();
        PROF_MARK(12);
        if (s_sc.abort_now) break;
    }

    // diagnostics
    cnt_near = __reduce_add_sync(kFullMask, cnt_near);
    cnt_soft = __reduce_add_sync(kFullMask, cnt_soft);
    cnt_key = __reduce_add_sync(kFullMask, cnt_key);
    my_max_bin = __reduce_max_sync(kFullMask, my_max_bin);
    if (lane == 0) {
        if (cnt_near) atomicAdd(&ctrl->near_ties, (unsigned long long)cnt_near);
        if (cnt_soft) atomicAdd(&ctrl->soft_ties, (unsigned long long)cnt_soft);
        if (cnt_key) atomicAdd(&ctrl->key_ties, (unsigned long long)(cnt_key / 2));
        atomicMax(&ctrl->max_bin, my_max_bin);
    }
    if (prof && tid == 0)
        for (int i = 0; i < kProf; ++i) a.prof[(size_t)c * kProf + i] += s_prof[i];
#undef GRID_FLAG
#undef GRID_ARRIVE
#undef GRID_WAIT
#undef PROF_MARK
}

// ---------------------------------------------------------------------------------------------
// after the persistent kernel: O(T G) reductions, the tail (:540-562, Q6), output assembly
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grid_reduce_kernel(const double* __restrict__ psum, int G,
                                                          double* __restrict__ sums) {
    const int t = blockIdx.x, k = threadIdx.x >> 5, lane = threadIdx.x & 31;   // 8 warps = 8 components
    double s = 0.0;
    for (int c = lane; c < G; c += 32) s = s + psum[((size_t)t * G + c) * 8 + k];
    s = warp_sum(s);
    if (lane == 0) sums[(size_t)t * 8 + k] = s;
}

// part[irel][block][0] = sum_p W_T[p] hist_idx[p];  [1..4] = sum_p W_i[p] g(hist_idx, hist_idx-1)
// with i = NOBS - L + irel, idx = L - 1 - irel; hist_k[p] = value of the ancestor k steps back of
// the particle at sorted position p of the final generation.
__global__ void __launch_bounds__(256) grid_tail_kernel(GridArgs a, const double* __restrict__ sums,
                                                        double* __restrict__ part, int nblk) {
    __shared__ double red[5 * 32];
    const int tid = threadIdx.x;
    const int L = a.LAG, N = a.N, T = a.NOBS - 1, RP = a.RP;
    const int irel = blockIdx.y, i = a.NOBS - L + irel, idx = L - 1 - irel;
    SvConst k;
    sv_const_init(k, a.params);
    const double y1 = obs_wrap(a.obs, i - 1, a.NOBS);
    const double ST = sums[(size_t)T * 8], Si = sums[(size_t)i * 8];
    const double* shT = a.shring + (size_t)(T % L) * N;
    const double* shi = a.shring + (size_t)(i % L) * N;
    const REntry* Rt = a.R + (size_t)(T & 1) * N;
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const bool live = a.ctrl->status == 0;   // an abandoned evaluation leaves stale rows behind
    for (int p = blockIdx.x * 256 + tid; live && p < N; p += nblk * 256) {
        const int b = min(max(a.perm[p], 0), N - 1);
        const double wT = shT[p] / ST;
        double curr;
        if (idx == 0) {
            curr = a.xs[p];
            acc[0] += wT * curr;
        } else {
            // entry of the ancestor idx-1 steps back holds (next = its value, curr = its parent's value)
            const int m = idx - 1;
            const int row = (m == 0) ? b : min(max(Rt[b].a[m - 1], 0), N - 1);
            const PEntry pe = a.P[(size_t)((T - m) % RP) * N + row];
            curr = pe.c;
            acc[0] += wT * curr;
            const double wi = shi[p] / Si;
            double sq, g[4];
            sv_score_tail_e(k, curr, pe.e, pe.n, y1, sq, g);
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[1 + q] += g[q] * wi;
        }
    }
    block_sum<5>(acc, red);
    if (tid < 5) part[((size_t)irel * nblk + blockIdx.x) * 8 + tid] = acc[tid];
}

__global__ void __launch_bounds__(256) grid_tail_reduce_kernel(const double* __restrict__ part, int nblk,
                                                               double* __restrict__ out) {
    __shared__ double red[5 * 32];
    const int irel = blockIdx.x, tid = threadIdx.x;
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int q = tid; q < nblk; q += 256)
#pragma unroll
        for (int k = 0; k < 5; ++k) acc[k] += part[((size_t)irel * nblk + q) * 8 + k];
    block_sum<5>(acc, red);
    if (tid < 5) out[irel * 8 + tid] = acc[tid];
}

// sums[t][0] = sum sh, [1] = sum sh x, [2] = sum sh curr, [3..6] = sum sh g;  tail[irel][0..4]
__global__ void grid_finish_kernel(const GridCtrl* __restrict__ ctrl, const double* __restrict__ sums,
                                   const double* __restrict__ shift, const double* __restrict__ xmin,
                                   const double* __restrict__ tail, const double* __restrict__ params,
                                   int nobs, int L, double n_total, double* __restrict__ log_like,
                                   double* __restrict__ filt, double* __restrict__ smo,
                                   double* __restrict__ grad, double* __restrict__ traj,
                                   long long* __restrict__ diag, long long* __restrict__ info) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid == 0) {
        double ll = 0.0;
        const double logn = log(n_total);
        for (int t = 1; t < nobs; ++t) ll += shift[t] + log(sums[(size_t)t * 8]) - logn;   // :537
        log_like[0] = ll;
        for (int k = 0; k < PMMH_DIAG_COUNT; ++k) diag[k] = 0;
        diag[PMMH_DIAG_NEAR_TIES] = (long long)ctrl->near_ties;
        diag[PMMH_DIAG_MAX_BIN] = ctrl->max_bin;
        diag[PMMH_DIAG_STATUS] = ctrl->status ? 1 : 0;
        diag[PMMH_DIAG_KEY_TIES] = (long long)ctrl->key_ties;
        diag[PMMH_DIAG_KERNEL] = 5;
        diag[PMMH_DIAG_FAST_INFO] = ctrl->status;
        if (info) {
            info[0] = (long long)ctrl->soft_ties;
            info[1] = ctrl->status;
        }
    }
    for (int t = tid; t < nobs; t += gridDim.x * blockDim.x) {
        filt[t] = sums[(size_t)t * 8 + 1] / sums[(size_t)t * 8];
        traj[t] = (t == 0) ? params[0] : xmin[t];   // Q10/Q11: traj[t] = X_t[0] for t >= 1, X_0 == mu
        double s = 0.0, g[4] = {0.0, 0.0, 0.0, 0.0};
        const int src = t + L - 1;   // main-loop terms land at tt = i - L + 1 (:445-470)
        if (t >= 1 && src < nobs) {
            const double S = sums[(size_t)src * 8];
            s = sums[(size_t)src * 8 + 2] / S;
            for (int q = 0; q < 4; ++q) g[q] = sums[(size_t)src * 8 + 3 + q] / S;
        }
        const int irel_s = t - (nobs - L);   // tail: i = nobs-L+irel adds smo[i], gradient[.][i-L+1]
        if (irel_s >= 0 && irel_s < L) s += tail[irel_s * 8];
        const int irel_g = t + L - 1 - (nobs - L);
        if (irel_g >= 0 && irel_g < L - 1)
            for (int q = 0; q < 4; ++q) g[q] += tail[irel_g * 8 + 1 + q];
        smo[t] = s;
        for (int q = 0; q < 4; ++q) grad[(size_t)q * nobs + t] = g[q];
    }
}

struct GridLayout {
    size_t ctrl, ghist, tilecnt, tinfo, H, Hcarry, xs, es, perm, mail, R, P, psum, shiftv, xminv, shring, parentpos,
        sums, tailpart, tail, info, total;
    int RP, nblk;
};

size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

GridLayout grid_layout(int nobs, int n, int lag, int G, int hist) {
    GridLayout L = GridLayout();
    L.RP = lag - 1 < 2 ? 2 : lag - 1;
    L.nblk = 296;
    size_t o = 0;
    const size_t N = (size_t)n;
    L.ctrl = o;      o += al256(sizeof(GridCtrl));
    L.ghist = o;     o += al256((size_t)2 * kNF * 4);
    L.tilecnt = o;   o += al256((size_t)2 * kMaxTiles * 4);
    L.tinfo = o;     o += al256((size_t)kMaxTiles * 4 * 8);
    L.Hcarry = o;    o += al256((size_t)kMaxTiles * 4);
    L.H = o;         o += al256(N * 4);
    L.xs = o;        o += al256(N * 8);
    L.es = o;        o += al256(N * 8);
    L.perm = o;      o += al256(N * 4);
    L.mail = o;      o += al256(N * 16);
    L.R = o;         o += al256(2 * N * 32);
    L.P = o;         o += al256((size_t)L.RP * N * 32);
    L.psum = o;      o += al256((size_t)nobs * G * 8 * 8);
    L.shiftv = o;    o += al256((size_t)nobs * 8);
    L.xminv = o;     o += al256((size_t)nobs * 8);
    L.shring = o;    o += al256((size_t)lag * N * 8);
    L.parentpos = o; o += al256(hist ? N * 4 : 256);
    L.sums = o;      o += al256((size_t)nobs * 8 * 8);
    L.tailpart = o;  o += al256((size_t)lag * L.nblk * 8 * 8);
    L.tail = o;      o += al256((size_t)lag * 8 * 8);
    L.info = o;      o += al256(8 * 8);
    L.total = o;
    (void)G;
    return L;
}

#define GRID_CUDA(call)                                                  \
    do {                                                                 \
        cudaError_t e__ = (call);                                        \
        if (e__ != cudaSuccess) return pmmh::set_cuda_error(e__, #call); \
    } while (0)

}  // namespace

int sv_grid_ctas(int n, int sm_count, int ctas) {
    int G = ctas > 0 ? ctas : (n + 2047) / 2048;
    if (G > sm_count) G = sm_count;
    if (G > kMaxTiles) G = kMaxTiles;
    if (G < 1) G = 1;
    return G;
}

// a tile holds N/G particles +- one histogram bin; 12 % head room below the shared-memory capacity
bool sv_grid_eligible(int nobs, int n, int lag, int G) {
    if (G < 1 || G > kMaxTiles) return false;
    if (lag < 2 || lag > 10 || nobs < 2 * lag || n < 32) return false;
    const int Wc = (n + G - 1) / G;
    return Wc <= kCap - kCap / 8;
}

size_t sv_grid_ws_bytes(int nobs, int n, int lag, int G, int hist) {
    return grid_layout(nobs, n, lag, G, hist).total;
}

int sv_grid_run(const double* d_obs, const double* d_params, const double* d_rvr, const double* d_u, int nobs,
                int n, int lag, int G, double* d_filt, double* d_smo, double* d_ll, double* d_grad, double* d_traj,
                long long* d_diag, double* d_xh, int* d_ah, void* d_ws, size_t ws_bytes, long long* d_prof,
                cudaStream_t st) {
    if (!sv_grid_eligible(nobs, n, lag, G)) return set_error(PMMH_ERR_INVALID, "grid kernel: sizes not eligible");
    const int hist = d_xh != nullptr;
    const GridLayout L = grid_layout(nobs, n, lag, G, hist);
    if (ws_bytes < L.total) return set_error(PMMH_ERR_WORKSPACE, "grid kernel: workspace too small");
    char* ws = (char*)d_ws;
    GridArgs a;
    memset(&a, 0, sizeof(a));
    a.N = n;
    a.NOBS = nobs;
    a.LAG = lag;
    a.G = G;
    a.Wc = (n + G - 1) / G;
    a.RP = L.RP;
    a.hist = hist;
    a.obs = d_obs;
    a.params = d_params;
    a.rvr = d_rvr;
    a.U = d_u;
    a.ctrl = (GridCtrl*)(ws + L.ctrl);
    a.ghist = (int*)(ws + L.ghist);
    a.tilecnt = (int*)(ws + L.tilecnt);
    a.tinfo = (double*)(ws + L.tinfo);
    a.H = (int*)(ws + L.H);
    a.Hcarry = (int*)(ws + L.Hcarry);
    a.xs = (double*)(ws + L.xs);
    a.es = (double*)(ws + L.es);
    a.perm = (int*)(ws + L.perm);
    a.mail = (MailEntry*)(ws + L.mail);
    a.R = (REntry*)(ws + L.R);
    a.P = (PEntry*)(ws + L.P);
    a.psum = (double*)(ws + L.psum);
    a.shiftv = (double*)(ws + L.shiftv);
    a.xminv = (double*)(ws + L.xminv);
    a.shring = (double*)(ws + L.shring);
    a.parentpos = (int*)(ws + L.parentpos);
    a.Xhist = d_xh;
    a.Ahist = d_ah;
    a.prof = d_prof;
    // control block, histograms, reservation counters, tile info: zero; head markers: -1
    GRID_CUDA(cudaMemsetAsync(ws + L.ctrl, 0, L.Hcarry - L.ctrl, st));
    GRID_CUDA(cudaMemsetAsync(ws + L.Hcarry, 0xff, (L.xs - L.Hcarry), st));
    static thread_local bool attr_set[64] = {false};
    int dev = 0;
    GRID_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        GRID_CUDA(cudaFuncSetAttribute(sv_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem));
        attr_set[dev] = true;
    }
    void* kargs[] = {(void*)&a};
    GRID_CUDA(cudaLaunchCooperativeKernel((void*)sv_grid_kernel, dim3(G), dim3(kGT), kargs, kDynSmem, st));
    double* sums = (double*)(ws + L.sums);
    double* tailpart = (double*)(ws + L.tailpart);
    double* tail = (double*)(ws + L.tail);
    grid_reduce_kernel<<<nobs, 256, 0, st>>>(a.psum, G, sums);
    grid_tail_kernel<<<dim3(L.nblk, lag), 256, 0, st>>>(a, sums, tailpart, L.nblk);
    grid_tail_reduce_kernel<<<lag, 256, 0, st>>>(tailpart, L.nblk, tail);
    grid_finish_kernel<<<(nobs + 255) / 256, 256, 0, st>>>(a.ctrl, sums, a.shiftv, a.xminv, tail, d_params, nobs, lag,
                                                           (double)n, d_ll, d_filt, d_smo, d_grad, d_traj, d_diag,
                                                           (long long*)(ws + L.info));
    GRID_CUDA(cudaGetLastError());
    return PMMH_OK;
}

// soft ties / raw status of the last evaluation that used this workspace (after synchronisation)
int sv_grid_read_info(const void* d_ws, int nobs, int n, int lag, int G, int hist, long long* h_info) {
    const GridLayout L = grid_layout(nobs, n, lag, G, hist);
    GRID_CUDA(cudaMemcpy(h_info, (const char*)d_ws + L.info, 2 * sizeof(long long), cudaMemcpyDeviceToHost));
    return PMMH_OK;
}

}  // namespace pmmh
