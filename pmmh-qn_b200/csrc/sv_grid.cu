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
#include <stdlib.h>
#include <string.h>

#include "../../include/pmmh_qn.h"
#include "common.cuh"
#include "sv_grid.cuh"
#include "sv_math.cuh"

namespace pmmh {

int set_error(int code, const char* what);           // capi.cu
int set_cuda_error(cudaError_t err, const char* where);

namespace {

constexpr int kCap = 8192;         // entries of one tile (shared-memory capacity)
constexpr int kNF = 8192;          // bins of the global value histogram
constexpr int kNSB = 8192;         // sub-bins of the in-tile counting sort
constexpr int kMaxSub = 1024;      // a sub-bin larger than this abandons the evaluation
constexpr int kMaxTiles = 160;     // >= SM count
constexpr double kZ = 6.5;         // histogram range: predicted mean +- 6.5 predicted sd
constexpr int kDynSmem = 192 * 1024;
constexpr int kProf = 16;

struct __align__(16) MailEntry {   // aliases one (x, exp(-x/2)) pair of the sorted generation
    double x;
    int j, a;
};
struct __align__(16) PEntry {
    double n, c;           // value, parent value
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
    int dbg;   // development (timing only, results wrong): 1 skip R records, 2 skip P store, 4 skip score gather
    const double *obs, *params, *rvr, *U;
    GridCtrl* ctrl;
    int* ghist;        // [2][kNF]
    int* tilecnt;      // [2][kMaxTiles]
    double* tinfo;     // [kMaxTiles][4]  tot, n, sum sh m, sum sh m^2
    int* H;            // [N] head markers (-1 = none)
    int* Hcarry;       // [kMaxTiles]
    double2* XE;       // [N] sorted generation: (x, exp(-x/2)); the mailbox of the next generation aliases it
    int* perm;         // [N] sorted position -> birth row
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
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void prefetch_l2_keep(const void* p) {
    asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p));
}
__device__ __forceinline__ unsigned long long policy_evict_last() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// 32-byte genealogy records: one 256-bit access, kept in L2 (evict_last)
__device__ __forceinline__ void ld_rec(const REntry* p, unsigned long long pol, int (&r)[8]) {
    asm volatile("ld.global.cg.L2::cache_hint.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p), "l"(pol));
}
__device__ __forceinline__ void st_rec(REntry* p, unsigned long long pol, int a0, int a1, int a2, int a3, int a4,
                                       int a5, int a6, int a7) {
    asm volatile("st.global.cg.L2::cache_hint.v8.s32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8}, %9;" ::"l"(p), "r"(a0),
                 "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5), "r"(a6), "r"(a7), "l"(pol)
                 : "memory");
}
// streaming 16-byte store that should leave L2 first (written once, read 8 steps later or never)
__device__ __forceinline__ void st_stream_f64x2(void* p, unsigned long long pol, double a, double b) {
    asm volatile("st.global.cs.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" ::"l"(p), "d"(a), "d"(b), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ double ld_stream_hint_f64(const double* p, unsigned long long pol) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ int warp_incl_max(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(kFullMask, v, d);
        if (lane >= d) v = max(v, o);
    }
    return v;
}
__device__ __forceinline__ int pick8(const int (&r)[8], int idx) {
    switch (idx) {
        case 0: return r[0];
        case 1: return r[1];
        case 2: return r[2];
        case 3: return r[3];
        case 4: return r[4];
        case 5: return r[5];
        case 6: return r[6];
        default: return r[7];
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
    const double fr = e - floor(e);
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

// Block-wide exclusive scans over one value per thread.  s_w: shared [32].  Every warp scans the
// warp totals itself (shuffles), so there are two barriers and no serial loop; the order of the
// additions is fixed (deterministic).  s_w may be reused right after the call.
template <int GT>
__device__ __forceinline__ int block_excl_scan_int(int v, int* s_w, int& total, int lane, int warp) {
    constexpr int NW = GT / 32;
    const int incl = warp_incl_scan(v, lane);
    __syncthreads();
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    const int wt = (lane < NW) ? s_w[lane] : 0;
    const int wincl = warp_incl_scan(wt, lane);
    total = __shfl_sync(kFullMask, wincl, 31);
    const int woff = __shfl_sync(kFullMask, wincl - wt, warp);
    return woff + incl - v;
}
template <int GT>
__device__ __forceinline__ int block_excl_max_int(int v, int init, int* s_w, int lane, int warp) {
    constexpr int NW = GT / 32;
    const int incl = warp_incl_max(v, lane);
    __syncthreads();
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    const int wt = (lane < NW) ? s_w[lane] : init;
    const int wincl = warp_incl_max(wt, lane);
    int wex = __shfl_up_sync(kFullMask, wincl, 1);
    if (lane == 0) wex = init;
    const int woff = max(init, __shfl_sync(kFullMask, wex, warp));
    int ex = __shfl_up_sync(kFullMask, incl, 1);
    if (lane == 0) ex = init;
    return max(woff, ex);
}
template <int GT>
__device__ __forceinline__ double block_excl_scan_f64(double v, double* s_w, int lane, int warp) {
    constexpr int NW = GT / 32;
    const double incl = warp_incl_scan(v, lane);
    __syncthreads();
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    const double wt = (lane < NW) ? s_w[lane] : 0.0;
    const double wincl = warp_incl_scan(wt, lane);
    double wex = __shfl_up_sync(kFullMask, wincl, 1);
    if (lane == 0) wex = 0.0;
    const double woff = __shfl_sync(kFullMask, wex, warp);
    double ex = __shfl_up_sync(kFullMask, incl, 1);
    if (lane == 0) ex = 0.0;
    return woff + ex;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int GT>
__global__ void __launch_bounds__(GT, 1) sv_grid_kernel(const GridArgs a) {
    constexpr int KPT = kCap / GT;          // entries per thread (strided assignment)
    constexpr int KCH = KPT | 1;            // longest chunk of the thread-contiguous passes (odd)
    constexpr int BPT = kNF / GT;           // histogram bins per thread
    static_assert(kNF == kNSB, "one scan shape for both histograms");
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
    const unsigned long long pol_keep = policy_evict_last(), pol_stream = policy_evict_first();
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
    for (int i = tid; i < kMaxTiles; i += GT) s_tcnt[i] = 0;
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
        for (int q = tid; q < n; q += GT) {
            __stcg(&a.XE[pstart + q], make_double2(mu, e0));
            __stcg(&a.perm[pstart + q], pstart + q);
            s_sh[q] = 1.0;
            st_rec(&a.R[pstart + q], pol_keep, 0, 0, 0, 0, 0, 0, 0, 0);
            if (a.hist) {
                a.Xhist[pstart + q] = mu;
                a.Ahist[pstart + q] = pstart + q;
            }
        }
        const int Lc = ((n + GT - 1) / GT) | 1;
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
        const double ur = a.rvr[t];
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
            __syncwarp();
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
                // child range end of the tiles in front of this one (running maximum, see below)
                int carry = 0;
                double fr;
                for (int k = max(0, c - 2); k < c; ++k)
                    carry = max(carry, count_le((s_off[k] + s_tot[k]) * invS, ur, N, dn, inv_n, pow2, fr));
                s_sc.carry = carry;
            }
        }
        __syncthreads();
        {
            const double invS = s_sc.invS, offk = s_off[c];
            const int Lc = ((n + GT - 1) / GT) | 1, q0 = tid * Lc;
            const double tol_soft = 2.220446049250313e-16 * dn * (4.0 + 2.0 * sqrt(dn));
            const double tol_near = 64.0 * 2.220446049250313e-16 * dn;
            int ubv[KCH];
            int rmax = 0;
            double run = 0.0;
#pragma unroll
            for (int kk = 0; kk < KCH; ++kk) {
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
            int prev = block_excl_max_int<GT>(rmax, s_sc.carry, s_wi, lane, warp);
            const float inv_wc = 1.0f / (float)Wc;
#pragma unroll
            for (int kk = 0; kk < KCH; ++kk) {
                const int q = q0 + kk;
                if (kk < Lc && q < n) {
                    const int ub = max(ubv[kk], prev);
                    if (ub > prev) {
                        const int P = pstart + q;
                        __stcg(&a.H[prev], P);
                        if (ub - prev > 1) {
                            // child-tile boundaries m * Wc strictly inside (prev, ub): the tile's
                            // first child has no marker of its own
                            int m = (int)((float)prev * inv_wc);
                            while (m * Wc > prev) --m;
                            while ((m + 1) * Wc <= prev) ++m;
                            ++m;
                            while (m < G && m * Wc < ub) {
                                __stcg(&a.Hcarry[m], P);
                                ++m;
                            }
                        }
                    }
                    prev = ub;
                }
            }
        }
        PROF_MARK(0);
        GRID_ARRIVE();   // ---- barrier 4: head markers complete
        for (int b = tid; b < kNF; b += GT) s_fhist[b] = 0;
        PROF_MARK(1);
        GRID_WAIT();
        PROF_MARK(2);
        if (s_sc.abort_now) break;

        // --------------------------------------------------------------------------------------
        // phase A: ancestors of this CTA's children, propagation (:354-358), value histogram
        // --------------------------------------------------------------------------------------
        {
            int hv[KPT];
#pragma unroll
            for (int kk = 0; kk < KPT; ++kk) {
                const int i = kk * GT + tid;
                hv[kk] = -1;
                if (i < nc) hv[kk] = __ldcg(&a.H[jb + i]);
            }
            if (tid == 0) s_sc.hc = __ldcg(&a.Hcarry[c]);
#pragma unroll
            for (int kk = 0; kk < KPT; ++kk) {
                const int i = kk * GT + tid;
                if (i < nc) {
                    s_par[i] = hv[kk];
                    if (hv[kk] >= 0) __stcg(&a.H[jb + i], -1);
                }
            }
            if (tid == 0) __stcg(&a.Hcarry[c], -1);
        }
        __syncthreads();
        {
            const int Lc2 = ((nc + GT - 1) / GT) | 1, i0 = tid * Lc2;
            int mx = -1;
#pragma unroll
            for (int kk = 0; kk < KCH; ++kk) {
                const int i = i0 + kk;
                if (kk < Lc2 && i < nc) mx = max(mx, s_par[i]);
            }
            int run = block_excl_max_int<GT>(mx, s_sc.hc, s_wi, lane, warp);
            bool orphan = false;
#pragma unroll
            for (int kk = 0; kk < KCH; ++kk) {
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
        double xn[KPT];
        int bp[KPT];
        {
            const double y1 = a.obs[t - 1];
            const double mhat = s_sc.mhat, inv_shat = s_sc.inv_shat;
            const double mu = s_k.mu, phi = s_k.phi, sr = s_k.sr, sd = s_k.sd;
            const double* Ut = a.U + (size_t)t * N;
            PEntry* Pt = a.P + (size_t)(t % RP) * N;
            const REntry* Rp = a.R + (size_t)((t - 1) & 1) * N;
            double vmin = INFINITY, vmax = -INFINITY;
            bool bad = false;
#pragma unroll
            for (int k0 = 0; k0 < KPT; k0 += 4) {
                // all loads of four children are in flight before the first one is used
                double2 xe[4];
                double uu[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = (k0 + u) * GT + tid;
                    xe[u] = make_double2(0.0, 0.0);
                    uu[u] = 0.0;
                    bp[k0 + u] = 0;
                    if (i < nc) {
                        const int p = s_par[i];
                        xe[u] = __ldcg(&a.XE[p]);
                        bp[k0 + u] = __ldcg(&a.perm[p]);
                        uu[u] = ld_stream_hint_f64(Ut + jb + i, pol_stream);
                        if (a.hist) a.parentpos[jb + i] = p;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = (k0 + u) * GT + tid;
                    xn[k0 + u] = 0.0;
                    if (i < nc) {
                        double mean = mu + phi * (xe[u].x - mu);     // :355
                        mean += (sr * xe[u].y) * y1;                  // :356
                        const double x = mean + sd * uu[u];           // :357-358
                        if (!isfinite(x)) bad = true;
                        xn[k0 + u] = x;
                        if (!(a.dbg & 1)) prefetch_l2_keep(&Rp[bp[k0 + u]]);
                        atomicAdd(&s_fhist[fine_bin(x, mhat, inv_shat)], 1);
                        vmin = fmin(vmin, x);
                        vmax = fmax(vmax, x);
                        if (!(a.dbg & 2)) st_stream_f64x2(&Pt[jb + i], pol_stream, x, xe[u].x);
                    }
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
#pragma unroll
            for (int kk = 0; kk < BPT; ++kk) {
                const int b = kk * GT + tid;
                const int cnt = s_fhist[b];
                if (cnt) atomicAdd(&gh[b], cnt);
            }
            if (warp == 0) {
                const double vmin = warp_min(lane < GT / 32 ? s_red[lane] : INFINITY);
                const double vmax = warp_max(lane < GT / 32 ? s_red[32 + lane] : -INFINITY);
                if (lane == 0 && nc > 0) {
                    atomicMin(&ctrl->mn[par], enc_f64(vmin));
                    atomicMax(&ctrl->mx[par], enc_f64(vmax));
                }
            }
        }
        PROF_MARK(3);
        GRID_ARRIVE();   // ---- barrier 1: global histogram complete
        if (!(a.dbg & 1)) {
            // genealogy records (only feed outputs): child = (parent row, parent's ancestors 1..7)
            const REntry* Rp = a.R + (size_t)((t - 1) & 1) * N;
            REntry* Rc = a.R + (size_t)(t & 1) * N;
            const PEntry* Pnow = a.P + (size_t)((t - (L - 2) + RP) % RP) * N;
#pragma unroll
            for (int k0 = 0; k0 < KPT; k0 += 2) {
                int r[2][8];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int i = (k0 + u) * GT + tid;
#pragma unroll
                    for (int z = 0; z < 8; ++z) r[u][z] = 0;
                    if (i < nc) ld_rec(&Rp[bp[k0 + u]], pol_keep, r[u]);
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int i = (k0 + u) * GT + tid;
                    if (i < nc) {
                        const int j = jb + i;
                        const int b = bp[k0 + u];
                        st_rec(&Rc[j], pol_keep, b, r[u][0], r[u][1], r[u][2], r[u][3], r[u][4], r[u][5], r[u][6]);
                        // row of the ancestor L-2 steps back (new record = (b, r[0..6]))
                        int anc = j;
                        if (L == 3) anc = b;
                        else if (L > 3) anc = pick8(r[u], L - 4);
                        bp[k0 + u] = anc;
                        if (t >= L && !(a.dbg & 4)) prefetch_l2(&Pnow[min(max(anc, 0), N - 1)]);
                    }
                }
            }
        }
        {
            // housekeeping for the next step
            const int zper = (kNF + G - 1) / G;
            int* ghn = a.ghist + (par ^ 1) * kNF;
            for (int b = c * zper + tid; b < min(kNF, (c + 1) * zper); b += GT) __stcg(&ghn[b], 0);
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
            int cnt[BPT];
#pragma unroll
            for (int i = 0; i < BPT; i += 4) {
                const int4 v = __ldcg((const int4*)(gh + BPT * tid + i));
                cnt[i] = v.x;
                cnt[i + 1] = v.y;
                cnt[i + 2] = v.z;
                cnt[i + 3] = v.w;
            }
            const int prevcnt = tid > 0 ? __ldcg(gh + BPT * tid - 1) : 0;
            int loc = 0, mxb = 0;
#pragma unroll
            for (int i = 0; i < BPT; ++i) {
                loc += cnt[i];
                mxb = max(mxb, cnt[i]);
            }
            my_max_bin = max(my_max_bin, mxb);
            int total;
            int start = block_excl_scan_int<GT>(loc, s_wi, total, lane, warp);
            int tprev = (tid == 0) ? -1 : min(G - 1, (start - prevcnt) / Wc);
            int tl = min(G - 1, start / Wc);
#pragma unroll
            for (int i = 0; i < BPT; ++i) {
                while (tl < G - 1 && start >= (tl + 1) * Wc) ++tl;
                s_tileof[BPT * tid + i] = (unsigned short)tl;
                for (int k = tprev + 1; k <= tl; ++k) s_tstart[k] = start;
                tprev = tl;
                start += cnt[i];
            }
            if (tid == GT - 1) {
                for (int k = tprev + 1; k <= G; ++k) s_tstart[k] = N;
                if (total != N) GRID_FLAG(5);
            }
            if (tid == 0) {
                s_sc.xmin = dec_f64(*(volatile unsigned long long*)&ctrl->mn[par]);
                s_sc.xmax = dec_f64(*(volatile unsigned long long*)&ctrl->mx[par]);
            }
        }
        __syncthreads();
        int kr[KPT];
        {
            const double mhat = s_sc.mhat, inv_shat = s_sc.inv_shat;
#pragma unroll
            for (int kk = 0; kk < KPT; ++kk) {
                const int i = kk * GT + tid;
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
        {
            MailEntry* mail = (MailEntry*)a.XE;
#pragma unroll
            for (int kk = 0; kk < KPT; ++kk) {
                const int i = kk * GT + tid;
                if (i < nc) {
                    const int pos = s_tbase[kr[kk] >> 16] + (kr[kk] & 0xffff);
                    const long long xb = __double_as_longlong(xn[kk]);
                    const int4 ent = make_int4((int)(xb & 0xffffffffll), (int)(xb >> 32), jb + i, bp[kk]);
                    if (pos >= 0 && pos < N) __stcg((int4*)&mail[pos], ent);
                }
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
        for (int b = tid; b < kNSB; b += GT) s_sub[b] = 0;
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
                    prefetch_l2_bulk(q, (unsigned)min((long long)16384, (long long)(q1 - q)));
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
            const MailEntry* mb = (const MailEntry*)a.XE + pstart;
            double ex[KPT];
            int ej[KPT], ea[KPT], er[KPT];
            double lmin = INFINITY, lmax = -INFINITY;
#pragma unroll
            for (int kk = 0; kk < KPT; ++kk) {
                const int e = kk * GT + tid;
                ex[kk] = 0.0;
                ej[kk] = ea[kk] = 0;
                if (e < n) {
                    const int4 raw = __ldcg((const int4*)(mb + e));
                    ex[kk] = __longlong_as_double(((long long)raw.y << 32) | (long long)(unsigned)raw.x);
                    ej[kk] = raw.z;
                    ea[kk] = raw.w;
                }
            }
#pragma unroll
            for (int kk = 0; kk < KPT; ++kk) {
                const int e = kk * GT + tid;
                if (e < n) {
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
                const double lo = warp_min(lane < GT / 32 ? s_red[lane] : INFINITY);
                const double hi = warp_max(lane < GT / 32 ? s_red[32 + lane] : -INFINITY);
                if (lane == 0) {
                    s_sc.lo = lo;
                    s_sc.hi = hi;
                }
            }
            __syncthreads();
            const double lo = s_sc.lo;
            const double scale = (s_sc.hi > lo) ? (double)kNSB / (s_sc.hi - lo) : 0.0;
#pragma unroll
            for (int kk = 0; kk < KPT; ++kk) {
                const int e = kk * GT + tid;
                er[kk] = 0;
                if (e < n) er[kk] = atomicAdd(&s_sub[sub_bin(ex[kk], lo, scale)], 1);
            }
            __syncthreads();
            {
                // exclusive scan of the sub-bin counters in place (BPT consecutive bins per thread)
                int cnt[BPT];
#pragma unroll
                for (int i = 0; i < BPT; i += 4) {
                    const int4 v = *(const int4*)(s_sub + BPT * tid + i);
                    cnt[i] = v.x;
                    cnt[i + 1] = v.y;
                    cnt[i + 2] = v.z;
                    cnt[i + 3] = v.w;
                }
                int loc = 0, mxb = 0;
#pragma unroll
                for (int i = 0; i < BPT; ++i) {
                    loc += cnt[i];
                    mxb = max(mxb, cnt[i]);
                }
                if (mxb > kMaxSub) GRID_FLAG(3);
                int total;
                int start = block_excl_scan_int<GT>(loc, s_wi, total, lane, warp);
#pragma unroll
                for (int i = 0; i < BPT; ++i) {
                    const int cn = cnt[i];
                    cnt[i] = start;
                    start += cn;
                }
#pragma unroll
                for (int i = 0; i < BPT; i += 4)
                    *(int4*)(s_sub + BPT * tid + i) = make_int4(cnt[i], cnt[i + 1], cnt[i + 2], cnt[i + 3]);
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < KPT; ++kk) {
                const int e = kk * GT + tid;
                if (e < n) {
                    const int pos = s_sub[sub_bin(ex[kk], lo, scale)] + er[kk];
                    s_x[pos] = ex[kk];
                    s_j[pos] = ej[kk];
                    s_a[pos] = ea[kk];
                }
            }
            __syncthreads();
            // exact order inside a sub-bin: by value, then by birth row (:32-35 never returns 0)
            int np[KPT];
#pragma unroll
            for (int kk = 0; kk < KPT; ++kk) {
                const int q = kk * GT + tid;
                np[kk] = q;
                ex[kk] = 0.0;
                if (q < n) {
                    const double x = s_x[q];
                    ex[kk] = x;
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
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < KPT; ++kk) {
                const int q = kk * GT + tid;
                if (q < n) {
                    s_x[np[kk]] = ex[kk];
                    ej[kk] = s_j[q];
                }
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < KPT; ++kk) {
                const int q = kk * GT + tid;
                if (q < n) {
                    s_j[np[kk]] = ej[kk];
                    ea[kk] = s_a[q];
                }
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < KPT; ++kk) {
                const int q = kk * GT + tid;
                if (q < n) s_a[np[kk]] = ea[kk];
            }
            __syncthreads();
        }
        PROF_MARK(9);
        // birth rows of the new sorted generation (the values follow with exp(-x/2), below)
#pragma unroll
        for (int kk = 0; kk < KPT; ++kk) {
            const int q = kk * GT + tid;
            if (q < n) {
                const int j = s_j[q];
                __stcg(&a.perm[pstart + q], j);
                if (a.hist) {
                    a.Xhist[(size_t)t * N + pstart + q] = s_x[q];
                    a.Ahist[(size_t)t * N + pstart + q] = __ldcg(&a.parentpos[j]);
                }
            }
        }
        {
            // weights (:427-437): lw = -0.9189 - x/2 - y^2 exp(-x) / 2, sh = exp(lw - shift);
            // thread = Lc consecutive sorted particles, sequential running sum
            const double y = a.obs[t], hy2 = 0.5 * y * y, shift = s_sc.shift;
            const double mu = s_k.mu, phi = s_k.phi, sr = s_k.sr;
            const int Lc = ((n + GT - 1) / GT) | 1, q0 = tid * Lc;
            double run = 0.0;
            double acc[3] = {0.0, 0.0, 0.0};
            bool bad = false;
#pragma unroll
            for (int kk = 0; kk < KCH; ++kk) {
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
                    // (x, e) of the new generation: this thread's chunk is contiguous in memory
                    __stcg(&a.XE[pstart + q], make_double2(x, e));
                    s_sh[q] = sh;
                    run = run + sh;
                    acc[0] += sh * x;
                    double m = mu + phi * (x - mu);
                    m += (sr * e) * y;
                    acc[1] += sh * m;
                    acc[2] += sh * (m * m);
                }
            }
            if (bad) GRID_FLAG(2);
            toff = block_excl_scan_f64<GT>(run, s_wd, lane, warp);
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
            // fixed-lag score terms (:445-470): ancestor pair (time t-L+1, t-L+2) from one half sector
            const int Lc = ((n + GT - 1) / GT) | 1, q0 = tid * Lc;
            double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
            if (t >= L && !(a.dbg & 4)) {
                const double ylag = a.obs[t - L];   // Q5: obs[i - LAG]
                const PEntry* Pg = a.P + (size_t)((t - (L - 2)) % RP) * N;
#pragma unroll
                for (int k0 = 0; k0 < KCH; k0 += 3) {
                    double2 pe[3];
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const int kk = k0 + u, q = q0 + kk;
                        pe[u] = make_double2(0.0, 0.0);
                        if (kk < KCH && kk < Lc && q < n) {
                            const int an = min(max(s_a[q], 0), N - 1);
                            pe[u] = __ldcg((const double2*)&Pg[an]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const int kk = k0 + u, q = q0 + kk;
                        if (kk < KCH && kk < Lc && q < n) {
                            const double sh = s_sh[q];
                            double sq, g[4];
                            sv_score_main_e(s_k, pe[u].y, exp(-0.5 * pe[u].y), pe[u].x, ylag, sq, g);
                            acc[0] += sh * pe[u].y;
#pragma unroll
                            for (int i = 0; i < 4; ++i) acc[1 + i] += g[i] * sh;
                        }
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
            if (t >= NOBS - L) {
#pragma unroll
                for (int kk = 0; kk < KPT; ++kk) {
                    const int q = kk * GT + tid;
                    if (q < n) a.shring[(size_t)(t % L) * N + pstart + q] = s_sh[q];
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
            curr = a.XE[p].x;
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
            sv_score_tail_e(k, curr, exp(-0.5 * curr), pe.n, y1, sq, g);
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
    size_t ctrl, ghist, tilecnt, tinfo, H, Hcarry, XE, perm, R, P, psum, shiftv, xminv, shring, parentpos,
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
    L.XE = o;        o += al256(N * 16);
    L.perm = o;      o += al256(N * 4);
    L.R = o;         o += al256(2 * N * 32);
    L.P = o;         o += al256((size_t)L.RP * N * 16);
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
    a.XE = (double2*)(ws + L.XE);
    a.perm = (int*)(ws + L.perm);
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
    {
        const char* e = getenv("PMMH_GRID_DEBUG");
        a.dbg = e ? atoi(e) : 0;
    }
    // control block, histograms, reservation counters, tile info: zero; head markers: -1
    GRID_CUDA(cudaMemsetAsync(ws + L.ctrl, 0, L.Hcarry - L.ctrl, st));
    GRID_CUDA(cudaMemsetAsync(ws + L.Hcarry, 0xff, (L.XE - L.Hcarry), st));
    static thread_local bool attr_set[64] = {false};
    static int threads = 0;
    if (!threads) {
        const char* e = getenv("PMMH_GRID_THREADS");
        threads = (e && atoi(e) == 1024) ? 1024 : 512;
    }
    int dev = 0;
    GRID_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        GRID_CUDA(cudaFuncSetAttribute(sv_grid_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem));
        GRID_CUDA(cudaFuncSetAttribute(sv_grid_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem));
        attr_set[dev] = true;
    }
    void* kargs[] = {(void*)&a};
    if (threads == 1024)
        GRID_CUDA(cudaLaunchCooperativeKernel((void*)sv_grid_kernel<1024>, dim3(G), dim3(1024), kargs, kDynSmem, st));
    else
        GRID_CUDA(cudaLaunchCooperativeKernel((void*)sv_grid_kernel<512>, dim3(G), dim3(512), kargs, kDynSmem, st));
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
