// sv_grid.cu -- the "grid kernel": ONE stochastic-volatility fixed-lag smoother evaluation
// (log-likelihood + gradient) as ONE persistent cooperative launch over all SMs.
//
// Restates flps_sv_corr (/root/reference/python/state/particle_methods/stochastic_volatility.pyx
// :205-655: correlated systematic resampling :694-715, propagation :354-358, argsort :392-424 /
// :23-52, weights :427-442, fixed-lag score terms :445-470, tail :540-562, log-likelihood :537,
// trajectory :630-633 with quirks Q10/Q11).  B200 design, 148 CTAs, one CTA per SM:
//
//   * the sorted generation is cut into G TILES of ~N/G particles (7 085 at N = 2^20, G = 148);
//     CTA c owns tile c while it is sorted and weighted (shared memory) and owns the CHILDREN
//     [c*Wc, (c+1)*Wc) while they are generated (equal work for every CTA whatever the weights
//     look like);
//   * four grid barriers per time step (one atomic counter, arrive / wait split so that work that
//     only feeds outputs sits between the two):
//       C  owner of a tile: cumulative weights -> PARENT-side child ranges in closed form
//          ub(p) = #{j : (u + j)/N <= cum(p)} (exact predicate re-checked), ancestor of every
//          child written to H (coalesced fill)                                     | barrier 4
//       A1 owner of a child range: parent (x, exp(-x/2), birth row), propagation, 8192-bin value
//          histogram (shared-memory atomics, merged into striped global copies)    | barrier 1
//          (after the arrive: genealogy records)
//       A2 scan of the global histogram -> tile boundaries on bin edges (every tile gets N/G
//          particles +- one bin), slot reservation per (CTA, tile), entries ordered by tile in
//          shared memory and copied out in runs (value, birth row, lagged ancestor row)
//                                                                                  | barrier 2
//       B  owner of a tile: counting sort over 4096 sub-bins of the tile's value range + exact
//          in-bin ranking by (value, birth row) = the reference's argsort; weights, fixed-lag
//          score terms (one gather from a generation that was bulk-prefetched into L2), sorted
//          generation written out, block scan of the weights, moments             | barrier 3
//   * genealogy: a generation is stored ONCE in birth order as P[t][j] = (value, parent value)
//     and R[t][j] = birth rows of the ancestors 1..8 steps back (one 32-byte sector each).  The
//     fixed-lag terms of step t need one random sector of P[t-lag+2]; a child copies its parent's
//     R with one random sector read.  No history is ever moved.  Both tables are older than the
//     L2 working set when they are needed again, so they are brought back with sequential bulk
//     prefetches (cp.async.bulk.prefetch.L2) a phase ahead of the random reads.
//   * one exp per weight, one exp per particle shared by the weight and the propagation mean of
//     the children (exp(-x/2) is stored next to x), one exp per score term.
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

constexpr int kGT = 1024;          // threads of a CTA = main warps + helper warps
constexpr int kCap = 8192;         // entries of one tile (shared-memory capacity)
constexpr int kNFMax = 12288;      // bins of the global value histogram (upper bound over the variants)
constexpr int kMaxSub = 1024;      // a sub-bin larger than this abandons the evaluation
constexpr int kMaxTiles = 160;     // >= SM count
constexpr int kCntStride = 32;     // ints between two slot counters (one 128-byte line each)
constexpr double kZ = 6.5;         // histogram range: predicted mean +- 6.5 predicted sd
constexpr int kDynSmem = 192 * 1024;
constexpr int kProf = 16;

struct __align__(16) MailEntry {   // aliases one (x, exp(-x/2)) pair of the sorted generation
    double x;
    int j, pad;
};
struct GridCtrl {
    unsigned bar;                  // arrival counter of the grid barrier of the main warps (monotone)
    int status;                    // 0, or (reason << 24) | first barrier index at which everybody stops
    int max_bin;
    unsigned hbar;                 // arrival counter of the helper warps (one arrival per CTA and time step)
    unsigned long long near_ties, soft_ties, key_ties;
};

struct GridArgs {
    int N, NOBS, LAG, G, Wc, RP, hist;
    int dbg;   // development (timing only, results wrong): 1 skip the helper's work, 4 skip the score terms
    const double *obs, *params, *rvr, *U;
    GridCtrl* ctrl;
    int* ghist;        // [2][kNFMax]
    int* tilecnt;      // [2][kMaxTiles * kCntStride]
    double* tinfo;     // [kMaxTiles][4]  tot, n, sum sh m, sum sh m^2
    int* H;            // [N] ancestor (sorted position) of every child
    double2* XE;       // [N] sorted generation: (x, exp(-x/2)); the mailbox of the next generation aliases it
    int* perm;         // [N] sorted position -> birth row
    int* BP;           // [kRingBP][N] birth row of the parent of row j of generation t % kRingBP (main -> helper warps)
    int *J2, *J4;      // [kRingJ][N] birth rows of the ancestors 2 / 4 generations back
    double2* Q;        // [2][N] lagged pair of row j of generation t (helper -> main warps), by step parity
    double2* P;        // [RP][N] (parent value, own value) of generation t % RP in birth order
    double* psum;      // [NOBS][G][8]
    double *shiftv, *xminv;   // [NOBS]
    double* shring;    // [LAG][N] sh of the last LAG generations (sorted order)
    int* parentpos;    // [N] (history dump only)
    double* Xhist;
    int* Ahist;
    long long* prof;
};

struct StepScalars {
    double S, invS, mhat, shat, inv_shat, shift, lo, scale, tot, offk;
    int carry, abort_now, binlo, binhi;
};

// ---- small device helpers ---------------------------------------------------------------------
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// [p0, p1) -> L2 in 16 KB pieces (sequential HBM reads issued by the copy unit, no SM cycles);
// the pieces are dealt to the callers round-robin: piece k goes to caller k % nlanes == lane
__device__ __forceinline__ void prefetch_range(const void* b0, const void* b1, int lane, int nlanes) {
    const char* q0 = (const char*)(((uintptr_t)b0 + 15) & ~(uintptr_t)15);
    const char* q1 = (const char*)((uintptr_t)b1 & ~(uintptr_t)15);
    for (const char* q = q0 + (size_t)lane * 16384; q < q1; q += (size_t)nlanes * 16384)
        prefetch_l2_bulk(q, (unsigned)min((long long)16384, (long long)(q1 - q)));
}
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
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
// the resampling point of child j, (u + j) / N as the reference computes it (:703)
__device__ __forceinline__ double cpoint(double u, int j, double dn, double inv_n, bool pow2) {
    const double s = u + (double)j;
    return pow2 ? s * inv_n : s / dn;
}
__device__ __noinline__ int count_le_slow(int est, double c, double u, int N, double dn, double inv_n, bool pow2) {
    while (est > 0 && cpoint(u, est - 1, dn, inv_n, pow2) > c) --est;
    while (est < N && cpoint(u, est, dn, inv_n, pow2) <= c) ++est;
    return est;
}
// #{ j in [0, N) : (u + j) / N <= c }: closed form, then the exact predicate of :703-711 on both
// neighbours.  frac_out = distance of c*N - u to the nearest integer (how close to a tie).
__device__ __forceinline__ int count_le(double c, double u, int N, double dn, double inv_n, bool pow2,
                                        double& frac_out) {
    const double e = c * dn - u;
    if (!(e >= 0.0) || e >= dn - 1.0) {   // the ends of the range: rare, exact search
        frac_out = 1.0;
        return count_le_slow(e >= 0.0 ? N : 0, c, u, N, dn, inv_n, pow2);
    }
    const int fi = (int)e;               // floor(e), 0 <= fi <= N - 2
    const double fl = (double)fi;
    const double fr = e - fl;
    frac_out = fmin(fr, 1.0 - fr);
    const double slo = u + fl, shi = u + (fl + 1.0);   // u + j for j = fi, fi + 1 (both exact conversions)
    const bool ok = pow2 ? (slo * inv_n <= c && !(shi * inv_n <= c)) : (slo / dn <= c && !(shi / dn <= c));
    return ok ? fi + 1 : count_le_slow(fi + 1, c, u, N, dn, inv_n, pow2);
}

template <int NF>
__device__ __forceinline__ int fine_bin(double x, double mhat, double inv_shat) {
    const double t = ((x - mhat) * inv_shat + kZ) * ((double)NF / (2.0 * kZ));
    if (!(t >= 0.0)) return 0;
    if (t >= (double)NF) return NF - 1;
    return (int)t;
}
template <int NSB>
__device__ __forceinline__ int sub_bin(double x, double lo, double scale) {
    const double t = (x - lo) * scale;
    if (!(t >= 0.0)) return 0;
    if (t >= (double)NSB) return NSB - 1;
    return (int)t;
}

// Scans over one value per thread of the GT main threads (named barrier 1).  s_w: shared [32].
// Every warp scans the warp totals itself (shuffles), so there are two barriers and no serial
// loop; the order of the additions is fixed (deterministic).  s_w may be reused right after.
template <int GT>
__device__ __forceinline__ int block_excl_scan_int(int v, int* s_w, int& total, int lane, int warp) {
    constexpr int NW = GT / 32;
    const int incl = warp_incl_scan(v, lane);
    bar_sync(1, GT);
    if (lane == 31) s_w[warp] = incl;
    bar_sync(1, GT);
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
    bar_sync(1, GT);
    if (lane == 31) s_w[warp] = incl;
    bar_sync(1, GT);
    const int wt = (lane < NW) ? s_w[lane] : init;
    const int wincl = warp_incl_max(wt, lane);
    int wex = __shfl_up_sync(kFullMask, wincl, 1);
    if (lane == 0) wex = init;
    const int woff = max(init, __shfl_sync(kFullMask, wex, warp));
    int ex = __shfl_up_sync(kFullMask, incl, 1);
    if (lane == 0) ex = init;
    return max(woff, ex);
}

// ---------------------------------------------------------------------------------------------
// helper warps: what only feeds outputs (genealogy, fixed-lag score terms)
// ---------------------------------------------------------------------------------------------
// Genealogy by pointer jumping.  BP[t][j] = birth row (generation t-1) of the parent of row j of
// generation t (written by the main warps), J2[t] = BP[t-1] o BP[t] (2 generations back), J4[t] =
// J2[t-2] o J2[t] (4 back).  Every random read goes to a table of an OLDER step, so one helper
// barrier per step orders everything; the tables are 4 bytes per particle and the hot ones stay in
// L2.  The ancestor LAG-2 generations back follows from at most three more hops; its (parent
// value, value) pair and the particle's own weight (same formula and shift as the main warps) give
// the fixed-lag score terms of step t (:445-470) directly: no atomics, fixed summation order.
// grid_finish_kernel assembles the gradient terms from the six sums.
constexpr int kRingBP = 12, kRingJ = 8;

// Helper warps, step t, rows [jb, jb + nc) of generation t: jump tables, then the (parent value,
// value) pair of the ancestor LAG-2 generations back is copied to Q[t & 1][j] (birth order), where
// the main warps read it sequentially one step later (score_pass).  No arithmetic here: the
// helper warps are few, all they have is memory-level parallelism (IT independent chains each).
template <int GTH>
__device__ __forceinline__ void helper_lineage(const GridArgs& a, int t, int jb, int nc, int htid) {
    constexpr int IT = 16;   // independent chains in flight per thread
    const int N = a.N, L = a.LAG, RP = a.RP;
    const int* BPt = a.BP + (size_t)(t % kRingBP) * N + jb;
    const int* BPm1 = a.BP + (size_t)((t - 1) % kRingBP) * N;
    int* J2t = a.J2 + (size_t)(t % kRingJ) * N + jb;
    const int* J2m2 = a.J2 + (size_t)((t + kRingJ - 2) % kRingJ) * N;
    int* J4t = a.J4 + (size_t)(t % kRingJ) * N + jb;
    const bool do2 = t >= 2, do4 = t >= 4, score = t >= L && !(a.dbg & 4);
    const int g = t - (L - 2);                               // generation of the lagged pair
    const double2* Pg = a.P + (size_t)((g > 0 ? g : 0) % RP) * N;
    double2* Qt = a.Q + (size_t)(t & 1) * N + jb;
    for (int i0 = 0; i0 < nc; i0 += IT * GTH) {
        int cur[IT], b1[IT];
#pragma unroll
        for (int u = 0; u < IT; ++u) {
            const int i = i0 + u * GTH + htid;
            b1[u] = (i < nc) ? __ldcg(&BPt[i]) : 0;
        }
#pragma unroll
        for (int u = 0; u < IT; ++u) cur[u] = do2 ? __ldcg(&BPm1[min(max(b1[u], 0), N - 1)]) : 0;
        if (do2) {
#pragma unroll
            for (int u = 0; u < IT; ++u) {
                const int i = i0 + u * GTH + htid;
                if (i < nc) __stcg(&J2t[i], cur[u]);
            }
        }
        int gen = t, rem = L - 2;
        if (rem >= 4 || !score) {
            // 4 generations back (the table is needed by later steps whatever the lag is)
            if (do4) {
#pragma unroll
                for (int u = 0; u < IT; ++u) cur[u] = __ldcg(&J2m2[min(max(cur[u], 0), N - 1)]);
#pragma unroll
                for (int u = 0; u < IT; ++u) {
                    const int i = i0 + u * GTH + htid;
                    if (i < nc) __stcg(&J4t[i], cur[u]);
                }
            }
            gen -= 4;
            rem -= 4;
        } else {
            if (do4) {
#pragma unroll
                for (int u = 0; u < IT; ++u) {
                    const int i = i0 + u * GTH + htid;
                    const int b4 = __ldcg(&J2m2[min(max(cur[u], 0), N - 1)]);
                    if (i < nc) __stcg(&J4t[i], b4);
                }
            }
            if (rem >= 2) {
                gen -= 2;
                rem -= 2;
            } else if (rem == 1) {
#pragma unroll
                for (int u = 0; u < IT; ++u) cur[u] = b1[u];
                gen -= 1;
                rem -= 1;
            } else {
#pragma unroll
                for (int u = 0; u < IT; ++u) cur[u] = jb + i0 + u * GTH + htid;
            }
        }
        if (score) {
            // the remaining hops to generation g (uniform over the threads)
            while (rem > 0) {
                const int h = rem >= 4 ? 4 : (rem >= 2 ? 2 : 1);
                const int* tab = (h == 4 ? a.J4 + (size_t)(gen % kRingJ) * N
                                         : (h == 2 ? a.J2 + (size_t)(gen % kRingJ) * N : a.BP + (size_t)(gen % kRingBP) * N));
#pragma unroll
                for (int u = 0; u < IT; ++u) cur[u] = __ldcg(&tab[min(max(cur[u], 0), N - 1)]);
                gen -= h;
                rem -= h;
            }
#pragma unroll
            for (int u0 = 0; u0 < IT; u0 += 8) {
                double2 pe[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) pe[u] = __ldcg(&Pg[min(max(cur[u0 + u], 0), N - 1)]);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + (u0 + u) * GTH + htid;
                    if (i < nc) __stcg(&Qt[i], pe[u]);
                }
            }
        }
    }
}

// Main warps: fixed-lag score terms of step ts (:445-470) over the rows [jb, jb + nc) of generation
// ts (birth order): the row's weight (same formula and shift as phase B) times the monomials of the
// lagged pair Q[ts & 1][j] the helper warps left behind.  Sequential reads, fixed summation order.
// grid_finish_kernel assembles the gradient terms from the six sums.
template <int GT, int KPT>
__device__ __forceinline__ void score_pass(const GridArgs& a, const SvConst& k, int ts, int jb, int nc, int c, int tid,
                                           double shift, double* s_red) {
    constexpr int NW = GT / 32;
    const int lane = tid & 31, warp = tid >> 5;
    const int N = a.N, L = a.LAG, g = ts - (L - 2);
    const double2* Pt = a.P + (size_t)(ts % a.RP) * N + jb;
    const double2* Qt = a.Q + (size_t)(ts & 1) * N + jb;
    const double y = a.obs[ts], hy2 = 0.5 * y * y;
    const double ylag = a.obs[g >= 2 ? g - 2 : 0];   // Q5: the score terms of step i use obs[i - LAG]
    const double mu = k.mu, phi = k.phi, sr = k.sr;
    double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int k0 = 0; k0 < KPT; k0 += 2) {
        double xv[2];
        double2 pe[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = (k0 + u) * GT + tid;
            xv[u] = 0.0;
            pe[u] = make_double2(0.0, 0.0);
            if (k0 + u < KPT && i < nc) {
                xv[u] = __ldcg(&Pt[i].y);
                pe[u] = __ldcg(&Qt[i]);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = (k0 + u) * GT + tid;
            if (k0 + u < KPT && i < nc) {
                const double e = exp(-0.5 * xv[u]);
                const double lw = (-0.91893853320467267 - 0.5 * xv[u]) - hy2 * (e * e);
                double sh = exp(lw - shift);
                if (!isfinite(sh)) sh = 0.0;   // phase B has abandoned the evaluation
                const double cc = pe[u].x, xa = pe[u].y;
                const double ec = exp(-0.5 * cc);
                // residual of the transition (value at g-1) -> (value at g) as the score terms use it (:452-453)
                double sq = xa - mu - phi * (cc - mu);
                sq -= sr * ec * ylag;
                const double ey = ec * ylag;
                const double ws = sh * sq;
                acc[0] += sh;
                acc[1] = fma(sh, cc, acc[1]);
                acc[2] += ws;
                acc[3] = fma(ws, cc, acc[3]);
                acc[4] = fma(ws, sq, acc[4]);
                acc[5] = fma(ws, ey, acc[5]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) acc[i] = warp_sum(acc[i]);
    bar_sync(1, GT);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 6; ++i) s_red[32 * i + warp] = acc[i];
    }
    bar_sync(1, GT);
    if (warp == 0) {
        double* ps = a.psum + ((size_t)ts * a.G + c) * 8;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const double v = warp_sum((lane < NW) ? s_red[32 * i + lane] : 0.0);
            if (lane == 0) ps[i == 0 ? 7 : 1 + i] = v;   // [7] = sum sh, [2..6] = the monomial sums
        }
    }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int GT, int BPT>
__global__ void __launch_bounds__(kGT, 1) sv_grid_kernel(const GridArgs a) {
    constexpr int GTH = kGT - GT;                    // helper threads
    constexpr int KPT = (kCap + GT - 1) / GT;        // entries per main thread (strided assignment)
    constexpr int KCH = KPT | 1;                     // longest chunk of the thread-contiguous passes (odd)
    constexpr int NF = GT * BPT;                     // bins of the global value histogram
    constexpr int SPT = 8;                           // sub-bins per thread
    constexpr int NSB = GT * SPT;                    // sub-bins of the in-tile counting sort
    constexpr int NW = GT / 32;
    constexpr int CH = KPT % 8 == 0 ? 8 : (KPT % 5 == 0 ? 5 : (KPT % 4 == 0 ? 4 : (KPT % 3 == 0 ? 3 : 1)));
    static_assert(BPT % 4 == 0 && NF <= kNFMax && NSB <= 8192 && GTH >= 64 && GT * KPT >= kCap, "shape");
    extern __shared__ __align__(16) unsigned char smem[];
    // phase B / C view
    double* s_xb = (double*)smem;                      // [kCap] values in sub-bin order
    int* s_jb = (int*)(smem + 65536);                  // [kCap] birth rows in sub-bin order
    double* s_sh = (double*)(smem + 98304);            // [kCap] unnormalised weights, sorted order
    int* s_sub = (int*)(smem + 163840);                // [NSB] sub-bin counters, then first positions
    int* s_ub = s_jb;                                  // [kCap] child range ends (phase C)
    // phase A view
    int* s_fhist = (int*)smem;                         // [NF] histogram of this CTA's children
    unsigned short* s_tileof = (unsigned short*)(smem + 49152);   // [NF] tile of a histogram bin

    __shared__ SvConst s_k;
    __shared__ StepScalars s_sc;
    __shared__ double s_tot[kMaxTiles], s_off[kMaxTiles + 1];
    __shared__ int s_tstart[kMaxTiles + 1], s_tbin[kMaxTiles + 1], s_tcnt[kMaxTiles], s_tbase[kMaxTiles];
    __shared__ double s_red[6 * 32];
    __shared__ int s_wi[32];
    __shared__ long long s_prof[kProf];
    // main warps -> helper warps: s_main_step = last step whose BP / P slices and shift are complete;
    // helper -> main: s_help_step = last step the helper has finished
    __shared__ volatile int s_main_step, s_help_step, s_abort, s_habort[2];
    __shared__ double s_shiftring[4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x, G = a.G, N = a.N, NOBS = a.NOBS, L = a.LAG, Wc = a.Wc, RP = a.RP;
    GridCtrl* ctrl = a.ctrl;
    const bool prof = a.prof != nullptr;
    const int jb = min(N, c * Wc), je = min(N, jb + Wc), nc = je - jb;   // this CTA's children

    if (tid == 0) {
        sv_const_init(s_k, a.params);
        s_sc.abort_now = 0;
        s_main_step = 0;
        s_help_step = 0;
        s_abort = 0;
        s_habort[0] = s_habort[1] = 0;
        for (int i = 0; i < kProf; ++i) s_prof[i] = 0;
    }
    for (int i = tid; i < kMaxTiles; i += kGT) s_tcnt[i] = 0;
    __syncthreads();

    if (tid >= GT) {
        // ======================================================================================
        // helper warps
        // ======================================================================================
        const int htid = tid - GT;
        long long hclk = 0, hbusy = 0, hwait = 0;
        if (prof && htid == 0) hclk = clock64();
        bool stopped = false;
        for (int t = 1; t < NOBS; ++t) {
            // step t needs the BP / P slices and the shift of step t from the main warps of this CTA
            // and the tables of every helper's step t-1
            if (htid == 0) {
                const unsigned tgt = (unsigned)(t - 1) * (unsigned)G;
                int ab = 0;
                while (!(ab = s_abort) && (s_main_step < t || ld_acquire_u32(&ctrl->hbar) < tgt)) {
                }
                __threadfence();
                s_habort[t & 1] = ab;
                if (prof) {
                    const long long now = clock64();
                    hwait += now - hclk;
                    hclk = now;
                }
            }
            bar_sync(2, GTH);
            if (s_habort[t & 1]) {
                stopped = true;
                break;
            }
            if (htid < 32 && t + 1 >= L && t + 1 < NOBS && nc > 0 && !(a.dbg & 4)) {
                // the generation the NEXT step's score terms gather from -> L2 (sequential HBM reads)
                const double2* Pg = a.P + (size_t)((t + 1 - (L - 2)) % RP) * N;
                prefetch_range(Pg + jb, Pg + je, htid, 32);
            }
            if (!(a.dbg & 1)) helper_lineage<GTH>(a, t, jb, nc, htid);
            __threadfence();
            bar_sync(2, GTH);
            if (htid == 0) {
                red_release_add(&ctrl->hbar, 1u);
                s_help_step = t;
                if (prof) {
                    const long long now = clock64();
                    hbusy += now - hclk;
                    hclk = now;
                }
            }
        }
        (void)stopped;
        if (prof && htid == 0) {
            s_prof[14] = hbusy;
            s_prof[15] = hwait;
        }
    } else {
        // ======================================================================================
        // main warps
        // ======================================================================================
        const double dn = (double)N, inv_n = 1.0 / dn;
        const bool pow2 = (N & (N - 1)) == 0;
        const unsigned long long pol_stream = policy_evict_first();
        constexpr double kBinW = 2.0 * kZ / (double)NF;    // width of a histogram bin in predicted sd
        unsigned epoch = 0;                 // arrives done so far
        unsigned cnt_near = 0, cnt_soft = 0, cnt_key = 0;
        int my_max_bin = 0;
        long long pclk = 0;

#define MSYNC() bar_sync(1, GT)
#define GRID_FLAG(reason) atomicCAS(&ctrl->status, 0, (int)((epoch + 1u) | ((unsigned)(reason) << 24)))
#define GRID_ARRIVE()                                \
    do {                                             \
        MSYNC();                                     \
        if (tid == 0) {                              \
            __threadfence();                         \
            red_release_add(&ctrl->bar, 1u);         \
        }                                            \
        ++epoch;                                     \
    } while (0)
#define GRID_WAIT(extra)                                                              \
    do {                                                                              \
        if (tid == 0) {                                                               \
            const unsigned tgt = epoch * (unsigned)G;                                 \
            while (ld_acquire_u32(&ctrl->bar) < tgt) {                                \
            }                                                                         \
            __threadfence();                                                          \
            const int stv = *(volatile int*)&ctrl->status;                            \
            s_sc.abort_now = (stv != 0 && (unsigned)(stv & 0xffffff) <= epoch) ? 1 : 0; \
            extra;                                                                    \
        }                                                                             \
        MSYNC();                                                                      \
    } while (0)
#define PROF_MARK(slot)                              \
    do {                                             \
        if (prof && tid == 0) {                      \
            const long long now__ = clock64();       \
            s_prof[slot] += now__ - pclk;            \
            pclk = now__;                            \
        }                                            \
    } while (0)

        // --------------------------------------------------------------------------------------
        // generation 0 (:306-323, Q1): every particle = mu, uniform weights, identity order
        // --------------------------------------------------------------------------------------
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
                }
            }
        }
        GRID_ARRIVE();
        GRID_WAIT((void)0);
        if (prof && tid == 0) pclk = clock64();

        for (int t = 1; t < NOBS; ++t) {
            const int par = t & 1;
            // ----------------------------------------------------------------------------------
            // phase C: totals of all tiles -> offsets; child ranges of this tile's parents (:694-715)
            // ----------------------------------------------------------------------------------
            const double ur = a.rvr[t];
            double mom1 = 0.0, mom2 = 0.0;   // warp 0: weighted moments of the propagation mean
            if (warp == 0) {
                constexpr int kPer = kMaxTiles / 32;
                double tv[kPer], m1v[kPer], m2v[kPer];
#pragma unroll
                for (int i = 0; i < kPer; ++i) {
                    const int k = lane * kPer + i;
                    tv[i] = m1v[i] = m2v[i] = 0.0;
                    if (k < G) {
                        const double2 t0 = __ldcg((const double2*)&a.tinfo[k * 4]);
                        const double2 t1 = __ldcg((const double2*)&a.tinfo[k * 4 + 2]);
                        tv[i] = t0.x;
                        m1v[i] = t1.x;
                        m2v[i] = t1.y;
                    }
                }
                double loc = 0.0, l1 = 0.0, l2 = 0.0;
#pragma unroll
                for (int i = 0; i < kPer; ++i) {
                    loc = loc + tv[i];
                    l1 = l1 + m1v[i];
                    l2 = l2 + m2v[i];
                }
                const double incl = warp_incl_scan(loc, lane);
                double run = __shfl_up_sync(kFullMask, incl, 1);
                if (lane == 0) run = 0.0;
#pragma unroll
                for (int i = 0; i < kPer; ++i) {
                    const int k = lane * kPer + i;
                    if (k < G) {
                        s_off[k] = run;
                        s_tot[k] = tv[i];
                        run = run + tv[i];
                    }
                }
                const double S = __shfl_sync(kFullMask, incl, 31);
                const double invS = 1.0 / S;
                __syncwarp();
                if (lane == 0) {
                    s_sc.S = S;
                    s_sc.invS = invS;
                    s_sc.offk = s_off[c];
                    if (!(S > 0.0) || !isfinite(S)) GRID_FLAG(2);
                }
                // child range end of the tiles in front of this one (running maximum, see below)
                {
                    int carry = 0;
                    const int k = c - 1 - lane;
                    if (lane < 2 && k >= 0) {
                        double fr;
                        carry = count_le((s_off[k] + s_tot[k]) * invS, ur, N, dn, inv_n, pow2, fr);
                    }
                    carry = max(carry, __shfl_down_sync(kFullMask, carry, 1));
                    if (lane == 0) s_sc.carry = carry;
                }
                mom1 = warp_sum(l1);
                mom2 = warp_sum(l2);
            }
            MSYNC();
            {
                const double invS = s_sc.invS, offk = s_sc.offk;
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
#pragma unroll
                for (int kk = 0; kk < KCH; ++kk) {
                    const int q = q0 + kk;
                    if (kk < Lc && q < n) {
                        prev = max(ubv[kk], prev);
                        s_ub[q] = prev;
                    }
                }
            }
            MSYNC();
            {
                // ancestor of every child: parent q owns the children [ub(q-1), ub(q)); consecutive
                // lanes hold consecutive parents, so the stores of a warp fall into a few lines
                const int carry = s_sc.carry;
#pragma unroll 1
                for (int kk = 0; kk < KPT; ++kk) {
                    const int q = kk * GT + tid;
                    if (kk * GT >= n) break;
                    int lo = 0, hi = 0;
                    if (q < n) {
                        hi = s_ub[q];
                        lo = q ? s_ub[q - 1] : carry;
                    }
                    const int P = pstart + q;
                    const bool longr = hi - lo > 8;
                    if (!longr)
                        for (int k = lo; k < hi; ++k) __stcg(&a.H[k], P);
                    unsigned m = __ballot_sync(kFullMask, longr);
                    while (m) {
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        const int l0 = __shfl_sync(kFullMask, lo, src), h0 = __shfl_sync(kFullMask, hi, src);
                        const int P0 = __shfl_sync(kFullMask, P, src);
                        for (int k = l0 + lane; k < h0; k += 32) __stcg(&a.H[k], P0);
                    }
                }
            }
            if (tid == 0) {
                // predicted mean / sd of the children: histogram range of phase A
                const double S = s_sc.S;
                const double mhat = mom1 / S;
                double var = mom2 / S - mhat * mhat;
                if (!(var > 0.0)) var = 0.0;
                var = var + s_k.sd * s_k.sd;
                const double shat = sqrt(var);
                s_sc.mhat = mhat;
                s_sc.shat = shat;
                s_sc.inv_shat = 1.0 / shat;
                if (!isfinite(mhat) || !(shat > 0.0) || !isfinite(shat)) GRID_FLAG(2);
            }
            PROF_MARK(0);   // C ranges+fill
            GRID_ARRIVE();   // ---- barrier 4: ancestors complete
            for (int b = tid; b < NF; b += GT) s_fhist[b] = 0;
            PROF_MARK(1);   // zero hist
            // (the helper has to be done with step t-2 before its BP / P slots are written again)
            GRID_WAIT(while (s_help_step < t - 2) {});
            PROF_MARK(2);   // wait 4
            if (s_sc.abort_now) break;

            // ----------------------------------------------------------------------------------
            // phase A1: parents of this CTA's children, propagation (:354-358), value histogram
            // ----------------------------------------------------------------------------------
            double xn[KPT];
            {
                int hp[KPT];
#pragma unroll
                for (int kk = 0; kk < KPT; ++kk) {
                    const int i = kk * GT + tid;
                    hp[kk] = 0;
                    if (i < nc) hp[kk] = __ldcg(&a.H[jb + i]);
                }
                const double y1 = a.obs[t - 1];
                const double mhat = s_sc.mhat, inv_shat = s_sc.inv_shat;
                const double mu = s_k.mu, phi = s_k.phi, sr = s_k.sr, sd = s_k.sd;
                const double* Ut = a.U + (size_t)t * N;
                double2* Pt = a.P + (size_t)(t % RP) * N;
                int* BPt = a.BP + (size_t)(t % kRingBP) * N;
                bool bad = false, orphan = false;
#pragma unroll
                for (int k0 = 0; k0 < KPT; k0 += CH) {
                    // all loads of CH children are in flight before the first one is used
                    double2 xe[CH];
                    double uu[CH];
                    int bpv[CH];
#pragma unroll
                    for (int u = 0; u < CH; ++u) {
                        const int i = (k0 + u) * GT + tid;
                        xe[u] = make_double2(0.0, 0.0);
                        uu[u] = 0.0;
                        bpv[u] = 0;
                        if (i < nc) {
                            int p = hp[k0 + u];
                            if ((unsigned)p >= (unsigned)N) {
                                orphan = true;
                                p = 0;
                            }
                            xe[u] = __ldcg(&a.XE[p]);
                            bpv[u] = __ldcg(&a.perm[p]);
                            uu[u] = ld_stream_hint_f64(Ut + jb + i, pol_stream);
                            if (a.hist) a.parentpos[jb + i] = p;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < CH; ++u) {
                        const int i = (k0 + u) * GT + tid;
                        xn[k0 + u] = 0.0;
                        if (i < nc) {
                            double mean = mu + phi * (xe[u].x - mu);     // :355
                            mean += (sr * xe[u].y) * y1;                  // :356
                            const double x = mean + sd * uu[u];           // :357-358
                            if (!isfinite(x)) bad = true;
                            xn[k0 + u] = x;
                            atomicAdd(&s_fhist[fine_bin<NF>(x, mhat, inv_shat)], 1);
                            __stcg(&Pt[jb + i], make_double2(xe[u].x, x));
                            __stcg(&BPt[jb + i], bpv[u]);
                        }
                    }
                }
                if (bad) GRID_FLAG(2);
                if (orphan) GRID_FLAG(4);
            }
            MSYNC();
            {
                int* gh = a.ghist + (size_t)par * kNFMax;
#pragma unroll
                for (int kk = 0; kk < BPT; ++kk) {
                    const int b = kk * GT + tid;
                    const int cnt = s_fhist[b];
                    if (cnt) atomicAdd(&gh[b], cnt);
                }
            }
            PROF_MARK(3);   // A1 children+hist
            GRID_ARRIVE();   // ---- barrier 1: global histogram complete
            {
                // housekeeping for the next step
                const int zper = (NF + G - 1) / G;
                int* ghn = a.ghist + (size_t)(par ^ 1) * kNFMax;
                for (int b = c * zper + tid; b < min(NF, (c + 1) * zper); b += GT) __stcg(&ghn[b], 0);
                if (tid == 0) __stcg(&a.tilecnt[((par ^ 1) * kMaxTiles + c) * kCntStride], 0);
            }
            PROF_MARK(4);   // housekeeping
            GRID_WAIT((void)0);
            PROF_MARK(5);   // wait 1
            if (s_sc.abort_now) break;

            // ----------------------------------------------------------------------------------
            // phase A2: scan of the global histogram: tile boundaries on bin edges, tile of every
            // bin; entries go to the mailbox of their tile in runs
            // ----------------------------------------------------------------------------------
            {
                const int* gh = a.ghist + (size_t)par * kNFMax;
                int cnt[BPT];
#pragma unroll
                for (int i = 0; i < BPT; i += 4) {
                    const int4 v = __ldcg((const int4*)(gh + BPT * tid + i));
                    cnt[i] = v.x;
                    cnt[i + 1] = v.y;
                    cnt[i + 2] = v.z;
                    cnt[i + 3] = v.w;
                }
                int prevcnt = 0;
                if (tid > 0) prevcnt = __ldcg(gh + BPT * tid - 1);
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
                    for (int k = tprev + 1; k <= tl; ++k) {
                        s_tstart[k] = start;
                        s_tbin[k] = BPT * tid + i;
                    }
                    tprev = tl;
                    if (cnt[i] > 0) {
                        if (start == 0) s_sc.binlo = BPT * tid + i;
                        if (start + cnt[i] == total) s_sc.binhi = BPT * tid + i;
                    }
                    start += cnt[i];
                }
                if (tid == GT - 1) {
                    for (int k = tprev + 1; k <= G; ++k) {
                        s_tstart[k] = N;
                        s_tbin[k] = NF;
                    }
                    if (total != N) GRID_FLAG(5);
                }
            }
            MSYNC();
            if (tid == 0) {
                // sub-bins of this tile: its histogram bins cover [lo, hi)
                const double mhat = s_sc.mhat, shat = s_sc.shat;
                const double lo = mhat + shat * ((double)s_tbin[c] * kBinW - kZ);
                const double hi = mhat + shat * ((double)s_tbin[c + 1] * kBinW - kZ);
                s_sc.lo = lo;
                s_sc.scale = (hi > lo) ? (double)NSB / (hi - lo) : 0.0;
                // shift = largest log-weight over the occupied bins (any shift cancels, Q4)
                const double y = a.obs[t];
                const double xmin = mhat + shat * ((double)s_sc.binlo * kBinW - kZ);
                const double xmax = mhat + shat * ((double)(s_sc.binhi + 1) * kBinW - kZ);
                double xs_ = (y != 0.0) ? 2.0 * log(fabs(y)) : xmin;
                xs_ = fmin(fmax(xs_, xmin), xmax);
                const double es_ = exp(-0.5 * xs_);
                s_sc.shift = (-0.91893853320467267 - 0.5 * xs_) - (0.5 * y * y) * (es_ * es_);
                if (c == 0) a.shiftv[t] = s_sc.shift;
                // hand step t to the helper warps: BP / P slices (written before the barriers above) + shift
                s_shiftring[t & 3] = s_sc.shift;
                __threadfence();
                s_main_step = t;
            }
            int kr[KPT];
            {
                const double mhat = s_sc.mhat, inv_shat = s_sc.inv_shat;
#pragma unroll
                for (int kk = 0; kk < KPT; ++kk) {
                    const int i = kk * GT + tid;
                    kr[kk] = 0;
                    if (i < nc) {
                        const int tl = s_tileof[fine_bin<NF>(xn[kk], mhat, inv_shat)];
                        const int r = atomicAdd(&s_tcnt[tl], 1);
                        kr[kk] = (tl << 16) | r;
                    }
                }
            }
            MSYNC();
            if (tid < G) {
                // slots of this CTA's run in every mailbox
                const int cnt = s_tcnt[tid];
                int base = s_tstart[tid];
                if (cnt) base += atomicAdd(&a.tilecnt[(par * kMaxTiles + tid) * kCntStride], cnt);
                s_tbase[tid] = base;
                s_tcnt[tid] = 0;
            }
            MSYNC();
            {
                MailEntry* mail = (MailEntry*)a.XE;
#pragma unroll
                for (int kk = 0; kk < KPT; ++kk) {
                    const int i = kk * GT + tid;
                    if (i < nc) {
                        const int pos = s_tbase[kr[kk] >> 16] + (kr[kk] & 0xffff);
                        const long long xb = __double_as_longlong(xn[kk]);
                        const int4 ent = make_int4((int)(xb & 0xffffffffll), (int)(xb >> 32), jb + i, 0);
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
            PROF_MARK(6);   // A2 scan+scatter
            GRID_ARRIVE();   // ---- barrier 2: mailboxes complete
            for (int b = tid; b < NSB; b += GT) s_sub[b] = 0;
            if (warp == 0 && nc > 0 && t + 1 < NOBS) {
                // the next step's slice of u -> L2
                const double* Un = a.U + (size_t)(t + 1) * N;
                prefetch_range(Un + jb, Un + je, lane, 32);
            }
            PROF_MARK(7);   // zero+prefetch
            GRID_WAIT((void)0);
            PROF_MARK(8);   // wait 2
            if (s_sc.abort_now) break;

            // ----------------------------------------------------------------------------------
            // phase B: sort this tile (:392-424 / :23-52), weights (:427-442), block scan
            // ----------------------------------------------------------------------------------
            double acc[3] = {0.0, 0.0, 0.0};   // sh * {x, m, m^2}
            {
                const MailEntry* mb = (const MailEntry*)a.XE + pstart;
                const double lo = s_sc.lo, scale = s_sc.scale;
                double bx[KPT];
                int bj[KPT], ber[KPT];
#pragma unroll
                for (int k0 = 0; k0 < KPT; k0 += CH) {
                    int4 raw[CH];
#pragma unroll
                    for (int u = 0; u < CH; ++u) {
                        const int e = (k0 + u) * GT + tid;
                        raw[u] = make_int4(0, 0, 0, 0);
                        if (e < n) raw[u] = __ldcg((const int4*)(mb + e));
                    }
#pragma unroll
                    for (int u = 0; u < CH; ++u) {
                        const int e = (k0 + u) * GT + tid;
                        const double x = __longlong_as_double(((long long)raw[u].y << 32) | (long long)(unsigned)raw[u].x);
                        bx[k0 + u] = x;
                        bj[k0 + u] = raw[u].z;
                        ber[k0 + u] = 0;
                        if (e < n) ber[k0 + u] = atomicAdd(&s_sub[sub_bin<NSB>(x, lo, scale)], 1);
                    }
                }
                MSYNC();
                {
                    // exclusive scan of the sub-bin counters in place (SPT consecutive bins per thread)
                    int cnt[SPT];
#pragma unroll
                    for (int i = 0; i < SPT; i += 4) {
                        const int4 v = *(const int4*)(s_sub + SPT * tid + i);
                        cnt[i] = v.x;
                        cnt[i + 1] = v.y;
                        cnt[i + 2] = v.z;
                        cnt[i + 3] = v.w;
                    }
                    int loc = 0, mxb = 0;
#pragma unroll
                    for (int i = 0; i < SPT; ++i) {
                        loc += cnt[i];
                        mxb = max(mxb, cnt[i]);
                    }
                    if (mxb > kMaxSub) GRID_FLAG(3);
                    int total;
                    int start = block_excl_scan_int<GT>(loc, s_wi, total, lane, warp);
#pragma unroll
                    for (int i = 0; i < SPT; ++i) {
                        const int cn = cnt[i];
                        cnt[i] = start;
                        start += cn;
                    }
#pragma unroll
                    for (int i = 0; i < SPT; i += 4)
                        *(int4*)(s_sub + SPT * tid + i) = make_int4(cnt[i], cnt[i + 1], cnt[i + 2], cnt[i + 3]);
                }
                MSYNC();
#pragma unroll
                for (int kk = 0; kk < KPT; ++kk) {
                    const int e = kk * GT + tid;
                    if (e < n) {
                        const int pos = min(s_sub[sub_bin<NSB>(bx[kk], lo, scale)] + ber[kk], kCap - 1);
                        s_xb[pos] = bx[kk];
                        s_jb[pos] = bj[kk];
                    }
                }
                MSYNC();
            }
            PROF_MARK(9);   // B bin sort
            {
                // sub-bin order -> exact order (by value, then by birth row: :32-35 never returns 0), in place
                const double lo = s_sc.lo, scale = s_sc.scale;
                double bx[KPT];
                int bj[KPT], np[KPT];
#pragma unroll
                for (int kk = 0; kk < KPT; ++kk) {
                    const int q = kk * GT + tid;
                    bx[kk] = 0.0;
                    bj[kk] = 0;
                    np[kk] = q;
                    if (q < n) {
                        const double x = s_xb[q];
                        const int j = s_jb[q];
                        bx[kk] = x;
                        bj[kk] = j;
                        const int kb = sub_bin<NSB>(x, lo, scale);
                        const int b0 = s_sub[kb];
                        const int b1 = (kb + 1 < NSB) ? s_sub[kb + 1] : n;
                        if (b1 - b0 > 1 && b1 - b0 <= kMaxSub) {
                            int rank = 0;
                            for (int m = b0; m < b1; ++m) {
                                const double xm = s_xb[m];
                                if (xm < x) ++rank;
                                else if (xm == x && m != q) {
                                    ++cnt_key;
                                    if (s_jb[m] < j) ++rank;
                                }
                            }
                            np[kk] = b0 + rank;
                        }
                    }
                }
                MSYNC();
#pragma unroll
                for (int kk = 0; kk < KPT; ++kk) {
                    const int q = kk * GT + tid;
                    if (q < n && np[kk] != q) {
                        s_xb[np[kk]] = bx[kk];
                        s_jb[np[kk]] = bj[kk];
                    }
                }
                MSYNC();
            }
            {
                // weights (:427-437): lw = -0.9189 - x/2 - y^2 exp(-x) / 2, sh = exp(lw - shift)
                const double y = a.obs[t], hy2 = 0.5 * y * y, shift = s_sc.shift;
                const double mu = s_k.mu, phi = s_k.phi, sr = s_k.sr;
                bool bad = false;
#pragma unroll
                for (int kk = 0; kk < KPT; ++kk) {
                    const int q = kk * GT + tid;
                    if (q < n) {
                        const double x = s_xb[q];
                        const int j = s_jb[q];
                        const double e = exp(-0.5 * x);
                        const double lw = (-0.91893853320467267 - 0.5 * x) - hy2 * (e * e);
                        double sh = exp(lw - shift);
                        if (!isfinite(sh)) {
                            bad = true;
                            sh = 0.0;
                        }
                        __stcg(&a.XE[pstart + q], make_double2(x, e));
                        __stcg(&a.perm[pstart + q], j);
                        s_sh[q] = sh;
                        if (pstart + q == 0) a.xminv[t] = x;   // Q10/Q11: traj[t] = X_t[0]
                        if (a.hist) {
                            a.Xhist[(size_t)t * N + pstart + q] = x;
                            a.Ahist[(size_t)t * N + pstart + q] = __ldcg(&a.parentpos[j]);
                        }
                        acc[0] = fma(sh, x, acc[0]);
                        double m = mu + phi * (x - mu);
                        m += (sr * e) * y;
                        const double shm = sh * m;
                        acc[1] += shm;
                        acc[2] = fma(shm, m, acc[2]);
                    }
                }
                if (bad) GRID_FLAG(2);
            }
            MSYNC();
            PROF_MARK(10);   // B rank+weights
            {
                // cumulative weights: thread = Lc consecutive sorted particles, sequential running sum;
                // one block-wide pass gives the offsets of the chunks and the weighted sums
                const int Lc = ((n + GT - 1) / GT) | 1, q0 = tid * Lc;
                double run = 0.0;
#pragma unroll
                for (int kk = 0; kk < KCH; ++kk) {
                    const int q = q0 + kk;
                    if (kk < Lc && q < n) run = run + s_sh[q];
                }
                const double incl = warp_incl_scan(run, lane);
#pragma unroll
                for (int i = 0; i < 3; ++i) acc[i] = warp_sum(acc[i]);
                if (lane == 31) s_red[warp] = incl;
                if (lane == 0) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) s_red[32 * (1 + i) + warp] = acc[i];
                }
                MSYNC();
                const double wt = (lane < NW) ? s_red[lane] : 0.0;
                const double wincl = warp_incl_scan(wt, lane);
                double wex = __shfl_up_sync(kFullMask, wincl, 1);
                if (lane == 0) wex = 0.0;
                const double woff = __shfl_sync(kFullMask, wex, warp);
                double ex = __shfl_up_sync(kFullMask, incl, 1);
                if (lane == 0) ex = 0.0;
                toff = woff + ex;
                double* ps = a.psum + ((size_t)t * G + c) * 8;
                if ((n > 0 && tid == (n - 1) / Lc) || (n == 0 && tid == 0)) {
                    const double tot = (n > 0) ? toff + run : 0.0;   // cumulative weight of the tile's last particle
                    __stcg(&a.tinfo[c * 4], tot);
                    ps[0] = tot;
                }
                if (warp == 0) {
                    double v[3];
#pragma unroll
                    for (int i = 0; i < 3; ++i) v[i] = warp_sum((lane < NW) ? s_red[32 * (1 + i) + lane] : 0.0);
                    if (lane == 0) {
                        __stcg(&a.tinfo[c * 4 + 1], (double)n);
                        __stcg((double2*)&a.tinfo[c * 4 + 2], make_double2(v[1], v[2]));
                        ps[1] = v[0];
                        if (t < L || (a.dbg & 4)) {   // later steps: the helper warps write ps[2..7]
#pragma unroll
                            for (int i = 2; i < 8; ++i) ps[i] = 0.0;
                        }
                    }
                }
                if (t >= NOBS - L) {
#pragma unroll
                    for (int kk = 0; kk < KPT; ++kk) {
                        const int q = kk * GT + tid;
                        if (q < n) a.shring[(size_t)(t % L) * N + pstart + q] = s_sh[q];
                    }
                }
            }
            PROF_MARK(11);   // B scan+publish
            GRID_ARRIVE();   // ---- barrier 3: tile totals published
            if (t - 1 >= L && !(a.dbg & 4)) {
                // score terms of the previous step (its lagged pairs come from the helper warps)
                if (tid == 0)
                    while (s_help_step < t - 1) {
                    }
                MSYNC();
                score_pass<GT, KPT>(a, s_k, t - 1, jb, nc, c, tid, s_shiftring[(t - 1) & 3], s_red);
            }
            PROF_MARK(13);   // score terms
            GRID_WAIT((void)0);
            PROF_MARK(12);   // wait 3
            if (s_sc.abort_now) break;
        }
        if (tid == 0 && s_sc.abort_now) s_abort = 1;
        if (!s_sc.abort_now && NOBS - 1 >= L && !(a.dbg & 4)) {
            if (tid == 0)
                while (s_help_step < NOBS - 1) {
                }
            MSYNC();
            score_pass<GT, KPT>(a, s_k, NOBS - 1, jb, nc, c, tid, s_shiftring[(NOBS - 1) & 3], s_red);
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
#undef MSYNC
#undef GRID_FLAG
#undef GRID_ARRIVE
#undef GRID_WAIT
#undef PROF_MARK
    }
    __syncthreads();
    if (prof && tid == 0)
        for (int i = 0; i < kProf; ++i) a.prof[(size_t)c * kProf + i] += s_prof[i];
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
            int row = b;   // birth row of the ancestor m generations back: m hops through the parent tables
            for (int h = 0; h < m; ++h) row = min(max(a.BP[(size_t)((T - h) % kRingBP) * N + row], 0), N - 1);
            const double2 pe = a.P[(size_t)((T - m) % RP) * N + row];
            curr = pe.x;
            const double pe_n = pe.y;
            acc[0] += wT * curr;
            const double wi = shi[p] / Si;
            double sq, g[4];
            sv_score_tail_e(k, curr, exp(-0.5 * curr), pe_n, y1, sq, g);   // the tail uses obs[i - 1] (Q6)
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

// sums[t][0] = sum sh, [1] = sum sh x (main warps); helper warps, w = fixed-point copy of sh: [7] = sum w,
// [2] = sum w c, [3] = sum w sq, [4] = sum w sq c, [5] = sum w sq^2, [6] = sum w sq ey (c = lagged ancestor
// value, sq = residual of its transition, ey = exp(-c/2) obs);  tail[irel][0..4]
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
            const double* sm = sums + (size_t)src * 8;
            const double S = sm[7];   // the score terms are normalised by the sum of the weights they were built from
            SvConst k;
            sv_const_init(k, params);
            if (S > 0.0) {
            s = sm[2] / S;
            // weighted means of g[0..3] of :454-465 written in the monomials sq, sq c, sq^2, sq ey
            g[0] = k.q * k.one_m_phi * sm[3] / S;
            g[1] = k.q * k.one_m_phi2 * (sm[4] - k.mu * sm[3]) / S;
            g[2] = (k.q * sm[5] + k.q * k.sr * sm[6] - S) / S;
            g[3] = (k.rho * S - k.q * k.rho * sm[5] + k.inv_sv * sm[6]) / S;
            }
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
    size_t ctrl, ghist, tilecnt, tinfo, H, XE, perm, BP, J2, J4, Q, P, psum, shiftv, xminv, shring, parentpos, sums,
        tailpart, tail, info, total;
    int RP, nblk;
};

size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

GridLayout grid_layout(int nobs, int n, int lag, int G, int hist) {
    GridLayout L = GridLayout();
    // ring depth of the (parent value, value) generations: the helper warps may run up to two steps
    // behind the main warps, and the sums of step t-1 read generation t-1-(lag-2): lag+2 slots keep
    // the writes of the main warps away from the reads of the helper warps
    L.RP = lag + 2;
    L.nblk = 296;
    size_t o = 0;
    const size_t N = (size_t)n;
    L.ctrl = o;      o += al256(sizeof(GridCtrl));
    L.ghist = o;     o += al256((size_t)2 * kNFMax * 4);
    L.tilecnt = o;   o += al256((size_t)2 * kMaxTiles * kCntStride * 4);
    L.tinfo = o;     o += al256((size_t)kMaxTiles * 4 * 8);
    L.H = o;         o += al256(N * 4);
    L.XE = o;        o += al256(N * 16);
    L.perm = o;      o += al256(N * 4);
    L.BP = o;        o += al256((size_t)kRingBP * N * 4);
    L.J2 = o;        o += al256((size_t)kRingJ * N * 4);
    L.J4 = o;        o += al256((size_t)kRingJ * N * 4);
    L.Q = o;         o += al256((size_t)2 * N * 16);
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
    if (lag < 2 || lag > 10 || nobs < 2 * lag || n < 32 || n > (1 << 21)) return false;
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
    a.XE = (double2*)(ws + L.XE);
    a.perm = (int*)(ws + L.perm);
    a.BP = (int*)(ws + L.BP);
    a.J2 = (int*)(ws + L.J2);
    a.J4 = (int*)(ws + L.J4);
    a.Q = (double2*)(ws + L.Q);
    a.P = (double2*)(ws + L.P);
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
    // control block, histograms, reservation counters, tile info: zero
    GRID_CUDA(cudaMemsetAsync(ws + L.ctrl, 0, L.H - L.ctrl, st));
    static thread_local bool attr_set[64] = {false};
    static int main_threads = 0;
    if (!main_threads) {
        const char* e = getenv("PMMH_GRID_MAIN");   // main threads of the 1024 (the rest are helper warps)
        main_threads = (e && atoi(e) == 768) ? 768 : 896;
    }
    int dev = 0;
    GRID_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        GRID_CUDA(cudaFuncSetAttribute(sv_grid_kernel<896, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem));
        GRID_CUDA(cudaFuncSetAttribute(sv_grid_kernel<768, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem));
        attr_set[dev] = true;
    }
    void* kargs[] = {(void*)&a};
    if (main_threads == 896)
        GRID_CUDA(cudaLaunchCooperativeKernel((void*)sv_grid_kernel<896, 8>, dim3(G), dim3(kGT), kargs, kDynSmem, st));
    else
        GRID_CUDA(cudaLaunchCooperativeKernel((void*)sv_grid_kernel<768, 12>, dim3(G), dim3(kGT), kargs, kDynSmem, st));
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
