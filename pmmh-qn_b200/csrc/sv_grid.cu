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
//          histogram (shared-memory atomics, merged into the global one), payload of the child
//          for the fixed-lag terms 8 steps later                                   | barrier 1
//          (after the arrive: genealogy records)
//       A2 scan of the global histogram -> tile boundaries on bin edges (every tile gets N/G
//          particles +- one bin), slot reservation per (CTA, tile), entries written to the
//          mailbox of their tile in runs (value, birth row, lagged ancestor row)  | barrier 2
//       B  owner of a tile: counting sort over 4096 sub-bins of the tile's value range + exact
//          in-bin ranking by (value, birth row) = the reference's argsort; weights, sorted
//          generation written out, block scan of the weights, moments             | barrier 3
//          (after the arrive: fixed-lag score terms, one random 32-byte sector per particle)
//   * genealogy: a generation is stored ONCE in birth order as P[t][j] = (parent value, residual
//     of the transition, exp(-parent/2) obs, value) and R[t][j] = birth rows of the ancestors 1..8
//     steps back (one 32-byte sector each).  The fixed-lag terms of step t need one random sector
//     of P[t-lag+2]; a child copies its parent's R with one random sector read.  No history is
//     ever moved.
//   * two exp per particle and step: exp(-x/2) is shared by the weight, the propagation mean of
//     the children and the payload; the score terms need none.
//   * L2 eviction hints on every access (createpolicy): data that is dead after its read (previous record
//     table, mailbox entries, H, the parents' (x, e) / perm entries, lagged payloads, u) is read
//     evict_first, what the next phase reads is stored evict_last: 114 -> 103 us per step.
//   * tried and measured slower (git history): descendant weights by integer atomics + sequential
//     payload pass (132 us per step), genealogy and score terms on dedicated helper warps (122 -
//     150 us: the helper's traffic slows the main warps as much as it saves them), pointer-jumping
//     tables instead of records, bulk L2 prefetch of the lagged generation, L2 prefetch of the record / payload
//     sectors a phase ahead, 16-byte payloads with an exp in the score terms, pointer-jumping tables by all
//     threads (DRAM traffic 189 -> 138 MB per step, but the three dependent 4-byte gathers still miss L2:
//     27 us against 14 us for the records, 118 us per step).  With the hints: 100 us; with the fence-free barrier
//     (release / acquire on the counter, abandon flag in its bit 31): 94 - 96 us.
//   * second instantiation HESS = true: the Hessian branch (:361-390, :472-534, :564-626), cumulative alpha in the second
//     sector of 64-byte record / payload entries; 162 us per step.
//   * host-resident u: particle-major slots filled by the copy engine, one 32-byte sector (4 steps) per particle re-laid
//     time-major every fourth step.
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
constexpr int kNCopy = 1;          // striped copies of the global histogram (CTA c adds into copy c % kNCopy)
constexpr int kNSB = 4096;         // sub-bins of the in-tile counting sort
constexpr int kMaxSub = 1024;      // a sub-bin larger than this abandons the evaluation
constexpr int kMaxTiles = 160;     // >= SM count
constexpr int kCntStride = 32;     // ints between two slot counters (one 128-byte line each)
constexpr double kZ = 6.5;         // histogram range: predicted mean +- 6.5 predicted sd
constexpr int kDynSmem = 208 * 1024;
constexpr int kProf = 32;

struct __align__(16) MailEntry {   // aliases one (x, exp(-x/2)) pair of the sorted generation
    double x;
    int j, a;
};
// Payload of a particle for the fixed-lag terms (one 32-byte sector, birth order): parent value c,
// residual sq of the transition (:452-453), ey = exp(-c/2) * obs[i - LAG], own value x (tail)
struct __align__(32) PEntry {
    double c, sq, ey, x;
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
};

struct GridArgs {
    int N, NOBS, LAG, G, Wc, RP, hist;
    int dbg;   // development (timing only, results wrong): 1 skip R records, 2 skip P store, 4 skip score gather, 8 bulk prefetch of the lagged generation, 64 no L2 eviction hints
    const double *obs, *params, *rvr, *U;
    // U[t][j] = U[(t / u_cs) * u_cstride + (t % u_cs) * u_tstride + j * u_jstride]: time-major resident array
    // (u_cs = 2^30, u_tstride = N, u_jstride = 1) or particle-major chunks of u_cs time steps as the copy
    // engine lays them down from the reference's host array (u_tstride = 1, u_jstride = u_cs)
    long long u_cstride, u_tstride, u_jstride;
    int u_cs;
    const int* u_flag;   // host-streamed u: number of time rows that have landed (written by the copy engine)
    double* ustash;      // [4][N] host-streamed u: the four time steps of one 32-byte sector per particle, time-major
    GridCtrl* ctrl;
    int* ghist;        // [2][kNCopy][kNF]
    int* tilecnt;      // [2][kMaxTiles * kCntStride]
    double* tinfo;     // [kMaxTiles][4]  tot, n, sum sh m, sum sh m^2
    int* H;            // [N] ancestor (sorted position) of every child
    double2* XE;       // [N] sorted generation: (x, exp(-x/2)); the mailbox of the next generation aliases it
    int* perm;         // [N] sorted position -> birth row
    REntry* R;         // [2][N]
    PEntry* P;         // [RP][N] payloads of generation t % RP
    double* psum;      // [NOBS][G][8]
    // Hessian branch (:361-390, :472-534, :564-626) only: the entries of R and P are 64 bytes (RS = 2 sectors), the
    // second sector is the cumulative alpha (4 parameters) of the particle -- next to its record for the children
    // (who gather both with one 64-byte access) and next to its payload for the descendants LAG-2 steps later
    int RS;            // 32-byte sectors per entry of R and P (1, Hessian branch 2)
    double* xlow;      // [SQ][NOBS] the first SQ sorted values of every generation, slot-major: consecutive parents read
                       // consecutive times (Q7), so the reads of a warp coalesce
    double* psumH;     // [NOBS][G][20] per-step sums: [0..3] monomials of hessian1, [10..19] hessian2 (upper triangle)
    int SQ;
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
__device__ __forceinline__ unsigned long long policy_evict_normal() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long policy_evict_last() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// 32-byte genealogy records: one 256-bit access, kept in the persisting part of L2 (evict_last; the
// host sets cudaLimitPersistingL2CacheSize to the size of the two record tables)
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
// streaming 16-byte store that should leave L2 first (written once, read lag - 2 steps later)
__device__ __forceinline__ void st_stream_f64x2(void* p, unsigned long long pol, double a, double b) {
    asm volatile("st.global.cs.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" ::"l"(p), "d"(a), "d"(b), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void st_stream_f64x4(void* p, unsigned long long pol, double a, double b, double c, double d) {
    asm volatile("st.global.cs.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d),
                 "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void st_stream_f64(double* p, unsigned long long pol, double a) {
    asm volatile("st.global.cs.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(a), "l"(pol) : "memory");
}
__device__ __forceinline__ double ld_stream_hint_f64(const double* p, unsigned long long pol) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
// L2 residency plan (dbg & 64 switches it off): what the next phase or the next step reads again -- the
// sorted (x, e) pairs / the mailbox that aliases them, perm, H, the genealogy records being written -- is
// stored evict_last (56 MB at N = 2^20); what is dead after its read -- the previous record table, the
// mailbox entries, H, u -- and what is not read for 8 steps -- the payloads -- goes evict_first
__device__ __forceinline__ void st_hint_b128(void* p, unsigned long long pol, int4 v) {
    asm volatile("st.global.cg.L2::cache_hint.v4.s32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ int4 ld_hint_b128(const void* p, unsigned long long pol) {
    int4 v;
    asm volatile("ld.global.cg.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_hint_b32(void* p, unsigned long long pol, int v) {
    asm volatile("st.global.cg.L2::cache_hint.s32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ int ld_hint_b32(const void* p, unsigned long long pol) {
    int v;
    asm volatile("ld.global.cg.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void ld_f64x4(const void* p, unsigned long long pol, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.cg.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                 : "=d"(a), "=d"(b), "=d"(c), "=d"(d)
                 : "l"(p), "l"(pol));
}
__device__ __forceinline__ void st_keep_f64x4(void* p, unsigned long long pol, double a, double b, double c, double d) {
    asm volatile("st.global.cg.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d),
                 "l"(pol)
                 : "memory");
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

// Hessian branch: the alpha terms whose Q7 read falls into the generation being built (listed by the records pass),
// one thread per entry
__device__ __forceinline__ void deferred_alpha_pass(const GridArgs& a, const SvConst& k, const int2* list, int nd, int t,
                                                    int jb, int tid, int nthreads, unsigned long long pol_rld,
                                                    unsigned long long pol_rst, unsigned long long pol_stream) {
    const int N = a.N, NOBS = a.NOBS, RS = a.RS;
    const double yi = a.obs[t], ylagH = obs_wrap(a.obs, t - a.LAG, NOBS);
    const REntry* Rp = a.R + (size_t)((t - 1) & 1) * N * RS;
    REntry* Rc = a.R + (size_t)(t & 1) * N * RS;
    PEntry* Pt = a.P + (size_t)(t % a.RP) * N * RS;
    for (int d = tid; d < nd; d += nthreads) {
        const int2 e = list[d];
        const int j = jb + e.x;
        const int b = min(max(e.y, 0), N - 1);
        const int ppos = min(max(__ldcg(&a.H[j]), 0), N - 1);
        const int sl = (int)(((unsigned)(t - 1) + (unsigned)ppos) / (unsigned)NOBS);
        const double x = __ldcg(&Pt[(size_t)j * RS].x);
        const double curr = (sl <= j && sl < N) ? __ldcg(&Pt[(size_t)sl * RS].x) : 0.0;
        double p0, p1, p2, p3;
        ld_f64x4(&Rp[(size_t)b * RS + 1], pol_rld, p0, p1, p2, p3);
        double al[4];
        sv_alpha_terms(k, x, curr, yi, ylagH, al);
        const double c0 = al[0] + p0, c1 = al[1] + p1, c2 = al[2] + p2, c3 = al[3] + p3;
        st_keep_f64x4(&Rc[(size_t)j * RS + 1], pol_rst, c0, c1, c2, c3);
        st_stream_f64x4((REntry*)&Pt[(size_t)j * RS] + 1, pol_stream, c0, c1, c2, c3);
    }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// HESS: the Hessian branch of the reference with its quirks Q7 / Q8 (second instantiation)
template <int GT, bool HESS>
__global__ void __launch_bounds__(GT, 1) sv_grid_kernel(const GridArgs a) {
    constexpr int KPT = kCap / GT;          // entries per thread (strided assignment)
    constexpr int KCH = KPT | 1;            // longest chunk of the thread-contiguous passes (odd)
    constexpr int BPT = kNF / GT;           // histogram bins per thread
    constexpr int SPT = kNSB / GT;          // sub-bins per thread
    constexpr int NW = GT / 32;
    constexpr int CH = KPT >= 16 ? 8 : 4;   // independent loads in flight per thread in the gather loops
#ifndef PMMH_GRID_HESS_RB
#define PMMH_GRID_HESS_RB 2
#endif
#ifndef PMMH_GRID_HESS_SB
#define PMMH_GRID_HESS_SB 1
#endif
    constexpr int RB = HESS ? (KPT >= 16 ? 2 * PMMH_GRID_HESS_RB : PMMH_GRID_HESS_RB) : (KPT >= 16 ? 4 : 2);   // genealogy records in flight per thread
    constexpr int RS = HESS ? 2 : 1;        // 32-byte sectors per entry of R and P
    static_assert(SPT % 4 == 0 && BPT % 4 == 0, "vector loads of the counters");
    extern __shared__ __align__(16) unsigned char smem[];
    // phase B / C view
    double* s_xb = (double*)smem;                      // [kCap] values in sub-bin order
    int* s_jb = (int*)(smem + 65536);                  // [kCap] birth rows in sub-bin order
    int* s_ab = (int*)(smem + 98304);                  // [kCap] lagged ancestor rows in sub-bin order
    double* s_sh = (double*)(smem + 131072);           // [kCap] unnormalised weights, sorted order
    int* s_sub = (int*)(smem + 196608);                // [kNSB] sub-bin counters, then first positions
    int* s_ub = s_jb;                                  // [kCap] child range ends (phase C)
    // phase A view (Hessian branch: list of deferred alpha terms = (child, parent row) pairs in the s_jb / s_ab region)
    int2* s_deflist = (int2*)(smem + 65536);         // [kCap]
    int* s_fhist = (int*)smem;                         // [kNF] histogram of this CTA's children
    unsigned short* s_tileof = (unsigned short*)(smem + 32768);   // [kNF] tile of a histogram bin

    __shared__ SvConst s_k;
    __shared__ StepScalars s_sc;
    __shared__ double s_tot[kMaxTiles], s_off[kMaxTiles + 1];
    __shared__ int s_tstart[kMaxTiles + 1], s_tbin[kMaxTiles + 1], s_tcnt[kMaxTiles], s_tbase[kMaxTiles];
    __shared__ double s_red[9 * 32];
    __shared__ double s_redH[HESS ? 20 * 32 : 1];   // Hessian branch: warp sums of the 20 hessian1 / hessian2 terms
    __shared__ int s_wi[32];
    __shared__ long long s_prof[kProf];
    __shared__ int s_ndef;   // Hessian branch: length of the list of deferred alpha terms

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x, G = a.G, N = a.N, NOBS = a.NOBS, L = a.LAG, Wc = a.Wc, RP = a.RP;
    GridCtrl* ctrl = a.ctrl;
    const double dn = (double)N, inv_n = 1.0 / dn;
    const bool pow2 = (N & (N - 1)) == 0;
    const unsigned long long pol_stream = policy_evict_first();
    const unsigned long long pol_keep = (a.dbg & 64) ? policy_evict_normal() : policy_evict_last();
    const unsigned long long pol_dead = (a.dbg & 64) ? policy_evict_normal() : policy_evict_first();
    // (measured one by one, N = 2^20: the previous record table, the lagged payloads and the parents' (x, e)
    // / perm entries read as evict_first are worth 8 + 3 us per step; evict_last on the stores alone is not)
    const unsigned long long pol_rld = pol_dead, pol_pld = pol_dead, pol_xld = pol_dead, pol_rst = pol_keep;
    constexpr double kBinW = 2.0 * kZ / (double)kNF;    // width of a histogram bin in predicted sd
    unsigned epoch = 0;                 // arrives done so far
    unsigned cnt_near = 0, cnt_soft = 0, cnt_key = 0;
    int my_max_bin = 0;
    long long pclk = 0;
    const bool prof = a.prof != nullptr;

// The arrive is bar.sync + red.release.gpu by one thread, the wait ld.acquire.gpu by one thread + bar.sync: release /
// acquire at gpu scope are cumulative over what bar.sync ordered before / after them, so no separate __threadfence is
// needed (measured: 99.5 -> 96.8 ms per evaluation; -DPMMH_GRID_FENCE puts the fences back)
#ifdef PMMH_GRID_FENCE
#define GRID_FENCE() __threadfence()
#else
#define GRID_FENCE()
#endif
#ifdef PMMH_GRID_BACKOFF
#define GRID_BACKOFF() __nanosleep(PMMH_GRID_BACKOFF)
#else
#define GRID_BACKOFF()
#endif
// Abandoning: the reason goes to ctrl->status (first one wins), and bit 31 of the barrier counter is set, so that the
// load every CTA polls the barrier with also tells it to stop -- at its next wait, whichever barrier that is (nothing
// after the time loop needs the grid); a second load of the status word after every barrier costs an L2 round trip
#ifdef PMMH_GRID_STATUS_LOAD
#define GRID_FLAG(reason) atomicCAS(&ctrl->status, 0, (int)((epoch + 1u) | ((unsigned)(reason) << 24)))
#else
#define GRID_FLAG(reason)                                                                        \
    do {                                                                                         \
        atomicCAS(&ctrl->status, 0, (int)((epoch + 1u) | ((unsigned)(reason) << 24)));           \
        atomicOr(&ctrl->bar, 0x80000000u);                                                       \
    } while (0)
#endif
#ifdef PMMH_GRID_STATUS_LOAD
#define GRID_POLL(tgt)                                                                \
    do {                                                                              \
        while (ld_acquire_u32(&ctrl->bar) < (tgt)) {                                  \
            GRID_BACKOFF();                                                           \
        }                                                                             \
        GRID_FENCE();                                                                 \
        const int stv = *(volatile int*)&ctrl->status;                                \
        s_sc.abort_now = (stv != 0 && (unsigned)(stv & 0xffffff) <= epoch) ? 1 : 0;   \
    } while (0)
#else
#define GRID_POLL(tgt)                                                                \
    do {                                                                              \
        unsigned v__;                                                                 \
        do {                                                                          \
            v__ = ld_acquire_u32(&ctrl->bar);                                         \
        } while ((v__ & 0x7fffffffu) < (tgt) && !(v__ & 0x80000000u));                \
        s_sc.abort_now = (int)(v__ >> 31);                                            \
    } while (0)
#endif
#define GRID_ARRIVE()                                \
    do {                                             \
        __syncthreads();                             \
        if (tid == 0) {                              \
            GRID_FENCE();                            \
            red_release_add(&ctrl->bar, 1u);         \
        }                                            \
        ++epoch;                                     \
    } while (0)
#define GRID_WAIT()                                                                   \
    do {                                                                              \
        if (tid == 0) {                                                               \
            const unsigned tgt = epoch * (unsigned)G;                                 \
            GRID_POLL(tgt);                                                           \
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
            st_rec(&a.R[(size_t)(pstart + q) * RS], pol_keep, 0, 0, 0, 0, 0, 0, 0, 0);
            if constexpr (HESS) {
                st_rec(&a.R[(size_t)(pstart + q) * RS + 1], pol_keep, 0, 0, 0, 0, 0, 0, 0, 0);   // alpha = 0
                if (pstart + q < a.SQ) a.xlow[(size_t)(pstart + q) * NOBS] = mu;
            }
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
    GRID_WAIT();
    if (prof && tid == 0) pclk = clock64();

    for (int t = 1; t < NOBS; ++t) {
        const int par = t & 1;
        // --------------------------------------------------------------------------------------
        // phase C: totals of all tiles -> offsets; child ranges of this tile's parents (:694-715)
        // --------------------------------------------------------------------------------------
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
        __syncthreads();
        PROF_MARK(16);   // C: totals -> offsets (warp 0)
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
            PROF_MARK(17);   // C: cumulative weights + child counts
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
        __syncthreads();
        PROF_MARK(18);   // C: running maximum
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
                    for (int k = lo; k < hi; ++k) st_hint_b32(&a.H[k], pol_keep, P);
                unsigned m = __ballot_sync(kFullMask, longr);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const int l0 = __shfl_sync(kFullMask, lo, src), h0 = __shfl_sync(kFullMask, hi, src);
                    const int P0 = __shfl_sync(kFullMask, P, src);
                    for (int k = l0 + lane; k < h0; k += 32) st_hint_b32(&a.H[k], pol_keep, P0);
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
        for (int b = tid; b < kNF; b += GT) s_fhist[b] = 0;
        PROF_MARK(1);   // zero hist
        if (a.u_flag && tid == 0) {
            // host-streamed u: row t has to have landed (the flag is written by the copy engine); at the first of
            // the four steps that share a 32-byte sector, all four (see the stash below)
            const int need = (((t % a.u_cs) & 3) == 0 || t == 1) ? min(t | 3, NOBS - 1) : t;
            while (ld_acquire_sys_s32(a.u_flag) <= need) {
            }
        }
        GRID_WAIT();
        PROF_MARK(2);   // wait 4
        if (s_sc.abort_now) break;

        // --------------------------------------------------------------------------------------
        // phase A1: parents of this CTA's children, propagation (:354-358), value histogram
        // --------------------------------------------------------------------------------------
        double xn[KPT];
        int bp[KPT];
        {
#pragma unroll
            for (int kk = 0; kk < KPT; ++kk) {
                const int i = kk * GT + tid;
                bp[kk] = 0;
                if (i < nc) bp[kk] = ld_hint_b32(&a.H[jb + i], pol_dead);
            }
            const double y1 = a.obs[t - 1];
            const double ylag = a.obs[t >= 2 ? t - 2 : 0];   // Q5: the score terms of step i use obs[i - LAG]; i = t + LAG - 2
            const double mhat = s_sc.mhat, inv_shat = s_sc.inv_shat;
            const double mu = s_k.mu, phi = s_k.phi, sr = s_k.sr, sd = s_k.sd;
            const double* Ut = a.U + (size_t)(t / a.u_cs) * a.u_cstride + (size_t)(t % a.u_cs) * a.u_tstride;
            long long ujs = a.u_jstride;
            if (a.u_flag) {
                // host-streamed u lies particle-major (rows of u_cs steps): a strided 8-byte read per particle and step
                // fetches a 64-byte DRAM burst for 8 bytes and costs the LSU one sector per lane (measured: 142 instead
                // of 94 us per step).  Every fourth step each thread reads the whole 32-byte sector of its particles once
                // and lays the four steps down time-major in a small stash (coalesced stores); all steps then read the
                // stash coalesced.  A thread reads back only what it wrote itself: no synchronisation.
                const int tin = t % a.u_cs;
                if ((tin & 3) == 0 || t == 1) {   // (the time loop starts at t = 1: that step fills the first sector)
                    const double* Ug = Ut - (tin & 3);
#pragma unroll
                    for (int kk = 0; kk < KPT; ++kk) {
                        const int i = kk * GT + tid;
                        if (i < nc) {
                            double v0, v1, v2, v3;
                            asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                                         : "=d"(v0), "=d"(v1), "=d"(v2), "=d"(v3)
                                         : "l"(Ug + (size_t)(jb + i) * ujs), "l"(pol_stream));
                            double* sp = a.ustash + jb + i;
                            __stcg(sp, v0);
                            __stcg(sp + N, v1);
                            __stcg(sp + 2 * (size_t)N, v2);
                            __stcg(sp + 3 * (size_t)N, v3);
                        }
                    }
                }
                Ut = a.ustash + (size_t)(tin & 3) * N;
                ujs = 1;
            }
            PEntry* Pt = a.P + (size_t)(t % RP) * N * RS;
            bool bad = false, orphan = false;
#pragma unroll
            for (int k0 = 0; k0 < KPT; k0 += CH) {
                // all loads of CH children are in flight before the first one is used
                double2 xe[CH];
                double uu[CH];
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    const int i = (k0 + u) * GT + tid;
                    xe[u] = make_double2(0.0, 0.0);
                    uu[u] = 0.0;
                    if (i < nc) {
                        int p = bp[k0 + u];
                        if ((unsigned)p >= (unsigned)N) {
                            orphan = true;
                            p = 0;
                        }
                        {
                            const int4 xr = ld_hint_b128(&a.XE[p], pol_xld);
                            xe[u].x = __longlong_as_double(((long long)xr.y << 32) | (long long)(unsigned)xr.x);
                            xe[u].y = __longlong_as_double(((long long)xr.w << 32) | (long long)(unsigned)xr.z);
                        }
                        bp[k0 + u] = ld_hint_b32(&a.perm[p], pol_xld);
                        // (the stash is rewritten every four steps: coherent load, not the read-only path)
                        uu[u] = a.u_flag ? __ldcg(Ut + (size_t)(jb + i) * ujs) : ld_stream_hint_f64(Ut + (size_t)(jb + i) * ujs, pol_stream);
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
                        atomicAdd(&s_fhist[fine_bin(x, mhat, inv_shat)], 1);
                        if (!(a.dbg & 2)) {
                            // residual of the transition parent -> child as the score terms use it (:452-453)
                            double sq = x - mu - phi * (xe[u].x - mu);
                            sq -= sr * xe[u].y * ylag;
                            st_stream_f64x4(&Pt[(size_t)(jb + i) * RS], pol_stream, xe[u].x, sq, xe[u].y * ylag, x);
                        }
                    }
                }
            }
            if (bad) GRID_FLAG(2);
            if (orphan) GRID_FLAG(4);
        }
        PROF_MARK(19);   // A1: gathers + propagation + histogram
        if (HESS && tid == 0) s_ndef = 0;
        __syncthreads();
        {
            int* gh = a.ghist + ((size_t)par * kNCopy + (c % kNCopy)) * kNF;
#pragma unroll
            for (int kk = 0; kk < BPT; ++kk) {
                const int b = kk * GT + tid;
                const int cnt = s_fhist[b];
                if (cnt) atomicAdd(&gh[b], cnt);
            }
        }
        PROF_MARK(3);   // A1 children+hist
        GRID_ARRIVE();   // ---- barrier 1: global histogram complete
        if (!(a.dbg & 1)) {
            // genealogy records (only feed outputs): child = (parent row, parent's ancestors 1..7)
            const REntry* Rp = a.R + (size_t)((t - 1) & 1) * N * RS;
            REntry* Rc = a.R + (size_t)(t & 1) * N * RS;
            // Hessian branch: alpha recursion (:361-390) in the same pass -- the parent's cumulative alpha is the
            // second sector of its record entry.  cumulative alpha of the child = its own term + its parent's.
            // Q7: the "current" state is particles[t - 1 + ancestor] read through the flat layout = sorted value
            // `sl` of time tq (table of the first SQ sorted values of every generation), or the unsorted new value
            // `sl` of this generation (complete only after barrier 1: deferred), or 0.  Q8: obs[t - LAG] wraps.
            const double yi = a.obs[t], ylagH = HESS ? obs_wrap(a.obs, t - L, NOBS) : 0.0;
            REntry* Pc = (REntry*)(a.P + (size_t)(t % RP) * N * RS);
#pragma unroll
            for (int k0 = 0; k0 < KPT; k0 += RB) {
                int r[RB][8];
                double2 pa0[HESS ? RB : 1], pa1[HESS ? RB : 1];
                double cur[HESS ? RB : 1];
#pragma unroll
                for (int u = 0; u < RB; ++u) {
                    const int i = (k0 + u) * GT + tid;
#pragma unroll
                    for (int z = 0; z < 8; ++z) r[u][z] = 0;
                    if constexpr (HESS) {
                        pa0[u] = pa1[u] = make_double2(0.0, 0.0);
                        cur[u] = 0.0;
                    }
                    if (i < nc) {
                        const REntry* src = &Rp[(size_t)((a.dbg & 8) ? jb + i : bp[k0 + u]) * RS];
                        ld_rec(src, pol_rld, r[u]);
                        if constexpr (HESS) {
                            ld_f64x4(src + 1, pol_rld, pa0[u].x, pa0[u].y, pa1[u].x, pa1[u].y);
                            const unsigned qq = (unsigned)(t - 1) + (unsigned)__ldcg(&a.H[jb + i]);
                            const unsigned sl = qq / (unsigned)NOBS, tq = qq - sl * (unsigned)NOBS;
                            if ((int)tq < t) cur[u] = ((int)sl < a.SQ) ? __ldcg(&a.xlow[(size_t)sl * NOBS + tq]) : 0.0;
                            else if ((int)tq == t) s_deflist[atomicAdd(&s_ndef, 1)] = make_int2(i, bp[k0 + u]);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < RB; ++u) {
                    const int i = (k0 + u) * GT + tid;
                    if (i < nc) {
                        const int j = jb + i;
                        const int b = bp[k0 + u];
                        st_rec(&Rc[(size_t)j * RS], pol_rst, b, r[u][0], r[u][1], r[u][2], r[u][3], r[u][4], r[u][5], r[u][6]);
                        // row of the ancestor L-2 steps back (new record = (b, r[0..6]))
                        int anc = j;
                        if (L == 3) anc = b;
                        else if (L > 3) anc = pick8(r[u], L - 4);
                        bp[k0 + u] = anc;
                        if constexpr (HESS) {
                            double al[4];
                            sv_alpha_terms(s_k, xn[k0 + u], cur[u], yi, ylagH, al);
                            const double c0 = al[0] + pa0[u].x, c1 = al[1] + pa0[u].y;
                            const double c2 = al[2] + pa1[u].x, c3 = al[3] + pa1[u].y;
                            st_keep_f64x4(&Rc[(size_t)j * RS + 1], pol_rst, c0, c1, c2, c3);
                            st_stream_f64x4(&Pc[(size_t)j * RS + 1], pol_stream, c0, c1, c2, c3);
                        }
                    }
                }
            }
        }
        {
            // housekeeping for the next step
            constexpr int kZero = kNCopy * kNF;
            const int zper = (kZero + G - 1) / G;
            int* ghn = a.ghist + (size_t)(par ^ 1) * kZero;
            for (int b = c * zper + tid; b < min(kZero, (c + 1) * zper); b += GT) __stcg(&ghn[b], 0);
            if (tid == 0) __stcg(&a.tilecnt[((par ^ 1) * kMaxTiles + c) * kCntStride], 0);
        }
        PROF_MARK(4);   // A1 records
        GRID_WAIT();
        PROF_MARK(5);   // wait 1
        if (s_sc.abort_now) break;
        // Hessian branch, Q7 reads that fall into this generation (ancestor position = 1 mod NOBS: one child in NOBS):
        // the unsorted new value `sl` (complete since barrier 1) if sl <= j, else 0.  The records pass listed these
        // children; one thread per entry redoes the term after the arrive of barrier 2 (outputs only: read by the
        // children of the next step and by the score terms LAG-2 steps later; LAG = 2: before the arrive)
        if constexpr (HESS) {
            if (L == 2) {
                __syncthreads();
                deferred_alpha_pass(a, s_k, s_deflist, min(s_ndef, kCap), t, jb, tid, GT, pol_rld, pol_rst, pol_stream);
            }
        }

        // --------------------------------------------------------------------------------------
        // phase A2: scan of the global histogram: tile boundaries on bin edges, tile of every bin;
        // entries ordered by tile in shared memory, copied to the mailboxes in runs
        // --------------------------------------------------------------------------------------
        {
            const int* gh = a.ghist + (size_t)par * kNCopy * kNF;
            int cnt[BPT];
#pragma unroll
            for (int i = 0; i < BPT; ++i) cnt[i] = 0;
            int prevcnt = 0;
#pragma unroll
            for (int cp = 0; cp < kNCopy; ++cp) {
#pragma unroll
                for (int i = 0; i < BPT; i += 4) {
                    const int4 v = __ldcg((const int4*)(gh + cp * kNF + BPT * tid + i));
                    cnt[i] += v.x;
                    cnt[i + 1] += v.y;
                    cnt[i + 2] += v.z;
                    cnt[i + 3] += v.w;
                }
                if (tid > 0) prevcnt += __ldcg(gh + cp * kNF + BPT * tid - 1);
            }
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
                    s_tbin[k] = kNF;
                }
                if (total != N) GRID_FLAG(5);
            }
        }
        __syncthreads();
        PROF_MARK(20);   // A2: histogram scan + tile map
        int kr[KPT];
        {
            const double mhat = s_sc.mhat, inv_shat = s_sc.inv_shat;
#pragma unroll
            for (int kk = 0; kk < KPT; ++kk) {
                const int i = kk * GT + tid;
                kr[kk] = 0;
                // (the children of a warp spread over ~25 tiles -- the innovation is wide against a tile -- so the plain
                // atomics hardly conflict: one atomic per (warp, tile) through __match_any_sync costs 13.8 -> 17.5 us for
                // this phase, through a ballot loop over the distinct tiles 2 -> 52 us for the counting alone)
                if (i < nc) {
                    const int tl = s_tileof[fine_bin(xn[kk], mhat, inv_shat)];
                    const int r = atomicAdd(&s_tcnt[tl], 1);
                    kr[kk] = (tl << 16) | r;
                }
            }
        }
        __syncthreads();
        PROF_MARK(21);   // A2: slot counting
        if (tid < G) {
            // slots of this CTA's run in every mailbox
            const int cnt = s_tcnt[tid];
            int base = s_tstart[tid];
            if (cnt) base += atomicAdd(&a.tilecnt[(par * kMaxTiles + tid) * kCntStride], cnt);
            s_tbase[tid] = base;
            s_tcnt[tid] = 0;
        }
        __syncthreads();
        PROF_MARK(22);   // A2: global reservation
        {
            MailEntry* mail = (MailEntry*)a.XE;
#pragma unroll
            for (int kk = 0; kk < KPT; ++kk) {
                const int i = kk * GT + tid;
                if (i < nc) {
                    const int pos = s_tbase[kr[kk] >> 16] + (kr[kk] & 0xffff);
                    const long long xb = __double_as_longlong(xn[kk]);
                    const int4 ent = make_int4((int)(xb & 0xffffffffll), (int)(xb >> 32), jb + i, bp[kk]);
                    if (pos >= 0 && pos < N) st_hint_b128(&mail[pos], pol_keep, ent);
                }
            }
        }
        pstart = s_tstart[c];
        n = s_tstart[c + 1] - pstart;
        if (n > kCap || n < 0) {
            if (tid == 0) GRID_FLAG(1);
            n = n < 0 ? 0 : kCap;
        }
        if (tid == 0) {
            // sub-bins of this tile: its histogram bins cover [lo, hi)
            const double mhat = s_sc.mhat, shat = s_sc.shat;
            const double lo = mhat + shat * ((double)s_tbin[c] * kBinW - kZ);
            const double hi = mhat + shat * ((double)s_tbin[c + 1] * kBinW - kZ);
            s_sc.lo = lo;
            s_sc.scale = (hi > lo) ? (double)kNSB / (hi - lo) : 0.0;
            // shift = largest log-weight over the occupied bins (any shift cancels, Q4)
            const double y = a.obs[t];
            const double xmin = mhat + shat * ((double)s_sc.binlo * kBinW - kZ);
            const double xmax = mhat + shat * ((double)(s_sc.binhi + 1) * kBinW - kZ);
            double xs_ = (y != 0.0) ? 2.0 * log(fabs(y)) : xmin;
            xs_ = fmin(fmax(xs_, xmin), xmax);
            const double es_ = exp(-0.5 * xs_);
            s_sc.shift = (-0.91893853320467267 - 0.5 * xs_) - (0.5 * y * y) * (es_ * es_);
            if (c == 0) a.shiftv[t] = s_sc.shift;
        }
        PROF_MARK(6);   // A2 scan+scatter
        GRID_ARRIVE();   // ---- barrier 2: mailboxes complete
        if constexpr (HESS) {
            if (L != 2) deferred_alpha_pass(a, s_k, s_deflist, min(s_ndef, kCap), t, jb, tid, GT, pol_rld, pol_rst, pol_stream);
            PROF_MARK(14);   // deferred alpha terms (Hessian branch)
        }
        for (int b = tid; b < kNSB; b += GT) s_sub[b] = 0;
        if (warp == 0 && nc > 0) {
            // the generation the fixed-lag terms of this step gather from, and the next step's slice of u -> L2
            if (t + 1 < NOBS && !a.u_flag) {
                const double* Un = a.U + (size_t)(t + 1) * N;
                prefetch_range(Un + jb, Un + je, lane, 32);
            }
            if (t >= L && (a.dbg & 8)) {   // (measured: the bulk prefetch of the lagged generation costs more than it saves)
                const PEntry* Pg = a.P + (size_t)((t - (L - 2)) % RP) * N;
                prefetch_range(Pg + jb, Pg + je, lane, 32);
            }
        }
        PROF_MARK(7);   // zero+prefetch
        GRID_WAIT();
        PROF_MARK(8);   // wait 2
        if (s_sc.abort_now) break;

        // --------------------------------------------------------------------------------------
        // phase B: sort this tile (:392-424 / :23-52), weights (:427-442), fixed-lag score terms
        // (:445-470), block scan
        // --------------------------------------------------------------------------------------
        double acc[3] = {0.0, 0.0, 0.0};   // sh * {x, m, m^2}
        {
            const MailEntry* mb = (const MailEntry*)a.XE + pstart;
            const double lo = s_sc.lo, scale = s_sc.scale;
            double bx[KPT];
            int bj[KPT], bae[KPT];
#pragma unroll
            for (int k0 = 0; k0 < KPT; k0 += CH) {
                int4 raw[CH];
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    const int e = (k0 + u) * GT + tid;
                    raw[u] = make_int4(0, 0, 0, 0);
                    if (e < n) raw[u] = ld_hint_b128(mb + e, pol_dead);
                }
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    const int e = (k0 + u) * GT + tid;
                    const double x = __longlong_as_double(((long long)raw[u].y << 32) | (long long)(unsigned)raw[u].x);
                    bx[k0 + u] = x;
                    bj[k0 + u] = raw[u].z;
                    bae[k0 + u] = raw[u].w & 0x1fffff;
                    if (e < n) {
                        const int er = atomicAdd(&s_sub[sub_bin(x, lo, scale)], 1);
                        bae[k0 + u] |= min(er, 2047) << 21;
                    }
                }
            }
            __syncthreads();
            PROF_MARK(23);   // B: mailbox load + sub-bin count
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
            __syncthreads();
            PROF_MARK(24);   // B: sub-bin scan
#pragma unroll
            for (int kk = 0; kk < KPT; ++kk) {
                const int e = kk * GT + tid;
                if (e < n) {
                    const int pos = min(s_sub[sub_bin(bx[kk], lo, scale)] + (int)((unsigned)bae[kk] >> 21), kCap - 1);
                    s_xb[pos] = bx[kk];
                    s_jb[pos] = bj[kk];
                    s_ab[pos] = bae[kk] & 0x1fffff;
                }
            }
            __syncthreads();
        }
        PROF_MARK(9);   // B bin sort
        {
            // sub-bin order -> exact order (by value, then by birth row: :32-35 never returns 0), in place
            const double lo = s_sc.lo, scale = s_sc.scale;
            double bx[KPT];
            int bj[KPT], ba[KPT], np[KPT];
#pragma unroll
            for (int kk = 0; kk < KPT; ++kk) {
                const int q = kk * GT + tid;
                bx[kk] = 0.0;
                bj[kk] = ba[kk] = 0;
                np[kk] = q;
                if (q < n) {
                    const double x = s_xb[q];
                    const int j = s_jb[q];
                    bx[kk] = x;
                    bj[kk] = j;
                    ba[kk] = s_ab[q];
                    const int kb = sub_bin(x, lo, scale);
                    const int b0 = s_sub[kb];
                    const int b1 = (kb + 1 < kNSB) ? s_sub[kb + 1] : n;
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
            __syncthreads();
            PROF_MARK(25);   // B: exact ranks
#pragma unroll
            for (int kk = 0; kk < KPT; ++kk) {
                const int q = kk * GT + tid;
                if (q < n && np[kk] != q) {
                    s_xb[np[kk]] = bx[kk];
                    s_jb[np[kk]] = bj[kk];
                    s_ab[np[kk]] = ba[kk];
                }
            }
            __syncthreads();
            PROF_MARK(26);   // B: reorder
        }
        {
            // weights (:427-437): lw = -0.9189 - x/2 - y^2 exp(-x) / 2, sh = exp(lw - shift);
            // fixed-lag score terms (:445-470): ancestor pair (time t-L+1, t-L+2) from one half sector
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
                    {
                        const long long xb = __double_as_longlong(x), eb = __double_as_longlong(e);
                        st_hint_b128(&a.XE[pstart + q], pol_keep,
                                     make_int4((int)(xb & 0xffffffffll), (int)(xb >> 32), (int)(eb & 0xffffffffll), (int)(eb >> 32)));
                    }
                    st_hint_b32(&a.perm[pstart + q], pol_keep, j);
                    if constexpr (HESS) {
                        if (pstart + q < a.SQ) a.xlow[(size_t)(pstart + q) * NOBS + t] = x;
                    }
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
        __syncthreads();
        PROF_MARK(10);   // B rank+weights+score
        {
            // cumulative weights: thread = Lc consecutive sorted particles, sequential running sum;
            // one block-wide pass gives the offsets of the chunks and the eight weighted sums
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
            __syncthreads();
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
                    ps[7] = 0.0;
                    if (t < L || (a.dbg & 4)) {   // later steps: the score terms below write ps[2..6]
#pragma unroll
                        for (int i = 2; i < 7; ++i) ps[i] = 0.0;
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
        if (t >= L && !(a.dbg & 4)) {
            // fixed-lag score terms (:445-470): weight x monomials of the payload of the ancestor
            // LAG-2 generations back (one random sector each), fixed summation order
            const PEntry* Pg = a.P + (size_t)((t - (L - 2)) % RP) * N * RS;
            double sc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
            // Hessian branch (:472-534): hessian1 is a polynomial in (curr - mu, sq, ey) with constant coefficients, so
            // four more weighted monomials (ey, curr^2, curr ey, ey^2 -> hacc[0..3]) next to the five of the gradient
            // give its ten entries afterwards (grid_hess_finish_kernel); hessian2 needs the cumulative alpha of the
            // same ancestor (the second sector of its payload entry) and is accumulated entry by entry in a second
            // pass (hacc[10..19]; the payloads come from L2 there).  Weights are the unnormalised ones (divided by S
            // afterwards).  The isfinite guards of :521-534 are kept for hessian2; a non-finite particle value
            // abandons the evaluation (phase A1), so hessian1 needs none.
            double hm[HESS ? 4 : 1];
            if constexpr (HESS) {
#pragma unroll
                for (int i = 0; i < 4; ++i) hm[i] = 0.0;
            }
#pragma unroll
            for (int k0 = 0; k0 < KPT; k0 += 4) {
                double pc[4], psq[4], pey[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int q = (k0 + u) * GT + tid;
                    pc[u] = psq[u] = pey[u] = 0.0;
                    if (q < n) {
                        const PEntry* pp = &Pg[(size_t)min(max(s_ab[q], 0), N - 1) * RS];
                        double d3;
                        asm volatile("ld.global.cg.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                                     : "=d"(pc[u]), "=d"(psq[u]), "=d"(pey[u]), "=d"(d3)
                                     : "l"(pp), "l"(HESS ? pol_keep : pol_pld));
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int q = (k0 + u) * GT + tid;
                    if (q < n) {
                        const double wd = s_sh[q];
                        const double ws = wd * psq[u];
                        sc[0] = fma(wd, pc[u], sc[0]);
                        sc[1] += ws;
                        sc[2] = fma(ws, pc[u], sc[2]);
                        sc[3] = fma(ws, psq[u], sc[3]);
                        sc[4] = fma(ws, pey[u], sc[4]);
                        if constexpr (HESS) {
                            const double we = wd * pey[u], wc = wd * pc[u];
                            hm[0] += we;
                            hm[1] = fma(wc, pc[u], hm[1]);
                            hm[2] = fma(wc, pey[u], hm[2]);
                            hm[3] = fma(we, pey[u], hm[3]);
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 5; ++i) sc[i] = warp_sum(sc[i]);
            __syncthreads();
            if constexpr (HESS) {
#pragma unroll
                for (int i = 0; i < 4; ++i) hm[i] = warp_sum(hm[i]);
            }
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < 5; ++i) s_red[32 * i + warp] = sc[i];
                if constexpr (HESS) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) s_redH[32 * i + warp] = hm[i];
                }
            }
            if constexpr (HESS) {
                constexpr int SB = PMMH_GRID_HESS_SB;   // 64-byte entries in flight per thread
                double h2[10];
#pragma unroll
                for (int i = 0; i < 10; ++i) h2[i] = 0.0;
#pragma unroll
                for (int k0 = 0; k0 < KPT; k0 += SB) {
                    double pc[SB], psq[SB], pey[SB];
                    double2 al0[SB], al1[SB];
#pragma unroll
                    for (int u = 0; u < SB; ++u) {
                        const int q = (k0 + u) * GT + tid;
                        pc[u] = psq[u] = pey[u] = 0.0;
                        al0[u] = al1[u] = make_double2(0.0, 0.0);
                        if (q < n) {
                            const PEntry* pp = &Pg[(size_t)min(max(s_ab[q], 0), N - 1) * RS];
                            double d3;
                            asm volatile("ld.global.cg.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                                         : "=d"(pc[u]), "=d"(psq[u]), "=d"(pey[u]), "=d"(d3)
                                         : "l"(pp), "l"(pol_pld));
                            asm volatile("ld.global.cg.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                                         : "=d"(al0[u].x), "=d"(al0[u].y), "=d"(al1[u].x), "=d"(al1[u].y)
                                         : "l"(pp + 1), "l"(pol_pld));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < SB; ++u) {
                        const int q = (k0 + u) * GT + tid;
                        if (q < n) {
                            const double wd = s_sh[q];
                            const double sq = psq[u], ey = pey[u];
                            // g of :454-465 from (curr, sq, ey)
                            const double g0 = s_k.q * sq * s_k.one_m_phi;
                            const double g1 = s_k.q * sq * (pc[u] - s_k.mu) * s_k.one_m_phi2;
                            double g2 = sq;
                            g2 += s_k.sr * ey;
                            g2 *= s_k.q * sq;
                            g2 -= 1.0;
                            double g3 = s_k.rho - s_k.q * s_k.rho * sq * sq;
                            g3 += s_k.inv_sv * sq * ey;
                            const double a0 = al0[u].x, a1 = al0[u].y, a2 = al1[u].x, a3 = al1[u].y;
                            double h;   // :505-519 (the '*' typos of :513,514,517 kept)
#define HESS2_ADD(slot, expr)                           \
    h = (expr);                                         \
    if (isfinite(h)) h2[slot - 10] += h * wd;
                            HESS2_ADD(10, g0 * g0 + 2.0 * a0 * g0)
                            HESS2_ADD(11, g0 * g1 + a0 * g1 + a1 * g0)
                            HESS2_ADD(12, g0 * g2 + a0 * g2 + a2 * g0)
                            HESS2_ADD(13, g0 * g3 + a0 * g3 + a3 * g0)
                            HESS2_ADD(14, g1 * g1 + 2.0 * a1 * g1)
                            HESS2_ADD(15, g1 * g2 + a1 * g2 * a2 * g1)
                            HESS2_ADD(16, g1 * g3 + a1 * g3 * a3 * g1)
                            HESS2_ADD(17, g2 * g2 + 2.0 * a2 * g2)
                            HESS2_ADD(18, g2 * g3 + a2 * g3 * a3 * g2)
                            HESS2_ADD(19, g3 * g3 + 2.0 * a3 * g3)
#undef HESS2_ADD
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 10; ++i) h2[i] = warp_sum(h2[i]);
                if (lane == 0) {
#pragma unroll
                    for (int i = 0; i < 10; ++i) s_redH[32 * (10 + i) + warp] = h2[i];
                }
            }
            __syncthreads();
            if constexpr (HESS) {
                for (int k = warp; k < 20; k += NW) {   // a warp sums component k over the warps (fixed order)
                    const double v = (k < 4 || k >= 10) ? warp_sum((lane < NW) ? s_redH[32 * k + lane] : 0.0) : 0.0;
                    if (lane == 0) a.psumH[((size_t)t * G + c) * 20 + k] = v;
                }
            }
            if (warp == 0) {
                double* ps = a.psum + ((size_t)t * G + c) * 8;
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    const double v = warp_sum((lane < NW) ? s_red[32 * i + lane] : 0.0);
                    if (lane == 0) ps[2 + i] = v;
                }
            }
        }
        PROF_MARK(13);   // score terms
        GRID_WAIT();
        PROF_MARK(12);   // wait 3
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
#undef GRID_POLL
#undef GRID_FENCE
#undef GRID_BACKOFF
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
// HESS: partH[irel][block][0..19] = the hessian1 / hessian2 terms of :564-626 (alpha of the ancestor LAG-2 steps
// back of the final particle, weights W_i, obs[i - LAG] with wrap-around, Q8)
template <bool HESS>
__global__ void __launch_bounds__(256) grid_tail_kernel(GridArgs a, const double* __restrict__ sums,
                                                        double* __restrict__ part, double* __restrict__ partH,
                                                        int nblk) {
    __shared__ double red[(HESS ? 20 : 5) * 32];
    const int tid = threadIdx.x;
    const int L = a.LAG, N = a.N, T = a.NOBS - 1, RP = a.RP;
    const int irel = blockIdx.y, i = a.NOBS - L + irel, idx = L - 1 - irel;
    SvConst k;
    sv_const_init(k, a.params);
    const double y1 = obs_wrap(a.obs, i - 1, a.NOBS);
    const double ST = sums[(size_t)T * 8], Si = sums[(size_t)i * 8];
    const double* shT = a.shring + (size_t)(T % L) * N;
    const double* shi = a.shring + (size_t)(i % L) * N;
    const int RS = a.RS;
    const REntry* Rt = a.R + (size_t)(T & 1) * N * RS;
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    double hacc[HESS ? 20 : 1];
    if constexpr (HESS) {
#pragma unroll
        for (int q = 0; q < 20; ++q) hacc[q] = 0.0;
    }
    const double ylagH = obs_wrap(a.obs, i - L, a.NOBS);
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
            const int row = (m == 0) ? b : min(max(Rt[(size_t)b * RS].a[m - 1], 0), N - 1);
            const PEntry pe = a.P[((size_t)((T - m) % RP) * N + row) * RS];
            curr = pe.c;
            const double pe_n = pe.x;
            acc[0] += wT * curr;
            const double wi = shi[p] / Si;
            double sq, g[4];
            const double ec = exp(-0.5 * curr);
            sv_score_tail_e(k, curr, ec, pe_n, y1, sq, g);   // the tail uses obs[i - 1] (Q6)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[1 + q] += g[q] * wi;
            if constexpr (HESS) {
                const int m2 = L - 2;
                const int row2 = (m2 == 0) ? b : min(max(Rt[(size_t)b * RS].a[m2 - 1], 0), N - 1);
                const double* a4 = (const double*)&a.P[((size_t)((T - m2) % RP) * N + row2) * RS + 1];
                const double al[4] = {a4[0], a4[1], a4[2], a4[3]};
                sv_hessian_terms_e(k, curr, ec, exp(-curr), sq, ylagH, g, al, wi, hacc);
            }
        }
    }
    block_sum<5>(acc, red);
    if (tid < 5) part[((size_t)irel * nblk + blockIdx.x) * 8 + tid] = acc[tid];
    if constexpr (HESS) {
        __syncthreads();
        block_sum<20>(hacc, red);
        if (tid < 20) partH[((size_t)irel * nblk + blockIdx.x) * 20 + tid] = hacc[tid];
    }
}

// Hessian branch: sumsH[t][k] = sum over the tiles of psumH[t][.][k] (t >= lag; 0 before)
__global__ void __launch_bounds__(640) grid_reduceH_kernel(const double* __restrict__ psumH, int G, int lag,
                                                           double* __restrict__ sumsH) {
    const int t = blockIdx.x, k = threadIdx.x >> 5, lane = threadIdx.x & 31;   // 20 warps = 20 components
    double s = 0.0;
    if (t >= lag)
        for (int c = lane; c < G; c += 32) s = s + psumH[((size_t)t * G + c) * 20 + k];
    s = warp_sum(s);
    if (lane == 0) sumsH[(size_t)t * 20 + k] = s;
}

// hessian1 / hessian2 (:521-534, :613-626): main-loop terms divided by the weight sum of their step + tail terms,
// upper triangles expanded into the symmetric 4 x 4 outputs
__global__ void __launch_bounds__(32) grid_hess_finish_kernel(const GridCtrl* __restrict__ ctrl,
                                                              const double* __restrict__ sums,
                                                              const double* __restrict__ sumsH,
                                                              const double* __restrict__ partH,
                                                              const double* __restrict__ params, int nblk, int nobs,
                                                              int lag, double* __restrict__ hess1,
                                                              double* __restrict__ hess2) {
    __shared__ double tri[20];
    const int k = threadIdx.x;
    if (ctrl->status != 0) return;   // abandoned: the general kernel writes the outputs
    if (k < 20) {
        SvConst c;
        sv_const_init(c, params);
        double h = 0.0;
        for (int t = lag; t < nobs; ++t) {
            const double* sm = sums + (size_t)t * 8;
            const double* sh = sumsH + (size_t)t * 20;
            const double S = sm[0];
            double v;
            if (k >= 10) {
                v = sh[k];
            } else {
                // weighted monomials in cm = curr - mu, sq, ey (sums[t][2..6] = c, sq, sq c, sq^2, sq ey; sumsH[t][0..3] =
                // ey, c^2, c ey, ey^2)
                const double A = c.one_m_phi2, B = c.one_m_phi;
                const double Mc = sm[2] - c.mu * S, Ms = sm[3], Me = sh[0];
                const double Mcc = sh[1] - 2.0 * c.mu * sm[2] + c.mu * c.mu * S;
                const double Mcs = sm[4] - c.mu * sm[3], Mce = sh[2] - c.mu * sh[0];
                const double Mss = sm[5], Mse = sm[6], Mee = sh[3];
                switch (k) {
                    case 0: v = -c.q * (B * B) * S; break;                                              // (0,0)
                    case 1: v = -c.q * A * (B * Mc + Ms); break;                                        // (0,1)
                    case 2: v = -2.0 * B * Ms - c.q * B * c.sr * Mse; break;                            // (0,2)
                    case 3: v = 2.0 * c.q * c.rho * B * Ms - c.inv_sv2 * B * c.sigmav * Me; break;      // (0,3)
                    case 4: v = -c.q * A * (2.0 * c.phi * Mcs + A * Mcc); break;                        // (1,1)
                    case 5: v = -c.q * A * (2.0 * Mcs + c.sr * Mce); break;                             // (1,2)
                    case 6: v = c.q * A * (2.0 * c.rho * Mcs - c.sigmav * c.rho_term * Mce); break;     // (1,3)
                    case 7: v = -2.0 * c.q * Mss - c.q * c.sr * Mse - c.q * (c.sr * c.sr) * Mee; break; // (2,2)
                    case 8:                                                                              // (2,3)
                        v = 2.0 * c.q * c.rho * Mss + (2.0 * (c.rho * c.rho) * c.q * c.sigmav + c.inv_sv) * Mse - c.rho * Mee;
                        break;
                    default:                                                                             // (3,3), "sigmav*(-2)" typo kept
                        v = c.rho_term * S - 2.0 * c.q * (c.rho * c.rho) * Mss + 2.0 * c.sigmav * Mss +
                            2.0 * c.inv_sv * c.rho * Mse - c.rho_term * Mee;
                        break;
                }
            }
            h += v / S;
        }
        for (int irel = 0; irel < lag - 1; ++irel)
            for (int q = 0; q < nblk; ++q) h += partH[((size_t)irel * nblk + q) * 20 + k];
        tri[k] = h;
    }
    __syncwarp();
    if (k < 16) {
        const int r = k >> 2, cidx = k & 3;
        const int lo = min(r, cidx), hi = max(r, cidx);
        const int ti = lo * 4 - (lo * (lo - 1)) / 2 + (hi - lo);
        hess1[k] = tri[ti];
        hess2[k] = tri[10 + ti];
    }
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

// sums[t][0] = sum sh, [1] = sum sh x, [2] = sum sh c, [3] = sum sh sq, [4] = sum sh sq c, [5] = sum sh sq^2,
// [6] = sum sh sq ey (c = lagged ancestor value, sq / ey as stored in P);  tail[irel][0..4]
__global__ void grid_finish_kernel(const GridCtrl* __restrict__ ctrl, const double* __restrict__ sums,
                                   const double* __restrict__ shift, const double* __restrict__ xmin,
                                   const double* __restrict__ tail, const double* __restrict__ params,
                                   int nobs, int L, double n_total, double* __restrict__ log_like,
                                   double* __restrict__ filt, double* __restrict__ smo,
                                   double* __restrict__ grad, double* __restrict__ traj,
                                   long long* __restrict__ diag, long long* __restrict__ info) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    // log-likelihood (:537): the terms are computed by the whole first block (the loads and logs of a single thread took
    // 0.29 ms at T = 1000), added up by thread 0 in time order -- the reference's summation order
    __shared__ double s_term[256];
    __shared__ double s_ll;
    if (blockIdx.x == 0) {
        const double logn = log(n_total);
        if (threadIdx.x == 0) s_ll = 0.0;
        for (int base = 1; base < nobs; base += (int)blockDim.x) {
            const int t = base + (int)threadIdx.x;
            if (t < nobs && threadIdx.x < 256) s_term[threadIdx.x] = shift[t] + log(sums[(size_t)t * 8]) - logn;
            __syncthreads();
            if (threadIdx.x == 0) {
                double ll = s_ll;
                const int cnt = min((int)blockDim.x, nobs - base);
                for (int k = 0; k < cnt; ++k) ll += s_term[k];
                s_ll = ll;
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        log_like[0] = s_ll;
        for (int k = 0; k < PMMH_DIAG_COUNT; ++k) diag[k] = 0;
        diag[PMMH_DIAG_NEAR_TIES] = (long long)ctrl->near_ties;
        diag[PMMH_DIAG_MAX_BIN] = ctrl->max_bin;
        diag[PMMH_DIAG_STATUS] = ctrl->status ? 1 : 0;
        diag[PMMH_DIAG_KEY_TIES] = (long long)ctrl->key_ties;
        diag[PMMH_DIAG_SOFT_TIES] = (long long)ctrl->soft_ties;
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
            const double S = sm[0];
            SvConst k;
            sv_const_init(k, params);
            s = sm[2] / S;
            // weighted means of g[0..3] of :454-465 written in the monomials sq, sq c, sq^2, sq ey
            g[0] = k.q * k.one_m_phi * sm[3] / S;
            g[1] = k.q * k.one_m_phi2 * (sm[4] - k.mu * sm[3]) / S;
            g[2] = (k.q * sm[5] + k.q * k.sr * sm[6] - S) / S;
            g[3] = (k.rho * S - k.q * k.rho * sm[5] + k.inv_sv * sm[6]) / S;
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
    size_t ctrl, ghist, tilecnt, tinfo, H, XE, perm, R, P, psum, shiftv, xminv, shring, parentpos, sums,
        tailpart, tail, info, ustash, xlow, psumH, sumsH, tailpartH, total;
    int RP, nblk, SQ;
};

size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

GridLayout grid_layout(int nobs, int n, int lag, int G, int hist, int hess) {
    GridLayout L = GridLayout();
    // ring depth of the payloads: step t writes generation t and gathers from generation t-(lag-2)
    L.RP = lag < 2 ? 2 : lag;
    L.nblk = 296;
    size_t o = 0;
    const size_t N = (size_t)n;
    L.ctrl = o;      o += al256(sizeof(GridCtrl));
    L.ghist = o;     o += al256((size_t)2 * kNCopy * kNF * 4);
    L.tilecnt = o;   o += al256((size_t)2 * kMaxTiles * kCntStride * 4);
    L.tinfo = o;     o += al256((size_t)kMaxTiles * 4 * 8);
    L.H = o;         o += al256(N * 4);
    L.XE = o;        o += al256(N * 16);
    L.perm = o;      o += al256(N * 4);
    const size_t RS = hess ? 2 : 1;   // Hessian branch: 64-byte entries (record | alpha, payload | alpha)
    L.R = o;         o += al256(2 * N * 32 * RS);
    L.P = o;         o += al256((size_t)L.RP * N * 32 * RS);
    L.psum = o;      o += al256((size_t)nobs * G * 8 * 8);
    L.shiftv = o;    o += al256((size_t)nobs * 8);
    L.xminv = o;     o += al256((size_t)nobs * 8);
    L.shring = o;    o += al256((size_t)lag * N * 8);
    L.parentpos = o; o += al256(hist ? N * 4 : 256);
    L.sums = o;      o += al256((size_t)nobs * 8 * 8);
    L.tailpart = o;  o += al256((size_t)lag * L.nblk * 8 * 8);
    L.tail = o;      o += al256((size_t)lag * 8 * 8);
    L.info = o;      o += al256(8 * 8);
    L.ustash = o;    o += al256(4 * N * 8);   // host-streamed u: four time steps, time-major
    // Hessian branch: the first SQ sorted values of every generation (Q7: particles[i - 1 + ancestor] read through
    // the flat layout), per-tile / per-step sums
    L.SQ = hess ? (nobs + n - 2) / nobs + 1 : 1;
    L.xlow = o;      o += al256(hess ? (size_t)nobs * L.SQ * 8 : 0);
    L.psumH = o;     o += al256(hess ? (size_t)nobs * G * 20 * 8 : 0);
    L.sumsH = o;     o += al256(hess ? (size_t)nobs * 20 * 8 : 0);
    L.tailpartH = o; o += al256(hess ? (size_t)lag * L.nblk * 20 * 8 : 0);
    L.total = o;
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

size_t sv_grid_ws_bytes(int nobs, int n, int lag, int G, int hist, int hess) {
    return grid_layout(nobs, n, lag, G, hist, hess).total;
}

int sv_grid_run(const double* d_obs, const double* d_params, const double* d_rvr, const double* d_u, int nobs,
                int n, int lag, int G, double* d_filt, double* d_smo, double* d_ll, double* d_grad, double* d_traj,
                long long* d_diag, double* d_xh, int* d_ah, void* d_ws, size_t ws_bytes, long long* d_prof,
                cudaStream_t st, int u_chunk_steps, const int* d_u_flag, double* d_hess1, double* d_hess2) {
    if (!sv_grid_eligible(nobs, n, lag, G)) return set_error(PMMH_ERR_INVALID, "grid kernel: sizes not eligible");
    const int hist = d_xh != nullptr;
    const int hess = d_hess1 != nullptr && d_hess2 != nullptr;
    const GridLayout L = grid_layout(nobs, n, lag, G, hist, hess);
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
    if (u_chunk_steps > 0) {   // particle-major chunks of u_chunk_steps time steps (host-streamed)
        a.u_cs = u_chunk_steps;
        a.u_cstride = (long long)n * u_chunk_steps;
        a.u_tstride = 1;
        a.u_jstride = u_chunk_steps;
        a.u_flag = d_u_flag;
        a.ustash = (double*)(ws + L.ustash);
    } else {
        a.u_cs = 1 << 30;
        a.u_cstride = 0;
        a.u_tstride = n;
        a.u_jstride = 1;
        a.u_flag = nullptr;
    }
    a.ctrl = (GridCtrl*)(ws + L.ctrl);
    a.ghist = (int*)(ws + L.ghist);
    a.tilecnt = (int*)(ws + L.tilecnt);
    a.tinfo = (double*)(ws + L.tinfo);
    a.H = (int*)(ws + L.H);
    a.XE = (double2*)(ws + L.XE);
    a.perm = (int*)(ws + L.perm);
    a.R = (REntry*)(ws + L.R);
    a.P = (PEntry*)(ws + L.P);
    a.psum = (double*)(ws + L.psum);
    a.shiftv = (double*)(ws + L.shiftv);
    a.xminv = (double*)(ws + L.xminv);
    a.shring = (double*)(ws + L.shring);
    a.parentpos = (int*)(ws + L.parentpos);
    a.xlow = (double*)(ws + L.xlow);
    a.psumH = (double*)(ws + L.psumH);
    a.SQ = L.SQ;
    a.RS = hess ? 2 : 1;
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
    static int threads = 0;
    if (!threads) {
        const char* e = getenv("PMMH_GRID_THREADS");
        threads = (e && atoi(e) == 512) ? 512 : 1024;
    }
    int dev = 0;
    GRID_CUDA(cudaGetDevice(&dev));
    {
        // the two genealogy record tables live in the persisting part of L2 (their reuse distance is a
        // whole time step); the limit is a device-wide setting, set once per device and size
        static size_t persist_set[64] = {0};
        size_t want = 0;
        const char* e = getenv("PMMH_GRID_L2_PERSIST_MB");
        if (e) want = (size_t)atoi(e) << 20;
        int maxp = 0;
        GRID_CUDA(cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, dev));
        if (want > (size_t)maxp) want = (size_t)maxp;
        if (dev >= 0 && dev < 64 && persist_set[dev] != want + 1) {
            GRID_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
            persist_set[dev] = want + 1;
        }
    }
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        GRID_CUDA(cudaFuncSetAttribute(sv_grid_kernel<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem));
        GRID_CUDA(cudaFuncSetAttribute(sv_grid_kernel<1024, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem));
        GRID_CUDA(cudaFuncSetAttribute(sv_grid_kernel<512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem));
        GRID_CUDA(cudaFuncSetAttribute(sv_grid_kernel<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem));
        attr_set[dev] = true;
    }
    void* kargs[] = {(void*)&a};
    const void* kern = hess ? (threads == 1024 ? (const void*)sv_grid_kernel<1024, true> : (const void*)sv_grid_kernel<512, true>)
                            : (threads == 1024 ? (const void*)sv_grid_kernel<1024, false> : (const void*)sv_grid_kernel<512, false>);
    GRID_CUDA(cudaLaunchCooperativeKernel(kern, dim3(G), dim3(threads), kargs, kDynSmem, st));
    double* sums = (double*)(ws + L.sums);
    double* tailpart = (double*)(ws + L.tailpart);
    double* tail = (double*)(ws + L.tail);
    grid_reduce_kernel<<<nobs, 256, 0, st>>>(a.psum, G, sums);
    if (hess) {
        double* sumsH = (double*)(ws + L.sumsH);
        double* tailpartH = (double*)(ws + L.tailpartH);
        grid_reduceH_kernel<<<nobs, 640, 0, st>>>(a.psumH, G, lag, sumsH);
        grid_tail_kernel<true><<<dim3(L.nblk, lag), 256, 0, st>>>(a, sums, tailpart, tailpartH, L.nblk);
        grid_hess_finish_kernel<<<1, 32, 0, st>>>(a.ctrl, sums, sumsH, tailpartH, d_params, L.nblk, nobs, lag, d_hess1, d_hess2);
    } else {
        grid_tail_kernel<false><<<dim3(L.nblk, lag), 256, 0, st>>>(a, sums, tailpart, nullptr, L.nblk);
    }
    grid_tail_reduce_kernel<<<lag, 256, 0, st>>>(tailpart, L.nblk, tail);
    grid_finish_kernel<<<(nobs + 255) / 256, 256, 0, st>>>(a.ctrl, sums, a.shiftv, a.xminv, tail, d_params, nobs, lag,
                                                           (double)n, d_ll, d_filt, d_smo, d_grad, d_traj, d_diag,
                                                           (long long*)(ws + L.info));
    GRID_CUDA(cudaGetLastError());
    return PMMH_OK;
}

// soft ties / raw status of the last evaluation that used this workspace (after synchronisation)
int sv_grid_read_info(const void* d_ws, int nobs, int n, int lag, int G, int hist, long long* h_info) {
    const GridLayout L = grid_layout(nobs, n, lag, G, hist, 0);   // (the Hessian buffers sit behind `info`)
    GRID_CUDA(cudaMemcpy(h_info, (const char*)d_ws + L.info, 2 * sizeof(long long), cudaMemcpyDeviceToHost));
    return PMMH_OK;
}

}  // namespace pmmh
