// sv_fast.cu -- "exchange" kernel for the SV fixed-lag particle smoother: log-likelihood +
// fixed-lag gradient (flps_sv_corr with compute_hessian = 0, stochastic_volatility.pyx:205-655;
// the call the quasi-Newton sampler makes twice per iteration, mh_quasi_newton.py:333,378).
//
// Why a second kernel: the ncu capture of the general kernel (profiles/r1_v0_*) and the B200
// micro-benchmarks (tools/microbench.cu, profiles/r1_microbench.txt) showed that a time step is
// bound by (1) grid-wide barriers (~1.3 us each, 7 per step), (2) global atomics on hot
// addresses (~100 ns per same-address op) and (3) un-coalesced global accesses, which cost one
// LSU wavefront per lane whether they hit L2 (4 us per 2^20) or HBM (25 us per 2^20).  This
// kernel is organised so that a step has TWO barriers, no global atomics and (apart from the
// fixed-lag look-ups) only coalesced global traffic:
//
//   phase A  (children, birth order; CTA c owns children [c N/G, (c+1) N/G))
//     * correlated systematic resampling (:694-715): the CTA stages the cumulative weights of
//       the parents it can need, every parent computes the index of its first child in closed
//       form (exact predicate re-checked), and a block-wide max-scan turns the marks into the
//       ancestor of every child
//     * propagation (:354-358), then the child is routed by value: the chip-wide sort is a
//       one-pass sample sort whose splitters are quantiles of the predicted child distribution
//       (all-gathered first and second moment of the propagation mean, normal shape).  The
//       child record is written to the mailbox of its destination chunk; slots come from
//       shared-memory atomics, the per-(chunk, source) run has a fixed base address
//   barrier 1
//   phase B  (chunks; CTA c owns chunks c, c+G, c+2G, ... so that shape errors of the splitters
//             average out)
//     * reads the runs addressed to the chunk (coalesced), sorts them entirely in shared
//       memory (fine counting sort + all-pairs inside a fine bin), evaluates the log-weights
//       (:427-437), writes the sorted generation as 32-byte records with 256-bit stores, the
//       chunk-local cumulative weights, and accumulates the filter mean, the moments for the
//       next splitters and the fixed-lag smoother terms (:445-470)
//   barrier 2 = all-gather of the chunk totals (weights, counts) and partial sums
//
// A particle's position is "virtual": chunk * kCap + rank inside the chunk.  Dense positions
// (the reference's 0..N-1) are only needed for the tail (:540-562) and for the optional
// history outputs, and follow from the all-gathered chunk counts.
// Every record carries the virtual positions of its ancestors 1..4 steps back, so the
// ancestor LAG-2 steps back is reached with ceil((LAG-2)/4) look-ups instead of LAG-2.
//
// Differences to the reference that stay inside the stated tolerances: sums over particles are
// fixed-order tree sums (deterministic); the weight shift is the maximum of the log-weight
// over the predicted range (the reference's my_max, Q4, picks another element; the shift
// cancels analytically); log N(y; 0, e^{x/2}) is evaluated as -0.9189.. - x/2 - y^2 e^{-x}/2.
// If a chunk or a run overflows its capacity (a degenerate cloud, or a child distribution far
// from normal) the evaluation is abandoned with status 1 and the host code re-runs the general
// kernel (sv_filter.cu) for that problem.
#include <math.h>

#include "common.cuh"
#include "sv_filter.cuh"
#include "sv_math.cuh"

namespace pmmh {

namespace {

constexpr int kThreads = kSvThreads;
constexpr int kCap = kFastCap;           // records per chunk (shared-memory sort capacity)
constexpr int kFineBins = 4096;
constexpr int kWinCap = 8192;            // staged window entries per piece (phase A)
constexpr int kChildBlock = 2 * kThreads;   // children routed per shared-memory staging pass
constexpr int kLutCells = 2048;
constexpr int kMaxLagF = 64;
constexpr int kSlotW = kMaxAllgather;    // doubles per CTA in the exchange buffers
constexpr int kNumSums = 10;             // fx, m1, m2, minx, lag[5], flag

struct __align__(32) Rec {   // one particle of one generation
    double x;      // value
    double xpar;   // value of its parent (time - 1)
    int b[4];      // virtual positions of its ancestors 1, 2, 3, 4 steps back
};
static_assert(sizeof(Rec) == 32, "a record is one 32-byte sector");

// 256-bit global accesses (sm_100: LDG.E.ENL2.256 / STG.E.ENL2.256).  Loads are .cg: the data
// was written by other CTAs of this launch, L1 must not serve it.
__device__ __forceinline__ void st_rec(Rec* p, const Rec& r) {
    const unsigned long long w2 = ((unsigned long long)(unsigned)r.b[1] << 32) | (unsigned)r.b[0];
    const unsigned long long w3 = ((unsigned long long)(unsigned)r.b[3] << 32) | (unsigned)r.b[2];
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(__double_as_longlong(r.x)),
                 "l"(__double_as_longlong(r.xpar)), "l"(w2), "l"(w3)
                 : "memory");
}
__device__ __forceinline__ Rec ld_rec(const Rec* p) {
    long long a, b, c, d;
    asm volatile("ld.global.cg.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    Rec r;
    r.x = __longlong_as_double(a);
    r.xpar = __longlong_as_double(b);
    r.b[0] = (int)(c & 0xffffffffll);
    r.b[1] = (int)(c >> 32);
    r.b[2] = (int)(d & 0xffffffffll);
    r.b[3] = (int)(d >> 32);
    return r;
}

struct FastWs {
    Rec* gen;        // [RING][NV]      sorted generations (virtual positions)
    double* cumloc;  // [NV]            chunk-local inclusive cumulative shifted weights
    double* shtail;  // [LAG][NV]       shifted weights of the last LAG generations (tail)
    Rec* mail;       // [G][NBLK * kChildBlock]  children of every source CTA, per block of
                     //                 kChildBlock children grouped by destination chunk
    int* tab;        // [G][NBLK][ND + 1]  start of every destination's run inside a block
    int* offs;       // [RING][ND + 1]  dense offset of every chunk, per generation
};

__host__ __device__ inline size_t fast_ws_carve(int ND, int G, int RING, int LAG, int NBLK, char* base,
                                                FastWs* w) {
    const size_t NV = (size_t)ND * kCap;
    size_t off = 0;
#define PMMH_CARVE(field, type, count)                   \
    do {                                                 \
        if (w) w->field = (type*)(base + off);           \
        off += sv_align((size_t)(count) * sizeof(type)); \
    } while (0)
    PMMH_CARVE(gen, Rec, (size_t)RING * NV);
    PMMH_CARVE(cumloc, double, NV);
    PMMH_CARVE(shtail, double, (size_t)LAG * NV);
    PMMH_CARVE(mail, Rec, (size_t)G * NBLK * kChildBlock);
    PMMH_CARVE(tab, int, (size_t)G * NBLK * (ND + 1));
    PMMH_CARVE(offs, int, (size_t)RING * (ND + 1));
#undef PMMH_CARVE
    return off;
}

// ---- team exchange: counter barrier + all-gather of K doubles per CTA -------------------
struct TeamF {
    int G, rank;
    unsigned epoch;
    unsigned* ctr;    // one counter per team (zeroed before launch)
    double* slots;    // [2][G][kSlotW]
};

// vals: K doubles in shared memory; out: shared [G * K].  Full barrier for the team; global
// writes made by any thread of the team before the call are visible to all after it.
__device__ __forceinline__ void team_exchange(TeamF& t, const double* vals, int K, double* out) {
    if (t.G == 1) {
        __syncthreads();
        if ((int)threadIdx.x < K) out[threadIdx.x] = vals[threadIdx.x];
        __syncthreads();
        return;
    }
    t.epoch++;
    double* buf = t.slots + (size_t)(t.epoch & 1u) * t.G * kSlotW;
    __syncthreads();
    if ((int)threadIdx.x < K) __stcg(&buf[t.rank * kSlotW + threadIdx.x], vals[threadIdx.x]);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(t.ctr, 1u);
        const unsigned target = t.epoch * (unsigned)t.G;
        while (ld_acquire_u32(t.ctr) < target) {
        }
        __threadfence();
    }
    __syncthreads();
    for (int q = threadIdx.x; q < t.G * K; q += blockDim.x) {
        const int c = q / K, k = q - c * K;
        out[q] = __ldcg(&buf[c * kSlotW + k]);
    }
    __syncthreads();
}

// ---- block-wide scans -------------------------------------------------------------------
// In-place exclusive scan of data[0..n), n <= 4 * blockDim.x; returns the total.
// s_w: shared int[33].
__device__ __forceinline__ int block_excl_scan4(int* data, int n, int* s_w) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int v[4], tsum = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int k = tid * 4 + q;
        v[q] = (k < n) ? data[k] : 0;
        tsum += v[q];
    }
    const int incl = warp_incl_scan(tsum, lane);
    __syncthreads();
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int tv = s_w[lane];
        const int ti = warp_incl_scan(tv, lane);
        __syncwarp();
        s_w[lane] = ti - tv;
        if (lane == 31) s_w[32] = ti;
    }
    __syncthreads();
    int run = s_w[warp] + (incl - tsum);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int k = tid * 4 + q;
        if (k < n) data[k] = run;
        run += v[q];
    }
    const int total = s_w[32];
    __syncthreads();
    return total;
}

// In-place inclusive max-scan of data[0..n) (any n).  s_w: shared int[33].
__device__ __forceinline__ void block_incl_maxscan(int* data, int n, int* s_w) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int carry = INT_MIN;
    for (int base = 0; base < n; base += 4 * kThreads) {
        int v[4];
        int m = INT_MIN;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = base + tid * 4 + q;
            v[q] = (k < n) ? data[k] : INT_MIN;
            m = max(m, v[q]);
            v[q] = m;
        }
        int incl = m;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(kFullMask, incl, d);
            if (lane >= d) incl = max(incl, o);
        }
        int excl = __shfl_up_sync(kFullMask, incl, 1);
        if (lane == 0) excl = INT_MIN;
        __syncthreads();
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int tv = s_w[lane];
            int ti = tv;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(kFullMask, ti, d);
                if (lane >= d) ti = max(ti, o);
            }
            int te = __shfl_up_sync(kFullMask, ti, 1);
            if (lane == 0) te = INT_MIN;
            __syncwarp();
            s_w[lane] = te;
            if (lane == 31) s_w[32] = ti;
        }
        __syncthreads();
        const int pre = max(carry, max(s_w[warp], excl));
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = base + tid * 4 + q;
            if (k < n) data[k] = max(pre, v[q]);
        }
        carry = max(carry, s_w[32]);
        __syncthreads();
    }
}

// Inverse standard normal cdf (Acklam's rational approximation, |rel err| < 1.2e-9).  Only
// used to place the sample-sort splitters; it has to be deterministic, not accurate.
__device__ double inv_norm_cdf(double p) {
    const double a0 = -3.969683028665376e+01, a1 = 2.209460984245205e+02, a2 = -2.759285104469687e+02,
                 a3 = 1.383577518672690e+02, a4 = -3.066479806614716e+01, a5 = 2.506628277459239e+00;
    const double b0 = -5.447609879822406e+01, b1 = 1.615858368580409e+02, b2 = -1.556989798598866e+02,
                 b3 = 6.680131188771972e+01, b4 = -1.328068155288572e+01;
    const double c0 = -7.784894002430293e-03, c1 = -3.223964580411365e-01, c2 = -2.400758277161838e+00,
                 c3 = -2.549732539343734e+00, c4 = 4.374664141464968e+00, c5 = 2.938163982698783e+00;
    const double d0 = 7.784695709041462e-03, d1 = 3.224671290700398e-01, d2 = 2.445134137142996e+00,
                 d3 = 3.754408661907416e+00;
    const double plow = 0.02425;
    if (p < plow) {
        const double q = sqrt(-2.0 * log(p));
        return (((((c0 * q + c1) * q + c2) * q + c3) * q + c4) * q + c5) /
               ((((d0 * q + d1) * q + d2) * q + d3) * q + 1.0);
    }
    if (p <= 1.0 - plow) {
        const double q = p - 0.5, r = q * q;
        return (((((a0 * r + a1) * r + a2) * r + a3) * r + a4) * r + a5) * q /
               (((((b0 * r + b1) * r + b2) * r + b3) * r + b4) * r + 1.0);
    }
    const double q = sqrt(-2.0 * log(1.0 - p));
    return -(((((c0 * q + c1) * q + c2) * q + c3) * q + c4) * q + c5) /
           ((((d0 * q + d1) * q + d2) * q + d3) * q + 1.0);
}

// strict weak order of the chunk sort: value first, then the rest of the record (records that
// compare equal on everything are bit-identical, so their relative order cannot matter)
__device__ __forceinline__ bool rec_less(double xa, double pa, const int4& ba, double xb, double pb,
                                         const int4& bb) {
    if (xa != xb) return xa < xb;
    if (ba.x != bb.x) return ba.x < bb.x;
    if (ba.y != bb.y) return ba.y < bb.y;
    if (ba.z != bb.z) return ba.z < bb.z;
    if (ba.w != bb.w) return ba.w < bb.w;
    return pa < pb;
}

// norm_logpdf(y, 0, exp(x/2)) (:428,659-664) with e = exp(-x/2): -0.5 log(2 pi) - x/2 - y^2 e^2 / 2
__device__ __forceinline__ double logw_e(double x, double e, double half_y2) {
    return (-0.91893853320467267 - 0.5 * x) - half_y2 * (e * e);
}

__device__ __forceinline__ int fine_bin(double x, double lo, double scale) {
    const double t = (x - lo) * scale;
    if (!(t >= 0.0)) return 0;
    if (t >= (double)kFineBins) return kFineBins - 1;
    return (int)t;
}

// development instrumentation: cycles per phase, accumulated by thread 0 of every CTA
#define PROF_MARK(slot)                          \
    do {                                         \
        if (a.prof && threadIdx.x == 0) {        \
            const long long now__ = clock64();   \
            prof_acc[slot] += now__ - prof_t;    \
            prof_t = now__;                      \
        }                                        \
    } while (0)

__global__ void __launch_bounds__(kThreads, 1) sv_fast_kernel(SvArgs a) {
    long long prof_acc[kProfSlots];
    long long prof_t = clock64();
#pragma unroll
    for (int q = 0; q < kProfSlots; ++q) prof_acc[q] = 0;
    extern __shared__ __align__(32) unsigned char dsm_raw[];
    const int N = a.N, NOBS = a.NOBS, LAG = a.LAG, G = a.G, RING = a.RING;
    const int S = a.NSUB, ND = S * G;
    const int KW = 2 * S + kNumSums;    // doubles per CTA per exchange
    const size_t NV = (size_t)ND * kCap;
    const int per_tile = (N + G - 1) / G;
    const int NBLK = (per_tile + kChildBlock - 1) / kChildBlock;

    // ---- dynamic shared memory
    double* s_cw = (double*)dsm_raw;                 // [ND]     chunk weight totals
    double* s_cP = s_cw + ND;                        // [ND + 1] exclusive prefix of s_cw
    double* s_z = s_cP + (ND + 1);                   // [ND + 1] splitters in z space ([0] unused)
    int* s_cc = (int*)(s_z + (ND + 1));              // [ND]     chunk counts
    int* s_coff = s_cc + ND;                         // [ND + 1] dense offsets (this generation)
    int* s_coffp = s_coff + (ND + 1);                // [ND + 1] dense offsets (previous generation)
    int* s_cnt = s_coffp + (ND + 1);                 // [ND]     per-destination counters (phase A)
    int* s_lut = s_cnt + ND;                         // [kLutCells]
    unsigned char* s_union = (unsigned char*)(((size_t)(s_lut + kLutCells) + 31) & ~(size_t)31);
    //   view G: exchange output
    double* s_gather = (double*)s_union;                               // [G * KW]
    //   view A (phase A)
    double* s_cn = (double*)s_union;                                   // [kWinCap + 1]
    Rec* s_stage = (Rec*)(s_cn + kWinCap + 4);                         // [kChildBlock]
    int* s_off = (int*)(s_stage + kChildBlock);                        // [ND + 1]
    int* s_par = s_off + (ND + 2);                                     // [per_tile]
    //   view B (phase B)
    double* s_x = (double*)s_union;                                    // [kCap]
    double* s_xp = s_x + kCap;                                         // [kCap]
    int4* s_b = (int4*)(s_xp + kCap);                                  // [kCap]
    int* s_fh = (int*)(s_b + kCap);                                    // [kFineBins + 8]
    unsigned short* s_slot = (unsigned short*)(s_fh + kFineBins + 8);  // [kCap]
    unsigned short* s_inv = s_slot + kCap;                             // [kCap]
    int* s_roff = (int*)(s_inv + kCap);                                // [G * NBLK + 8] run starts
    int* s_rbase = s_roff + (G * NBLK + 8);                            // [G * NBLK] run base in mail

    __shared__ double s_vals[kSlotW];
    __shared__ double s_tot[16];
    __shared__ double s_red[12 * 32];
    __shared__ double s_w[33], s_wx[33];
    __shared__ int s_iw[33];
    __shared__ int s_misc[8];
    __shared__ double s_dmisc[4];
    __shared__ double s_S[kMaxLagF];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarp = kThreads / 32;

    TeamF tm;
    tm.G = G;
    tm.rank = blockIdx.x % G;
    tm.epoch = 0;
    const int team_id = blockIdx.x / G;
    tm.ctr = (unsigned*)(a.ws + (size_t)team_id * 128);
    tm.slots = (double*)(a.ws + sv_align((size_t)a.n_teams * 128)) + (size_t)team_id * 2 * G * kSlotW;
    char* wsbase = a.ws + a.ws_sync_bytes + (size_t)team_id * a.ws_team_stride;
    const bool lead = (tm.rank == 0);
    const int p0 = min(N, tm.rank * per_tile), p1 = min(N, p0 + per_tile);

    FastWs w;
    fast_ws_carve(ND, G, RING, LAG, NBLK, wsbase, &w);
#define GEN(t) (w.gen + (size_t)((t) % RING) * NV)
#define OFFS(t) (w.offs + (size_t)((t) % RING) * (ND + 1))

    // ---- splitters in z space (constant for the whole launch) and their look-up table
    for (int k = tid; k <= ND; k += kThreads)
        s_z[k] = (k == 0) ? -INFINITY : ((k == ND) ? INFINITY : inv_norm_cdf((double)k / (double)ND));
    __syncthreads();

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

        SvConst c;
        sv_const_init(c, a.params + (size_t)prob * 4);
        const double logN = log((double)N);
        const double invN_exact = 1.0 / (double)N;
        const bool n_pow2 = (N & (N - 1)) == 0;   // then (u + j) / N == (u + j) * (1 / N) exactly

        // ---------------- time 0 (:306-323, Q1): every particle equals mu + stDev * 0.0
        const double stdev0 = c.sigmav / sqrt(1.0 - (c.phi * c.phi));
        const double x0 = c.mu + stdev0 * 0.0;
        for (int l = 0; l < S; ++l) {
            const int k = l * G + tm.rank;
            const int cntk = N / ND + (k < N % ND ? 1 : 0);
            const int offk = k * (N / ND) + min(k, N % ND);
            Rec r;
            r.x = x0;
            r.xpar = x0;
            r.b[0] = r.b[1] = r.b[2] = r.b[3] = 0;
            for (int q = tid; q < cntk; q += kThreads) {
                st_rec(&GEN(0)[(size_t)k * kCap + q], r);
                w.cumloc[(size_t)k * kCap + q] = (double)(q + 1);
                if (Xh) {
                    Xh[offk + q] = x0;
                    Ah[offk + q] = offk + q;
                }
            }
            if (tid == 0) {
                s_vals[2 * l] = (double)cntk;       // weight total (all shifted weights are 1)
                s_vals[2 * l + 1] = (double)cntk;   // count
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
        }
        // moment shift: the propagation mean of the step-1 children is the same for all
        double cshift = (c.mu + c.phi * (x0 - c.mu)) + c.sr * exp(-0.5 * x0) * obs[0];
        for (int k = tid; k <= ND; k += kThreads) {
            s_coff[k] = 0;
            if (prob != team_id)   // a previous problem adapted the splitters: start again from normal quantiles
                s_z[k] = (k == 0) ? -INFINITY : ((k == ND) ? INFINITY : inv_norm_cdf((double)k / (double)ND));
        }
        if (tid == 0) {
            double cn = 0.0;
            for (int l = 0; l < S; ++l) cn += s_vals[2 * l + 1];
            s_vals[2 * S + 0] = cn * x0;   // sum sh * x
            s_vals[2 * S + 1] = 0.0;       // sum sh * (f - cshift)
            s_vals[2 * S + 2] = 0.0;       // sum sh * (f - cshift)^2
            s_vals[2 * S + 3] = x0;        // min x
            for (int q = 4; q < kNumSums; ++q) s_vals[2 * S + q] = 0.0;
        }
        double loglike = 0.0, shift = 0.0;
        long long near_ties = 0, key_ties2 = 0;
        int max_chunk = 0, status = 0;
        long long fail_info = 0;
        team_exchange(tm, s_vals, KW, s_gather);

        for (int i = 0; i < NOBS; ++i) {
            PROF_MARK(0);   // barrier 2 (wait + gather)
            // =========== bookkeeping for generation i (just gathered)
            // chunk tables; chunk k = l * G + cta
            for (int k = tid; k < ND; k += kThreads) {
                const int cta = k % G, l = k / G;
                s_cw[k] = s_gather[cta * KW + 2 * l];
                s_cc[k] = (int)s_gather[cta * KW + 2 * l + 1];
                s_coffp[k] = s_coff[k];
            }
            if (tid == 0) s_coffp[ND] = s_coff[ND];
            for (int q = warp; q < kNumSums; q += nwarp) {
                double s;
                if (q == 3) s = gathered_min(s_gather, KW, 2 * S + q, G, lane);
                else if (q == 9) s = gathered_max(s_gather, KW, 2 * S + q, G, lane);
                else s = gathered_sum(s_gather, KW, 2 * S + q, G, lane);
                if (lane == 0) s_tot[q] = s;
            }
            __syncthreads();
            if (warp == 0) {
                // prefixes over the chunks in chunk order (one warp, sequential chunks of 32:
                // a fixed order, identical in every CTA)
                double carry = 0.0;
                int icarry = 0;
                for (int base = 0; base < ND; base += 32) {
                    const int k = base + lane;
                    const double v = (k < ND) ? s_cw[k] : 0.0;
                    const int iv = (k < ND) ? s_cc[k] : 0;
                    const double incl = warp_incl_scan(v, lane);
                    const int iincl = warp_incl_scan(iv, lane);
                    if (k < ND) {
                        s_cP[k] = carry + (incl - v);
                        s_coff[k] = icarry + (iincl - iv);
                    }
                    carry = carry + __shfl_sync(kFullMask, incl, 31);
                    icarry = icarry + __shfl_sync(kFullMask, iincl, 31);
                }
                if (lane == 0) {
                    s_cP[ND] = carry;
                    s_coff[ND] = icarry;
                }
            }
            __syncthreads();
            if (s_tot[9] > 0.0) {   // a chunk overflowed in phase B (uniform decision)
                status = 1;
                fail_info = 2 | ((long long)i << 8);
                break;
            }
            const double S_i = s_cP[ND];
            if (i >= 1) loglike += shift + log(S_i) - logN;   // :537
            if (tid == 0) s_S[i % kMaxLagF] = S_i;
            if (lead && tid == 0) {
                o_filt[i] = s_tot[0] / S_i;
                o_traj[i] = s_tot[3];   // Q11: traj[i] = X_i[0] (time 0: x0)
                if (i >= LAG) {
                    const int tt = i - LAG + 1;
                    o_smo[tt] = s_tot[4] / S_i;
                    o_grad[tt] = s_tot[5] / S_i;
                    o_grad[NOBS + tt] = s_tot[6] / S_i;
                    o_grad[2 * NOBS + tt] = s_tot[7] / S_i;
                    o_grad[3 * NOBS + tt] = s_tot[8] / S_i;
                }
            }
            if (lead)
                for (int k = tid; k <= ND; k += kThreads) OFFS(i)[k] = s_coff[k];
            {
                int mc = 0;
                for (int k = tid; k < ND; k += kThreads) mc = max(mc, s_cc[k]);
                max_chunk = max(max_chunk, mc);
            }
            if (Xh && i >= 1) {
                // history outputs: dense layout, ancestors as dense positions of generation i-1
                for (int l = 0; l < S; ++l) {
                    const int k = l * G + tm.rank;
                    const int cntk = s_cc[k], offk = s_coff[k];
                    for (int q = tid; q < cntk; q += kThreads) {
                        const Rec r = ld_rec(&GEN(i)[(size_t)k * kCap + q]);
                        Xh[(size_t)i * N + offk + q] = r.x;
                        const int pk = r.b[0] / kCap, pr = r.b[0] - pk * kCap;
                        Ah[(size_t)i * N + offk + q] = s_coffp[pk] + pr;
                    }
                }
            }
            if (i == NOBS - 1) break;

            // moments of the propagation mean => splitters and weight shift of step i + 1
            const double m1 = cshift + s_tot[1] / S_i;
            double var_c = s_tot[2] / S_i - (s_tot[1] / S_i) * (s_tot[1] / S_i);
            if (!(var_c > 0.0)) var_c = 0.0;
            var_c += c.sd * c.sd;
            const double sdc = sqrt(var_c);
            double inv_sdc = 1.0 / sdc;
            if (!isfinite(inv_sdc) || !isfinite(m1)) inv_sdc = 0.0;
            const int inext = i + 1;
            const double y1 = obs[inext - 1], yi = obs[inext];
            const double half_y2 = 0.5 * (yi * yi);
            {
                // shift: maximum of the (concave) log-weight over the predicted range
                double xs = log(yi * yi);
                const double lo = m1 - 6.5 * sdc, hi = m1 + 6.5 * sdc;
                if (!(xs >= lo)) xs = lo;
                if (xs > hi) xs = hi;
                if (!isfinite(xs)) xs = isfinite(m1) ? m1 : 0.0;
                shift = logw_e(xs, exp(-0.5 * xs), half_y2);
            }
            cshift = m1;
            __syncthreads();   // s_gather (view G) is dead from here on

            // ----- splitters of step inext.  In standardised space z = (x - m1) / sdc they start
            // as normal quantiles; afterwards they follow the shape the cloud really has: the
            // chunk counts of generation i give the empirical cdf at the current splitters, and
            // the new splitters are its ND-quantiles (piecewise linear inside a chunk, normal
            // tails in the two unbounded chunks).  Every CTA computes the same values.
            if (i >= 1 && ND > 1) {
                double* s_zn = s_cn;
                for (int k = 1 + tid; k < ND; k += kThreads) {
                    const long long tgt = (long long)k * N;
                    int lo2 = 0, hi2 = ND - 1;   // largest m with coff[m] / N <= k / ND
                    while (lo2 < hi2) {
                        const int mid = (lo2 + hi2 + 1) >> 1;
                        if ((long long)s_coff[mid] * ND <= tgt) lo2 = mid;
                        else hi2 = mid - 1;
                    }
                    const int m = lo2;
                    const double Fm = (double)s_coff[m] / (double)N, Fm1 = (double)s_coff[m + 1] / (double)N;
                    double frac = ((double)k / (double)ND - Fm) / (Fm1 - Fm);
                    if (!(frac >= 0.0)) frac = 0.0;
                    if (frac > 1.0) frac = 1.0;
                    double zn;
                    if (m == 0) {
                        const double pz = 0.5 * erfc(-s_z[1] * 0.70710678118654752);
                        zn = inv_norm_cdf(fmax(frac * pz, 1e-300));
                        if (zn > s_z[1]) zn = s_z[1];
                    } else if (m == ND - 1) {
                        const double pz = 0.5 * erfc(s_z[ND - 1] * 0.70710678118654752);
                        zn = -inv_norm_cdf(fmax((1.0 - frac) * pz, 1e-300));
                        if (zn < s_z[ND - 1]) zn = s_z[ND - 1];
                    } else {
                        zn = s_z[m] + frac * (s_z[m + 1] - s_z[m]);
                    }
                    s_zn[k] = zn;
                }
                __syncthreads();
                for (int k = 1 + tid; k < ND; k += kThreads) s_z[k] = s_zn[k];
                __syncthreads();
            }
            double lut_lo = -1.0, lut_scale = 1.0;
            if (ND > 1) {
                const double zlo = s_z[1] - 1e-9, zhi = s_z[ND - 1] + 1e-9;
                lut_lo = zlo;
                lut_scale = (double)kLutCells / (zhi - zlo);
                if (!isfinite(lut_scale) || !(lut_scale > 0.0)) lut_scale = 0.0;
                for (int cidx = tid; cidx < kLutCells; cidx += kThreads) {
                    const double edge = (lut_scale > 0.0) ? lut_lo + (double)cidx / lut_scale : -INFINITY;
                    int lo2 = 0, hi2 = ND - 1;   // largest d in [0, ND-1] with s_z[d] <= edge (s_z[0] = -inf)
                    while (lo2 < hi2) {
                        const int mid = (lo2 + hi2 + 1) >> 1;
                        if (s_z[mid] <= edge) lo2 = mid;
                        else hi2 = mid - 1;
                    }
                    s_lut[cidx] = lo2;
                }
            }
            __syncthreads();

            PROF_MARK(1);   // bookkeeping, splitters, LUT
            // =========== phase A: resample (:694-715) + propagate (:354-358) + route, step inext
            const double u = rvr[inext];
            const Rec* Gi = GEN(i);
            for (int k = tid; k < ND; k += kThreads) s_cnt[k] = 0;
            for (int k = tid; k < p1 - p0; k += kThreads) s_par[k] = -1;
            int pair_over = 0;
            if (p1 > p0) {
                const double cp_first = n_pow2 ? (u + (double)p0) * invN_exact : (u + (double)p0) / (double)N;
                const double cp_last =
                    n_pow2 ? (u + (double)(p1 - 1)) * invN_exact : (u + (double)(p1 - 1)) / (double)N;
                // chunk-level bracket: first chunk whose inclusive cumulative weight reaches cp
                if (warp == 0) {
                    int k_lo = ND - 1, k_hi = ND - 1;
                    for (int k = lane; k < ND; k += 32) {
                        const double endc = (s_cP[k] + s_cw[k]) / S_i;
                        if (s_cc[k] > 0 && endc >= cp_first) k_lo = min(k_lo, k);
                        if (s_cc[k] > 0 && endc >= cp_last) k_hi = min(k_hi, k);
                    }
                    k_lo = -warp_max(-k_lo);
                    k_hi = -warp_max(-k_hi);
                    if (lane == 0) {
                        s_misc[0] = k_lo;
                        s_misc[1] = k_hi;
                    }
                }
                __syncthreads();
                const int k_lo = s_misc[0], k_hi = s_misc[1];
                // walk the window chunk by chunk, in pieces of kWinCap entries
                for (int k = k_lo; k <= k_hi; ++k) {
                    const int cntk = s_cc[k];
                    if (cntk == 0) continue;
                    const double Pk = s_cP[k];
                    const double* cl = w.cumloc + (size_t)k * kCap;
                    for (int e0 = 0; e0 < cntk; e0 += kWinCap) {
                        const int len = min(kWinCap, cntk - e0);
                        __syncthreads();
                        // s_cn[0] = normalised cumulative weight just before entry e0
                        if (tid == 0) s_cn[0] = (e0 == 0) ? (Pk / S_i) : ((Pk + __ldcg(&cl[e0 - 1])) / S_i);
                        for (int q = tid; q < len; q += kThreads)
                            s_cn[q + 1] = (Pk + __ldcg(&cl[e0 + q])) / S_i;
                        __syncthreads();
                        for (int q = tid; q < len; q += kThreads) {
                            // entry (k, e0 + q): index of its first possible child
                            const double cprev = s_cn[q];
                            // the very first particle of the generation serves every child whose
                            // cp is below its cumulative weight: its first child is child 0
                            const bool first_ever = (e0 + q == 0) && (s_coff[k] == 0);
                            int fc;
                            if (first_ever) {
                                fc = 0;
                            } else if (!(cprev == cprev)) {
                                fc = N;   // NaN weights: no children (the evaluation fails below)
                            } else {
                                // smallest j >= 0 with cp_j > cprev  (i.e. NOT cum >= cp_j)
                                double guess = cprev * (double)N - u;
                                if (!(guess > -1.0)) guess = -1.0;
                                if (guess > (double)N) guess = (double)N;
                                fc = (int)floor(guess) + 1;
                                if (fc < 0) fc = 0;
                                if (fc > N) fc = N;
                                while (fc > 0) {
                                    const double cpm = n_pow2 ? (u + (double)(fc - 1)) * invN_exact
                                                              : (u + (double)(fc - 1)) / (double)N;
                                    if (cpm > cprev) --fc;
                                    else break;
                                }
                                while (fc < N) {
                                    const double cpj = n_pow2 ? (u + (double)fc) * invN_exact
                                                              : (u + (double)fc) / (double)N;
                                    if (cpj > cprev) break;
                                    ++fc;
                                }
                                // diagnostics: decisions within 64 ulp of a cumulative-weight tie,
                                // counted by the CTA that owns the child
                                const double tol = 64.0 * 2.220446049250313e-16 * cprev;
                                if (fc - 1 >= p0 && fc - 1 < p1) {
                                    const double cpa = n_pow2 ? (u + (double)(fc - 1)) * invN_exact
                                                              : (u + (double)(fc - 1)) / (double)N;
                                    if (fabs(cprev - cpa) <= tol) near_ties++;
                                }
                                if (fc >= p0 && fc < p1) {
                                    const double cpb = n_pow2 ? (u + (double)fc) * invN_exact
                                                              : (u + (double)fc) / (double)N;
                                    if (fabs(cpb - cprev) <= tol) near_ties++;
                                }
                            }
                            if (fc < p1) atomicMax(&s_par[max(fc, p0) - p0], k * kCap + e0 + q);
                        }
                    }
                }
                __syncthreads();
                PROF_MARK(2);   // window: first-child marks
                block_incl_maxscan(s_par, p1 - p0, s_iw);
            }
            PROF_MARK(3);   // max-scan
            // children, in blocks of kChildBlock: propagate, pick the destination chunk, group the
            // block by destination in shared memory, write it out coalesced together with the
            // table of run starts
            {
                const double* Ui = U + (size_t)inext * N;
                Rec* my_mail = w.mail + (size_t)tm.rank * NBLK * kChildBlock;
                int* my_tab = w.tab + (size_t)tm.rank * NBLK * (ND + 1);
                for (int blk = 0; blk < NBLK; ++blk) {
                    const int jb = p0 + blk * kChildBlock;
                    const int nb = max(0, min(kChildBlock, p1 - jb));
                    Rec r[2];
                    int dst[2], slot[2];
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        const int j = jb + m * kThreads + tid;
                        dst[m] = -1;
                        slot[m] = 0;
                        if (j < p1) {
                            int vp = s_par[j - p0];
                            if (vp < 0) {   // no parent found (non-finite weights): abandon
                                vp = 0;
                                pair_over = 1;
                            }
                            const double un = ld_stream_f64(&Ui[j]);
                            const Rec pr = ld_rec(&Gi[vp]);
                            const double xp = pr.x;
                            double mean = c.mu + c.phi * (xp - c.mu);
                            mean += c.sr * exp(-0.5 * xp) * y1;
                            const double xn = mean + c.sd * un;
                            int d = 0;
                            if (ND > 1) {
                                const double zx = (xn - m1) * inv_sdc;
                                const double tt = (zx - lut_lo) * lut_scale;
                                const int cell =
                                    (tt >= 0.0) ? ((tt < (double)kLutCells) ? (int)tt : kLutCells - 1) : 0;
                                d = s_lut[cell];
                                while (d + 1 < ND && zx >= s_z[d + 1]) ++d;
                                while (d > 0 && zx < s_z[d]) --d;
                            }
                            dst[m] = d;
                            slot[m] = atomicAdd(&s_cnt[d], 1);
                            r[m].x = xn;
                            r[m].xpar = xp;
                            r[m].b[0] = vp;
                            r[m].b[1] = pr.b[0];
                            r[m].b[2] = pr.b[1];
                            r[m].b[3] = pr.b[2];
                        }
                    }
                    __syncthreads();
                    // run starts of this block = exclusive scan of the destination counters
                    for (int k = tid; k < ND; k += kThreads) s_off[k] = s_cnt[k];
                    __syncthreads();
                    block_excl_scan4(s_off, ND, s_iw);
                    if (tid == 0) s_off[ND] = nb;
                    __syncthreads();
#pragma unroll
                    for (int m = 0; m < 2; ++m)
                        if (dst[m] >= 0) s_stage[s_off[dst[m]] + slot[m]] = r[m];
                    for (int k = tid; k < ND; k += kThreads) s_cnt[k] = 0;
                    __syncthreads();
                    for (int q = tid; q < nb; q += kThreads) st_rec(&my_mail[(size_t)blk * kChildBlock + q], s_stage[q]);
                    for (int k = tid; k <= ND; k += kThreads) my_tab[(size_t)blk * (ND + 1) + k] = s_off[k];
                    __syncthreads();
                }
            }
            pair_over = __syncthreads_or(pair_over);
            PROF_MARK(4);   // children: propagate + route + write
            if (tid == 0) s_vals[0] = (double)pair_over;
            team_exchange(tm, s_vals, 1, s_gather);   // barrier 1
            {
                double f = 0.0;
                for (int cc = tid; cc < G; cc += kThreads) f = fmax(f, s_gather[cc]);
                const int any = __syncthreads_or(f > 0.0);
                if (any) {
                    status = 1;   // a run overflowed (uniform decision)
                    if (lead && tid == 0) {   // diagnostics of the failing step
                        a.hess1[(size_t)prob * 16 + 0] = m1;
                        a.hess1[(size_t)prob * 16 + 1] = sdc;
                        a.hess1[(size_t)prob * 16 + 2] = S_i;
                        a.hess1[(size_t)prob * 16 + 3] = s_tot[1];
                        a.hess1[(size_t)prob * 16 + 4] = s_tot[2];
                        a.hess1[(size_t)prob * 16 + 5] = s_tot[0] / S_i;
                        a.hess1[(size_t)prob * 16 + 6] = c.sd;
                        for (int q = 0; q < 9 && q < ND; ++q) a.hess2[(size_t)prob * 16 + q] = (double)s_cnt[q * (ND / 9 > 0 ? ND / 9 : 1)];
                    }
                    fail_info = 1 | ((long long)inext << 8);
                    break;
                }
            }

            PROF_MARK(5);   // barrier 1
            // =========== phase B: sort my chunks of generation inext in shared memory
            Rec* Gn = GEN(inext);
            const double yl = (inext >= LAG) ? obs[inext - LAG] : 0.0;   // Q5
            const int K = LAG - 2;
            double acc[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) acc[q] = 0.0;
            double minx = INFINITY;
            int chunk_over = 0;
            for (int l = 0; l < S; ++l) {
                const int k = l * G + tm.rank;
                // (1) the runs addressed to chunk k: one per (source CTA, child block)
                __syncthreads();
                const int NR = G * NBLK;
                for (int q = tid; q < NR; q += kThreads) {
                    const int* tb = w.tab + (size_t)q * (ND + 1) + k;
                    const int st = __ldcg(&tb[0]), en = __ldcg(&tb[1]);
                    s_roff[q] = en - st;
                    s_rbase[q] = q * kChildBlock + st;
                }
                __syncthreads();
                int cnt = block_excl_scan4(s_roff, NR, s_iw);
                if (tid == 0) s_roff[NR] = cnt;
                if (cnt > kCap) {
                    chunk_over = 1;
                    cnt = 0;
                }
                __syncthreads();
                // (2) load the records, key range
                double kmn = INFINITY, kmx = -INFINITY;
                for (int q = tid; q < cnt; q += kThreads) {
                    int lo2 = 0, hi2 = NR - 1;   // largest run with s_roff[run] <= q
                    while (lo2 < hi2) {
                        const int mid = (lo2 + hi2 + 1) >> 1;
                        if (s_roff[mid] <= q) lo2 = mid;
                        else hi2 = mid - 1;
                    }
                    const Rec r = ld_rec(&w.mail[(size_t)s_rbase[lo2] + (q - s_roff[lo2])]);
                    s_x[q] = r.x;
                    s_xp[q] = r.xpar;
                    s_b[q] = make_int4(r.b[0], r.b[1], r.b[2], r.b[3]);
                    kmn = fmin(kmn, r.x);
                    kmx = fmax(kmx, r.x);
                }
                PROF_MARK(6);   // phase B: run table + record load
                for (int q = tid; q <= kFineBins; q += kThreads) s_fh[q] = 0;
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
                if (cnt > 0) minx = fmin(minx, fmin_k);
                // (3) fine counting sort
                int fb[kCap / kThreads], rf[kCap / kThreads];
#pragma unroll
                for (int m = 0; m < kCap / kThreads; ++m) {
                    const int q = tid + m * kThreads;
                    fb[m] = 0;
                    rf[m] = 0;
                    if (q < cnt) {
                        fb[m] = fine_bin(s_x[q], fmin_k, fscale);
                        rf[m] = atomicAdd(&s_fh[fb[m]], 1);
                    }
                }
                __syncthreads();
                block_excl_scan4(s_fh, kFineBins, s_iw);
                if (tid == 0) s_fh[kFineBins] = cnt;
                __syncthreads();
#pragma unroll
                for (int m = 0; m < kCap / kThreads; ++m) {
                    const int q = tid + m * kThreads;
                    if (q < cnt) s_slot[s_fh[fb[m]] + rf[m]] = (unsigned short)q;
                }
                __syncthreads();
                // (4) order inside each fine bin (all pairs; ~1 record per bin on average)
                for (int s = tid; s < cnt; s += kThreads) {
                    const int q = s_slot[s];
                    const double key = s_x[q];
                    const int f = fine_bin(key, fmin_k, fscale);
                    const int st = s_fh[f], en = s_fh[f + 1];
                    int rank = 0;
                    if (en - st > 1) {
                        const double kp = s_xp[q];
                        const int4 kb = s_b[q];
                        for (int o = st; o < en; ++o) {
                            if (o == s) continue;
                            const int q2 = s_slot[o];
                            const double x2 = s_x[q2];
                            bool lt = rec_less(x2, s_xp[q2], s_b[q2], key, kp, kb);
                            if (x2 == key) {
                                key_ties2++;
                                // bit-identical records: any fixed order will do
                                if (!lt && !rec_less(key, kp, kb, x2, s_xp[q2], s_b[q2])) lt = (o < s);
                            }
                            if (lt) rank++;
                        }
                    }
                    s_inv[st + rank] = (unsigned short)q;
                }
                __syncthreads();
                PROF_MARK(7);   // phase B: shared-memory sort
                // (5) sorted order: weights (:427-437), record + cumulative weight writes, sums
                double carry = 0.0;
                for (int rowb = 0; rowb < cnt; rowb += kThreads) {
                    const int r = rowb + tid;
                    const bool valid = r < cnt;
                    double shv = 0.0;
                    if (valid) {
                        const int q = s_inv[r];
                        Rec rec;
                        rec.x = s_x[q];
                        rec.xpar = s_xp[q];
                        const int4 bb = s_b[q];
                        rec.b[0] = bb.x;
                        rec.b[1] = bb.y;
                        rec.b[2] = bb.z;
                        rec.b[3] = bb.w;
                        const double e = exp(-0.5 * rec.x);
                        shv = exp(logw_e(rec.x, e, half_y2) - shift);
                        if (!isfinite(shv)) shv = 0.0;
                        const size_t vp = (size_t)k * kCap + r;
                        st_rec(&Gn[vp], rec);
                        if (inext >= NOBS - LAG) w.shtail[(size_t)(inext - (NOBS - LAG)) * NV + vp] = shv;
                        const double sx = shv * rec.x;
                        if (isfinite(sx)) acc[0] += sx;
                        // propagation mean of the next step (for the splitters)
                        const double df = ((c.mu + c.phi * (rec.x - c.mu)) + c.sr * e * yi) - cshift;
                        const double sdf = shv * df;
                        if (isfinite(sdf)) {
                            acc[1] += sdf;
                            acc[2] += sdf * df;
                        }
                        if (inext >= LAG) {
                            // fixed-lag smoother terms (:445-470): ancestor LAG-2 steps back
                            Rec cur = rec;
                            int tcur = inext, rem = K;
                            while (rem > 4) {
                                tcur -= 4;
                                cur = ld_rec(&GEN(tcur)[cur.b[3]]);
                                rem -= 4;
                            }
                            if (rem > 0) {
                                tcur -= rem;
                                cur = ld_rec(&GEN(tcur)[cur.b[rem - 1]]);
                            }
                            double sq, g[4];
                            sv_score_main(c, cur.xpar, cur.x, yl, sq, g);
                            acc[4] += shv * cur.xpar;
                            acc[5] += g[0] * shv;
                            acc[6] += g[1] * shv;
                            acc[7] += g[2] * shv;
                            acc[8] += g[3] * shv;
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
                    if (valid) w.cumloc[(size_t)k * kCap + r] = carry + (s_wx[warp] + incl);
                    // the chunk total must equal the last cumulative value bit for bit
                    const int last = min(cnt - 1 - rowb, kThreads - 1);
                    if (tid == last) s_dmisc[3] = carry + (s_wx[warp] + incl);
                    __syncthreads();
                    carry = s_dmisc[3];
                    __syncthreads();
                }
                if (tid == 0) {
                    s_vals[2 * l] = carry;
                    s_vals[2 * l + 1] = (double)cnt;
                }
                PROF_MARK(8);   // phase B: weights, record writes, fixed-lag look-ups
            }
            {
                double sums[8] = {acc[0], acc[1], acc[2], acc[4], acc[5], acc[6], acc[7], acc[8]};
                block_sum<8>(sums, s_red);
                minx = warp_min(minx);
                if (lane == 0) s_w[warp] = minx;
                chunk_over = __syncthreads_or(chunk_over);
                if (warp == 0) {
                    double v = s_w[lane];
                    v = warp_min(v);
                    if (lane == 0) s_vals[2 * S + 3] = v;
                }
                if (tid == 0) {
                    s_vals[2 * S + 0] = sums[0];
                    s_vals[2 * S + 1] = sums[1];
                    s_vals[2 * S + 2] = sums[2];
                    s_vals[2 * S + 4] = sums[3];
                    s_vals[2 * S + 5] = sums[4];
                    s_vals[2 * S + 6] = sums[5];
                    s_vals[2 * S + 7] = sums[6];
                    s_vals[2 * S + 8] = sums[7];
                    s_vals[2 * S + 9] = (double)chunk_over;
                }
            }
            PROF_MARK(9);   // phase B: block sums
            team_exchange(tm, s_vals, KW, s_gather);   // barrier 2
        }   // time loop
        PROF_MARK(10);

        // ---------------- tail (:540-562, Q6), dense positions
        if (status == 0) {
            const int T = NOBS - 1;
            const double S_T = s_S[T % kMaxLagF];
            int* s_offT = (int*)s_union;   // [ND + 2] dense chunk offsets of generation T
            int* s_offI = s_offT + (ND + 2);
            for (int k = 0; k < LAG; ++k) {
                const int ip = T - k;
                __syncthreads();
                for (int q = tid; q <= ND; q += kThreads) {
                    s_offT[q] = __ldcg(&OFFS(T)[q]);
                    s_offI[q] = __ldcg(&OFFS(ip)[q]);
                }
                __syncthreads();
                double tacc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
                const double S_ip = s_S[ip % kMaxLagF];
                const double y1 = obs_wrap(obs, ip - 1, NOBS);
                const double* shT = w.shtail + (size_t)(T - (NOBS - LAG)) * NV;
                const double* shI = w.shtail + (size_t)(ip - (NOBS - LAG)) * NV;
                for (int j = p0 + tid; j < p1; j += kThreads) {
                    // dense -> virtual at time T and at time ip
                    int lo2 = 0, hi2 = ND - 1;
                    while (lo2 < hi2) {
                        const int mid = (lo2 + hi2 + 1) >> 1;
                        if (s_offT[mid] <= j) lo2 = mid;
                        else hi2 = mid - 1;
                    }
                    const int vpT = lo2 * kCap + (j - s_offT[lo2]);
                    lo2 = 0;
                    hi2 = ND - 1;
                    while (lo2 < hi2) {
                        const int mid = (lo2 + hi2 + 1) >> 1;
                        if (s_offI[mid] <= j) lo2 = mid;
                        else hi2 = mid - 1;
                    }
                    const int vpI = lo2 * kCap + (j - s_offI[lo2]);
                    int b = vpT, bprev = vpT;
                    for (int h = 0; h < k; ++h) {
                        bprev = b;
                        b = ld_rec(&GEN(T - h)[b]).b[0];
                    }
                    const double curr = ld_rec(&GEN(ip)[b]).x;
                    double sT = __ldcg(&shT[vpT]);
                    if (!isfinite(sT)) sT = 0.0;
                    tacc[0] += (sT / S_T) * curr;
                    if (k >= 1) {
                        const double next = ld_rec(&GEN(ip + 1)[bprev]).x;
                        double sq, g[4];
                        sv_score_tail(c, curr, next, y1, sq, g);
                        double si = __ldcg(&shI[vpI]);
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
                double* s_g2 = (double*)(s_offI + (ND + 2));
                s_g2 = (double*)(((size_t)s_g2 + 15) & ~(size_t)15);
                team_exchange(tm, s_vals, 5, s_g2);
                if (warp < 5) {
                    const double s = gathered_sum(s_g2, 5, warp, G, lane);
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
                o_diag[kDiagStatus] = status;
                o_diag[kDiagWavefront] = 0;
                o_diag[kDiagTrajIdx] = 0;
                o_diag[kDiagKernel] = 2;
                o_diag[kDiagFastInfo] = fail_info;
            }
            if (tid < 16 && status == 0) {
                a.hess1[(size_t)prob * 16 + tid] = 0.0;
                a.hess2[(size_t)prob * 16 + tid] = 0.0;
            }
        }
        {
            double nt[3] = {(double)near_ties, (double)key_ties2, 0.0};
            block_sum<3>(nt, s_red);
            int mc = max_chunk;
            mc = warp_max(mc);
            __syncthreads();
            if (lane == 0) s_iw[warp] = mc;
            __syncthreads();
            if (tid == 0) {
                int m2 = 0;
                for (int q = 0; q < nwarp; ++q) m2 = max(m2, s_iw[q]);
                s_vals[0] = nt[0];
                s_vals[1] = nt[1];
                s_vals[2] = (double)m2;
                s_vals[3] = 0.0;
            }
            double* s_g2 = (double*)s_union;
            team_exchange(tm, s_vals, 4, s_g2);
            if (warp < 2) {
                const double s = gathered_sum(s_g2, 4, warp, G, lane);
                if (lane == 0 && lead)
                    o_diag[warp == 0 ? kDiagNearTies : kDiagKeyTies] = (long long)(warp == 0 ? s : s * 0.5);
            } else if (warp == 2) {
                const double s = gathered_max(s_g2, 4, 2, G, lane);
                if (lane == 0 && lead) o_diag[kDiagMaxBin] = (long long)s;
            } else if (warp == 3) {
                const double s = gathered_max(s_g2, 4, 3, G, lane);
                if (lane == 0 && lead) o_diag[kDiagFastInfo] = fail_info | ((long long)s << 32);
            }
            __syncthreads();
        }
    }   // problem loop
    PROF_MARK(11);   // tail + outputs
    if (a.prof && threadIdx.x == 0)
        for (int q = 0; q < kProfSlots; ++q) a.prof[(size_t)blockIdx.x * kProfSlots + q] += prof_acc[q];
#undef GEN
#undef OFFS
}

}  // namespace

int sv_fast_nsub(int N, int G) {
    const int per = (N + G - 1) / G;
    int s = (per + kFastFill - 1) / kFastFill;
    return s < 1 ? 1 : s;
}

static int fast_nblk(int N, int G) {
    const int per_tile = (N + G - 1) / G;
    return (per_tile + kChildBlock - 1) / kChildBlock;
}

size_t sv_fast_ws_bytes(int N, int G, int S, int RING, int LAG) {
    return fast_ws_carve(S * G, G, RING, LAG, fast_nblk(N, G), nullptr, nullptr);
}

size_t sv_fast_sync_bytes(int G, int n_teams) {
    return sv_align((size_t)n_teams * 128) + sv_align((size_t)n_teams * 2 * G * kSlotW * sizeof(double));
}

int sv_fast_smem_bytes(int N, int G, int S) {
    const int ND = S * G;
    const int KW = 2 * S + kNumSums;
    const int per_tile = (N + G - 1) / G;
    size_t head = (size_t)(ND + (ND + 1) + (ND + 1)) * sizeof(double) +
                  (size_t)(ND + (ND + 1) + (ND + 1) + ND + kLutCells) * sizeof(int) + 32;
    const size_t viewG = (size_t)G * KW * sizeof(double);
    const int NR = G * fast_nblk(N, G);
    const size_t viewA = (size_t)(kWinCap + 4) * sizeof(double) + (size_t)kChildBlock * sizeof(Rec) +
                         (size_t)(ND + 2 + per_tile) * sizeof(int);
    const size_t viewB = (size_t)kCap * (8 + 8 + 16) + (size_t)(kFineBins + 8) * sizeof(int) +
                         (size_t)2 * kCap * sizeof(unsigned short) + (size_t)(2 * NR + 8) * sizeof(int);
    const size_t viewT = (size_t)2 * (ND + 2) * sizeof(int) + 16 + (size_t)G * 5 * sizeof(double);
    size_t v = viewG;
    if (viewA > v) v = viewA;
    if (viewB > v) v = viewB;
    if (viewT > v) v = viewT;
    return (int)(head + v + 64);
}

cudaError_t sv_fast_launch(const SvArgs& a, int grid, cudaStream_t stream) {
    const int smem = sv_fast_smem_bytes(a.N, a.G, a.NSUB);
    cudaError_t err = cudaFuncSetAttribute(sv_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return err;
    void* kargs[] = {(void*)&a};
    return cudaLaunchCooperativeKernel((void*)sv_fast_kernel, dim3(grid), dim3(kThreads), kargs, smem, stream);
}

}  // namespace pmmh
