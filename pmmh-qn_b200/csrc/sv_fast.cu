// sv_fast.cu -- "exchange" kernel for the SV fixed-lag particle smoother: log-likelihood +
// fixed-lag gradient (flps_sv_corr with compute_hessian = 0, stochastic_volatility.pyx:205-655;
// the call the quasi-Newton sampler makes twice per iteration, mh_quasi_newton.py:333,378).
//
// One persistent cooperative launch runs all T time steps.  A team of G CTAs (one per SM) owns
// one problem.  A time step has TWO team-wide exchanges and no global atomics:
//
//   phase A  (parents -> children; every CTA works on the parents it already holds)
//     The sorted generation is cut into ND = S * G value-range chunks; CTA c owns S of them,
//     interleaved over the range (mirrored on odd rounds so that linear trends of the weight
//     function cancel) -- its share of the children is then close to 1/G whatever the weights.
//     After the exchange every CTA knows the weight total of every chunk, hence the global
//     cumulative weight in front of its own chunks.  Correlated systematic resampling
//     (:694-715) is evaluated parent-side in closed form: parent p owns the children
//     [ub(p-1), ub(p)), ub(p) = number of thresholds (u + j) / N at or below its cumulative
//     weight (exact predicate re-checked).  Children are generated in birth order from
//     coalesced reads of u, propagated (:354-358) and routed by value to the CTA that will
//     sort them: the chip-wide sort is a one-pass sample sort whose splitters follow the
//     empirical cdf of the previous generation blended with the predicted weight (a chunk holds
//     neither much more than its share of the particles nor of the weight).  The child record
//     (32 bytes = one sector: value, parent value, compact ids of the ancestors 1..4 steps back)
//     is written straight into the destination's mailbox region; its slot is deterministic:
//     (source CTA, source warp, stable rank among that warp's children for that destination,
//     from __match_any_sync) -- the phase has no block-wide synchronisation.
//   exchange 1 (barrier; the per-(destination, source) counts travel through global memory)
//   phase B  (arrivals -> sorted generation; CTA c works on what was routed to it)
//     Records never move again: a particle's id is (destination, arrival index).  The CTA reads
//     its runs, writes the dense tables later steps look up (id 4 steps back; value + parent
//     value), evaluates log-weights (:427-437), the fixed-lag score terms (:445-470; the ancestor
//     LAG-2 steps back is reached through the carried ids and the id table in ceil((LAG-2)/4)
//     look-ups) and all weighted sums in arrival order, and sorts (key, arrival
//     index) pairs in shared memory: counting sort on the chunk-relative position followed by
//     an all-pairs pass inside a bin.  A block scan in sorted order gives the cumulative
//     weights.
//   exchange 2 = all-gather of the chunk totals (weights, counts) and partial sums
//
// Differences to the reference that stay inside the stated tolerances: sums over particles are
// fixed-order tree sums (deterministic); the weight shift is the maximum of the log-weight
// over the predicted range (the reference's my_max, Q4, picks another element; the shift
// cancels analytically); log N(y; 0, e^{x/2}) is evaluated as -0.9189.. - x/2 - y^2 e^{-x}/2.
// If the arrivals of a CTA exceed its shared-memory capacity or a (destination, source, warp) run
// overflows (a degenerate cloud), the evaluation is abandoned with status 1 and the host code
// re-runs the general kernel (sv_filter.cu) for that problem.
//
// u is read either time-major from device memory or, for host-resident u, from the chunked
// staging buffer the copy engine fills while the kernel runs (pmmh_flps_sv_corr_streamed).
#include <math.h>

#include "common.cuh"
#include "sv_filter.cuh"
#include "sv_math.cuh"

namespace pmmh {

namespace {

constexpr int kT = kFastThreads;
constexpr int kNW = kT / 32;
constexpr int kCap = kFastCap;           // arrivals per CTA and generation (shared-memory capacity)
constexpr int kBinBits = 12;
constexpr int kBins = 1 << kBinBits;     // counting-sort bins per CTA (all its chunks together)
constexpr int kSubBits = 18;             // key bits below the bin index
constexpr int kKeyBits = kBinBits + kSubBits;   // 30
constexpr int kLutCells = 2048;
constexpr int kMaxLagF = 64;
constexpr int kSlotW = kMaxAllgather;    // doubles per CTA in the exchange buffers
constexpr int kNumSums = 10;             // fx, m1, m2, minx, lag[5], flag
constexpr int kRounds = 1;   // children per thread in flight (phase A)
constexpr int kB1 = 1;       // arrivals per thread in flight (phase B pass 1)
constexpr int kB2 = 1;       // look-ups per thread in flight (fixed-lag terms)
// (measured on B200 at N = 2^20: 4/4/8 in flight 210 ms, 2/2/4 202 ms, 1/1/1 199 ms per evaluation -- the
//  kernel is bound by instruction issue and code size, not by memory-level parallelism)
constexpr int kBinOccMax = 1024;         // a fuller bin means a degenerate cloud: abandon
constexpr int kMaxSub = kFastMaxSub;
constexpr int kPHint = 2 * kCap / 32;   // parent hints cover this many blocks of 32 children

static_assert(kCap <= 16384, "arrival indices are packed into 14 bits");
static_assert(kBins % kT == 0, "bin scan layout");
static_assert(kLutCells % kT == 0, "LUT build layout");

struct __align__(32) Rec {   // one particle of one generation
    double x;      // value
    double xpar;   // value of its parent (time - 1)
    int b[4];      // ids of its ancestors 1, 2, 3, 4 steps back
};
static_assert(sizeof(Rec) == 32, "a record is one 32-byte sector");

// 256-bit global accesses (sm_100: LDG.E.ENL2.256 / STG.E.ENL2.256).  Loads are .cg: the data
// was written by other CTAs of this launch, L1 must not serve it.
__device__ __forceinline__ void st_rec(Rec* p, double x, double xpar, int b0, int b1, int b2, int b3) {
    const unsigned long long w2 = ((unsigned long long)(unsigned)b1 << 32) | (unsigned)b0;
    const unsigned long long w3 = ((unsigned long long)(unsigned)b3 << 32) | (unsigned)b2;
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(__double_as_longlong(x)),
                 "l"(__double_as_longlong(xpar)), "l"(w2), "l"(w3)
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
__device__ __forceinline__ double ld_rec_x(const Rec* p) { return __ldcg(&p->x); }

struct FastWs {
    Rec* mail;       // [2][G * NR * CW]   child records by generation parity; NR = G * kNW runs per
                     //                    destination: slot = ((dest * G + source) * kNW + warp) * CW + rank
    int* cnt;        // [2][G][NR]         arrivals per (destination, source, warp)
    int* b4tab;      // [RB4][G * kCap]    compact tables written by the destination in arrival order
    double2* xptab;  // [RXP][G * kCap]    (index = compact id = dest * kCap + arrival): id 4 steps back,
    int* b1tab;      // [LAG][G * kCap]    (value, parent value), parent id (last LAG generations)
    int* did;        // [LAG][N]           compact id of the particle at every dense position (last LAG gens)
    double* dsh;     // [LAG][N]           its shifted weight
    int* dpos;       // [2][G * kCap]      dense position of every compact id (history output only)
    double* lev;     // [ND + 2]           target cumulative mass at every chunk boundary
};

__host__ __device__ inline int fast_rb4(int LAG) {
    const int K = LAG - 2;
    return (K > 4 ? K - 4 : 0) + 2;
}

__host__ __device__ inline size_t fast_ws_carve(int N, int G, int S, int CW, int LAG, int hist, char* base,
                                                FastWs* w) {
    const size_t NVM = (size_t)G * G * kNW * CW;
    const size_t NC = (size_t)G * kCap;
    size_t off = 0;
#define PMMH_CARVE(field, type, count)                   \
    do {                                                 \
        if (w) w->field = (type*)(base + off);           \
        off += sv_align((size_t)(count) * sizeof(type)); \
    } while (0)
    PMMH_CARVE(mail, Rec, 2 * NVM);
    PMMH_CARVE(cnt, int, (size_t)2 * G * G * kNW);
    PMMH_CARVE(b4tab, int, (size_t)fast_rb4(LAG) * NC);
    PMMH_CARVE(xptab, double2, (size_t)(LAG + 1) * NC);
    PMMH_CARVE(b1tab, int, (size_t)LAG * NC);
    PMMH_CARVE(did, int, (size_t)LAG * N);
    PMMH_CARVE(dsh, double, (size_t)LAG * N);
    PMMH_CARVE(dpos, int, hist ? 2 * NC : 1);
    PMMH_CARVE(lev, double, (size_t)S * G + 2);
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
    // all loads of a thread are issued before the first one is used (one L2 round trip)
    const int n = t.G * K;
    for (int q0 = threadIdx.x; q0 < n; q0 += 8 * blockDim.x) {
        double v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int q = q0 + r * blockDim.x;
            const int c = q / K, k = q - c * K;
            v[r] = (q < n) ? __ldcg(&buf[c * kSlotW + k]) : 0.0;
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int q = q0 + r * blockDim.x;
            if (q < n) out[q] = v[r];
        }
    }
    __syncthreads();
}

// The same exchange in two halves, so that work that does not depend on the other CTAs can run
// between publishing (team_arrive) and waiting for everybody (team_finish).
__device__ __forceinline__ void team_arrive(TeamF& t, const double* vals, int K) {
    if (t.G == 1) {
        __syncthreads();
        return;
    }
    t.epoch++;
    double* buf = t.slots + (size_t)(t.epoch & 1u) * t.G * kSlotW;
    __syncthreads();
    if ((int)threadIdx.x < K) __stcg(&buf[t.rank * kSlotW + threadIdx.x], vals[threadIdx.x]);
    __syncthreads();
    // the last warp signals (and later polls): it takes no part in the work between the two
    // halves, so the fence (which drains the CTA's stores) and the atomic overlap with that work
    if (threadIdx.x == blockDim.x - 32) {
        __threadfence();
        atomicAdd(t.ctr, 1u);
    }
}
__device__ __forceinline__ void team_finish(TeamF& t, const double* vals, int K, double* out) {
    if (t.G == 1) {
        __syncthreads();
        if ((int)threadIdx.x < K) out[threadIdx.x] = vals[threadIdx.x];
        __syncthreads();
        return;
    }
    const double* buf = t.slots + (size_t)(t.epoch & 1u) * t.G * kSlotW;
    if (threadIdx.x == blockDim.x - 32) {
        const unsigned target = t.epoch * (unsigned)t.G;
        while (ld_acquire_u32(t.ctr) < target) {
        }
        __threadfence();
    }
    __syncthreads();
    const int n = t.G * K;
    for (int q0 = threadIdx.x; q0 < n; q0 += 8 * blockDim.x) {
        double v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int q = q0 + r * blockDim.x;
            const int c = q / K, k = q - c * K;
            v[r] = (q < n) ? __ldcg(&buf[c * kSlotW + k]) : 0.0;
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int q = q0 + r * blockDim.x;
            if (q < n) out[q] = v[r];
        }
    }
    __syncthreads();
}

// ---- block-wide scans (kT threads; fixed order => deterministic) ---------------------------
// exclusive prefix of one value per thread; *total = sum over the block.  s_w: shared [kNW + 1].
__device__ __forceinline__ double block_excl_scan_d(double v, double* s_w, double* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double incl = warp_incl_scan(v, lane);
    __syncthreads();
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const double tv = (lane < kNW) ? s_w[lane] : 0.0;
        const double ti = warp_incl_scan(tv, lane);
        __syncwarp();
        if (lane < kNW) s_w[lane] = ti - tv;
        if (lane == kNW - 1) s_w[kNW] = ti;
    }
    __syncthreads();
    const double r = s_w[warp] + (incl - v);
    *total = s_w[kNW];
    return r;
}
__device__ __forceinline__ int block_excl_scan_i(int v, int* s_w, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int incl = warp_incl_scan(v, lane);
    __syncthreads();
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int tv = (lane < kNW) ? s_w[lane] : 0;
        const int ti = warp_incl_scan(tv, lane);
        __syncwarp();
        if (lane < kNW) s_w[lane] = ti - tv;
        if (lane == kNW - 1) s_w[kNW] = ti;
    }
    __syncthreads();
    const int r = s_w[warp] + (incl - v);
    *total = s_w[kNW];
    return r;
}
// exclusive prefixes of one double and one int per thread in the same three barriers
__device__ __forceinline__ void block_excl_scan_di(double v, int c, double* s_w, int* s_iw, double& vpre,
                                                   int& cpre, double& vtot, int& ctot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double incl = warp_incl_scan(v, lane);
    const int iincl = warp_incl_scan(c, lane);
    __syncthreads();
    if (lane == 31) {
        s_w[warp] = incl;
        s_iw[warp] = iincl;
    }
    __syncthreads();
    if (warp == 0) {
        const double tv = (lane < kNW) ? s_w[lane] : 0.0;
        const int tc = (lane < kNW) ? s_iw[lane] : 0;
        const double ti = warp_incl_scan(tv, lane);
        const int tci = warp_incl_scan(tc, lane);
        __syncwarp();
        if (lane < kNW) {
            s_w[lane] = ti - tv;
            s_iw[lane] = tci - tc;
        }
        if (lane == kNW - 1) {
            s_w[kNW] = ti;
            s_iw[kNW] = tci;
        }
    }
    __syncthreads();
    vpre = s_w[warp] + (incl - v);
    cpre = s_iw[warp] + (iincl - c);
    vtot = s_w[kNW];
    ctot = s_iw[kNW];
}

// exclusive running maximum (identity = lowest)
__device__ __forceinline__ int block_excl_maxscan_i(int v, int* s_w, int lowest) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(kFullMask, incl, d);
        if (lane >= d) incl = max(incl, o);
    }
    int excl = __shfl_up_sync(kFullMask, incl, 1);
    if (lane == 0) excl = lowest;
    __syncthreads();
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int ti = (lane < kNW) ? s_w[lane] : lowest;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(kFullMask, ti, d);
            if (lane >= d) ti = max(ti, o);
        }
        int te = __shfl_up_sync(kFullMask, ti, 1);
        if (lane == 0) te = lowest;
        __syncwarp();
        if (lane < kNW) s_w[lane] = te;
    }
    __syncthreads();
    return max(s_w[warp], excl);
}
__device__ __forceinline__ double block_excl_maxscan_d(double v, double* s_w) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double o = __shfl_up_sync(kFullMask, incl, d);
        if (lane >= d) incl = fmax(incl, o);
    }
    double excl = __shfl_up_sync(kFullMask, incl, 1);
    if (lane == 0) excl = 0.0;
    __syncthreads();
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        double ti = (lane < kNW) ? s_w[lane] : 0.0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double o = __shfl_up_sync(kFullMask, ti, d);
            if (lane >= d) ti = fmax(ti, o);
        }
        double te = __shfl_up_sync(kFullMask, ti, 1);
        if (lane == 0) te = 0.0;
        __syncwarp();
        if (lane < kNW) s_w[lane] = te;
    }
    __syncthreads();
    return fmax(s_w[warp], excl);
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

// norm_logpdf(y, 0, exp(x/2)) (:428,659-664) with e = exp(-x/2): -0.5 log(2 pi) - x/2 - y^2 e^2 / 2
__device__ __forceinline__ double logw_e(double x, double e, double half_y2) {
    return (-0.91893853320467267 - 0.5 * x) - half_y2 * (e * e);
}

// chunk k of the sorted generation -> (owning CTA, local chunk index); mirrored on odd rounds
// k / G for 0 <= k < 65536 with gm = ceil(2^32 / G)
// (gm = 0 stands for G = 1, whose multiplier does not fit 32 bits)
__device__ __forceinline__ int div_g(int k, unsigned gm) { return gm ? (int)__umulhi((unsigned)k, gm) : k; }
__device__ __forceinline__ int chunk_owner(int k, int G, unsigned gm) {
    const int l = div_g(k, gm), r = k - l * G;
    return (l & 1) ? (G - 1 - r) : r;
}
__device__ __forceinline__ int chunk_of(int l, int rank, int G) {
    return l * G + ((l & 1) ? (G - 1 - rank) : rank);
}

// Number of thresholds (u + j) / N (:711) at or below c = smallest j in [0, N] whose threshold
// exceeds c, written on integer-valued doubles (one conversion at the end); decisions within 64 ulp
// of a tie are counted on the way (diagnostics).  dN = (double)N.
__device__ __forceinline__ int first_child_above_nt(double c, double u, double dN, bool pow2, double invN,
                                                    long long& near_ties) {
    if (!(c == c)) return (int)dN;
    double jd = floor(c * dN - u);   // candidate for the last child whose threshold is <= c
    if (!(jd >= -1.0)) jd = -1.0;
    if (jd > dN - 1.0) jd = dN - 1.0;
    double cpl = pow2 ? (u + jd) * invN : (u + jd) / dN;            // threshold of child jd
    while (jd >= 0.0 && cpl > c) {
        jd -= 1.0;
        cpl = pow2 ? (u + jd) * invN : (u + jd) / dN;
    }
    double cph = pow2 ? (u + (jd + 1.0)) * invN : (u + (jd + 1.0)) / dN;   // threshold of child jd + 1
    while (jd < dN - 1.0 && !(cph > c)) {
        jd += 1.0;
        cpl = cph;
        cph = pow2 ? (u + (jd + 1.0)) * invN : (u + (jd + 1.0)) / dN;
    }
    const double tol = 64.0 * 2.220446049250313e-16 * c;
    if (jd >= 0.0 && fabs(c - cpl) <= tol) near_ties++;
    if (jd < dN - 1.0 && fabs(cph - c) <= tol) near_ties++;
    return (int)jd + 1;
}

// arrival index -> slot inside the destination region (NR runs, CW apart).  s_hint[e >> 3] is
// the run that holds arrival (e & ~7), so the walk below is a step or two.
__device__ __forceinline__ int arrival_slot(const unsigned short* s_off, const unsigned short* s_hint, int CW,
                                            int e) {
    int r = s_hint[e >> 3];
    while ((int)s_off[r + 1] <= e) ++r;
    return r * CW + (e - (int)s_off[r]);
}
// builds s_hint for arrivals [0, n_d) (all threads; the caller synchronises)
__device__ __forceinline__ void build_hints(const unsigned short* s_off, unsigned short* s_hint, int NR, int n_d) {
    for (int h = threadIdx.x; h <= (n_d >> 3); h += blockDim.x) {
        const int e = h << 3;
        int lo = 0, hi = NR - 1;   // largest r with s_off[r] <= e
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if ((int)s_off[mid] <= e) lo = mid;
            else hi = mid - 1;
        }
        s_hint[h] = (unsigned short)lo;
    }
}

// development instrumentation: cycles per phase, accumulated by thread 0 of every CTA
#define PROF_MARK(slot)                                                          \
    do {                                                                         \
        if (a.prof && threadIdx.x == 0) {                                        \
            const long long now__ = clock64();                                   \
            a.prof[(size_t)blockIdx.x * kProfSlots + (slot)] += now__ - prof_t;  \
            prof_t = now__;                                                      \
        }                                                                        \
    } while (0)

__global__ void __launch_bounds__(kT, 1) sv_fast_kernel(SvArgs a) {
    long long prof_t = clock64();
    extern __shared__ __align__(32) unsigned char dsm_raw[];
    const int N = a.N, NOBS = a.NOBS, LAG = a.LAG, G = a.G;
    const int S = a.NSUB, ND = S * G, CW = a.CP, NR = G * kNW;
    const int KW = 2 * S + kNumSums;    // doubles per CTA per exchange
    const size_t NVM = (size_t)G * NR * CW;   // mailbox records per generation
    const size_t NC = (size_t)G * kCap;       // compact ids per generation
    const int K = LAG - 2;                    // the fixed-lag terms need the ancestor K steps back
    const int RB4 = fast_rb4(LAG), RXP = LAG + 1;
    const int per_tile = (N + G - 1) / G;
    const unsigned gmagic = (G == 1) ? 0u : (unsigned)((0x100000000ull + (unsigned)G - 1u) / (unsigned)G);
    int lbits = 0;
    while ((1 << lbits) < S) ++lbits;
    const int kb = kKeyBits - lbits;           // key bits inside a chunk
    const double key_span = (double)(1u << kb);

    // ---- dynamic shared memory
    double* s_sh = (double*)dsm_raw;                                  // [kCap]  arrival order: shifted weight
    unsigned* s_karr = (unsigned*)(s_sh + kCap);                      // [kCap]  arrival order: key
    unsigned* s_k32 = s_karr + kCap;                                  // [kCap]  bin order: (sub key, arrival)
    int* s_hist = (int*)(s_k32 + kCap);                               // [kBins + 8]
    unsigned short* s_e = (unsigned short*)(s_hist + kBins + 8);      // [kCap]  sorted position -> arrival
    unsigned short* s_lut = s_e + kCap;                               // [kLutCells]
    unsigned short* s_hint = s_lut + kLutCells;                       // [kCap / 8 + 8] arrival -> run hints
    unsigned short* s_phint = s_hint + (kCap / 8 + 8);                // [kPHint + 8] child block -> parent hints
    double* s_z = (double*)(s_phint + (kPHint + 8));                  // [ND + 2] splitters in z space
    unsigned short* s_off = (unsigned short*)(s_z + (ND + 2));        // [NR + 2] arrival runs
    unsigned char* s_own = (unsigned char*)(s_off + (NR + 2));        // [ND] chunk -> owning CTA
    //   views of the s_karr region (dead between pass 2 of phase B and the next pass 1)
    double* s_gather = (double*)s_karr;                               // [G * KW] exchange output
    int* s_fc = (int*)s_karr;                                         // [kCap]  sorted position -> end of its children
    //   views of the s_k32 region (dead between pass 3 of phase B and the next pass 2)
    double* s_P = (double*)s_k32;                                     // [ND + 2] cumulative weight in front of every chunk
    double* s_zn = s_P + (ND + 2);                                    // [ND + 2]
    int* s_coff = (int*)(s_zn + (ND + 2));                            // [ND + 2] dense offset of every chunk
    double* s_H = (double*)(s_coff + ((ND + 3) & ~1));                // [ND + 2] blended cdf at the splitters
    double* s_g1 = (double*)s_k32;                                    // [G * 8]  small exchanges
    //   views of the s_hist region (phase A)
    unsigned short* s_wh = (unsigned short*)s_hist;                   // [kNW][G] per-warp destination counts

    __shared__ double s_vals[kSlotW];
    __shared__ double s_tot[16];
    __shared__ double s_red[12 * 32];
    __shared__ double s_w[kNW + 1];
    __shared__ int s_iw[kNW + 1];
    __shared__ double s_S[kMaxLagF];
    __shared__ double s_lag[5];              // fixed-lag sums of the previous generation (published one step late)
    // per local chunk
    __shared__ int s_lstart[kMaxSub + 1];    // first sorted position
    __shared__ double s_Cb[kMaxSub + 1];     // CTA-wide cumulative weight in front of it
    __shared__ double s_lP0[kMaxSub];        // global cumulative weight in front of it
    __shared__ int s_lLB[kMaxSub], s_lUB[kMaxSub], s_lco[kMaxSub + 1], s_lcoff[kMaxSub];
    __shared__ double s_lzlo[kMaxSub], s_lzsc[kMaxSub];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    TeamF tm;
    tm.G = G;
    tm.rank = blockIdx.x % G;
    tm.epoch = 0;
    const int team_id = blockIdx.x / G;
    tm.ctr = (unsigned*)(a.ws + (size_t)team_id * 128);
    tm.slots = (double*)(a.ws + sv_align((size_t)a.n_teams * 128)) + (size_t)team_id * 2 * G * kSlotW;
    char* wsbase = a.ws + a.ws_sync_bytes + (size_t)team_id * a.ws_team_stride;
    const bool lead = (tm.rank == 0);
    const int me = tm.rank;
    const int p0 = min(N, me * per_tile), p1 = min(N, p0 + per_tile);

    FastWs w;
    fast_ws_carve(N, G, S, CW, LAG, a.Xhist != nullptr, wsbase, &w);
#define MAIL(t) (w.mail + (size_t)((t) & 1) * NVM)
#define B4T(t) (w.b4tab + (size_t)((t) % RB4) * NC)
#define XPT(t) (w.xptab + (size_t)((t) % RXP) * NC)
#define B1T(t) (w.b1tab + (size_t)((t) - (NOBS - LAG)) * NC)
    const size_t my_base = (size_t)me * NR * CW;   // first mailbox slot of my region
    const int my_cid = me * kCap;                  // first compact id of my region

    for (int k = tid; k < ND; k += kT) s_own[k] = (unsigned char)chunk_owner(k, G, gmagic);   // G <= 255
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

        // ---- chunk boundaries in standardised space z.  In the bulk the chunks hold equal mass;
        // towards the tails a chunk is at most wmax wide, wmax = a fraction of the spread of the
        // children of one parent.  A parent chunk then scatters its children over many chunks
        // (bounded arrivals per (destination, source) run) and a tail that suddenly carries most
        // of the weight is shared by many CTAs.  lev[k] = target cumulative mass at boundary k;
        // the boundaries start at the normal quantiles of the levels.
        {
            const double zmax = 5.0, phi0 = 0.3989422804014327;
            const double varx = (c.sigmav * c.sigmav) / (1.0 - c.phi * c.phi);
            double sigz = c.sd / sqrt(c.phi * c.phi * varx + c.sd * c.sd);
            if (!(sigz > 0.0) || !(sigz <= 1.0)) sigz = 1.0;
            double wmax = 0.3 * sigz;
            if (ND > 2 && wmax < 8.0 * zmax / (double)ND) wmax = 8.0 * zmax / (double)ND;
            // bulk density M phi(z), tails 1 / wmax: find M so that ND - 2 chunks cover [-zmax, zmax]
            double mlo = 0.0, mhi = 2.0 * ND + 8.0, Mc = 0.0, zc = 0.0;
            for (int it = 0; it < 60; ++it) {
                Mc = 0.5 * (mlo + mhi);
                const double q = Mc * wmax * phi0;
                zc = (q > 1.0) ? sqrt(2.0 * log(q)) : 0.0;
                if (zc > zmax) zc = zmax;
                const double cover = Mc * (1.0 - erfc(zc * 0.70710678118654752)) + 2.0 * (zmax - zc) / wmax;
                if (cover > (double)(ND - 2)) mhi = Mc;
                else mlo = Mc;
            }
            const double Lt = (zmax - zc) / wmax;                                   // chunks per tail
            const double Lb = Mc * (1.0 - erfc(zc * 0.70710678118654752));          // chunks in the bulk
            const double Plo = 0.5 * erfc(zc * 0.70710678118654752);                // mass below -zc
            for (int k = tid; k <= ND; k += kT) {
                double z;
                if (k == 0) z = -INFINITY;
                else if (k == ND) z = INFINITY;
                else {
                    const double t = (double)(k - 1);
                    if (t <= Lt) z = -zmax + t * wmax;
                    else if (t <= Lt + Lb) z = inv_norm_cdf(Plo + (t - Lt) / Mc);
                    else z = zc + (t - Lt - Lb) * wmax;
                    if (z > zmax) z = zmax;
                    if (z < -zmax) z = -zmax;
                }
                s_z[k] = z;
                w.lev[k] = (k == 0) ? 0.0 : ((k == ND) ? 1.0 : 0.5 * erfc(-z * 0.70710678118654752));
            }
        }

        // the levels of the chunks this thread handles in the bookkeeping (fixed for the problem)
        double levr[5];
        __syncthreads();
        {
            const int PERT = (ND + kT - 1) / kT;
            const int k0 = min(ND, tid * PERT);
#pragma unroll
            for (int q = 0; q < 5; ++q) levr[q] = __ldcg(&w.lev[min(ND, k0 + q)]);
        }
        // ---------------- time 0 (:306-323, Q1): every particle equals mu + stDev * 0.0
        const double stdev0 = c.sigmav / sqrt(1.0 - (c.phi * c.phi));
        const double x0 = c.mu + stdev0 * 0.0;
        int n_d = 0;
        if (tid == 0) {
            int acc0 = 0;
            for (int l = 0; l < S; ++l) {
                const int k = chunk_of(l, me, G);
                s_lstart[l] = acc0;
                acc0 += N / ND + (k < N % ND ? 1 : 0);
            }
            s_lstart[S] = acc0;
        }
        __syncthreads();
        n_d = s_lstart[S];
        for (int r = tid; r <= NR; r += kT) s_off[r] = (r == 0) ? 0 : (unsigned short)n_d;   // generation 0: one run
        for (int h = tid; h <= (n_d >> 3); h += kT) s_hint[h] = 0;
        for (int q = tid; q < n_d; q += kT) {
            s_sh[q] = 1.0;
            s_e[q] = (unsigned short)q;
            st_rec(&MAIL(0)[my_base + q], x0, x0, 0, 0, 0, 0);
            B4T(0)[my_cid + q] = 0;
            XPT(0)[my_cid + q] = make_double2(x0, x0);
        }
        __syncthreads();
        if (lead) {
            for (int t = tid; t < NOBS; t += kT) {
                o_smo[t] = 0.0;
                o_grad[t] = 0.0;
                o_grad[NOBS + t] = 0.0;
                o_grad[2 * NOBS + t] = 0.0;
                o_grad[3 * NOBS + t] = 0.0;
            }
        }
        // moment shift: the propagation mean of the step-1 children is the same for all
        double cshift = (c.mu + c.phi * (x0 - c.mu)) + c.sr * exp(-0.5 * x0) * obs[0];
        double acc[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) acc[q] = 0.0;
        double minx = x0;
        if (tid == 0) acc[0] = (double)n_d * x0;   // sum sh * x
        int chunk_over = 0;
        double tpre = 0.0;                          // cumulative weight in front of my sorted positions
        double loglike = 0.0, shift = 0.0;
        long long near_ties = 0, key_ties = 0;
        int max_occ = 0, status = 0, max_arr = 0;
        long long fail_info = 0;

        // ---- fixed-lag terms (:445-470) of generation g for my arrivals [e_lo, e_hi): the ancestor
        // K = LAG - 2 steps back through the compact tables (the carried id reaches generation
        // g - ((K-1) % 4 + 1), then 4 at a time).  These sums only feed outputs, so they are
        // evaluated while the CTA waits for the exchanges and published one step late.
        double lacc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        if (tid < 5) s_lag[tid] = 0.0;
        auto lag_sweep = [&](int g, int e_lo, int e_hi) {
            if (g < LAG) return;
            const double ylg = obs[g - LAG];   // Q5
            const int hop0 = (K >= 1) ? ((K - 1) & 3) : 0;
            const int g1 = g - (hop0 + 1);
            const int nh = (K >= 1) ? ((K - 1) >> 2) : 0;
            const double2* xpk = XPT(g - K);
            const int* b4g = B4T(g);
            const Rec* Mg = MAIL(g);
            // (the last warp is busy with the exchange when a team has more than one CTA)
            const int nsw = (G > 1) ? kT - 32 : kT;
            if (tid >= nsw) return;
            for (int e = e_lo + tid; e < e_hi; e += nsw) {
                int id;
                if (K == 0) id = my_cid + e;   // the pair is the record itself
                else if (hop0 == 3) id = __ldcg(&b4g[my_cid + e]);
                else id = __ldcg(&Mg[my_base + arrival_slot(s_off, s_hint, CW, e)].b[hop0]);
                for (int h = 0; h < nh; ++h) id = __ldcg(&B4T(g1 - 4 * h)[id]);
                const double2 pv = __ldcg(&xpk[id]);
                const double shv = s_sh[e];
                double sq, gq[4];
                sv_score_main(c, pv.y, pv.x, ylg, sq, gq);
                lacc[0] += shv * pv.y;
                lacc[1] += gq[0] * shv;
                lacc[2] += gq[1] * shv;
                lacc[3] += gq[2] * shv;
                lacc[4] += gq[3] * shv;
            }
        };
        for (int i = 0; i < NOBS; ++i) {
            // =========== cumulative weights of generation i in sorted order (CTA-local), publish
            const int R = (n_d + kT - 1) / kT;
            {
                const int q0 = min(n_d, tid * R), q1 = min(n_d, q0 + R);
                double tsum = 0.0;
                for (int q = q0; q < q1; ++q) tsum += s_sh[s_e[q]];
                double total;
                tpre = block_excl_scan_d(tsum, s_w, &total);
                if (tid == 0)
                    for (int l = 0; l <= S; ++l)
                        if (s_lstart[l] == 0) s_Cb[l] = 0.0;
                __syncthreads();
                int lc = 0;
                while (lc < S && s_lstart[lc + 1] <= q0) ++lc;   // chunk of my first position
                double run = tpre;
                for (int q = q0; q < q1; ++q) {
                    run += s_sh[s_e[q]];
                    while (lc < S && s_lstart[lc + 1] == q + 1) {
                        s_Cb[lc + 1] = run;
                        ++lc;
                    }
                }
                __syncthreads();
                double sums[3] = {acc[0], acc[1], acc[2]};
                block_sum<3>(sums, s_red);
                minx = warp_min(minx);
                if (lane == 0) s_w[warp] = minx;
                chunk_over = __syncthreads_or(chunk_over);
                if (warp == 0) {
                    double v = (lane < kNW) ? s_w[lane] : INFINITY;
                    v = warp_min(v);
                    if (lane == 0) s_vals[2 * S + 3] = v;
                }
                if (tid < S) {
                    s_vals[2 * tid] = s_Cb[tid + 1] - s_Cb[tid];
                    s_vals[2 * tid + 1] = (double)(s_lstart[tid + 1] - s_lstart[tid]);
                }
                if (tid == 0) {
                    s_vals[2 * S + 0] = sums[0];
                    s_vals[2 * S + 1] = sums[1];
                    s_vals[2 * S + 2] = sums[2];
                    s_vals[2 * S + 4] = s_lag[0];   // fixed-lag sums of generation i - 1
                    s_vals[2 * S + 5] = s_lag[1];
                    s_vals[2 * S + 6] = s_lag[2];
                    s_vals[2 * S + 7] = s_lag[3];
                    s_vals[2 * S + 8] = s_lag[4];
                    s_vals[2 * S + 9] = (double)chunk_over;
                }
            }
            PROF_MARK(0);   // cumulative weights + block sums
            team_arrive(tm, s_vals, KW);   // exchange 2 ...
            lag_sweep(i, 0, n_d >> 1);     // ... first half of the fixed-lag terms while the others arrive
            team_finish(tm, s_vals, KW, s_gather);
            PROF_MARK(1);   // exchange 2 (wait + gather) + half of the fixed-lag terms

            // =========== bookkeeping for generation i (just gathered)
            for (int q = warp; q < kNumSums; q += kNW) {
                double s;
                if (q == 3) s = gathered_min(s_gather, KW, 2 * S + q, G, lane);
                else if (q == 9) s = gathered_max(s_gather, KW, 2 * S + q, G, lane);
                else s = gathered_sum(s_gather, KW, 2 * S + q, G, lane);
                if (lane == 0) s_tot[q] = s;
            }
            {
                // prefixes over the chunks in chunk order: thread t owns PERT (<= 4) consecutive chunks
                const int PERT = (ND + kT - 1) / kT;
                const int k0 = min(ND, tid * PERT), k1 = min(ND, k0 + PERT);
                double wv4[4];
                int cv4[4];
                double wsum = 0.0;
                int csum = 0, cmax = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    wv4[q] = 0.0;
                    cv4[q] = 0;
                    const int k = k0 + q;
                    if (k < k1) {
                        const int l = div_g(k, gmagic);
                        const int cta = s_own[k];
                        wv4[q] = s_gather[cta * KW + 2 * l];
                        cv4[q] = (int)s_gather[cta * KW + 2 * l + 1];
                    }
                    wsum += wv4[q];
                    csum += cv4[q];
                    cmax = max(cmax, cv4[q]);
                }
                double wtot, wpre;
                int ctot, cpre;
                block_excl_scan_di(wsum, csum, s_w, s_iw, wpre, cpre, wtot, ctot);
                // enforce a non-decreasing prefix across threads (tree sums may dip by an ulp)
                double last = wpre;
                {
                    double runw = wpre;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (k0 + q < k1) {
                            last = runw;
                            runw += wv4[q];
                        }
                }
                const double floorw = block_excl_maxscan_d((k1 > k0) ? last : 0.0, s_w);
                double runw = wpre;
                int runc = cpre;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (k0 + q < k1) {
                        s_P[k0 + q] = fmax(runw, floorw);
                        s_coff[k0 + q] = runc;
                        runw += wv4[q];
                        runc += cv4[q];
                    }
                if (tid == kT - 1) {
                    s_P[ND] = fmax(wtot, fmax(runw, floorw));
                    s_coff[ND] = ctot;
                }
                max_occ = max(max_occ, cmax);
            }
            __syncthreads();
            PROF_MARK(10);   // bk: sums + prefixes
            if (s_tot[9] > 0.0) {   // a CTA overflowed in phase B (uniform decision)
                status = 1;
                fail_info = 2 | ((long long)i << 8);
                break;
            }
            const double S_i = s_P[ND];
            if (!(S_i > 0.0) || !isfinite(S_i) || s_coff[ND] != N) {
                status = 1;
                fail_info = 3 | ((long long)i << 8);
                break;
            }
            if (i >= 1) loglike += shift + log(S_i) - logN;   // :537
            if (tid == 0) s_S[i % kMaxLagF] = S_i;
            if (lead && tid == 0) {
                o_filt[i] = s_tot[0] / S_i;
                o_traj[i] = s_tot[3];   // Q11: traj[i] = X_i[0] (time 0: x0)
                if (i - 1 >= LAG) {   // the fixed-lag sums travel one step late
                    const int tt = i - LAG;
                    const double S_p = s_S[(i - 1) % kMaxLagF];
                    o_smo[tt] = s_tot[4] / S_p;
                    o_grad[tt] = s_tot[5] / S_p;
                    o_grad[NOBS + tt] = s_tot[6] / S_p;
                    o_grad[2 * NOBS + tt] = s_tot[7] / S_p;
                    o_grad[3 * NOBS + tt] = s_tot[8] / S_p;
                }
            }
            if (tid < S) s_lcoff[tid] = s_coff[chunk_of(tid, me, G)];
            __syncthreads();
            // dense-position maps: history outputs, and the last LAG generations for the tail
            if (Xh != nullptr || i >= NOBS - LAG) {
                const Rec* Gi = MAIL(i);
                int* dposc = w.dpos + (size_t)(i & 1) * NC;
                const int* dposp = w.dpos + (size_t)((i + 1) & 1) * NC;
                for (int q = tid; q < n_d; q += kT) {
                    int l = 0;
                    while (l + 1 < S && s_lstart[l + 1] <= q) ++l;
                    const int dense = s_lcoff[l] + (q - s_lstart[l]);
                    const int e = s_e[q];
                    if (i >= NOBS - LAG) {
                        const size_t sl = (size_t)(i - (NOBS - LAG)) * N + dense;
                        w.did[sl] = my_cid + e;
                        w.dsh[sl] = s_sh[e];
                    }
                    if (Xh) {
                        const Rec r = ld_rec(&Gi[my_base + arrival_slot(s_off, s_hint, CW, e)]);
                        Xh[(size_t)i * N + dense] = r.x;
                        Ah[(size_t)i * N + dense] = (i == 0) ? dense : __ldcg(&dposp[r.b[0]]);
                        dposc[my_cid + e] = dense;
                    }
                }
            }
            if (i == NOBS - 1) break;
            PROF_MARK(11);   // bk: outputs + dense maps

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

            // ----- per local chunk: cumulative weight in front, range of children
            const double u = rvr[inext];
            if (tid < S) {
                const int l = tid, k = chunk_of(l, me, G);
                const int cntl = s_lstart[l + 1] - s_lstart[l];
                int lb = 0, ub = 0;
                long long dummy_nt = 0;   // chunk boundaries are counted by the sweep below
                s_lP0[l] = s_P[k];
                if (cntl > 0) {
                    lb = (s_coff[k] == 0) ? 0 : first_child_above_nt(s_P[k] / S_i, u, (double)N, n_pow2, invN_exact, dummy_nt);
                    int kn = k + 1;   // next chunk that holds particles
                    while (kn < ND && s_coff[kn + 1] == s_coff[kn]) ++kn;
                    ub = (kn >= ND) ? N : first_child_above_nt(s_P[kn] / S_i, u, (double)N, n_pow2, invN_exact, dummy_nt);
                    if (ub < lb) ub = lb;
                }
                s_lLB[l] = lb;
                s_lUB[l] = ub;
            }
            __syncthreads();
            if (tid == 0) {
                int accn = 0;
                for (int l = 0; l < S; ++l) {
                    s_lco[l] = accn;
                    accn += s_lUB[l] - s_lLB[l];
                }
                s_lco[S] = accn;
            }
            PROF_MARK(12);   // bk: moments + child ranges per chunk
            // ----- splitters of step inext.  In standardised space z = (x - m1) / sdc they start
            // as normal quantiles; afterwards they follow the shape the cloud really has: the
            // chunk counts of generation i give the empirical cdf F at the current splitters.
            // A chunk should hold neither much more than its share of the particles (arrivals
            // per CTA) nor much more than its share of the weight (children per CTA when an
            // outlying observation puts all the weight into one tail), so the new splitters are
            // the quantiles, at the levels, of H = (F + W) / 2, W = cdf of the predicted weight
            // w_{i+1}(x) dF (piecewise linear inside a chunk, normal tails in the two unbounded
            // chunks).  Every CTA computes the same values.
            if (ND > 1) {
                const int PERT = (ND + kT - 1) / kT;
                const int k0 = min(ND, tid * PERT), k1 = min(ND, k0 + PERT);
                double wsum = 0.0;
                double wm4[4], F4[5];
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    const int k = min(ND, k0 + q);
                    F4[q] = (i == 0) ? levr[q] : (double)s_coff[k] / (double)N;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int k = k0 + q;
                    wm4[q] = 0.0;
                    if (k < k1) {
                        const double zl = (k == 0) ? s_z[1] - 0.5 : s_z[k];
                        const double zr = (k == ND - 1) ? s_z[ND - 1] + 0.5 : s_z[k + 1];
                        const double xm = m1 + sdc * (0.5 * (zl + zr));
                        // single precision is plenty here: the weights only steer the splitters
                        // (every CTA evaluates the same instructions on the same numbers)
                        const double eh = (double)__expf((float)(-0.5 * xm));
                        double wv = (double)__expf((float)(logw_e(xm, eh, half_y2) - shift));
                        if (!isfinite(wv)) wv = 0.0;
                        wm4[q] = (F4[q + 1] - F4[q]) * wv;
                    }
                    wsum += wm4[q];
                }
                double wtot;
                const double wpre = block_excl_scan_d(wsum, s_w, &wtot);
                const bool use_w = (wtot > 0.0) && isfinite(wtot);
                const double inv_wtot = use_w ? 1.0 / wtot : 0.0;
                double runw = wpre;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (k0 + q < k1) {
                        s_H[k0 + q] = use_w ? 0.5 * (F4[q] + runw * inv_wtot) : F4[q];
                        runw += wm4[q];
                    }
                if (tid == 0) s_H[ND] = 1.0;
                __syncthreads();
                PROF_MARK(13);   // bk: blended cdf
                int mw = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {   // consecutive levels: one search, then walk
                    const int k = k0 + q;
                    if (k >= 1 && k < k1) {
                        const double Tk = levr[q];
                        if (q == 0 || k == 1) {
                            int hi2 = ND - 1;   // largest m with H[m] <= lev[k]
                            while (mw < hi2) {
                                const int mid = (mw + hi2 + 1) >> 1;
                                if (s_H[mid] <= Tk) mw = mid;
                                else hi2 = mid - 1;
                            }
                        } else {
                            while (mw + 1 < ND && s_H[mw + 1] <= Tk) ++mw;
                        }
                        const int m = mw;
                        const double Fm = s_H[m], Fm1 = s_H[m + 1];
                        double frac = (Tk - Fm) / (Fm1 - Fm);
                        if (!(frac >= 0.0)) frac = 0.0;
                        if (frac > 1.0) frac = 1.0;
                        double zn;
                        if (m == 0) {   // unbounded chunk: exponential tail with the decay rate of a normal tail
                            const double z1 = s_z[1];
                            zn = z1 + log(fmax(frac, 1e-300)) / fmax(1.0, fabs(z1));
                        } else if (m == ND - 1) {
                            const double z9 = s_z[ND - 1];
                            zn = z9 - log(fmax(1.0 - frac, 1e-300)) / fmax(1.0, fabs(z9));
                        } else {
                            zn = s_z[m] + frac * (s_z[m + 1] - s_z[m]);
                        }
                        s_P[k] = zn;   // s_P is free: the per-chunk values were copied out above
                    }
                }
                __syncthreads();
                for (int k = 1 + tid; k < ND; k += kT) s_z[k] = s_P[k];
                __syncthreads();
            }
            PROF_MARK(14);   // bk: new splitters
            double lut_lo = -1.0, lut_scale = 0.0;
            if (ND > 1) {
                const double zlo = s_z[1] - 1e-9, zhi = s_z[ND - 1] + 1e-9;
                lut_lo = zlo;
                lut_scale = (double)kLutCells / (zhi - zlo);
                if (!isfinite(lut_scale) || !(lut_scale > 0.0)) lut_scale = 0.0;
                constexpr int PC = kLutCells / kT;   // consecutive cells per thread: one search, then walk
                int lo2 = 0;
#pragma unroll
                for (int q = 0; q < PC; ++q) {
                    const int cidx = tid * PC + q;
                    const double edge = (lut_scale > 0.0) ? lut_lo + (double)cidx / lut_scale : -INFINITY;
                    if (q == 0) {
                        int hi2 = ND - 1;   // largest d in [0, ND-1] with s_z[d] <= edge (s_z[0] = -inf)
                        while (lo2 < hi2) {
                            const int mid = (lo2 + hi2 + 1) >> 1;
                            if (s_z[mid] <= edge) lo2 = mid;
                            else hi2 = mid - 1;
                        }
                    } else {
                        while (lo2 + 1 < ND && s_z[lo2 + 1] <= edge) ++lo2;
                    }
                    s_lut[cidx] = (unsigned short)lo2;
                }
            } else {
                for (int cidx = tid; cidx < kLutCells; cidx += kT) s_lut[cidx] = 0;
            }
            PROF_MARK(2);   // bookkeeping, splitters, LUT
            // ----- end of children per sorted position (s_fc aliases the gather buffer: all
            //       reads of s_gather are behind the barriers above)
            {
                const int q0 = min(n_d, tid * R), q1 = min(n_d, q0 + R);
                int lc = 0;
                while (lc < S && s_lstart[lc + 1] <= q0) ++lc;
                double run = tpre;
                int m = 0;
                for (int q = q0; q < q1; ++q) {
                    run += s_sh[s_e[q]];
                    while (lc < S - 1 && s_lstart[lc + 1] <= q) ++lc;
                    int ub;
                    if (q + 1 == s_lstart[lc + 1]) {
                        ub = s_lUB[lc];
                    } else {
                        const double cc = (s_lP0[lc] + (run - s_Cb[lc])) / S_i;
                        ub = first_child_above_nt(cc, u, (double)N, n_pow2, invN_exact, near_ties);
                    }
                    m = max(m, ub);
                    s_fc[q] = m;
                }
                const int mt = block_excl_maxscan_i((q1 > q0) ? m : 0, s_iw, 0);
                lc = 0;
                while (lc < S && s_lstart[lc + 1] <= q0) ++lc;
                for (int q = q0; q < q1; ++q) {
                    while (lc < S - 1 && s_lstart[lc + 1] <= q) ++lc;
                    int v = max(s_fc[q], mt);
                    v = max(v, s_lLB[lc]);
                    v = min(v, s_lUB[lc]);
                    if (q + 1 == s_lstart[lc + 1]) v = s_lUB[lc];
                    s_fc[q] = v;
                }
            }
            for (int q = tid; q < kNW * G; q += kT) s_wh[q] = 0;
            __syncthreads();
            // parent hints: s_phint[b] = sorted position of the parent of child 32 b (CTA order)
            const bool use_ph = s_lco[S] <= 32 * kPHint;
            if (use_ph) {
                for (int q = tid; q < n_d; q += kT) {
                    int l = 0;
                    while (l + 1 < S && s_lstart[l + 1] <= q) ++l;
                    const int ca = (q == s_lstart[l]) ? s_lLB[l] : s_fc[q - 1];
                    const int cb = s_fc[q];
                    if (cb > ca) {
                        const int ta = s_lco[l] + (ca - s_lLB[l]), tb = s_lco[l] + (cb - s_lLB[l]);
                        for (int bb = (ta + 31) >> 5; (bb << 5) < tb; ++bb) s_phint[bb] = (unsigned short)q;
                    }
                }
            }
            __syncthreads();
            PROF_MARK(3);   // children ranges

            // =========== phase A: children of my parents, in birth order: propagate (:354-358),
            //             route, write the record into the destination's region.  Every warp owns
            //             a contiguous share of the children and its own run per destination:
            //             no block-wide synchronisation in this phase.
            int pair_over = 0;
            {
                int n_c = s_lco[S];
                if (n_c > 60000 * kNW) {   // per-warp run counters are 16 bit: a degenerate cloud, abandon
                    pair_over = 1;
                    n_c = 0;
                }
                const double* Ui = U + (size_t)inext * N;
                // streamed u (host copies still in flight): wait until the chunk that holds time
                // step inext has landed.  Chunk layout: [chunk][particle][UC steps].
                const int UC = a.u_chunk;
                const double* Uc = nullptr;
                if (UC > 0) {
                    if (tid == 0 && a.u_ready != nullptr) {
                        long long spins = 0;
                        while (ld_acquire_sys_s32(a.u_ready) < inext + 1) {
                            if (++spins > (1ll << 28)) {   // ~ seconds: the copies never came, give up
                                pair_over = 1;
                                break;
                            }
                        }
                    }
                    __syncthreads();
                    const int ch = inext / UC;
                    Uc = U + (size_t)ch * N * UC + (inext - ch * UC);
                }
                const Rec* Gi = MAIL(i);
                Rec* Gn = MAIL(inext);
                unsigned short* my_wh = s_wh + warp * G;
                // rounds of 32 consecutive children are dealt to the warps round-robin: the runs a
                // warp writes then do not follow the value range of one stretch of parents
                const int tw1 = n_c;
                for (int t0 = warp * 32; t0 < n_c; t0 += 32 * kNW * kRounds) {
                    double un[kRounds];
                    Rec pr[kRounds];
                    int pcid[kRounds];
#pragma unroll
                    for (int r = 0; r < kRounds; ++r) {
                        const int t = t0 + r * 32 * kNW + lane;
                        pcid[r] = -1;
                        un[r] = 0.0;
                        pr[r].x = 0.0;
                        pr[r].b[0] = pr[r].b[1] = pr[r].b[2] = 0;
                        if (t < tw1) {
                            int l = 0;
                            while (l + 1 < S && s_lco[l + 1] <= t) ++l;
                            const int j = s_lLB[l] + (t - s_lco[l]);
                            un[r] = (UC > 0) ? __ldcg(&Uc[(size_t)j * UC]) : ld_stream_f64(&Ui[j]);
                            // parent: first sorted position of chunk l whose children end beyond j
                            int lo2 = s_lstart[l], hi2 = s_lstart[l + 1] - 1;
                            if (use_ph) {
                                const int bb = t >> 5;
                                lo2 = max(lo2, (int)s_phint[bb]);
                                if (((bb + 1) << 5) < n_c) hi2 = min(hi2, (int)s_phint[bb + 1]);
                            }
                            while (lo2 < hi2) {
                                const int mid = (lo2 + hi2) >> 1;
                                if (s_fc[mid] > j) hi2 = mid;
                                else lo2 = mid + 1;
                            }
                            const int e = s_e[lo2];
                            pcid[r] = my_cid + e;
                            pr[r] = ld_rec(&Gi[my_base + arrival_slot(s_off, s_hint, CW, e)]);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < kRounds; ++r) {
                        const bool valid = pcid[r] >= 0;
                        int d = G;   // inactive lanes never match a destination
                        double xnew = 0.0;
                        if (valid) {
                            double mean = c.mu + c.phi * (pr[r].x - c.mu);
                            mean += c.sr * exp(-0.5 * pr[r].x) * y1;
                            xnew = mean + c.sd * un[r];
                            int kc = 0;
                            if (ND > 1) {
                                const double zx = (xnew - m1) * inv_sdc;
                                const double tt = (zx - lut_lo) * lut_scale;
                                const int cell =
                                    (tt >= 0.0) ? ((tt < (double)kLutCells) ? (int)tt : kLutCells - 1) : 0;
                                kc = s_lut[cell];
                                while (kc + 1 < ND && zx >= s_z[kc + 1]) ++kc;
                                while (kc > 0 && zx < s_z[kc]) --kc;
                            }
                            d = s_own[kc];
                        }
                        // stable rank inside the warp's run: lanes with the same destination
                        const unsigned peers = __match_any_sync(kFullMask, d);
                        if (valid) {
                            const int before = __popc(peers & ((1u << lane) - 1u));
                            const int base = my_wh[d];
                            __syncwarp(peers);
                            if (before == 0) my_wh[d] = (unsigned short)(base + __popc(peers));
                            const int rank = base + before;
                            if (rank < CW)
                                st_rec(&Gn[(((size_t)d * G + me) * kNW + warp) * CW + rank], xnew, pr[r].x, pcid[r],
                                       pr[r].b[0], pr[r].b[1], pr[r].b[2]);
                            else
                                pair_over = 1;
                        }
                        __syncwarp();
                    }
                }
                __syncthreads();
                int* cn = w.cnt + (size_t)(inext & 1) * G * NR;
                for (int q = tid; q < kNW * G; q += kT) {
                    const int wv = div_g(q, gmagic), d = q - wv * G;
                    cn[(size_t)d * NR + me * kNW + wv] = min((int)s_wh[q], CW);
                }
            }
            pair_over = __syncthreads_or(pair_over);
            PROF_MARK(4);   // children: propagate + route + write
            if (tid == 0) s_vals[0] = (double)pair_over;
            team_arrive(tm, s_vals, 1);   // exchange 1 ...
            // ... second half of the fixed-lag terms of generation i while the others arrive
            lag_sweep(i, n_d >> 1, n_d);
            block_sum<5>(lacc, s_red);
            if (tid < 5) s_lag[tid] = s_red[tid * 32];   // block_sum leaves the totals there
#pragma unroll
            for (int q = 0; q < 5; ++q) lacc[q] = 0.0;
            team_finish(tm, s_vals, 1, s_g1);
            {
                double f = 0.0;
                for (int cc = tid; cc < G; cc += kT) f = fmax(f, s_g1[cc]);
                const int any = __syncthreads_or(f > 0.0);
                if (any) {
                    status = 1;   // a run overflowed (uniform decision)
                    fail_info = 1 | ((long long)inext << 8);
                    break;
                }
            }
            PROF_MARK(5);   // exchange 1

            // =========== phase B: my arrivals of generation inext
            {
                const int* cn = w.cnt + (size_t)(inext & 1) * G * NR + (size_t)me * NR;
                // runs of the (source, warp) pairs: thread t owns PR consecutive runs
                const int PR = (NR + kT - 1) / kT;
                const int r0 = min(NR, tid * PR), r1 = min(NR, r0 + PR);
                int csum = 0;
                int cv[8];   // PR <= 8: at most 8 * kT runs
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    cv[q] = (r0 + q < r1) ? __ldcg(&cn[r0 + q]) : 0;
                    csum += cv[q];
                }
                int total = 0;
                int run = block_excl_scan_i(csum, s_iw, &total);
                const bool fits = total <= kCap;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (r0 + q < r1) s_off[r0 + q] = (unsigned short)(fits ? run : 0);
                    run += cv[q];
                }
                if (tid == 0) s_off[NR] = (unsigned short)(fits ? total : 0);
                for (int q = tid; q < kBins + 8; q += kT) s_hist[q] = 0;
                n_d = total;
                chunk_over = 0;
                if (!fits) {
                    chunk_over = 1;
                    n_d = 0;
                }
                max_arr = max(max_arr, total);
                __syncthreads();
                build_hints(s_off, s_hint, NR, n_d);
            }
            if (tid < S) {
                const int k = chunk_of(tid, me, G);
                double zlo = (k == 0) ? ((ND > 1) ? s_z[1] - 2.0 : -8.0) : s_z[k];
                double zhi = (k == ND - 1) ? ((ND > 1) ? s_z[ND - 1] + 2.0 : 8.0) : s_z[k + 1];
                double sc = key_span / (zhi - zlo);
                if (!isfinite(sc) || !(sc > 0.0)) sc = 0.0;
                s_lzlo[tid] = zlo;
                s_lzsc[tid] = sc;
            }
            __syncthreads();
            const Rec* Gn = MAIL(inext);
#pragma unroll
            for (int q = 0; q < 9; ++q) acc[q] = 0.0;
            minx = INFINITY;
            // ---- pass 1 (arrival order): compact tables, key + histogram, weight, sums
            {
                int* b4n = B4T(inext);
                double2* xpn = XPT(inext);
                int* b1n = (inext >= NOBS - LAG) ? B1T(inext) : nullptr;
                for (int e0 = 0; e0 < n_d; e0 += kB1 * kT) {
                    Rec rc[kB1];
#pragma unroll
                    for (int r = 0; r < kB1; ++r) {
                        const int e = e0 + r * kT + tid;
                        rc[r].x = 0.0;
                        if (e < n_d) rc[r] = ld_rec(&Gn[my_base + arrival_slot(s_off, s_hint, CW, e)]);
                    }
#pragma unroll
                    for (int r = 0; r < kB1; ++r) {
                        const int e = e0 + r * kT + tid;
                        if (e < n_d) {
                            const double xv = rc[r].x;
                            b4n[my_cid + e] = rc[r].b[3];
                            __stcs(&xpn[my_cid + e], make_double2(xv, rc[r].xpar));   // read 8 steps later: stream out
                            if (b1n) __stcs(&b1n[my_cid + e], rc[r].b[0]);
                            // chunk and key
                            int kc = 0;
                            const double zx = (xv - m1) * inv_sdc;
                            if (ND > 1) {
                                const double tt = (zx - lut_lo) * lut_scale;
                                const int cell =
                                    (tt >= 0.0) ? ((tt < (double)kLutCells) ? (int)tt : kLutCells - 1) : 0;
                                kc = s_lut[cell];
                                while (kc + 1 < ND && zx >= s_z[kc + 1]) ++kc;
                                while (kc > 0 && zx < s_z[kc]) --kc;
                            }
                            const int l = div_g(kc, gmagic);
                            const double tq = (zx - s_lzlo[l]) * s_lzsc[l];
                            unsigned kq = 0;
                            if (tq >= 0.0) kq = (tq < key_span) ? (unsigned)tq : ((1u << kb) - 1u);
                            const unsigned key = ((unsigned)l << kb) | kq;
                            s_karr[e] = key;
                            atomicAdd(&s_hist[key >> kSubBits], 1);
                            // weight (:427-437)
                            const double eh = exp(-0.5 * xv);
                            double shv = exp(logw_e(xv, eh, half_y2) - shift);
                            if (!isfinite(shv)) shv = 0.0;
                            s_sh[e] = shv;
                            minx = fmin(minx, xv);
                            const double sx = shv * xv;
                            if (isfinite(sx)) acc[0] += sx;
                            // propagation mean of the next step (for the splitters)
                            const double df = ((c.mu + c.phi * (xv - c.mu)) + c.sr * eh * yi) - cshift;
                            const double sdf = shv * df;
                            if (isfinite(sdf)) {
                                acc[1] += sdf;
                                acc[2] += sdf * df;
                            }
                        }
                    }
                }
            }
            __syncthreads();
            PROF_MARK(6);   // phase B pass 1
            // ---- bin offsets (exclusive scan of the histogram)
            {
                constexpr int PB = kBins / kT;
                int v[PB], tsum = 0, occ = 0;
#pragma unroll
                for (int q = 0; q < PB; ++q) {
                    v[q] = s_hist[tid * PB + q];
                    tsum += v[q];
                    occ = max(occ, v[q]);
                }
                int total;
                int run = block_excl_scan_i(tsum, s_iw, &total);
#pragma unroll
                for (int q = 0; q < PB; ++q) {
                    s_hist[tid * PB + q] = run;
                    run += v[q];
                }
                if (occ > kBinOccMax) chunk_over = 1;
                __syncthreads();
                if (tid <= S) {
                    // first sorted position of every local chunk = start of its first bin
                    const int b = tid << (kBinBits - lbits);
                    s_lstart[tid] = (tid == S || b >= kBins) ? n_d : s_hist[b];
                }
                __syncthreads();
                if (__syncthreads_or(chunk_over)) {
                    chunk_over = 1;
                    n_d = 0;   // abandon: skip the sort, the flag travels with the next exchange
                    if (tid <= S) s_lstart[tid] = 0;
                }
            }
            // ---- pass 2: scatter (sub key, arrival) into bin order
            for (int e = tid; e < n_d; e += kT) {
                const unsigned key = s_karr[e];
                const int slot = atomicAdd(&s_hist[key >> kSubBits], 1);
                s_k32[slot] = ((key & ((1u << kSubBits) - 1u)) << 14) | (unsigned)e;
            }
            __syncthreads();
            // ---- pass 3: order inside each bin (all pairs; ~2 records per bin on average)
            for (int e = tid; e < n_d; e += kT) {
                const unsigned key = s_karr[e];
                const int bin = (int)(key >> kSubBits);
                const unsigned sub = key & ((1u << kSubBits) - 1u);
                const int st = (bin > 0) ? s_hist[bin - 1] : 0, en = s_hist[bin];
                int rank = 0;
                for (int o = st; o < en; ++o) {
                    const unsigned v = s_k32[o];
                    const unsigned sub2 = v >> 14;
                    const int e2 = (int)(v & 0x3fffu);
                    if (e2 == e) continue;
                    bool lt = sub2 < sub;
                    if (sub2 == sub) {
                        // same 30-bit key: decide on the exact values, then on the arrival index
                        const double xa = ld_rec_x(&Gn[my_base + arrival_slot(s_off, s_hint, CW, e)]);
                        const double xb = ld_rec_x(&Gn[my_base + arrival_slot(s_off, s_hint, CW, e2)]);
                        if (xb == xa) {
                            key_ties++;
                            lt = e2 < e;
                        } else {
                            lt = xb < xa;
                        }
                    }
                    if (lt) rank++;
                }
                s_e[st + rank] = (unsigned short)e;
            }
            __syncthreads();
            PROF_MARK(7);   // phase B sort
        }   // time loop
        PROF_MARK(8);

        // ---------------- fixed-lag terms of the last generation (they travel one step late)
        if (status == 0) {
            const int T = NOBS - 1;
            lag_sweep(T, n_d >> 1, n_d);   // the first half ran while waiting for the last exchange
            block_sum<5>(lacc, s_red);
            if (tid < 5) s_vals[tid] = s_red[tid * 32];
            team_exchange(tm, s_vals, 5, s_g1);
            if (warp < 5) {
                const double sv = gathered_sum(s_g1, 5, warp, G, lane);
                if (lane == 0) s_tot[warp] = sv;
            }
            __syncthreads();
            if (lead && tid == 0 && T >= LAG) {
                const int tt = T - LAG + 1;
                const double S_T = s_S[T % kMaxLagF];
                o_smo[tt] = s_tot[0] / S_T;
                o_grad[tt] = s_tot[1] / S_T;
                o_grad[NOBS + tt] = s_tot[2] / S_T;
                o_grad[2 * NOBS + tt] = s_tot[3] / S_T;
                o_grad[3 * NOBS + tt] = s_tot[4] / S_T;
            }
            __syncthreads();
        }
        // ---------------- tail (:540-562, Q6), dense positions
        if (status == 0) {
            const int T = NOBS - 1;
            const double S_T = s_S[T % kMaxLagF];
            __syncthreads();
            if (tid == 0) s_vals[0] = 0.0;
            team_exchange(tm, s_vals, 1, s_g1);   // the dense maps of generation T are complete
            const int* didT = w.did + (size_t)(LAG - 1) * N;
            const double* shT = w.dsh + (size_t)(LAG - 1) * N;
            for (int k = 0; k < LAG; ++k) {
                const int ip = T - k;
                double tacc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
                const double S_ip = s_S[ip % kMaxLagF];
                const double y1 = obs_wrap(obs, ip - 1, NOBS);
                const double* shI = w.dsh + (size_t)(ip - (NOBS - LAG)) * N;
                for (int j = p0 + tid; j < p1; j += kT) {
                    int b = __ldcg(&didT[j]), bprev = b;
                    for (int h = 0; h < k; ++h) {
                        bprev = b;
                        b = __ldcg(&B1T(T - h)[b]);
                    }
                    const double curr = __ldcg(&XPT(ip)[b]).x;
                    double sT = __ldcg(&shT[j]);
                    if (!isfinite(sT)) sT = 0.0;
                    tacc[0] += (sT / S_T) * curr;
                    if (k >= 1) {
                        const double next = __ldcg(&XPT(ip + 1)[bprev]).x;
                        double sq, g[4];
                        sv_score_tail(c, curr, next, y1, sq, g);
                        double si = __ldcg(&shI[j]);
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
                team_exchange(tm, s_vals, 5, s_g1);
                if (warp < 5) {
                    const double s = gathered_sum(s_g1, 5, warp, G, lane);
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
            }
            if (tid < 16 && status == 0) {
                a.hess1[(size_t)prob * 16 + tid] = 0.0;
                a.hess2[(size_t)prob * 16 + tid] = 0.0;
            }
        }
        {
            double nt[3] = {(double)near_ties, (double)key_ties, 0.0};
            block_sum<3>(nt, s_red);
            int mo = warp_max(max_occ);
            if (lane == 0) s_iw[warp] = mo;
            __syncthreads();
            if (tid == 0) {
                for (int q = 0; q < kNW; ++q) mo = max(mo, s_iw[q]);
                s_vals[0] = nt[0];
                s_vals[1] = nt[1];
                s_vals[2] = (double)mo;
                s_vals[3] = (double)max_arr;
            }
            team_exchange(tm, s_vals, 4, s_g1);
            if (warp < 2) {
                const double s = gathered_sum(s_g1, 4, warp, G, lane);
                if (lane == 0 && lead)
                    o_diag[warp == 0 ? kDiagNearTies : kDiagKeyTies] = (long long)(warp == 0 ? s : s * 0.5);
            } else if (warp == 2) {
                const double s = gathered_max(s_g1, 4, 2, G, lane);
                if (lane == 0 && lead) o_diag[kDiagMaxBin] = (long long)s;
            } else if (warp == 3) {
                const double s = gathered_max(s_g1, 4, 3, G, lane);
                if (lane == 0 && lead) o_diag[kDiagFastInfo] = fail_info | ((long long)s << 32);
            }
            __syncthreads();
        }
    }   // problem loop
    PROF_MARK(9);   // tail + outputs
#undef MAIL
#undef B4T
#undef XPT
#undef B1T
}

}  // namespace

// chunks per CTA: as many as possible (interleaving them over the value range is what balances
// the children per CTA and the arrivals per (destination, source) run), at least 32 particles each
int sv_fast_nsub(int N, int G) {
    const int per = (N + G - 1) / G;
    int s = per / 32;
    if (s < 1) s = 1;
    if (s > kFastMaxSub) s = kFastMaxSub;
    return s;
}

// capacity of one (destination, source CTA, source warp) run
int sv_fast_pair_cap(int N, int G) {
    const int per = (N + G - 1) / G;
    const int perw = (per + kNW - 1) / kNW;   // children of one warp when the CTAs are balanced
    long long cw;
    if (G == 1) {
        cw = perw + 40;
    } else {
        const double mean = (double)N / ((double)G * (double)G * (double)kNW);
        cw = (long long)(4.0 * mean + 12.0 * sqrt(mean) + 32.0);
        if (G <= 8 && cw < 2ll * perw + 40) cw = 2ll * perw + 40;   // few CTAs: runs follow the parents' range
    }
    if (cw * G * kNW < per + 1) cw = (per + G * kNW) / (G * kNW) + 1;   // generation 0 lives in one region
    return (int)((cw + 7) & ~7ll);
}

// is the problem eligible (shared-memory capacity with 30 % head room)?
int sv_fast_eligible(int N, int G) {
    const int per = (N + G - 1) / G;
    return (long long)per * 13 <= (long long)kCap * 10 && G <= kT;
}

size_t sv_fast_ws_bytes(int N, int G, int S, int CW, int LAG, int hist) {
    return fast_ws_carve(N, G, S, CW, LAG, hist, nullptr, nullptr);
}

size_t sv_fast_sync_bytes(int G, int n_teams) {
    return sv_align((size_t)n_teams * 128) + sv_align((size_t)n_teams * 2 * G * kSlotW * sizeof(double));
}

int sv_fast_smem_bytes(int N, int G, int S) {
    const int ND = S * G;
    size_t b = (size_t)kCap * 8 + (size_t)kCap * 4 * 2 + (size_t)(kBins + 8) * 4 + (size_t)kCap * 2 +
               (size_t)kLutCells * 2 + (size_t)(kCap / 8 + 8) * 2 + (size_t)(kPHint + 8) * 2 +
               (size_t)(ND + 2) * 8 + (size_t)(G * kNW + 2) * 2 + (size_t)ND;
    (void)N;
    return (int)(b + 64);
}

cudaError_t sv_fast_launch(const SvArgs& a, int grid, cudaStream_t stream) {
    const int smem = sv_fast_smem_bytes(a.N, a.G, a.NSUB);
    cudaError_t err = cudaFuncSetAttribute(sv_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return err;
    void* kargs[] = {(void*)&a};
    return cudaLaunchCooperativeKernel((void*)sv_fast_kernel, dim3(grid), dim3(kT), kargs, smem, stream);
}

}  // namespace pmmh
