// common.cuh -- device primitives shared by the pmmh-qn B200 kernels.
//
// * "team" = the G co-resident CTAs that cooperate on ONE likelihood evaluation inside the
//   persistent kernel.  team_allgather() is the only inter-CTA primitive: every CTA
//   publishes K doubles and receives everybody's, which doubles as a full team barrier
//   (release/acquire at gpu scope).  All cross-CTA reductions are then summed in a FIXED
//   order, so results are deterministic and independent of scheduling.
// * scans use a warp-contiguous layout (each warp owns a contiguous, 32-aligned segment and
//   walks it in coalesced rounds of 32), one block-level combine per pass.
//
// Compiled with -fmad=false: the parity oracle is plain IEEE fp64 without contraction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pmmh {

constexpr int kMaxAllgather = 48;   // doubles per CTA per all-gather
constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_acquire_sys_s32(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// read-once streaming load (the auxiliary variables u): do not pollute L1
__device__ __forceinline__ double ld_stream_f64(const double* p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

struct Team {
    int G;              // CTAs in the team
    int rank;           // this CTA's rank in the team
    unsigned epoch;     // all-gather counter (uniform across the team)
    unsigned* stamps;   // [G]     (global, zeroed before launch)
    double* slots;      // [2][G][kMaxAllgather] (global)
};

// vals: K doubles in SHARED memory; out: shared [G*K].  Full barrier for the team.
__device__ __forceinline__ void team_allgather(Team& t, const double* vals, int K, double* out) {
    __syncthreads();
    if (t.G == 1) {
        if (vals != out && (int)threadIdx.x < K) out[threadIdx.x] = vals[threadIdx.x];
        __syncthreads();
        return;
    }
    t.epoch++;
    double* buf = t.slots + (size_t)(t.epoch & 1u) * t.G * kMaxAllgather;
    if (threadIdx.x == 0) {
        for (int k = 0; k < K; ++k) __stcg(&buf[t.rank * kMaxAllgather + k], vals[k]);
        __threadfence();
        st_release_u32(&t.stamps[t.rank], t.epoch);
    }
    for (int c = threadIdx.x; c < t.G; c += blockDim.x) {
        while (ld_acquire_u32(&t.stamps[c]) < t.epoch) {
        }
        for (int k = 0; k < K; ++k) out[c * K + k] = __ldcg(&buf[c * kMaxAllgather + k]);
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ void team_barrier(Team& t, double* scratch) {
    team_allgather(t, scratch, 0, scratch);
}

// ---- warp helpers --------------------------------------------------------------------
__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        double o = __shfl_up_sync(kFullMask, v, d);
        if (lane >= d) v = v + o;
    }
    return v;
}
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(kFullMask, v, d);
        if (lane >= d) v = v + o;
    }
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = v + __shfl_xor_sync(kFullMask, v, d);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmax(v, __shfl_xor_sync(kFullMask, v, d));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmin(v, __shfl_xor_sync(kFullMask, v, d));
    return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(kFullMask, v, d));
    return v;
}

// Deterministic sum over the first `count` CTAs of quantity k of an all-gather result
// (layout g[c*K + k]).  Must be called by one full warp; every lane gets the result.
__device__ __forceinline__ double gathered_sum(const double* g, int K, int k, int count, int lane) {
    double s = 0.0;
    for (int c = lane; c < count; c += 32) s = s + g[c * K + k];
    return warp_sum(s);
}
__device__ __forceinline__ double gathered_max(const double* g, int K, int k, int count, int lane) {
    double s = -INFINITY;
    for (int c = lane; c < count; c += 32) s = fmax(s, g[c * K + k]);
    return warp_max(s);
}
__device__ __forceinline__ double gathered_min(const double* g, int K, int k, int count, int lane) {
    double s = INFINITY;
    for (int c = lane; c < count; c += 32) s = fmin(s, g[c * K + k]);
    return warp_min(s);
}

// Block-wide deterministic sum of NV per-thread values.  red: shared [NV * 32].
// Result for value v is left in red[v] after the call (valid for all threads).
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = warp_sum(v[i]);
        if (lane == 0) red[i * 32 + warp] = s;
    }
    __syncthreads();
    for (int i = warp; i < NV; i += nwarp) {
        double s = (lane < nwarp) ? red[i * 32 + lane] : 0.0;
        s = warp_sum(s);
        __syncwarp();
        if (lane == 0) red[i * 32] = s;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = red[i * 32];
    __syncthreads();
}

// Geometry of a CTA's tile and of the warp segments inside it.
struct Tile {
    int p0, p1;     // CTA tile [p0, p1)
    int wseg;       // warp segment length (multiple of 32)
    __device__ __forceinline__ void init(int n, int G, int rank) {
        int per = (n + G - 1) / G;
        p0 = min(n, rank * per);
        p1 = min(n, p0 + per);
        int nwarp = blockDim.x >> 5;
        int len = p1 - p0;
        wseg = ((len + nwarp - 1) / nwarp + 31) & ~31;
    }
    __device__ __forceinline__ int seg_begin(int warp) const { return min(p1, p0 + warp * wseg); }
    __device__ __forceinline__ int seg_end(int warp) const { return min(p1, p0 + (warp + 1) * wseg); }
};

}  // namespace pmmh
