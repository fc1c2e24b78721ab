// microbench2.cu -- access patterns of the grid kernel (sv_grid.cu) in isolation, to see what each
// phase could cost: genealogy records (random 32-byte read + coalesced 32-byte write), the score
// gather (random 16 bytes out of a 16 MB generation), 16-byte scatters, atomics on striped
// histograms, the cost of the release fence after a burst of stores.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/microbench2 tools/microbench2.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e = (x);                                                       \
        if (e != cudaSuccess) {                                                    \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                               \
        }                                                                          \
    } while (0)

constexpr int kThreads = 1024;
constexpr int KPT = 7;

__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
struct __align__(32) Rec {
    int a[8];
};
__device__ __forceinline__ void ld_rec8(const Rec* p, int (&r)[8]) {
    asm volatile("ld.global.cg.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}
__device__ __forceinline__ void st_rec8(Rec* p, const int (&r)[8], int b) {
    asm volatile("st.global.cg.v8.s32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(b), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6])
                 : "memory");
}
__device__ __forceinline__ void ld_rec4(const Rec* p, int (&r)[8]) {
    const int4 x = __ldcg((const int4*)p), y = __ldcg((const int4*)p + 1);
    r[0] = x.x; r[1] = x.y; r[2] = x.z; r[3] = x.w; r[4] = y.x; r[5] = y.y; r[6] = y.z; r[7] = y.w;
}
__device__ __forceinline__ void st_rec4(Rec* p, const int (&r)[8], int b) {
    __stcg((int4*)p, make_int4(b, r[0], r[1], r[2]));
    __stcg((int4*)p + 1, make_int4(r[3], r[4], r[5], r[6]));
}

// records: thread = KPT children (strided), batches of RB: random (or coalesced) 32-byte read from
// src, coalesced 32-byte write to dst; src / dst swap every rep (ping-pong like the two parities)
template <int RB, int V8>
__global__ void __launch_bounds__(kThreads, 1) k_records(Rec* A, Rec* B, unsigned n_mask, int reps, int coalesced,
                                                         int* sink) {
    const int tid = threadIdx.x, nc = KPT * kThreads, jb = blockIdx.x * nc;
    int s = 0;
    for (int r = 0; r < reps; ++r) {
        const Rec* src = (r & 1) ? B : A;
        Rec* dst = (r & 1) ? A : B;
        for (int k0 = 0; k0 < KPT; k0 += RB) {
            int rr[RB][8];
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                const int i = (k0 + u) * kThreads + tid;
                if (k0 + u < KPT) {
                    const unsigned h = coalesced ? (unsigned)(jb + i) & n_mask
                                                 : hash32((unsigned)(jb + i) * 131u + r * 104729u) & n_mask;
                    if (V8) ld_rec8(&src[h], rr[u]);
                    else ld_rec4(&src[h], rr[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                const int i = (k0 + u) * kThreads + tid;
                if (k0 + u < KPT) {
                    if (V8) st_rec8(&dst[(unsigned)(jb + i) & n_mask], rr[u], i);
                    else st_rec4(&dst[(unsigned)(jb + i) & n_mask], rr[u], i);
                    s += rr[u][7];
                }
            }
        }
    }
    if (s == -12345) *sink = s;
}

// score gather: 16 bytes at a random row of a table, KPT in flight
__global__ void __launch_bounds__(kThreads, 1) k_gather16(const double2* P, unsigned n_mask, int reps, double* sink) {
    const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.0;
    for (int r = 0; r < reps; ++r) {
        double2 v[KPT];
#pragma unroll
        for (int k = 0; k < KPT; ++k) v[k] = __ldcg(&P[hash32(gid * 131u + k * 7919u + r * 104729u) & n_mask]);
#pragma unroll
        for (int k = 0; k < KPT; ++k) acc += v[k].x + v[k].y;
    }
    if (acc == -12345.0) *sink = acc;
}
// 16-byte scatter: mode 0 fully random rows; mode 1 runs (48 consecutive children share a destination region)
__global__ void __launch_bounds__(kThreads, 1) k_scatter16(int4* M, unsigned n_mask, int reps, int mode) {
    const int tid = threadIdx.x, nc = KPT * kThreads, jb = blockIdx.x * nc;
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < KPT; ++k) {
            const int i = k * kThreads + tid;
            unsigned h;
            if (mode == 0) h = hash32((unsigned)(jb + i) * 131u + r * 104729u) & n_mask;
            else h = ((hash32((unsigned)(jb + i) / 48u * 131u + r * 104729u) & n_mask) & ~63u) + (unsigned)(i % 48);
            __stcg(&M[h & n_mask], make_int4(i, r, k, tid));
        }
    }
}
// histogram merge: every CTA adds `nnz` counts into a striped global histogram (ncopy copies of nb bins)
__global__ void __launch_bounds__(kThreads, 1) k_histmerge(int* gh, int nb, int ncopy, int nnz, int reps) {
    const int tid = threadIdx.x;
    int* g = gh + (size_t)(blockIdx.x % ncopy) * nb;
    for (int r = 0; r < reps; ++r) {
        for (int k = tid; k < nnz; k += kThreads) {
            const unsigned b = (hash32(blockIdx.x * 977u + r * 13u) + (unsigned)k * 2u) % (unsigned)nb;
            atomicAdd(&g[b], 1);
        }
        __syncthreads();
        if (tid == 0) __threadfence();
        __syncthreads();
    }
}
// fence after a burst of coalesced 16-byte stores (KPT per thread): stores + __syncthreads + fence by thread 0
__global__ void __launch_bounds__(kThreads, 1) k_store_fence(int4* M, int reps, int fence, long long* cyc) {
    const int tid = threadIdx.x, nc = KPT * kThreads, jb = blockIdx.x * nc;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < KPT; ++k) __stcg(&M[jb + k * kThreads + tid], make_int4(r, k, tid, 0));
        __syncthreads();
        if (fence && tid == 0) __threadfence();
        __syncthreads();
    }
    if (tid == 0 && blockIdx.x == 0) *cyc = clock64() - t0;
}

template <class F>
float time_ms(F f) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int SM = prop.multiProcessorCount;
    printf("device %s, %d SMs\n", prop.name, SM);
    const int REPS = 40;
    const double per_rep = (double)SM * KPT * kThreads;   // records per rep
    int* sink;
    CK(cudaMalloc(&sink, 64));
    Rec *A, *B;
    const size_t nrec = (size_t)1 << 22;   // 4 Mi records = 128 MiB each
    CK(cudaMalloc(&A, nrec * sizeof(Rec)));
    CK(cudaMalloc(&B, nrec * sizeof(Rec)));
    CK(cudaMemset(A, 0, nrec * sizeof(Rec)));
    CK(cudaMemset(B, 0, nrec * sizeof(Rec)));
    unsigned masks[3] = {(1u << 18) - 1, (1u << 20) - 1, (1u << 22) - 1};
    for (int m = 0; m < 3; ++m) {
        const double mb = (masks[m] + 1.0) * 32 / 1048576.0;
        for (int co = 0; co < 2; ++co) {
            float ms;
            ms = time_ms([&] { k_records<2, 1><<<SM, kThreads>>>(A, B, masks[m], REPS, co, sink); });
            printf("records RB=2 v8  %s 2x%4.0f MiB: %.2f us per 2^20\n", co ? "coalesced" : "random   ", mb, ms * 1e3 / (REPS * per_rep) * 1048576.0);
            ms = time_ms([&] { k_records<4, 1><<<SM, kThreads>>>(A, B, masks[m], REPS, co, sink); });
            printf("records RB=4 v8  %s 2x%4.0f MiB: %.2f us per 2^20\n", co ? "coalesced" : "random   ", mb, ms * 1e3 / (REPS * per_rep) * 1048576.0);
            ms = time_ms([&] { k_records<7, 1><<<SM, kThreads>>>(A, B, masks[m], REPS, co, sink); });
            printf("records RB=7 v8  %s 2x%4.0f MiB: %.2f us per 2^20\n", co ? "coalesced" : "random   ", mb, ms * 1e3 / (REPS * per_rep) * 1048576.0);
            ms = time_ms([&] { k_records<2, 0><<<SM, kThreads>>>(A, B, masks[m], REPS, co, sink); });
            printf("records RB=2 2v4 %s 2x%4.0f MiB: %.2f us per 2^20\n", co ? "coalesced" : "random   ", mb, ms * 1e3 / (REPS * per_rep) * 1048576.0);
            ms = time_ms([&] { k_records<4, 0><<<SM, kThreads>>>(A, B, masks[m], REPS, co, sink); });
            printf("records RB=4 2v4 %s 2x%4.0f MiB: %.2f us per 2^20\n", co ? "coalesced" : "random   ", mb, ms * 1e3 / (REPS * per_rep) * 1048576.0);
        }
    }
    {
        unsigned pm[2] = {(1u << 20) - 1, (1u << 22) - 1};
        for (int m = 0; m < 2; ++m) {
            float ms = time_ms([&] { k_gather16<<<SM, kThreads>>>((const double2*)A, pm[m], REPS, (double*)sink); });
            printf("gather 16 B random, %3.0f MiB table, 7 in flight: %.2f us per 2^20\n", (pm[m] + 1.0) * 16 / 1048576.0,
                   ms * 1e3 / (REPS * per_rep) * 1048576.0);
        }
        for (int mode = 0; mode < 2; ++mode) {
            float ms = time_ms([&] { k_scatter16<<<SM, kThreads>>>((int4*)A, (1u << 20) - 1, REPS, mode); });
            printf("scatter 16 B %s, 16 MiB table: %.2f us per 2^20\n", mode ? "runs of 48" : "random    ",
                   ms * 1e3 / (REPS * per_rep) * 1048576.0);
        }
    }
    {
        int* gh;
        CK(cudaMalloc(&gh, 16 * 8192 * 4));
        CK(cudaMemset(gh, 0, 16 * 8192 * 4));
        const int ncs[4] = {1, 4, 8, 16};
        for (int q = 0; q < 4; ++q) {
            float ms = time_ms([&] { k_histmerge<<<SM, kThreads>>>(gh, 8192, ncs[q], 3000, REPS); });
            printf("hist merge 3000 atomics per CTA into 8192 bins x %2d copies + fence: %.2f us per step\n", ncs[q], ms * 1e3 / REPS);
        }
        float ms = time_ms([&] { k_histmerge<<<SM, kThreads>>>(gh, 8192 * 4, 1, 3000, REPS); });
        printf("hist merge 3000 atomics per CTA into 32768 bins x 1 copy + fence: %.2f us per step\n", ms * 1e3 / REPS);
        ms = time_ms([&] { k_histmerge<<<SM, kThreads>>>(gh, 8192, 4, 0, REPS); });
        printf("(empty: sync + fence only: %.2f us per step)\n", ms * 1e3 / REPS);
    }
    {
        long long* cyc;
        CK(cudaMalloc(&cyc, 8));
        for (int f = 0; f < 2; ++f) {
            float ms = time_ms([&] { k_store_fence<<<SM, kThreads>>>((int4*)A, 200, f, cyc); });
            printf("7 coalesced 16 B stores per thread + sync%s: %.2f us per round\n", f ? " + fence" : "        ", ms * 1e3 / 200);
        }
    }
    return 0;
}
