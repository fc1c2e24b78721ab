// microbench.cu -- B200 primitives that size the SV particle-filter kernel design:
// grid barrier latency, global / shared atomics, random 32-byte sector gathers, scattered
// 32-byte record writes, coalesced L2 reads and fp64 exp throughput.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <math.h>
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

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

// ---- 1a: flag all-gather barrier (every CTA publishes a stamp, everyone polls all stamps)
__global__ void __launch_bounds__(kThreads, 1) k_barrier_flags(unsigned* stamps, int iters, long long* cyc) {
    const int G = gridDim.x;
    long long t0 = clock64();
    for (int it = 1; it <= iters; ++it) {
        __syncthreads();
        if (threadIdx.x == 0) st_release(&stamps[blockIdx.x], (unsigned)it);
        for (int c = threadIdx.x; c < G; c += blockDim.x)
            while (ld_acquire(&stamps[c]) < (unsigned)it) {
            }
        __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = clock64() - t0;
}
// ---- 1b: counter barrier (one atomic per CTA, spin on a generation word)
__global__ void __launch_bounds__(kThreads, 1) k_barrier_counter(unsigned* ctr, int iters, long long* cyc) {
    const unsigned G = gridDim.x;
    long long t0 = clock64();
    for (int it = 1; it <= iters; ++it) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(ctr, 1u);
            while (ld_acquire(ctr) < G * (unsigned)it) {
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = clock64() - t0;
}
// ---- 1c: flag barrier, stamps padded to 128 B each
__global__ void __launch_bounds__(kThreads, 1) k_barrier_flags_pad(unsigned* stamps, int iters, long long* cyc) {
    const int G = gridDim.x;
    long long t0 = clock64();
    for (int it = 1; it <= iters; ++it) {
        __syncthreads();
        if (threadIdx.x == 0) st_release(&stamps[blockIdx.x * 32], (unsigned)it);
        for (int c = threadIdx.x; c < G; c += blockDim.x)
            while (ld_acquire(&stamps[c * 32]) < (unsigned)it) {
            }
        __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = clock64() - t0;
}

// ---- 2: global atomics with return value
__global__ void __launch_bounds__(kThreads, 1) k_gatomic(int* cnt, unsigned nb_mask, int per_thread, int reps, int* sink) {
    int acc = 0;
    const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int r = 0; r < reps; ++r) {
#pragma unroll 7
        for (int k = 0; k < per_thread; ++k) {
            const unsigned h = hash32(gid * 131u + k * 7919u + r * 104729u);
            acc += atomicAdd(&cnt[h & nb_mask], 1);
        }
    }
    if (acc == -12345) *sink = acc;
}
// ---- 2b: global reductions without return
__global__ void __launch_bounds__(kThreads, 1) k_gred(int* cnt, unsigned nb_mask, int per_thread, int reps) {
    const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int r = 0; r < reps; ++r) {
#pragma unroll 7
        for (int k = 0; k < per_thread; ++k) {
            const unsigned h = hash32(gid * 131u + k * 7919u + r * 104729u);
            atomicAdd(&cnt[h & nb_mask], 1);
        }
    }
}
// ---- 3: shared-memory atomics with return value
__global__ void __launch_bounds__(kThreads, 1) k_satomic(unsigned nb_mask, int per_thread, int reps, int* sink) {
    extern __shared__ int s_cnt[];
    for (int k = threadIdx.x; k <= (int)nb_mask; k += blockDim.x) s_cnt[k] = 0;
    __syncthreads();
    int acc = 0;
    const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int r = 0; r < reps; ++r) {
#pragma unroll 7
        for (int k = 0; k < per_thread; ++k) {
            const unsigned h = hash32(gid * 131u + k * 7919u + r * 104729u);
            acc += atomicAdd(&s_cnt[h & nb_mask], 1);
        }
    }
    if (acc == -12345) *sink = acc;
}
// ---- 4: random 32-byte record gathers (16 bytes used)
struct __align__(32) Rec {
    double a, b, c, d;
};
__global__ void __launch_bounds__(kThreads, 1) k_gather(const Rec* recs, unsigned n_mask, int per_thread, int reps, double* sink, int dependent) {
    double acc = 0.0;
    const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int r = 0; r < reps; ++r) {
        if (!dependent) {
#pragma unroll 7
            for (int k = 0; k < per_thread; ++k) {
                const unsigned h = hash32(gid * 131u + k * 7919u + r * 104729u) & n_mask;
                const double2 v = *reinterpret_cast<const double2*>(&recs[h]);
                acc += v.x + v.y;
            }
        } else {
            unsigned h = hash32(gid * 131u + r * 104729u) & n_mask;
            for (int k = 0; k < per_thread; ++k) {
                const double2 v = *reinterpret_cast<const double2*>(&recs[h]);
                acc += v.y;
                h = (unsigned)(__double_as_longlong(v.x)) & n_mask;
            }
        }
    }
    if (acc == -12345.0) *sink = acc;
}
// ---- 5: scattered 32-byte record writes
__global__ void __launch_bounds__(kThreads, 1) k_scatter(Rec* recs, unsigned n_mask, int per_thread, int reps) {
    const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int r = 0; r < reps; ++r) {
#pragma unroll 7
        for (int k = 0; k < per_thread; ++k) {
            const unsigned h = hash32(gid * 131u + k * 7919u + r * 104729u) & n_mask;
            Rec v;
            v.a = (double)h;
            v.b = (double)k;
            v.c = 1.0;
            v.d = 2.0;
            double4* p = reinterpret_cast<double4*>(&recs[h]);
            *p = make_double4(v.a, v.b, v.c, v.d);
        }
    }
}
// ---- 6: coalesced reads (double2 per thread)
__global__ void __launch_bounds__(kThreads, 1) k_stream(const double2* src, size_t n2, int reps, double* sink) {
    double acc = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
            const double2 v = src[i];
            acc += v.x + v.y;
        }
    if (acc == -12345.0) *sink = acc;
}
// ---- 6b: coalesced copy
__global__ void __launch_bounds__(kThreads, 1) k_copy(const double2* src, double2* dst, size_t n2, int reps) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) dst[i] = src[i];
}
// ---- 7: fp64 exp throughput
__global__ void __launch_bounds__(kThreads, 1) k_exp(int per_thread, int reps, double* sink) {
    double acc = 0.0;
    double x = -0.5 + 1e-6 * threadIdx.x;
    for (int r = 0; r < reps; ++r)
#pragma unroll 4
        for (int k = 0; k < per_thread; ++k) {
            acc += exp(x);
            x += 1e-3;
        }
    if (acc == -12345.0) *sink = acc;
}
__global__ void __launch_bounds__(kThreads, 1) k_fma64(int per_thread, int reps, double* sink) {
    double a0 = 1.0 + threadIdx.x, a1 = 2.0, a2 = 3.0, a3 = 4.0;
    const double m = 1.0000001, c = 1e-9;
    for (int r = 0; r < reps; ++r)
#pragma unroll 8
        for (int k = 0; k < per_thread; ++k) {
            a0 = fma(a0, m, c);
            a1 = fma(a1, m, c);
            a2 = fma(a2, m, c);
            a3 = fma(a3, m, c);
        }
    if (a0 + a1 + a2 + a3 == -12345.0) *sink = a0;
}

template <typename F>
float time_ms(F f, int warm = 1, int runs = 3) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < warm; ++i) f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < runs; ++i) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int SM = prop.multiProcessorCount;
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("device %s, %d SMs, clock %d kHz, L2 %d MB\n", prop.name, SM, clk_khz, prop.l2CacheSize >> 20);
    const int PER = 7, REPS = 20;
    const double ops = (double)SM * kThreads * PER * REPS;

    // 1: barriers
    {
        unsigned* d;
        long long* dc;
        CK(cudaMalloc(&d, 64 * 1024));
        CK(cudaMalloc(&dc, 8));
        const int iters = 2000;
        void* fns[3] = {(void*)k_barrier_flags, (void*)k_barrier_counter, (void*)k_barrier_flags_pad};
        const char* names[3] = {"flags", "counter", "flags_padded"};
        for (int which = 0; which < 3; ++which) {
            CK(cudaMemset(d, 0, 64 * 1024));
            int it = iters;
            void* args[] = {&d, &it, &dc};
            float ms = time_ms([&] {
                CK(cudaMemset(d, 0, 64 * 1024));
                CK(cudaLaunchCooperativeKernel(fns[which], dim3(SM), dim3(kThreads), args, 0, 0));
            }, 1, 2);
            printf("barrier %-13s grid=%d: %.3f us per barrier\n", names[which], SM, ms * 1e3 / iters);
        }
        // smaller thread count barrier for reference (256 threads)
        {
            CK(cudaMemset(d, 0, 64 * 1024));
            int it = iters;
            void* args[] = {&d, &it, &dc};
            float ms = time_ms([&] {
                CK(cudaMemset(d, 0, 64 * 1024));
                CK(cudaLaunchCooperativeKernel((void*)k_barrier_flags, dim3(SM), dim3(256), args, 0, 0));
            }, 1, 2);
            printf("barrier flags 256thr  grid=%d: %.3f us per barrier\n", SM, ms * 1e3 / iters);
        }
        cudaFree(d);
        cudaFree(dc);
    }
    int* sink_i;
    double* sink_d;
    CK(cudaMalloc(&sink_i, 8));
    CK(cudaMalloc(&sink_d, 8));
    // 2: global atomics
    {
        int* cnt;
        CK(cudaMalloc(&cnt, (size_t)4 << 20));
        CK(cudaMemset(cnt, 0, (size_t)4 << 20));
        unsigned masks[4] = {1023, 4095, 65535, (1u << 20) - 1};
        for (int m = 0; m < 4; ++m) {
            float ms = time_ms([&] { k_gatomic<<<SM, kThreads>>>(cnt, masks[m], PER, REPS, sink_i); });
            printf("global atomicAdd(ret) bins=%-8u: %.3f us per 2^20 ops\n", masks[m] + 1, ms * 1e3 / ops * 1048576.0);
            ms = time_ms([&] { k_gred<<<SM, kThreads>>>(cnt, masks[m], PER, REPS); });
            printf("global red (no ret)   bins=%-8u: %.3f us per 2^20 ops\n", masks[m] + 1, ms * 1e3 / ops * 1048576.0);
        }
        cudaFree(cnt);
    }
    // 3: shared atomics
    {
        unsigned masks[3] = {255, 4095, 16383};
        for (int m = 0; m < 3; ++m) {
            const int smem = (masks[m] + 1) * 4;
            CK(cudaFuncSetAttribute(k_satomic, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
            float ms = time_ms([&] { k_satomic<<<SM, kThreads, smem>>>(masks[m], PER, REPS, sink_i); });
            printf("shared atomicAdd(ret) bins=%-8u: %.3f us per 7168 ops per CTA\n", masks[m] + 1, ms * 1e3 / REPS);
        }
    }
    // 4/5: gathers and scatters
    {
        const size_t big = (size_t)1 << 25;   // 32 Mi records = 1 GiB
        Rec* recs;
        CK(cudaMalloc(&recs, big * sizeof(Rec)));
        CK(cudaMemset(recs, 0, big * sizeof(Rec)));
        unsigned masks[3] = {(1u << 20) - 1, (1u << 22) - 1, (1u << 25) - 1};
        for (int m = 0; m < 3; ++m) {
            const double mb = (masks[m] + 1.0) * 32 / 1048576.0;
            float ms = time_ms([&] { k_gather<<<SM, kThreads>>>(recs, masks[m], PER, REPS, sink_d, 0); });
            printf("random 32B-sector gather, %7.0f MiB table: %.3f us per 2^20 (%.0f GB/s of sectors)\n", mb,
                   ms * 1e3 / ops * 1048576.0, ops * 32 / (ms * 1e-3) / 1e9);
            ms = time_ms([&] { k_gather<<<SM, kThreads>>>(recs, masks[m], PER, REPS, sink_d, 1); });
            printf("dependent chase (7 hops),  %7.0f MiB table: %.3f us per 2^20 hops\n", mb, ms * 1e3 / ops * 1048576.0);
            ms = time_ms([&] { k_scatter<<<SM, kThreads>>>(recs, masks[m], PER, REPS); });
            printf("random 32B-sector scatter, %7.0f MiB table: %.3f us per 2^20 (%.0f GB/s of sectors)\n", mb,
                   ms * 1e3 / ops * 1048576.0, ops * 32 / (ms * 1e-3) / 1e9);
        }
        // 6: coalesced
        size_t sizes[3] = {(size_t)32 << 20, (size_t)96 << 20, (size_t)1 << 30};
        for (int s = 0; s < 3; ++s) {
            const size_t n2 = sizes[s] / 16;
            const int reps = sizes[s] > ((size_t)256 << 20) ? 4 : 40;
            float ms = time_ms([&] { k_stream<<<SM * 2, kThreads>>>((const double2*)recs, n2, reps, sink_d); });
            printf("coalesced read  %5zu MiB x%d: %.0f GB/s\n", sizes[s] >> 20, reps, (double)sizes[s] * reps / (ms * 1e-3) / 1e9);
        }
        {
            const size_t bytes = (size_t)16 << 20;
            float ms = time_ms([&] { k_copy<<<SM * 2, kThreads>>>((const double2*)recs, (double2*)((char*)recs + ((size_t)64 << 20)), bytes / 16, 40); });
            printf("coalesced copy 16+16 MiB (L2) x40: %.0f GB/s (read+write)\n", 2.0 * bytes * 40 / (ms * 1e-3) / 1e9);
            const size_t bytes2 = (size_t)400 << 20;
            ms = time_ms([&] { k_copy<<<SM * 2, kThreads>>>((const double2*)recs, (double2*)((char*)recs + ((size_t)512 << 20)), bytes2 / 16, 4); });
            printf("coalesced copy 400+400 MiB (HBM) x4: %.0f GB/s (read+write)\n", 2.0 * bytes2 * 4 / (ms * 1e-3) / 1e9);
        }
        cudaFree(recs);
    }
    // 7: fp64
    {
        const int per = 64, reps = 10;
        float ms = time_ms([&] { k_exp<<<SM, kThreads>>>(per, reps, sink_d); });
        const double n = (double)SM * kThreads * per * reps;
        printf("fp64 exp: %.3f us per 2^20 (%.1f G exp/s)\n", ms * 1e3 / n * 1048576.0, n / (ms * 1e-3) / 1e9);
        const int perf = 4096;
        ms = time_ms([&] { k_fma64<<<SM, kThreads>>>(perf, reps, sink_d); });
        const double nf = (double)SM * kThreads * perf * reps * 4;
        printf("fp64 fma: %.2f TFLOP/s\n", 2.0 * nf / (ms * 1e-3) / 1e12);
    }
    return 0;
}
