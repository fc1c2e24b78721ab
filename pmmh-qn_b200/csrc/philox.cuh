// philox.cuh -- Philox4x32-10 counter-based generator (Salmon et al. 2011) shared by the
// Crank-Nicolson kernel and the split particle filter: element k of a stream is a pure
// function of (seed, offset, k), so any rank can regenerate any u[t][j].
#pragma once
#include <stdint.h>

namespace pmmh {

__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
        const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0;
        c[1] = n1;
        c[2] = n2;
        c[3] = n3;
        k0 += W0;
        k1 += W1;
    }
}

__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    const unsigned long long v = (((unsigned long long)hi << 32) | lo) >> 11;
    return ((double)v + 0.5) * 1.1102230246251565e-16;   // (0, 1)
}

// Standard normal number k of the stream (seed, offset): Box-Muller on counter offset + k/2,
// cosine branch for even k, sine branch for odd k -- the same numbers pmmh_crank_nicolson
// draws for element k when d_xi == NULL.
__device__ __forceinline__ double philox_normal(unsigned long long seed, unsigned long long offset,
                                                unsigned long long k) {
    const unsigned long long ctr = offset + (k >> 1);
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u1 = u53(c[0], c[1]), u2 = u53(c[2], c[3]);
    const double r = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    return (k & 1ull) ? r * sn : r * cs;
}

}  // namespace pmmh
