// sv_math.cuh -- per-particle arithmetic of the SV model shared by the SV kernels
// (propagation constants, log-weight, score and Hessian terms).  Every function cites the
// reference lines it restates (/root/reference/python/state/particle_methods/
// stochastic_volatility.pyx).  fp64, reference operation order, compiled with -fmad=false.
#pragma once
#include <math.h>

namespace pmmh {

struct SvConst {
    double mu, phi, sigmav, rho;
    double sd, q, rho_term, one_m_phi, one_m_phi2, inv_sv, inv_sv2, sr;
};

__device__ __forceinline__ void sv_const_init(SvConst& c, const double* par) {
    c.mu = par[0];
    c.phi = par[1];
    c.sigmav = par[2];
    c.rho = par[3];
    c.rho_term = 1.0 - c.rho * c.rho;
    c.q = 1.0 / (c.sigmav * c.sigmav * (1.0 - c.rho * c.rho));
    c.sd = sqrt(c.rho_term) * c.sigmav;
    c.one_m_phi = 1.0 - c.phi;
    c.one_m_phi2 = 1.0 - c.phi * c.phi;
    c.inv_sv = 1.0 / c.sigmav;
    c.inv_sv2 = 1.0 / (c.sigmav * c.sigmav);
    c.sr = c.sigmav * c.rho;
}

// Range [lo, hi] that holds the children of parents in [xmin, xmax] up to `nsd` innovation
// standard deviations: exact extrema of the propagation mean
//   f(x) = mu + phi (x - mu) + sigma_v rho exp(-x/2) y_prev        (:355-356)
// over the interval (end points and, if inside, the stationary point) widened by nsd * sd.
__device__ __forceinline__ void sv_child_range(const SvConst& c, double xmin, double xmax, double y1,
                                               double nsd, double& lo, double& hi) {
    const double cc = c.sr * y1;
    const double fa = (c.mu + c.phi * (xmin - c.mu)) + cc * exp(-0.5 * xmin);
    const double fb = (c.mu + c.phi * (xmax - c.mu)) + cc * exp(-0.5 * xmax);
    double fmn = fmin(fa, fb), fmx = fmax(fa, fb);
    if (c.phi * cc > 0.0) {
        const double xs = -2.0 * log(2.0 * c.phi / cc);
        if (xs > xmin && xs < xmax) {
            const double fs = (c.mu + c.phi * (xs - c.mu)) + cc * exp(-0.5 * xs);
            fmn = fmin(fmn, fs);
            fmx = fmax(fmx, fs);
        }
    }
    lo = fmn - nsd * c.sd;
    hi = fmx + nsd * c.sd;
}

// norm_logpdf(y, 0, exp(x/2)), stochastic_volatility.pyx:428,659-664 (literal)
__device__ __forceinline__ double sv_logw(double x, double y) {
    double s = exp(0.5 * x);
    double part2 = -log(s);
    double part3 = -0.5 * (y - 0.0) * (y - 0.0) / (s * s);
    return -0.91893853320467267 + part2 + part3;
}

__device__ __forceinline__ int sv_bin(double x, double lo, double scale, int NB) {
    double t = (x - lo) * scale;
    if (!(t >= 0.0)) return 0;
    if (t >= (double)NB) return NB - 1;
    return (int)t;
}

// main-loop score terms, stochastic_volatility.pyx:452-465; e = exp(-0.5 * curr)
__device__ __forceinline__ void sv_score_main_e(const SvConst& c, double curr, double e, double next,
                                                double y, double& sq, double g[4]) {
    sq = next - c.mu - c.phi * (curr - c.mu);
    sq -= c.sr * e * y;
    g[0] = c.q * sq * c.one_m_phi;
    g[1] = c.q * sq * (curr - c.mu) * c.one_m_phi2;
    double g2 = sq;
    g2 += c.sr * e * y;
    g2 *= c.q * sq;
    g2 -= 1.0;
    g[2] = g2;
    double g3 = c.rho - c.q * c.rho * sq * sq;
    g3 += c.inv_sv * sq * e * y;
    g[3] = g3;
}
__device__ __forceinline__ void sv_score_main(const SvConst& c, double curr, double next, double y,
                                              double& sq, double g[4]) {
    sv_score_main_e(c, curr, exp(-0.5 * curr), next, y, sq, g);
}

// tail score terms, stochastic_volatility.pyx:548-557 (different operation order)
__device__ __forceinline__ void sv_score_tail_e(const SvConst& c, double curr, double e, double next,
                                                double y, double& sq, double g[4]) {
    sq = next - c.mu - c.phi * (curr - c.mu);
    sq -= c.sr * e * y;
    g[0] = c.q * sq * c.one_m_phi;
    g[1] = c.q * sq * (curr - c.mu) * c.one_m_phi2;
    double g2 = c.q * sq * sq - 1.0;
    g2 += c.q * sq * c.sigmav * c.rho * e * y;
    g[2] = g2;
    double g3 = c.rho;
    g3 -= c.q * c.rho * sq * sq;
    g3 += c.q * sq * c.sigmav * e * y * c.rho_term;
    g[3] = g3;
}
__device__ __forceinline__ void sv_score_tail(const SvConst& c, double curr, double next, double y,
                                              double& sq, double g[4]) {
    sv_score_tail_e(c, curr, exp(-0.5 * curr), next, y, sq, g);
}

// Upper-triangular sub_hessian1 / sub_hessian2 terms, stochastic_volatility.pyx:473-519,
// accumulated with the isfinite guards of :521-534.  Order: (0,0)(0,1)(0,2)(0,3)(1,1)(1,2)
// (1,3)(2,2)(2,3)(3,3); acc[0..9] = hessian1, acc[10..19] = hessian2.
// (e = exp(-curr / 2) and e2 = exp(-curr) are passed in: the chain kernel shares e with the score terms and
// uses e * e for e2)
__device__ __forceinline__ void sv_hessian_terms_e(const SvConst& c, double curr, double e, double e2, double sq,
                                                   double yl, const double g[4], const double al[4], double w,
                                                   double* acc);
__device__ __forceinline__ void sv_hessian_terms(const SvConst& c, double curr, double sq, double yl,
                                                 const double g[4], const double al[4], double w,
                                                 double* acc) {
    sv_hessian_terms_e(c, curr, exp(-0.5 * curr), exp(-curr), sq, yl, g, al, w, acc);
}
__device__ __forceinline__ void sv_hessian_terms_e(const SvConst& c, double curr, double e, double e2, double sq,
                                                   double yl, const double g[4], const double al[4], double w,
                                                   double* acc) {
    const double cm = curr - c.mu;
    double h1[10], h2[10];
    // (0,0)
    h1[0] = -c.q * (c.one_m_phi * c.one_m_phi);
    // (1,1)
    double t = 2.0 * c.phi * sq + cm * c.one_m_phi2;
    t *= -c.q * cm * c.one_m_phi2;
    h1[4] = t;
    // (2,2)
    t = -2.0 * c.q * sq * sq;
    t -= 2.0 * c.q * sq * c.rho * c.sigmav * e * yl;
    {
        double r = c.rho * c.sigmav * e * yl;
        t -= c.q * (r * r);
    }
    t += c.q * sq * c.rho * c.sigmav * e * yl;
    h1[7] = t;
    // (3,3)   note the reference's "sigmav*(-2)" typo
    t = c.rho_term - 2.0 * c.q * (c.rho * c.rho) * (sq * sq) - c.sigmav * (-2.0) * (sq * sq);
    t += 2.0 * c.inv_sv * c.rho * sq * e * yl;
    t -= e2 * (yl * yl) * c.rho_term;
    h1[9] = t;
    // (0,1)
    t = -c.q * cm * c.one_m_phi - c.q * sq;
    t *= c.one_m_phi2;
    h1[1] = t;
    // (0,2)
    t = -2.0 * sq * c.one_m_phi;
    t -= c.q * c.one_m_phi * c.sigmav * c.rho * sq * e * yl;
    h1[2] = t;
    // (0,3)
    t = 2.0 * c.q * c.rho * sq * c.one_m_phi;
    t -= c.inv_sv2 * c.one_m_phi * c.sigmav * e * yl;
    h1[3] = t;
    // (1,2)
    t = -2.0 * sq - c.rho * c.sigmav * e * yl;
    t *= c.q * cm * c.one_m_phi2;
    h1[5] = t;
    // (1,3)
    t = 2.0 * c.rho * sq - c.sigmav * e * yl * c.rho_term;
    t *= c.q * cm * c.one_m_phi2;
    h1[6] = t;
    // (2,3)
    t = 2.0 * c.q * (sq * sq) * c.rho;
    t += 2.0 * (c.rho * c.rho) * c.q * sq * c.sigmav * e * yl;
    t -= c.rho * e2 * (yl * yl);
    t += c.inv_sv * sq * e * yl;
    h1[8] = t;

    h2[0] = g[0] * g[0] + 2.0 * al[0] * g[0];
    h2[1] = g[0] * g[1] + al[0] * g[1] + al[1] * g[0];
    h2[2] = g[0] * g[2] + al[0] * g[2] + al[2] * g[0];
    h2[3] = g[0] * g[3] + al[0] * g[3] + al[3] * g[0];
    h2[4] = g[1] * g[1] + 2.0 * al[1] * g[1];
    h2[5] = g[1] * g[2] + al[1] * g[2] * al[2] * g[1];   // '*' typos of :513,514,517 kept
    h2[6] = g[1] * g[3] + al[1] * g[3] * al[3] * g[1];
    h2[7] = g[2] * g[2] + 2.0 * al[2] * g[2];
    h2[8] = g[2] * g[3] + al[2] * g[3] * al[3] * g[2];
    h2[9] = g[3] * g[3] + 2.0 * al[3] * g[3];
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        if (isfinite(h1[k])) acc[k] += h1[k] * w;
        if (isfinite(h2[k])) acc[10 + k] += h2[k] * w;
    }
}

// the same terms from ey = exp(-curr / 2) * yl (what the grid kernel keeps in its payload): e and yl only ever
// appear as that product
__device__ __forceinline__ void sv_hessian_terms_ey(const SvConst& c, double curr, double ey, double sq,
                                                    const double g[4], const double al[4], double w,
                                                    double* acc) {
    const double cm = curr - c.mu;
    double h1[10], h2[10];
    // (0,0)
    h1[0] = -c.q * (c.one_m_phi * c.one_m_phi);
    // (1,1)
    double t = 2.0 * c.phi * sq + cm * c.one_m_phi2;
    t *= -c.q * cm * c.one_m_phi2;
    h1[4] = t;
    // (2,2)
    t = -2.0 * c.q * sq * sq;
    t -= 2.0 * c.q * sq * c.rho * c.sigmav * ey;
    {
        double r = c.rho * c.sigmav * ey;
        t -= c.q * (r * r);
    }
    t += c.q * sq * c.rho * c.sigmav * ey;
    h1[7] = t;
    // (3,3)   note the reference's "sigmav*(-2)" typo
    t = c.rho_term - 2.0 * c.q * (c.rho * c.rho) * (sq * sq) - c.sigmav * (-2.0) * (sq * sq);
    t += 2.0 * c.inv_sv * c.rho * sq * ey;
    t -= (ey * ey) * c.rho_term;
    h1[9] = t;
    // (0,1)
    t = -c.q * cm * c.one_m_phi - c.q * sq;
    t *= c.one_m_phi2;
    h1[1] = t;
    // (0,2)
    t = -2.0 * sq * c.one_m_phi;
    t -= c.q * c.one_m_phi * c.sigmav * c.rho * sq * ey;
    h1[2] = t;
    // (0,3)
    t = 2.0 * c.q * c.rho * sq * c.one_m_phi;
    t -= c.inv_sv2 * c.one_m_phi * c.sigmav * ey;
    h1[3] = t;
    // (1,2)
    t = -2.0 * sq - c.rho * c.sigmav * ey;
    t *= c.q * cm * c.one_m_phi2;
    h1[5] = t;
    // (1,3)
    t = 2.0 * c.rho * sq - c.sigmav * ey * c.rho_term;
    t *= c.q * cm * c.one_m_phi2;
    h1[6] = t;
    // (2,3)
    t = 2.0 * c.q * (sq * sq) * c.rho;
    t += 2.0 * (c.rho * c.rho) * c.q * sq * c.sigmav * ey;
    t -= c.rho * (ey * ey);
    t += c.inv_sv * sq * ey;
    h1[8] = t;

    h2[0] = g[0] * g[0] + 2.0 * al[0] * g[0];
    h2[1] = g[0] * g[1] + al[0] * g[1] + al[1] * g[0];
    h2[2] = g[0] * g[2] + al[0] * g[2] + al[2] * g[0];
    h2[3] = g[0] * g[3] + al[0] * g[3] + al[3] * g[0];
    h2[4] = g[1] * g[1] + 2.0 * al[1] * g[1];
    h2[5] = g[1] * g[2] + al[1] * g[2] * al[2] * g[1];   // '*' typos of :513,514,517 kept
    h2[6] = g[1] * g[3] + al[1] * g[3] * al[3] * g[1];
    h2[7] = g[2] * g[2] + 2.0 * al[2] * g[2];
    h2[8] = g[2] * g[3] + al[2] * g[3] * al[3] * g[2];
    h2[9] = g[3] * g[3] + 2.0 * al[3] * g[3];
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        if (isfinite(h1[k])) acc[k] += h1[k] * w;
        if (isfinite(h2[k])) acc[10 + k] += h2[k] * w;
    }
}

// alpha recursion, own term of a child (:361-390): x = the new state, curr = the "current" state as the reference
// reads it (Q7), yi = obs[i], yl = obs[i - LAG] (Q8)
__device__ __forceinline__ void sv_alpha_terms(const SvConst& c, double x, double curr, double yi, double yl,
                                               double al[4]) {
    double sq = x - c.mu - c.phi * (curr - c.mu);
    const double ec = exp(-0.5 * curr);
    sq -= c.sr * ec * yl;
    al[0] = c.q * sq * c.one_m_phi;
    al[1] = c.q * sq * (curr - c.mu) * c.one_m_phi2;
    double a2 = sq;
    a2 += c.sr * ec * yi;
    a2 *= c.q * sq;
    a2 -= 1.0;
    al[2] = a2;
    double a3 = c.rho - c.q * c.rho * sq * sq;
    a3 += c.inv_sv * sq * ec * yi;
    al[3] = a3;
}

__device__ __forceinline__ double obs_wrap(const double* obs, int k, int nobs) {
    return obs[k < 0 ? k + nobs : k];   // Cython memoryview wraparound (Q8)
}

}  // namespace pmmh
