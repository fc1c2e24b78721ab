// pf_model.cuh -- the device-function interface a scalar-state model implements to run on the
// chain kernel (sv_chain.cu): the three things the reference tells a user to change when porting
// its Cython particle smoother to another model (/root/reference/python/README.md:73-76:
// "how the propagation and weighting of particles is carried out as well as the gradients of the
// log joint distribution of states and measurements").  Everything else -- sorted correlated
// systematic resampling, the counting sort, cumulative weights, fixed-lag bookkeeping, the tail,
// the reference's indexing quirks Q5 / Q6 -- is model independent and lives in the kernel.
//
// A model M provides (all __device__, fp64, no contraction):
//   struct M::Const                      constants derived from the parameter vector (<= 4 doubles)
//   M::init(Const&, const double* par)
//   M::initial_state(c)                  value of every particle of generation 0 (the reference's
//                                        flps starts all particles at one point, Q1)
//   M::propagate(c, x_parent, y_prev, u) child value; u = the particle's standard normal
//                                        (SV: stochastic_volatility.pyx:354-358)
//   M::child_range(c, xmin, xmax, y_prev, nsd, lo, hi)
//                                        an interval holding the children of parents in
//                                        [xmin, xmax] up to nsd innovation sd (sort-bin range)
//   M::logw(c, x, y)                     log observation density up to the reference's constant
//                                        (SV: :427-437)
//   M::logw_max(c, lo, hi, y)            an upper bound of logw on [lo, hi] that is attained or
//                                        nearly so (the shift of the weights; any value cancels)
//   M::score_main(c, curr, next, y, g)   the 4 score terms of the pair (curr, next) in the main
//                                        loop; y = obs[time(curr) - 1] as the reference hands it
//                                        over (Q5) (SV: :452-465)
//   M::score_tail(c, curr, next, y, g)   the same in the tail loop, y = obs[time(next) - 1]
//                                        wrapped (Q6) (SV: :548-557, another operation order)
// Unused gradient slots stay 0.  Model ids are part of the C ABI (include/pmmh_qn.h).
#pragma once
#include <math.h>

#include "sv_math.cuh"

namespace pmmh {

// ---- 0: stochastic volatility with leverage (the reference's model; models/stochastic_volatility.py)
struct SvLeverageModel {
    typedef SvConst Const;
    static __device__ __forceinline__ void init(Const& c, const double* par) { sv_const_init(c, par); }
    static __device__ __forceinline__ double initial_state(const Const& c) {
        const double stdev0 = c.sigmav / sqrt(1.0 - (c.phi * c.phi));
        return c.mu + stdev0 * 0.0;   // :309 (Q1)
    }
    static __device__ __forceinline__ double propagate(const Const& c, double xp, double y1, double u) {
        double mean = c.mu + c.phi * (xp - c.mu);
        mean += c.sr * exp(-0.5 * xp) * y1;
        return mean + c.sd * u;
    }
    static __device__ __forceinline__ void child_range(const Const& c, double xmin, double xmax, double y1, double nsd,
                                                       double& lo, double& hi) {
        sv_child_range(c, xmin, xmax, y1, nsd, lo, hi);
    }
    static __device__ __forceinline__ double logw(const Const&, double x, double y) {
        const double e = exp(-0.5 * x);
        return (-0.91893853320467267 - 0.5 * x) - (0.5 * (y * y)) * (e * e);
    }
    static __device__ __forceinline__ double logw_max(const Const& c, double lo, double hi, double y) {
        double xs = log(y * y);   // the log-weight is concave in x with its maximum at log y^2
        if (!(xs >= lo)) xs = lo;
        if (xs > hi) xs = hi;
        if (!isfinite(xs)) xs = 0.0;
        return logw(c, xs, y);
    }
    static __device__ __forceinline__ void score_main(const Const& c, double curr, double next, double y, double g[4]) {
        double sq;
        sv_score_main(c, curr, next, y, sq, g);
    }
    static __device__ __forceinline__ void score_tail(const Const& c, double curr, double next, double y, double g[4]) {
        double sq;
        sv_score_tail(c, curr, next, y, sq, g);
    }
};

// ---- 1: linear Gaussian state space model  x_t = phi x_{t-1} + sigma_v v_t,  y_t = x_t + sigma_e e_t
// (parameters phi, sigma_v, sigma_e; the fourth slot is unused).  Not a model of the reference: it is
// here to show that a second model plugs into the same kernel, and because its exact likelihood is
// known (Kalman filter), which the tests use.
struct LinearGaussianModel {
    struct Const {
        double phi, sv, se, inv_sv2, inv_se2, log_se;
    };
    static __device__ __forceinline__ void init(Const& c, const double* par) {
        c.phi = par[0];
        c.sv = par[1];
        c.se = par[2];
        c.inv_sv2 = 1.0 / (c.sv * c.sv);
        c.inv_se2 = 1.0 / (c.se * c.se);
        c.log_se = log(c.se);
    }
    static __device__ __forceinline__ double initial_state(const Const&) { return 0.0; }
    static __device__ __forceinline__ double propagate(const Const& c, double xp, double, double u) {
        return c.phi * xp + c.sv * u;
    }
    static __device__ __forceinline__ void child_range(const Const& c, double xmin, double xmax, double, double nsd,
                                                       double& lo, double& hi) {
        const double a = c.phi * xmin, b = c.phi * xmax;
        lo = fmin(a, b) - nsd * c.sv;
        hi = fmax(a, b) + nsd * c.sv;
    }
    static __device__ __forceinline__ double logw(const Const& c, double x, double y) {
        const double r = y - x;
        return (-0.91893853320467267 - c.log_se) - (0.5 * (r * r)) * c.inv_se2;
    }
    static __device__ __forceinline__ double logw_max(const Const& c, double lo, double hi, double y) {
        double xs = y;
        if (!(xs >= lo)) xs = lo;
        if (xs > hi) xs = hi;
        if (!isfinite(xs)) xs = 0.0;
        return logw(c, xs, y);
    }
    // d/d(phi, sigma_v, sigma_e) of log N(next; phi curr, sigma_v^2) + log N(y; curr, sigma_e^2)
    static __device__ __forceinline__ void score_main(const Const& c, double curr, double next, double y, double g[4]) {
        const double r = next - c.phi * curr;
        const double ry = y - curr;
        g[0] = (r * curr) * c.inv_sv2;
        g[1] = ((r * r) * c.inv_sv2 - 1.0) / c.sv;
        g[2] = ((ry * ry) * c.inv_se2 - 1.0) / c.se;
        g[3] = 0.0;
    }
    static __device__ __forceinline__ void score_tail(const Const& c, double curr, double next, double y, double g[4]) {
        score_main(c, curr, next, y, g);
    }
};

// ---- 2: the same linear Gaussian model run as a FULLY ADAPTED particle filter (Pitt & Shephard): particles are
// drawn from p(x_t | x_{t-1}, y_t) and weighted with the predictive density p(y_{t+1} | x_t), both in closed form
// for this model.  The kernel's algorithm is untouched: the caller hands over the observations shifted by one
// (obs'[t] = obs[t + 1]), so `propagate` sees y_t where the SV model sees y_{t-1} (leverage) and `logw` sees
// y_{t+1}.  The reference has no fully adapted filter (SURVEY 8a note): parity is unpinned, the check is the exact
// Kalman likelihood.  Outputs: log-likelihood of y_2 .. y_T given y_1 (the caller adds log p(y_1)), and
// E[x_t | y_1 .. y_{t+1}] in `filt`; the score terms are not defined for this variant (gradient rows 0).
struct LinearGaussianFullyAdapted {
    struct Const {
        double phi, sv, se, k, sdp, inv_s2, log_s;   // gain, proposal sd, predictive variance
    };
    static __device__ __forceinline__ void init(Const& c, const double* par) {
        c.phi = par[0];
        c.sv = par[1];
        c.se = par[2];
        const double s2 = c.sv * c.sv + c.se * c.se;
        c.k = (c.sv * c.sv) / s2;
        c.sdp = sqrt((c.sv * c.sv) * (c.se * c.se) / s2);
        c.inv_s2 = 1.0 / s2;
        c.log_s = 0.5 * log(s2);
    }
    static __device__ __forceinline__ double initial_state(const Const&) { return 0.0; }
    // x_t | x_{t-1}, y_t ~ N(phi x + k (y_t - phi x), sv^2 se^2 / (sv^2 + se^2))
    static __device__ __forceinline__ double propagate(const Const& c, double xp, double y_now, double u) {
        const double m = c.phi * xp;
        return (m + c.k * (y_now - m)) + c.sdp * u;
    }
    static __device__ __forceinline__ void child_range(const Const& c, double xmin, double xmax, double y_now, double nsd,
                                                       double& lo, double& hi) {
        const double a = c.phi * xmin, b = c.phi * xmax;
        const double ma = a + c.k * (y_now - a), mb = b + c.k * (y_now - b);
        lo = fmin(ma, mb) - nsd * c.sdp;
        hi = fmax(ma, mb) + nsd * c.sdp;
    }
    // log p(y_{t+1} | x_t) = log N(y; phi x, sv^2 + se^2)
    static __device__ __forceinline__ double logw(const Const& c, double x, double y_next) {
        const double r = y_next - c.phi * x;
        return (-0.91893853320467267 - c.log_s) - (0.5 * (r * r)) * c.inv_s2;
    }
    static __device__ __forceinline__ double logw_max(const Const& c, double lo, double hi, double y_next) {
        double xs = (c.phi != 0.0) ? y_next / c.phi : lo;
        const double l2 = fmin(lo, hi), h2 = fmax(lo, hi);
        if (!(xs >= l2)) xs = l2;
        if (xs > h2) xs = h2;
        if (!isfinite(xs)) xs = 0.0;
        return logw(c, xs, y_next);
    }
    static __device__ __forceinline__ void score_main(const Const&, double, double, double, double g[4]) {
        g[0] = g[1] = g[2] = g[3] = 0.0;
    }
    static __device__ __forceinline__ void score_tail(const Const&, double, double, double, double g[4]) {
        g[0] = g[1] = g[2] = g[3] = 0.0;
    }
};

}  // namespace pmmh
