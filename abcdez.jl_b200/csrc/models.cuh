// models.cuh -- the dist!(theta, ve) -> (d, blob) plugins as registered CUDA device functors.
//
// In the reference `dist!` is an arbitrary Julia closure (src/abcdez_smc.jl:215, contract in
// docs/src/index.md:288-324).  Here a model is a struct with compile-time D (= length(prior)),
// BLOB (bytes carried with the particle, multiple of 8) and a static
//     __device__ double run(const double* theta /*push_p-ed*/, const double* data, SimRng&, double* blob)
// instantiated into the fused sweep kernels (sweep.cuh).  `ve` (per-thread scratch,
// src/abcdez_smc.jl:111) is the thread's registers.  Normative model definitions: DESIGN.md
// "Models"; the CPU oracle restates them independently.
#pragma once
#include "common.cuh"

namespace abcdez {

struct ModelData { double v[ABCDEZ_MAXDATA]; };

// examples/minimal_example.jl:17-24, test/runtests.jl:138: y ~ N(theta, sigma), d = |y - data|
struct Gauss1D {
    static constexpr int D = 1, BLOB = 0, NOISE = 0;
    static constexpr const char* name = "gauss1d";
    __device__ static __forceinline__ double run(const double* th, const double* data, SimRng& r, double*)
    {
        return fabs(th[0] + data[1] * r.n() - data[0]);
    }
};

// same simulator, the simulated y is carried as the particle's blob (docs/src/index.md:298-324)
struct Gauss1DBlob {
    static constexpr int D = 1, BLOB = 8, NOISE = 0;
    static constexpr const char* name = "gauss1d_blob";
    __device__ static __forceinline__ double run(const double* th, const double* data, SimRng& r, double* blob)
    {
        double y = th[0] + data[1] * r.n();
        blob[0] = y;
        return fabs(y - data[0]);
    }
};

// config 2: y ~ N(theta, Sigma), Sigma_ij = rho^|i-j| (stationary AR(1) noise), d = ||y - y_obs||_2
// data = y_obs[10], rho, and data[11] = sqrt(1 - rho^2), filled in by abcdez_model_bind (model_prepare_data, api.cu)
struct GaussCorr10 {
    static constexpr int D = 10, BLOB = 0, NOISE = 0;
    static constexpr const char* name = "gauss_corr10";
    // the five noise pairs, then the AR(1) recursion and the distance as one straight loop: written this way the
    // compiler schedules the (branch-free) Box-Muller pairs and the scoring freely -- 102.7 us per sweep against 108.3
    // with the scoring statements interleaved by hand after every pair; explicit two-pair lockstep code 103.5, drawing
    // the noise before the partner gathers 106.8-115.8, parking it in shared memory slower still (profiles/README.md)
    __device__ static __forceinline__ void draw(SimRng& r, double* z)
    {
#pragma unroll
        for (int k = 0; k < 10; k += 2) r.n2(z[k], z[k + 1]);
    }
    __device__ static __forceinline__ double score(const double* th, const double* data, const double* z, double*)
    {
        double rho = data[10], sr = data[11], e = 0.0, acc = 0.0;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            e = (k == 0) ? z[0] : rho * e + sr * z[k];
            double dy = th[k] + e - data[k];
            acc += dy * dy;
        }
        return sqrt(acc);
    }
    __device__ static __forceinline__ double run(const double* th, const double* data, SimRng& r, double* blob)
    {
        double z[10];
        draw(r, z);
        return score(th, data, z, blob);
    }
};

// test/runtests.jl:496-497
struct Dirac {
    static constexpr int D = 1, BLOB = 0, NOISE = 0;
    static constexpr const char* name = "dirac";
    __device__ static __forceinline__ double run(const double* th, const double* data, SimRng&, double*)
    {
        return fabs(th[0] * th[0] + 1.0 - data[0]);
    }
};

// test/runtests.jl:524-525
struct NormDU {
    static constexpr int D = 2, BLOB = 0, NOISE = 0;
    static constexpr const char* name = "normdu";
    __device__ static __forceinline__ double run(const double* th, const double* data, SimRng& r, double*)
    {
        return fabs((th[0] * th[0] + th[1]) * (th[0] + r.n() * 0.01) - data[0]);
    }
};

// test/runtests.jl:603 (dist1!) and :614 (dist2!, returns Inf with probability 1/2)
template <bool WITH_INF>
struct TwoD {
    static constexpr int D = 2, BLOB = 0, NOISE = 0;
    static constexpr const char* name = WITH_INF ? "twod_inf" : "twod";
    __device__ static __forceinline__ double run(const double* th, const double*, SimRng& r, double*)
    {
        double z1, z2;
        r.n2(z1, z2);
        double t1 = th[0] + z1 * 0.01 - th[1] * th[1];
        double t2 = th[1] - 1.0 + z2 * 0.01;
        double v = 50.0 * (t1 * t1) + t2 * t2;
        if (WITH_INF) { double u1, u2; r.u2(u1, u2); if (!(u1 < 0.5)) v = INFINITY; }
        return v;
    }
};

// test/runtests.jl:582-583
struct Mixture {
    static constexpr int D = 1, BLOB = 0, NOISE = 0;
    static constexpr const char* name = "mixture";
    __device__ static __forceinline__ double run(const double* th, const double* data, SimRng& r, double*)
    {
        double z1, z2, u1, u2;
        r.n2(z1, z2);
        r.u2(u1, u2);
        double noise = (u1 < 0.5) ? z1 * 0.1 : z2;
        return fabs(th[0] + noise - data[0]);
    }
};

// test/runtests.jl:537-549: 31-point drifted-Wiener RMS summary, mean absolute difference
struct Wiener {
    static constexpr int D = 2, BLOB = 0, NOISE = 0;
    static constexpr const char* name = "wiener";
    __device__ static __forceinline__ double run(const double* th, const double* data, SimRng& r, double*)
    {
        // one scalar noise factor per simulation: the reference's `@.(...) .* (0.95 + 0.1 * rand())`
        double acc = 0.0, f = 0.95 + 0.1 * r.u();
        for (int t = 0; t <= 30; ++t) {
            double tt = (double)t;
            double v = sqrt(th[0] * th[0] * tt * tt + th[1] * th[1] * tt) * f;
            acc += fabs(v - data[t]);
        }
        return acc / 31.0;
    }
};

// config 4: Lotka-Volterra, fixed-step RK4 (see DESIGN.md for the data layout).  Two competing models for the
// evidence comparison: the classical one (predators grow with the encounter rate x y) and a variant whose
// predators grow with the prey density alone (LINEAR = true, "lotka_volterra_lin").
template <bool LINEAR>
struct LotkaVolterraT {
    static constexpr int D = 4, BLOB = 0, NOISE = 0;
    static constexpr bool SPLIT = true, STEPPED = false;    // heavy simulator: proposals outside the prior's support must not idle lanes
    static constexpr const char* name = LINEAR ? "lotka_volterra_lin" : "lotka_volterra";
    __device__ static __forceinline__ void rhs(const double* th, double x, double y, double& dx, double& dy)
    {
        dx = th[0] * x - th[1] * x * y;
        dy = LINEAR ? th[2] * x - th[3] * y : th[2] * x * y - th[3] * y;
    }
    __device__ static double run(const double* th, const double* data, SimRng& r, double*)
    {
        double x = data[0], y = data[1], dt = data[2];
        int sub = (int)data[3], nobs = (int)data[4];
        double sig = data[5], acc = 0.0;
        for (int j = 0; j < nobs; ++j) {
            for (int s = 0; s < sub; ++s) {
                double k1x, k1y, k2x, k2y, k3x, k3y, k4x, k4y;
                rhs(th, x, y, k1x, k1y);
                rhs(th, x + 0.5 * dt * k1x, y + 0.5 * dt * k1y, k2x, k2y);
                rhs(th, x + 0.5 * dt * k2x, y + 0.5 * dt * k2y, k3x, k3y);
                rhs(th, x + dt * k3x, y + dt * k3y, k4x, k4y);
                x = x + dt / 6.0 * (k1x + 2.0 * k2x + 2.0 * k3x + k4x);
                y = y + dt / 6.0 * (k1y + 2.0 * k2y + 2.0 * k3y + k4y);
            }
            double z1, z2;
            r.n2(z1, z2);
            double rx = x + sig * z1 - data[6 + 2 * j];
            double ry = y + sig * z2 - data[7 + 2 * j];
            acc += rx * rx + ry * ry;
        }
        return sqrt(acc / (2.0 * nobs));
    }
};
typedef LotkaVolterraT<false> LotkaVolterra;
typedef LotkaVolterraT<true> LotkaVolterraLin;

// config 5: linear birth-death process, Gillespie SSA (divergent trajectory lengths).
// The simulator is written as begin / step / finish -- one step is one event attempt or the closing of one observation
// window, and consumes at most one Philox block -- so that the queue-driven sweep (sweep.cuh, SPLIT models) can refill a
// lane with the next pending simulation the moment its trajectory ends instead of idling until the longest trajectory of
// its warp is through.  run() is the same three functions in a loop: one definition, identical arithmetic.
struct BirthDeath {
    static constexpr int D = 2, BLOB = 16, NOISE = 0;
    static constexpr bool SPLIT = true, STEPPED = true;
    static constexpr const char* name = "birth_death";
    struct State { double n, t, acc, events, lam_mu, th0, dt, maxev; int j, nobs; };
    __device__ static __forceinline__ void begin(State& s, const double* th, const double* data)
    {
        s.n = data[0]; s.nobs = (int)data[1]; s.dt = data[2]; s.maxev = data[3];
        s.t = 0.0; s.acc = 0.0; s.events = 0.0; s.lam_mu = th[0] + th[1]; s.th0 = th[0]; s.j = 0;
    }
    // false when the trajectory is complete
    __device__ static __forceinline__ bool step(State& s, const double* data, SimRng& r)
    {
        if (s.j >= s.nobs) return false;
        const double tobs = s.dt * (double)(s.j + 1);
        bool close = true;
        if (s.n > 0.0 && s.events < s.maxev) {
            double rate = s.lam_mu * s.n, u1, u2;
            r.u2(u1, u2);
            double tn = s.t + (-plog_unit(1.0 - u1)) / rate;    // 1 - u1 in [2^-53, 1]
            if (!(tn > tobs)) {                                   // (a pending event beyond the observation time is discarded: memoryless)
                s.t = tn;
                s.n += (u2 * s.lam_mu < s.th0) ? 1.0 : -1.0;
                s.events += 1.0;
                close = false;
            }
        }
        if (close) {
            s.t = tobs;
            double dn = s.n - data[4 + s.j];
            s.acc += dn * dn;
            s.j += 1;
        }
        return s.j < s.nobs;
    }
    __device__ static __forceinline__ double finish(const State& s, double* blob)
    {
        blob[0] = s.n; blob[1] = s.events;
        return sqrt(s.acc / (double)s.nobs);
    }
    __device__ static double run(const double* th, const double* data, SimRng& r, double* blob)
    {
        State s;
        begin(s, th, data);
        while (step(s, data, r)) {}
        return finish(s, blob);
    }
    // scheduling hint for the queue-driven sweep (never affects a result): trajectories expected to be long are handed out
    // first, so that no lane starts a 5000-event trajectory when the rest of the queue is already drained.
    // E[events] ~ n0 (lambda + mu) (e^{rT} - 1) / r with r = lambda - mu, T = nobs dt
    __device__ static __forceinline__ bool heavy(const double* th, const double* data)
    {
        const double r = th[0] - th[1], T = data[1] * data[2];
        const double growth = fabs(r) > 1e-9 ? (exp(r * T) - 1.0) / r : T;
        return data[0] * (th[0] + th[1]) * growth > 0.2 * data[3];
    }
};

// test/runtests.jl:427-437 (socks): sequential picks without replacement
struct Socks {
    static constexpr int D = 2, BLOB = 0, NOISE = 0;
    static constexpr const char* name = "socks";
    __device__ static double run(const double* th, const double* data, SimRng& r, double*)
    {
        double n_socks = th[0], prop = th[1];
        double n_pairs = rint(prop * floor(n_socks / 2.0));
        double n_odd = n_socks - 2.0 * n_pairs;
        double n_pick = n_socks < 11.0 ? n_socks : 11.0;
        double p2 = n_pairs, p1 = 0.0, o = n_odd, got_pairs = 0.0;
        for (int t = 0; t < (int)n_pick; ++t) {
            double tot = 2.0 * p2 + p1 + o;
            double u = r.u() * tot;
            if (u < 2.0 * p2) { p2 -= 1.0; p1 += 1.0; }
            else if (u < 2.0 * p2 + p1) { p1 -= 1.0; got_pairs += 1.0; }
            else { o -= 1.0; }
        }
        double sample_odds = n_pick - 2.0 * got_pairs;
        return fabs(got_pairs - data[0]) + fabs(sample_odds - data[1]);
    }
};

// model ids: keep in sync with MODEL_TABLE in registry.cu (and, independently, the oracle)
enum {
    M_GAUSS1D = 0, M_GAUSS1D_BLOB = 1, M_GAUSS_CORR10 = 2, M_DIRAC = 3, M_NORMDU = 4, M_TWOD = 5,
    M_TWOD_INF = 6, M_MIXTURE = 7, M_WIENER = 8, M_LOTKA_VOLTERRA = 9, M_BIRTH_DEATH = 10, M_GK = 11,
    M_SOCKS = 12, M_GK_F32 = 13, M_LOTKA_VOLTERRA_LIN = 14, M_COUNT = 15
};

}  // namespace abcdez
