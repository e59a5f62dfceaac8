/*
 * abcdez_oracle.c -- CPU restatement of the ABCdeZ.jl particle hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (abcdez.jl_b200/, csrc/,
 * include/) may import, link or call this file.  It is used by tests/, by
 * __graft_entry__.smoke() and by the cpu_baseline / --impl reference legs of
 * bench.py as the checker and the CPU baseline.
 *
 * What it restates (paths relative to /root/reference, ABCdeZ.jl v0.6.0):
 *   src/abcdez_priors.jl:27-61      Factored pdf/logpdf/rand/length
 *   src/abcdez_types.jl:3-23        Particle algebra (op) and push_p
 *   src/abcdez_types.jl:26-73       the four ABC kernels
 *   src/abcdez_init.jl:2-22         abcde_init!
 *   src/abcdez_smc.jl:8             get_ess
 *   src/abcdez_smc.jl:15-56         wsample_stratified!
 *   src/abcdez_smc.jl:59-83         abcdesmc_update_ws!
 *   src/abcdez_smc.jl:85-104        abcdesmc_resample!
 *   src/abcdez_smc.jl:106-153       abcdesmc_swarm!
 *   src/abcdez_smc.jl:215-394       abcdesmc!
 *   src/abcdez_mc.jl:5-61           abcdemc_swarm!
 *   src/abcdez_mc.jl:102-172        abcdemc!
 *
 * Third-party arithmetic that is NOT in /root/reference (Project.toml:12-16;
 * no Manifest, so versions are unpinned) is restated from its published
 * algorithm at the reference's call sites:
 *   Statistics.quantile (type 7)          at src/abcdez_smc.jl:301
 *   StatsBase.wsample(rng, 1:N, alive)    at src/abcdez_smc.jl:121,125
 *   Distributions rand(Uniform(a,b))      at src/abcdez_smc.jl:47
 *   Distributions logpdf/rand of marginals at src/abcdez_priors.jl:41-54
 * PARITY PINNING: the reference's own tests pin Factored pdf/logpdf, push_p,
 * the four kernels and the analytic evidences (test/runtests.jl:21-121,176,
 * 336); tests/test_oracle_golden.py checks this file against all of them.
 * quantile / wsample / stratified indices / accept decisions are NOT pinned
 * by any reference test ("parity unpinned" at those third-party boundaries);
 * they are cross-checked against numpy/scipy instead (tests/).
 *
 * Randomness.  The reference consumes a variable number of draws from a
 * task-local Xoshiro stream, so "same seed" parity is impossible.  Every
 * random decision here is therefore taken from a *provider*: either arrays
 * injected by the caller (a, b, z, u per particle; u per stratum; s for
 * abcdemc!) or the counter-based Philox4x32-10 contract documented in
 * DESIGN.md ("Randomness contract"), which the CUDA library implements
 * independently.  Philox is pinned by the Random123 known-answer vectors.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXD 16
#define ORC_MAXDATA 64
#define ORC_MAXBLOB 64

/* ------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. 2011; Random123 KAT pinned in tests)   */
/* ------------------------------------------------------------------ */
static inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    philox4x32_10(ctr, key, out);
}

/* ------------------------------------------------------------------ */
/* Portable math contract (DESIGN.md "Numerics"): log, exp, lgamma and  */
/* sin/cos(2 pi u) defined operation by operation in IEEE-754 binary64  */
/* with +,-,*,/ and explicit fma() in the Horner steps (no implicit contraction).  The differential-evolution  */
/* move amplifies a 1-ulp perturbation of theta by ~2.6x per accepted   */
/* move, so two implementations whose libm differ in the last bit       */
/* diverge within ~10 SMC iterations; with these definitions the CUDA   */
/* library and this oracle produce bit-identical runs.  Accuracy is     */
/* <= 2 ulp (checked against glibc in tests/test_oracle_golden.py).     */
/* ------------------------------------------------------------------ */
#define PM_LN2      0x1.62e42fefa39efp-1
#define PM_LN2_HI   0x1.62e42fee00000p-1
#define PM_LN2_LO   0x1.a39ef35793c76p-33
#define PM_INV_LN2  0x1.71547652b82fep+0
#define PM_SQRT2    0x1.6a09e667f3bcdp+0
#define PM_TWO_PI   0x1.921fb54442d18p+2
#define PM_HALF_LOG_2PI 0x1.d67f1c864beb5p-1

static inline double plog(double x)
{
    if (x != x || x < 0.0) return NAN;
    if (x == 0.0) return -INFINITY;
    if (x == INFINITY) return x;
    uint64_t b; memcpy(&b, &x, 8);
    int e = (int)((b >> 52) & 0x7ff);
    if (e == 0) { x = x * 0x1p54; memcpy(&b, &x, 8); e = (int)((b >> 52) & 0x7ff) - 54; }
    e -= 1023;
    b = (b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m; memcpy(&m, &b, 8);
    if (m > PM_SQRT2) { m = m * 0.5; e += 1; }
    double f = m - 1.0;
    double s = f / (2.0 + f);
    double z = s * s;
    /* log(m) = 2 atanh(s) = 2s + s*z*(2/3 + z*(2/5 + ... + z*2/23)) */
    double p = 2.0 / 23.0;
    p = fma(p, z, 2.0 / 21.0);
    p = fma(p, z, 2.0 / 19.0);
    p = fma(p, z, 2.0 / 17.0);
    p = fma(p, z, 2.0 / 15.0);
    p = fma(p, z, 2.0 / 13.0);
    p = fma(p, z, 2.0 / 11.0);
    p = fma(p, z, 2.0 / 9.0);
    p = fma(p, z, 2.0 / 7.0);
    p = fma(p, z, 2.0 / 5.0);
    p = fma(p, z, 2.0 / 3.0);
    double r = (s * z) * p;
    double lm = 2.0 * s + r;
    return ((double)e * PM_LN2_HI + lm) + (double)e * PM_LN2_LO;
}

static inline double pexp(double x)
{
    if (x != x) return x;
    if (x > 709.78) return INFINITY;
    if (x < -745.2) return 0.0;
    double k = floor(x * PM_INV_LN2 + 0.5);
    double r = (x - k * PM_LN2_HI) - k * PM_LN2_LO;
    double p = 1.0 / 6227020800.0;            /* 1/13! */
    p = fma(p, r, 1.0 / 479001600.0);
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    /* scale by 2^k in two exact steps (k in [-1075, 1024]) */
    int ki = (int)k, k1 = ki / 2, k2 = ki - k1;
    uint64_t b1 = (uint64_t)(k1 + 1023) << 52, b2 = (uint64_t)(k2 + 1023) << 52;
    double s1, s2; memcpy(&s1, &b1, 8); memcpy(&s2, &b2, 8);
    return (p * s1) * s2;
}

/* sin and cos of 2*pi*u for u in [0, 1] */
static inline void psincos2pi(double u, double* sn, double* cs)
{
    double q = floor(4.0 * u + 0.5);
    double r = u - 0.25 * q;                   /* exact, |r| <= 1/8 */
    double x = r * PM_TWO_PI;
    double x2 = x * x;
    double ps = -1.0 / 355687428096000.0;      /* -1/17! */
    ps = fma(ps, x2, 1.0 / 1307674368000.0);
    ps = fma(ps, x2, -(1.0 / 6227020800.0));
    ps = fma(ps, x2, 1.0 / 39916800.0);
    ps = fma(ps, x2, -(1.0 / 362880.0));
    ps = fma(ps, x2, 1.0 / 5040.0);
    ps = fma(ps, x2, -(1.0 / 120.0));
    ps = fma(ps, x2, 1.0 / 6.0);
    double s = x - x * (x2 * ps);
    double pc = 1.0 / 6402373705728000.0;      /* 1/18! */
    pc = fma(pc, x2, -(1.0 / 20922789888000.0));
    pc = fma(pc, x2, 1.0 / 87178291200.0);
    pc = fma(pc, x2, -(1.0 / 479001600.0));
    pc = fma(pc, x2, 1.0 / 3628800.0);
    pc = fma(pc, x2, -(1.0 / 40320.0));
    pc = fma(pc, x2, 1.0 / 720.0);
    pc = fma(pc, x2, -(1.0 / 24.0));
    pc = fma(pc, x2, 0.5);
    double c = 1.0 - x2 * pc;
    int k = (int)q & 3;
    if (k == 0) { *sn = s; *cs = c; }
    else if (k == 1) { *sn = c; *cs = -s; }
    else if (k == 2) { *sn = -s; *cs = -c; }
    else { *sn = -c; *cs = s; }
}

/* log Gamma(z), z > 0: upward recurrence to z >= 10, then Stirling's series */
static inline double plgamma(double z)
{
    if (!(z > 0.0)) return (z == 0.0) ? INFINITY : NAN;
    if (z == INFINITY) return z;
    double prod = 1.0;
    while (z < 10.0) { prod = prod * z; z = z + 1.0; }
    double zi = 1.0 / z, z2 = zi * zi;
    double t = -691.0 / 360360.0;
    t = fma(t, z2, 1.0 / 1188.0);
    t = fma(t, z2, -(1.0 / 1680.0));
    t = fma(t, z2, 1.0 / 1260.0);
    t = fma(t, z2, -(1.0 / 360.0));
    t = fma(t, z2, 1.0 / 12.0);
    double st = ((z - 0.5) * plog(z) - z) + PM_HALF_LOG_2PI + t * zi;
    return st - plog(prod);
}

double orc_plog(double x) { return plog(x); }
double orc_pexp(double x) { return pexp(x); }
double orc_plgamma(double x) { return plgamma(x); }
void orc_psincos2pi(double u, double* s, double* c) { psincos2pi(u, s, c); }

/* Checker for an evaluation shortcut of the CUDA library (csrc/common.cuh pdiv_r): the z-score of a Normal
 * marginal, (x - mu) / sigma, is evaluated there through the host-rounded reciprocal rb = RN(1/sigma) as
 * q = RN(a*rb); r = fma(-sigma, q, a); z = RN(q + r*rb)  (Markstein's correction step; correctly rounded when
 * rb is).  The oracle itself keeps the plain division; this routine counts, over n pseudo-random numerators
 * (xorshift64*, magnitudes spread over 2^-40..2^40, including values just below powers of two), how often
 * the shortcut differs from a / sigma.  Expected: 0. */
int64_t orc_pdiv_mismatches(int64_t n, uint64_t seed, double sigma)
{
    const double rb = 1.0 / sigma;
    uint64_t st = seed ? seed : 0x9E3779B97F4A7C15ull;
    int64_t bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        st ^= st >> 12; st ^= st << 25; st ^= st >> 27;
        uint64_t r = st * 0x2545F4914F6CDD1Dull;
        uint64_t mant = r & 0x000fffffffffffffull;
        if ((i & 7) == 7) mant |= 0x000ffffffffff000ull;            /* significand close to 2 */
        if ((i & 15) == 8) mant &= 0x0000000000000fffull;           /* significand close to 1 */
        uint64_t ex = 1023u - 40u + ((r >> 52) % 81u);
        uint64_t bits = (r & 0x8000000000000000ull) | (ex << 52) | mant;
        double a; memcpy(&a, &bits, 8);
        double q = a * rb;
        double rr = fma(-sigma, q, a);
        double z = fma(rr, rb, q);
        if (z != a / sigma) bad++;
    }
    return bad;
}

/* Randomness contract: stream tags (DESIGN.md "Randomness contract"). */
enum {
    TAG_PRIOR = 1,      /* prior draws; c3 = (dim<<16) | block              */
    TAG_PARTNER = 2,    /* DE partner attempts; c3 = attempt; u1->a, u2->b   */
    TAG_MOVE = 3,       /* c3=0: (u1,u2)->z Box-Muller; c3=1: u1->accept u   */
    TAG_MODEL = 4,      /* simulator noise in sweeps; c3 = block             */
    TAG_RESAMPLE = 5,   /* c0 = stratum; u1                                  */
    TAG_MC = 6,         /* abcdemc!: c3=0: u1->s ; c3=1: u1->prior MH u      */
    TAG_INIT_MODEL = 7  /* simulator noise in abcde_init!                    */
};

typedef struct {
    uint32_t key[2];
    uint32_t c0, c1, c2;
} stream_t;

static inline stream_t mk_stream(uint64_t seed, uint32_t particle, uint32_t epoch, uint32_t tag)
{
    stream_t s;
    s.key[0] = (uint32_t)seed; s.key[1] = (uint32_t)(seed >> 32);
    s.c0 = particle; s.c1 = epoch; s.c2 = tag;
    return s;
}

/* one Philox block -> two uniforms in [0,1) with 53 random bits each */
static inline void stream_u2(const stream_t* s, uint32_t block, double* u1, double* u2)
{
    uint32_t ctr[4] = { s->c0, s->c1, s->c2, block }, o[4];
    philox4x32_10(ctr, s->key, o);
    uint64_t a = ((uint64_t)o[1] << 32) | o[0];
    uint64_t b = ((uint64_t)o[3] << 32) | o[2];
    *u1 = (double)(a >> 11) * 0x1.0p-53;
    *u2 = (double)(b >> 11) * 0x1.0p-53;
}


/* Box-Muller pair from one block: z1 = r cos(2 pi u2), z2 = r sin(2 pi u2),
 * r = sqrt(-2 plog(1-u1)).  (1-u1) is in (0,1] and exactly representable. */
static inline void stream_n2(const stream_t* s, uint32_t block, double* z1, double* z2)
{
    double u1, u2;
    stream_u2(s, block, &u1, &u2);
    double r = sqrt(-2.0 * plog(1.0 - u1)), sn, cs;
    psincos2pi(u2, &sn, &cs);
    *z1 = r * cs;
    *z2 = r * sn;
}

/* sequential RNG view handed to the simulators */
typedef struct {
    stream_t s;
    uint32_t blk;
} simrng_t;

static inline void sim_u2(simrng_t* r, double* u1, double* u2) { stream_u2(&r->s, r->blk++, u1, u2); }
static inline void sim_n2(simrng_t* r, double* z1, double* z2) { stream_n2(&r->s, r->blk++, z1, z2); }
static inline double sim_u(simrng_t* r) { double a, b; sim_u2(r, &a, &b); return a; }
static inline double sim_n(simrng_t* r) { double a, b; sim_n2(r, &a, &b); return a; }

/* ------------------------------------------------------------------ */
/* Priors: Factored of univariate marginals                            */
/* src/abcdez_priors.jl:18-61, Distributions.jl marginals              */
/* ------------------------------------------------------------------ */
enum {
    FAM_NORMAL = 0,           /* p0=mu, p1=sigma                    */
    FAM_UNIFORM = 1,          /* p0=a, p1=b                         */
    FAM_DISCRETE_UNIFORM = 2, /* p0=a, p1=b (integers)              */
    FAM_LOGNORMAL = 3,        /* p0=mu, p1=sigma                    */
    FAM_EXPONENTIAL = 4,      /* p0=scale theta                     */
    FAM_GAMMA = 5,            /* p0=shape alpha, p1=scale theta     */
    FAM_BETA = 6,             /* p0=alpha, p1=beta                  */
    FAM_NEGBIN = 7            /* p0=r, p1=p                         */
};

typedef struct {
    int d;
    int family[ORC_MAXD];
    double p[ORC_MAXD][4];
} prior_t;

static int fam_is_discrete(int f) { return f == FAM_DISCRETE_UNIFORM || f == FAM_NEGBIN; }

/* push_p, src/abcdez_types.jl:20-23: continuous -> float(p); discrete ->
 * round(Int, p), Julia's round is ties-to-even == rint() in the default
 * rounding mode. */
static inline void push_p(const prior_t* pr, const double* th, double* out)
{
    for (int k = 0; k < pr->d; ++k)
        out[k] = fam_is_discrete(pr->family[k]) ? rint(th[k]) : th[k];
}

#define ORC_LOG2PI 1.8378770664093454835606594728112

/* additive constant of the log density, evaluated once with the host libm (the CUDA library
 * evaluates the same expressions with the same libm on the host side) */
static double marginal_const(int fam, const double* p)
{
    switch (fam) {
    case FAM_NORMAL: return log(p[1]);
    case FAM_UNIFORM: return -log(p[1] - p[0]);
    case FAM_DISCRETE_UNIFORM: return log(1.0 / (p[1] - p[0] + 1.0));
    case FAM_LOGNORMAL: return log(p[1]);
    case FAM_EXPONENTIAL: return log(p[0]);
    case FAM_GAMMA: return lgamma(p[0]) + p[0] * log(p[1]);
    case FAM_BETA: return lgamma(p[0]) + lgamma(p[1]) - lgamma(p[0] + p[1]);
    case FAM_NEGBIN: return p[0] * log(p[1]) - lgamma(p[0]);
    }
    return NAN;
}

static double marginal_logpdf(int fam, const double* p, double x)
{
    double c = marginal_const(fam, p);
    switch (fam) {
    case FAM_NORMAL: {            /* Distributions normlogpdf: -(z^2+log2pi)/2 - log(sigma) */
        double z = (x - p[0]) / p[1];
        return -(z * z + ORC_LOG2PI) / 2.0 - c;
    }
    case FAM_UNIFORM:             /* insupport [a,b] ? -log(b-a) : -Inf */
        return (x >= p[0] && x <= p[1]) ? c : -INFINITY;
    case FAM_DISCRETE_UNIFORM:    /* log(1/(b-a+1)) on integers a..b */
        return (x >= p[0] && x <= p[1] && x == rint(x)) ? c : -INFINITY;
    case FAM_LOGNORMAL: {
        if (!(x > 0.0)) return -INFINITY;
        double lx = plog(x);
        double z = (lx - p[0]) / p[1];
        return -(z * z + ORC_LOG2PI) / 2.0 - c - lx;
    }
    case FAM_EXPONENTIAL:         /* scale parametrisation */
        return (x >= 0.0) ? -x / p[0] - c : -INFINITY;
    case FAM_GAMMA: {
        if (!(x >= 0.0)) return -INFINITY;
        if (x == 0.0) return p[0] == 1.0 ? -plog(p[1]) : (p[0] < 1.0 ? INFINITY : -INFINITY);
        return (p[0] - 1.0) * plog(x) - x / p[1] - c;
    }
    case FAM_BETA: {
        if (!(x >= 0.0 && x <= 1.0)) return -INFINITY;
        double t1 = (p[0] == 1.0) ? 0.0 : (p[0] - 1.0) * plog(x);
        double t2 = (p[1] == 1.0) ? 0.0 : (p[1] - 1.0) * plog(1.0 - x);
        return t1 + t2 - c;
    }
    case FAM_NEGBIN: {            /* k failures before r successes, success prob p */
        if (!(x >= 0.0) || x != rint(x)) return -INFINITY;
        return plgamma(x + p[0]) - plgamma(x + 1.0) + c + x * plog(1.0 - p[1]);
    }
    }
    return NAN;
}

/* logpdf(::Factored, x), src/abcdez_priors.jl:40-46: left-to-right sum */
static double prior_logpdf(const prior_t* pr, const double* xpushed)
{
    double s = marginal_logpdf(pr->family[0], pr->p[0], xpushed[0]);
    for (int k = 1; k < pr->d; ++k)
        s += marginal_logpdf(pr->family[k], pr->p[k], xpushed[k]);
    return s;
}

/* Marsaglia-Tsang gamma(shape a>=1... generalised), unit scale; sequential
 * blocks from the dim's sub-stream.  Not reference arithmetic (Distributions
 * uses its own samplers; the reference's tests only constrain the
 * distribution), shared with the CUDA library through DESIGN.md. */
static double gamma_draw(const stream_t* s, uint32_t dimbase, uint32_t* blk, double a)
{
    double boost = 1.0;
    if (a < 1.0) {
        double u1, u2;
        stream_u2(s, dimbase | (*blk)++, &u1, &u2);
        boost = pexp(plog(1.0 - u1) / a);
        a += 1.0;
    }
    double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (int it = 0; it < 1000; ++it) {
        double z, z2, u1, u2;
        stream_n2(s, dimbase | (*blk)++, &z, &z2);
        double v = 1.0 + c * z;
        if (v <= 0.0) continue;
        v = v * v * v;
        stream_u2(s, dimbase | (*blk)++, &u1, &u2);
        if (plog(1.0 - u1) < 0.5 * z * z + d - d * v + d * plog(v)) return boost * d * v;
    }
    return boost * d;
}

/* rand(rng, ::Factored), src/abcdez_priors.jl:53-54 + op(float, .),
 * src/abcdez_smc.jl:242.  Draw k uses the sub-stream c3=(k<<16)|block. */
static void prior_sample(const prior_t* pr, uint64_t seed, uint32_t particle, uint32_t epoch, double* out)
{
    stream_t s = mk_stream(seed, particle, epoch, TAG_PRIOR);
    for (int k = 0; k < pr->d; ++k) {
        const double* p = pr->p[k];
        uint32_t base = (uint32_t)k << 16, blk = 0;
        double u1, u2, z1, z2;
        switch (pr->family[k]) {
        case FAM_NORMAL:
            stream_n2(&s, base, &z1, &z2); out[k] = p[0] + p[1] * z1; break;
        case FAM_UNIFORM:
            stream_u2(&s, base, &u1, &u2); out[k] = p[0] + (p[1] - p[0]) * u1; break;
        case FAM_DISCRETE_UNIFORM:
            stream_u2(&s, base, &u1, &u2); out[k] = p[0] + floor(u1 * (p[1] - p[0] + 1.0)); break;
        case FAM_LOGNORMAL:
            stream_n2(&s, base, &z1, &z2); out[k] = pexp(p[0] + p[1] * z1); break;
        case FAM_EXPONENTIAL:
            stream_u2(&s, base, &u1, &u2); out[k] = -p[0] * plog(1.0 - u1); break;
        case FAM_GAMMA:
            out[k] = p[1] * gamma_draw(&s, base, &blk, p[0]); break;
        case FAM_BETA: {
            double g1 = gamma_draw(&s, base, &blk, p[0]);
            double g2 = gamma_draw(&s, base, &blk, p[1]);
            out[k] = g1 / (g1 + g2); break;
        }
        case FAM_NEGBIN: {   /* gamma-Poisson mixture: lambda ~ Gamma(r, (1-p)/p) */
            double lam = gamma_draw(&s, base, &blk, p[0]) * (1.0 - p[1]) / p[1];
            /* Poisson by sequential inversion on the exponential clock */
            double acc = 0.0; long cnt = -1;
            do {
                stream_u2(&s, base | blk++, &u1, &u2);
                acc += -plog(1.0 - u1); cnt++;
            } while (acc <= lam && cnt < 100000);
            out[k] = (double)cnt; break;
        }
        default: out[k] = NAN;
        }
    }
}

/* ------------------------------------------------------------------ */
/* ABC kernels, src/abcdez_types.jl:26-73                              */
/* ------------------------------------------------------------------ */
enum { K_INDICATOR = 0, K_INDICATOR_STRICT = 1, K_EPA = 2, K_EPA_STRICT = 3 };

static inline int abck_insupport(int kind, double eps, double x)
{
    if (kind == K_INDICATOR || kind == K_EPA) return (0.0 <= x && x <= eps);   /* :34,:59 */
    return (0.0 <= x && x < eps);                                              /* :46,:71 */
}

double orc_kernel_pdf(int kind, double eps, double x)
{
    if (!abck_insupport(kind, eps, x)) return 0.0;
    if (kind == K_INDICATOR || kind == K_INDICATOR_STRICT) return 1.0;
    double q = x / eps;
    return 1.0 - q * q;                                                        /* :60,:72 */
}

double orc_kernel_logpdf(int kind, double eps, double x)
{
    if (!abck_insupport(kind, eps, x)) return -INFINITY;
    if (kind == K_INDICATOR || kind == K_INDICATOR_STRICT) return 0.0;
    double q = x / eps;
    return plog(1.0 - q * q);                                                  /* :61,:73 */
}

/* ------------------------------------------------------------------ */
/* Simulators (the dist!(theta, ve) -> (d, blob) plugins).             */
/* Definitions are normative in DESIGN.md "Models"; the CUDA functors  */
/* are written independently against the same text.                    */
/* ------------------------------------------------------------------ */
enum {
    M_GAUSS1D = 0, M_GAUSS1D_BLOB = 1, M_GAUSS_CORR10 = 2, M_DIRAC = 3, M_NORMDU = 4,
    M_TWOD = 5, M_TWOD_INF = 6, M_MIXTURE = 7, M_WIENER = 8, M_LOTKA_VOLTERRA = 9,
    M_BIRTH_DEATH = 10, M_GK = 11, M_SOCKS = 12, M_GK_F32 = 13, M_LOTKA_VOLTERRA_LIN = 14, M_COUNT
};

typedef struct { const char* name; int d; int blob; } model_info_t;
static const model_info_t MODELS[M_COUNT] = {
    { "gauss1d", 1, 0 }, { "gauss1d_blob", 1, 8 }, { "gauss_corr10", 10, 0 }, { "dirac", 1, 0 },
    { "normdu", 2, 0 }, { "twod", 2, 0 }, { "twod_inf", 2, 0 }, { "mixture", 1, 0 },
    { "wiener", 2, 0 }, { "lotka_volterra", 4, 0 }, { "birth_death", 2, 16 }, { "gk", 4, 0 },
    { "socks", 2, 0 }, { "gk_f32", 4, 0 }, { "lotka_volterra_lin", 4, 0 }
};

int orc_model_count(void) { return M_COUNT; }
const char* orc_model_name(int id) { return (id >= 0 && id < M_COUNT) ? MODELS[id].name : NULL; }
int orc_model_dim(int id) { return (id >= 0 && id < M_COUNT) ? MODELS[id].d : -1; }
int orc_model_blob(int id) { return (id >= 0 && id < M_COUNT) ? MODELS[id].blob : -1; }

/* --- Lotka-Volterra, fixed-step RK4 (config 4) --------------------- */
/* theta = (a, b, c, e): x' = a x - b x y ; y' = e b x y - c y... see DESIGN.md:
 *   dx/dt = th0*x - th1*x*y ;  dy/dt = th2*x*y - th3*y
 * data[0]=x0, data[1]=y0, data[2]=dt, data[3]=steps per obs, data[4]=nobs (<=14),
 * data[5]=obs noise sigma, data[6+2j], data[7+2j] = observed (x,y) at obs j.
 * distance = sqrt(mean squared residual) over 2*nobs values; observation
 * noise sigma*N(0,1) is added to each simulated observation. */
/* the competing model of the evidence comparison ("lotka_volterra_lin"): predators grow with the prey density alone,
 *   dy/dt = th2*x - th3*y */
static int lv_linear = 0;
#pragma omp threadprivate(lv_linear)
static inline void lv_rhs(const double* th, double x, double y, double* dx, double* dy)
{
    *dx = th[0] * x - th[1] * x * y;
    *dy = lv_linear ? th[2] * x - th[3] * y : th[2] * x * y - th[3] * y;
}

static double model_lv(const double* th, const double* data, simrng_t* r)
{
    double x = data[0], y = data[1], dt = data[2];
    int sub = (int)data[3], nobs = (int)data[4];
    double sig = data[5], acc = 0.0;
    for (int j = 0; j < nobs; ++j) {
        for (int s = 0; s < sub; ++s) {
            double k1x, k1y, k2x, k2y, k3x, k3y, k4x, k4y;
            lv_rhs(th, x, y, &k1x, &k1y);
            lv_rhs(th, x + 0.5 * dt * k1x, y + 0.5 * dt * k1y, &k2x, &k2y);
            lv_rhs(th, x + 0.5 * dt * k2x, y + 0.5 * dt * k2y, &k3x, &k3y);
            lv_rhs(th, x + dt * k3x, y + dt * k3y, &k4x, &k4y);
            x = x + dt / 6.0 * (k1x + 2.0 * k2x + 2.0 * k3x + k4x);
            y = y + dt / 6.0 * (k1y + 2.0 * k2y + 2.0 * k3y + k4y);
        }
        double z1, z2;
        sim_n2(r, &z1, &z2);
        double rx = x + sig * z1 - data[6 + 2 * j];
        double ry = y + sig * z2 - data[7 + 2 * j];
        acc += rx * rx + ry * ry;
    }
    return sqrt(acc / (2.0 * nobs));
}

/* --- linear birth-death Gillespie SSA (config 5) ------------------- */
/* theta = (lambda, mu).  data[0]=n0, data[1]=nobs (<=16), data[2]=dt between
 * observations, data[3]=max events, data[4+j]=observed population at obs j.
 * distance = sqrt(mean squared diff of populations).  blob = (final pop,
 * event count) as two doubles. */
static double model_bd(const double* th, const double* data, simrng_t* r, double* blob)
{
    double n = data[0];
    int nobs = (int)data[1];
    double dt = data[2], maxev = data[3];
    double t = 0.0, acc = 0.0, events = 0.0;
    double lam_mu = th[0] + th[1];
    for (int j = 0; j < nobs; ++j) {
        double tobs = dt * (double)(j + 1);
        while (n > 0.0 && events < maxev) {
            double rate = lam_mu * n, u1, u2;
            sim_u2(r, &u1, &u2);
            double tn = t + (-plog(1.0 - u1)) / rate;
            if (tn > tobs) break;            /* memoryless: pending event discarded at tobs */
            t = tn;
            n += (u2 * lam_mu < th[0]) ? 1.0 : -1.0;
            events += 1.0;
        }
        t = tobs;
        double dn = n - data[4 + j];
        acc += dn * dn;
    }
    if (blob) { blob[0] = n; blob[1] = events; }
    return sqrt(acc / (double)nobs);
}

/* --- g-and-k (config 3) ------------------------------------------- */
/* theta = (A, B, g, k), c = 0.8.  data[0] = number of draws n (<= 16384), data[1..7] = observed octiles.
 * Draw n standard normals (Philox block b of the simulator's stream -> one Box-Muller pair: draws 2b, 2b+1),
 * map each through the g-and-k quantile function
 *     x = A + B (1 + 0.8 (1 - e^{-g z}) / (1 + e^{-g z})) (1 + z^2)^k z
 * evaluated as written below (FP64, the portable exp / log of this file: the arithmetic of a Julia Float64
 * dist!), take the 7 octiles as order statistics x_(ceil(n*j/8)) (1-based), distance = sqrt(mean squared diff).
 * The CUDA functor performs the same operations in the same order: bit-identical distances. */
static int cmp_dbl(const void* a, const void* b)
{
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

static double model_gk(const double* th, const double* data, simrng_t* r)
{
    int n = (int)data[0];
    n = n < 8 ? 8 : (n > 16384 ? 16384 : n);
    double* x = (double*)malloc(sizeof(double) * (size_t)(n + 2));
    const double A = th[0], B = th[1], g = th[2], k = th[3];
    for (int i = 0; i < n; i += 2) {
        double z[2];
        stream_n2(&r->s, (uint32_t)(i / 2), &z[0], &z[1]);
        for (int j = 0; j < 2; ++j) {
            const double zz = z[j];
            const double gz = g * zz;
            const double e = pexp(-fabs(gz));
            double t = (1.0 - e) / (1.0 + e);
            t = gz < 0.0 ? -t : t;
            const double c = 1.0 + 0.8 * t;
            const double l = plog(1.0 + zz * zz);
            const double p = pexp(k * l);
            x[i + j] = A + ((B * c) * p) * zz;
        }
    }
    qsort(x, (size_t)n, sizeof(double), cmp_dbl);
    double acc = 0.0;
    for (int j = 1; j <= 7; ++j) {
        int idx = (n * j + 7) / 8; /* ceil(n*j/8), 1-based */
        double dq = x[idx - 1] - data[j];
        acc += dq * dq;
    }
    free(x);
    return sqrt(acc / 7.0);
}

/* --- g-and-k, relaxed-precision mode "gk_f32"  -------------------------------- */
/* The same definition with the draws in FP32: Philox block b -> four 24-bit uniforms -> two Box-Muller pairs
 * (draws 4b .. 4b+3), portable FP32 log / exp / sin / cos below (IEEE single operations and fmaf only;
 * -ffp-contract=off), octiles of the FP32 draws, distance in FP64.  Restated operation by operation in
 * abcdez.jl_b200/csrc/gk.cu. */
static int cmp_float(const void* a, const void* b)
{
    float x = *(const float*)a, y = *(const float*)b;
    return (x > y) - (x < y);
}

static inline void philox_f4(const stream_t* s, uint32_t block, float u[4])
{
    uint32_t ctr[4] = { s->c0, s->c1, s->c2, block }, o[4];
    philox4x32_10(ctr, s->key, o);
    for (int i = 0; i < 4; ++i) u[i] = (float)(o[i] >> 8) * 0x1.0p-24f;
}

static inline float gk_logf_pos(float x)
{
    uint32_t b; memcpy(&b, &x, 4);
    int e = (int)((b >> 23) & 0xffu) - 127;
    uint32_t mb = (b & 0x007fffffu) | 0x3f800000u;
    float m; memcpy(&m, &mb, 4);
    if (m > 0x1.6a09e6p+0f) { m = m * 0.5f; e += 1; }
    const float f = m - 1.0f;
    const float s = f / (2.0f + f);
    const float z = s * s;
    float p = 2.0f / 9.0f;
    p = fmaf(p, z, 2.0f / 7.0f);
    p = fmaf(p, z, 2.0f / 5.0f);
    p = fmaf(p, z, 2.0f / 3.0f);
    const float r = (s * z) * p;
    const float lm = 2.0f * s + r;
    return ((float)e * 0x1.62e4p-1f + lm) + (float)e * 0x1.7f7d1cp-20f;
}

static inline float gk_expf(float x)
{
    if (x > 88.0f) return INFINITY;
    if (x < -87.0f) return 0.0f;
    const float k = floorf(x * 0x1.715476p+0f + 0.5f);
    const float r = (x - k * 0x1.62e4p-1f) - k * 0x1.7f7d1cp-20f;
    float p = 1.0f / 5040.0f;
    p = fmaf(p, r, 1.0f / 720.0f);
    p = fmaf(p, r, 1.0f / 120.0f);
    p = fmaf(p, r, 1.0f / 24.0f);
    p = fmaf(p, r, 1.0f / 6.0f);
    p = fmaf(p, r, 0.5f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    uint32_t sb = (uint32_t)((int)k + 127) << 23;
    float sc; memcpy(&sc, &sb, 4);
    return p * sc;
}

static inline void gk_sincos2pif(float u, float* sn, float* cs)
{
    const float q = floorf(4.0f * u + 0.5f);
    const float r = u - 0.25f * q;
    const float x = r * 0x1.921fb6p+2f;
    const float x2 = x * x;
    float ps = -1.0f / 39916800.0f;
    ps = fmaf(ps, x2, 1.0f / 362880.0f);
    ps = fmaf(ps, x2, -1.0f / 5040.0f);
    ps = fmaf(ps, x2, 1.0f / 120.0f);
    ps = fmaf(ps, x2, -1.0f / 6.0f);
    float pc = -1.0f / 3628800.0f;
    pc = fmaf(pc, x2, 1.0f / 40320.0f);
    pc = fmaf(pc, x2, -1.0f / 720.0f);
    pc = fmaf(pc, x2, 1.0f / 24.0f);
    pc = fmaf(pc, x2, -0.5f);
    const float s = x + x * (x2 * ps);
    const float c = 1.0f + x2 * pc;
    const int k = (int)q & 3;
    const int swap = (k & 1) != 0;
    const float a = swap ? c : s, b = swap ? s : c;
    *sn = (k & 2) ? -a : a;
    *cs = ((k + 1) & 2) ? -b : b;
}

static double model_gk_f32(const double* th, const double* data, simrng_t* r)
{
    int n = (int)data[0];
    n = n < 8 ? 8 : (n > 16384 ? 16384 : n);
    float* x = (float*)malloc(sizeof(float) * (size_t)(n + 4));
    const float A = (float)th[0], B = (float)th[1], g = (float)th[2], k = (float)th[3];
    for (int i = 0; i < n; i += 4) {
        float u[4], s1, c1, s2, c2;
        philox_f4(&r->s, (uint32_t)(i / 4), u);
        const float r1 = sqrtf(-2.0f * gk_logf_pos(1.0f - u[0])), r2 = sqrtf(-2.0f * gk_logf_pos(1.0f - u[2]));
        gk_sincos2pif(u[1], &s1, &c1);
        gk_sincos2pif(u[3], &s2, &c2);
        const float z[4] = { r1 * c1, r1 * s1, r2 * c2, r2 * s2 };
        for (int j = 0; j < 4; ++j) {
            const float zz = z[j];
            const float gz = g * zz;
            const float e = gk_expf(-fabsf(gz));
            float t = (1.0f - e) / (1.0f + e);
            t = gz < 0.0f ? -t : t;
            const float c = 1.0f + 0.8f * t;
            const float l = gk_logf_pos(1.0f + zz * zz);
            const float p = gk_expf(k * l);
            x[i + j] = A + ((B * c) * p) * zz;
        }
    }
    qsort(x, (size_t)n, sizeof(float), cmp_float);
    double acc = 0.0;
    for (int j = 1; j <= 7; ++j) {
        int idx = (n * j + 7) / 8;
        double dq = (double)x[idx - 1] - data[j];
        acc += dq * dq;
    }
    free(x);
    return sqrt(acc / 7.0);
}

/* --- socks (test/runtests.jl:427-437) ------------------------------ */
/* theta = (n_socks, prop_pairs) pushed (n_socks integer).  Sequentially pick
 * min(n_socks, 11) socks without replacement; identical in distribution to
 * the reference's randperm-and-take.  distance = |pairs-0| + |odds-11| with
 * data = (0, 11) in data[0..1]. */
static double model_socks(const double* th, const double* data, simrng_t* r)
{
    double n_socks = th[0], prop = th[1];
    double n_pairs = rint(prop * floor(n_socks / 2.0));
    double n_odd = n_socks - 2.0 * n_pairs;
    double n_pick = n_socks < 11.0 ? n_socks : 11.0;
    /* state: pairs with both socks unpicked (p2), pairs with one picked (p1), odd unpicked (o) */
    double p2 = n_pairs, p1 = 0.0, o = n_odd, got_pairs = 0.0;
    for (int t = 0; t < (int)n_pick; ++t) {
        double tot = 2.0 * p2 + p1 + o;
        double u = sim_u(r) * tot;
        if (u < 2.0 * p2) { p2 -= 1.0; p1 += 1.0; }
        else if (u < 2.0 * p2 + p1) { p1 -= 1.0; got_pairs += 1.0; }
        else { o -= 1.0; }
    }
    double sample_pairs = got_pairs;
    double sample_odds = n_pick - 2.0 * got_pairs;
    return fabs(sample_pairs - data[0]) + fabs(sample_odds - data[1]);
}

/* dist!(theta, ve) -> (d, blob).  theta is already push_p-ed. */
static double simulate(int model, const double* th, const double* data, simrng_t* r, void* blob)
{
    double z1, z2, u1, u2;
    switch (model) {
    case M_GAUSS1D:           /* examples/minimal_example.jl:17-24; test/runtests.jl:138 */
        return fabs(th[0] + data[1] * sim_n(r) - data[0]);
    case M_GAUSS1D_BLOB: {    /* same, blob = simulated y (docs/src/index.md:298-324) */
        double y = th[0] + data[1] * sim_n(r);
        if (blob) memcpy(blob, &y, 8);
        return fabs(y - data[0]);
    }
    case M_GAUSS_CORR10: {    /* config 2; AR(1) noise: Sigma_ij = rho^|i-j| */
        double rho = data[10], sr = sqrt(1.0 - rho * rho), e = 0.0, acc = 0.0, z[10];
        for (int k = 0; k < 10; k += 2) sim_n2(r, &z[k], &z[k + 1]);
        for (int k = 0; k < 10; ++k) {
            e = (k == 0) ? z[0] : rho * e + sr * z[k];
            double dy = th[k] + e - data[k];
            acc += dy * dy;
        }
        return sqrt(acc);
    }
    case M_DIRAC:             /* test/runtests.jl:496-497 */
        return fabs(th[0] * th[0] + 1.0 - data[0]);
    case M_NORMDU:            /* test/runtests.jl:524-525 */
        return fabs((th[0] * th[0] + th[1]) * (th[0] + sim_n(r) * 0.01) - data[0]);
    case M_TWOD: case M_TWOD_INF: {   /* test/runtests.jl:603,614 */
        sim_n2(r, &z1, &z2);
        double t1 = th[0] + z1 * 0.01 - th[1] * th[1];
        double t2 = th[1] - 1.0 + z2 * 0.01;
        double v = 50.0 * (t1 * t1) + t2 * t2;
        if (model == M_TWOD_INF) { sim_u2(r, &u1, &u2); if (!(u1 < 0.5)) v = INFINITY; }
        return v;
    }
    case M_MIXTURE: {         /* test/runtests.jl:582-583 */
        sim_n2(r, &z1, &z2);
        sim_u2(r, &u1, &u2);
        double noise = (u1 < 0.5) ? z1 * 0.1 : z2;
        return fabs(th[0] + noise - data[0]);
    }
    case M_WIENER: {          /* test/runtests.jl:537-549; data[0..30] = tdata.  The reference's
                               * `@.(sqrt(...)) .* (0.95 + 0.1 * rand())` draws ONE scalar factor per
                               * simulation (the @. does not reach the rand()). */
        double acc = 0.0, f = 0.95 + 0.1 * sim_u(r);
        for (int t = 0; t <= 30; ++t) {
            double tt = (double)t;
            double v = sqrt(th[0] * th[0] * tt * tt + th[1] * th[1] * tt) * f;
            acc += fabs(v - data[t]);
        }
        return acc / 31.0;
    }
    case M_LOTKA_VOLTERRA: lv_linear = 0; return model_lv(th, data, r);
    case M_LOTKA_VOLTERRA_LIN: lv_linear = 1; return model_lv(th, data, r);
    case M_GK_F32:         return model_gk_f32(th, data, r);
    case M_BIRTH_DEATH:    return model_bd(th, data, r, (double*)blob);
    case M_GK:             return model_gk(th, data, r);
    case M_SOCKS:          return model_socks(th, data, r);
    }
    return NAN;
}

/* ------------------------------------------------------------------ */
/* exported stage functions                                            */
/* ------------------------------------------------------------------ */
static void mk_prior(prior_t* pr, int d, const int32_t* family, const double* params)
{
    pr->d = d;
    for (int k = 0; k < d; ++k) {
        pr->family[k] = family[k];
        for (int j = 0; j < 4; ++j) pr->p[k][j] = params[4 * k + j];
    }
}

int orc_push(int d, const int32_t* family, int64_t N, const double* theta, double* out)
{
    prior_t pr; double zero[4 * ORC_MAXD] = { 0 };
    mk_prior(&pr, d, family, zero);
    for (int64_t i = 0; i < N; ++i) push_p(&pr, theta + i * d, out + i * d);
    return 0;
}

/* logpdf(prior, push_p(prior, theta)) per particle, src/abcdez_smc.jl:243 */
int orc_prior_logpdf(int d, const int32_t* family, const double* params, int64_t N,
                     const double* theta, double* out)
{
    prior_t pr; mk_prior(&pr, d, family, params);
    for (int64_t i = 0; i < N; ++i) {
        double x[ORC_MAXD];
        push_p(&pr, theta + i * d, x);
        out[i] = prior_logpdf(&pr, x);
    }
    return 0;
}

/* logpdf of an already-pushed point (no rounding), for the Factored tests */
int orc_prior_logpdf_raw(int d, const int32_t* family, const double* params, int64_t N,
                         const double* x, double* out)
{
    prior_t pr; mk_prior(&pr, d, family, params);
    for (int64_t i = 0; i < N; ++i) out[i] = prior_logpdf(&pr, x + i * d);
    return 0;
}

int orc_prior_sample(int d, const int32_t* family, const double* params, int64_t N,
                     uint64_t seed, uint32_t epoch, int64_t id0, double* theta)
{
    prior_t pr; mk_prior(&pr, d, family, params);
    for (int64_t i = 0; i < N; ++i) prior_sample(&pr, seed, (uint32_t)(id0 + i), epoch, theta + i * d);
    return 0;
}

/* one simulate-and-score evaluation per particle (for model parity tests) */
int orc_simulate(int model, const double* data, int d, int64_t N, const double* theta_pushed,
                 uint64_t seed, uint32_t epoch, uint32_t tag, int64_t id0, double* dist, uint8_t* blobs)
{
    int B = MODELS[model].blob;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        simrng_t r; r.s = mk_stream(seed, (uint32_t)(id0 + i), epoch, tag); r.blk = 0;
        uint8_t blob[ORC_MAXBLOB];
        dist[i] = simulate(model, theta_pushed + i * d, data, &r, blob);
        if (B && blobs) memcpy(blobs + i * B, blob, (size_t)B);
    }
    return 0;
}

#define ORC_INIT_MAX_ATTEMPTS 100000

/* abcde_init!, src/abcdez_init.jl:2-22.  theta/logpi come in as the prior
 * draws of src/abcdez_smc.jl:242-243 (or are drawn here when draw_prior!=0,
 * attempt 0), particles with non-finite logpi or distance are redrawn
 * (attempt a uses prior epoch a and TAG_INIT_MODEL epoch a). */
int orc_init(int d, const int32_t* family, const double* params, int model, const double* data,
             int64_t N, uint64_t seed, int64_t id0, int draw_prior,
             double* theta, double* logpi, double* delta, uint8_t* blobs, int64_t* nredraw)
{
    prior_t pr; mk_prior(&pr, d, family, params);
    int B = MODELS[model].blob;
    int64_t redraws = 0; int fail = 0;
#pragma omp parallel for schedule(static) reduction(+:redraws) reduction(|:fail)
    for (int64_t i = 0; i < N; ++i) {
        double x[ORC_MAXD]; uint8_t blob[ORC_MAXBLOB];
        double* th = theta + i * d;
        uint32_t pid = (uint32_t)(id0 + i);
        if (draw_prior) {
            prior_sample(&pr, seed, pid, 0, th);
            push_p(&pr, th, x);
            logpi[i] = prior_logpdf(&pr, x);
        }
        double dl = NAN;
        if (isfinite(logpi[i])) {                                   /* init.jl:9-13 */
            simrng_t r; r.s = mk_stream(seed, pid, 0, TAG_INIT_MODEL); r.blk = 0;
            push_p(&pr, th, x);
            dl = simulate(model, x, data, &r, blob);
        }
        uint32_t attempt = 0;
        while (!isfinite(dl) || !isfinite(logpi[i])) {              /* init.jl:14-20 */
            if (++attempt >= ORC_INIT_MAX_ATTEMPTS) { fail = 1; break; }
            prior_sample(&pr, seed, pid, attempt, th);
            push_p(&pr, th, x);
            logpi[i] = prior_logpdf(&pr, x);
            simrng_t r; r.s = mk_stream(seed, pid, attempt, TAG_INIT_MODEL); r.blk = 0;
            dl = simulate(model, x, data, &r, blob);
            redraws++;
        }
        delta[i] = dl;
        if (B && blobs) memcpy(blobs + i * B, blob, (size_t)B);
    }
    if (nredraw) *nredraw = redraws;
    return fail ? 5 : 0;
}

/* StatsBase.wsample(rng, 1:N, alive) at src/abcdez_smc.jl:121,125, faithful
 * O(N) form: t = u*sum(alive); i=1; cw=alive[1]; while cw<t && i<N: i+=1;
 * cw+=alive[i].  Returns a 0-based index. */
static int64_t wsample_faithful(const uint8_t* alive, int64_t N, double u)
{
    double sum = 0.0;
    for (int64_t i = 0; i < N; ++i) sum += alive[i] ? 1.0 : 0.0;
    double t = u * sum;
    int64_t i = 0;
    double cw = alive[0] ? 1.0 : 0.0;
    while (cw < t && i < N - 1) { i++; cw += alive[i] ? 1.0 : 0.0; }
    return i;
}

/* same answer in O(1) from the compacted list of alive indices:
 * the ceil(t)-th alive particle, and index 0 when t == 0. */
static inline int64_t wsample_list(const uint32_t* alive_list, int64_t n_alive, double u, int64_t first)
{
    double t = u * (double)n_alive;
    int64_t k = (int64_t)ceil(t);
    if (k <= 0) return first;      /* index 1 of the (sub)population, alive or not */
    if (k > n_alive) k = n_alive;
    return (int64_t)alive_list[k - 1];
}

static int64_t build_alive_list(const uint8_t* alive, int64_t N, uint32_t* list)
{
    int64_t n = 0;
    for (int64_t i = 0; i < N; ++i) if (alive[i]) list[n++] = (uint32_t)i;
    return n;
}

#define FLAG_SIM 1
#define FLAG_ACC 2
#define ORC_PARTNER_MAX_ATTEMPTS 100000

/* abcdesmc_swarm!, src/abcdez_smc.jl:106-153.  Jacobi sweep: reads generation
 * g (theta, logpi, delta), writes generation g+1 (n*), which the caller has
 * pre-filled with a copy of g (smc.jl:337-340).  inj_* may be NULL (Philox
 * contract) or per-particle arrays (indices 0-based).  flags[i] gets
 * FLAG_SIM / FLAG_ACC. */
/* islands > 1 (sharded runs of the CUDA library): the population is cut into `islands`
 * contiguous blocks [floor(r N / R), floor((r+1) N / R)) and the DE partners of a particle are drawn from
 * the alive particles of its own block (the north-star's rank-local subpopulations); islands = 1 is the
 * reference's global pool (src/abcdez_smc.jl:121). */
#define ORC_MAX_ISLANDS 64
int orc_smc_sweep_islands(int d, const int32_t* family, const double* params, int model, const double* data,
                  int64_t N, const double* theta, const double* logpi, const double* delta,
                  const uint8_t* blobs, const uint8_t* alive,
                  double eps, int kind, double gamma0, double gsig,
                  uint64_t seed, uint32_t epoch, int64_t id0,
                  const int32_t* inj_a, const int32_t* inj_b, const double* inj_z, const double* inj_u,
                  int faithful_wsample, int islands,
                  double* ntheta, double* nlogpi, double* ndelta, uint8_t* nblobs, uint8_t* flags,
                  int64_t* nsims_out, int64_t* naccs_out)
{
    prior_t pr; mk_prior(&pr, d, family, params);
    int B = MODELS[model].blob;
    if (islands < 1) islands = 1;
    if (islands > ORC_MAX_ISLANDS) return 1;
    uint32_t* alist = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(N > 0 ? N : 1));
    int64_t isl_lo[ORC_MAX_ISLANDS + 1], isl_off[ORC_MAX_ISLANDS + 1];
    {   /* per-island alive lists, stored back to back */
        int64_t n = 0;
        for (int r = 0; r < islands; ++r) {
            isl_lo[r] = (int64_t)(((__int128)N * r) / islands);
            int64_t hi = (int64_t)(((__int128)N * (r + 1)) / islands);
            isl_off[r] = n;
            for (int64_t i = isl_lo[r]; i < hi; ++i) if (alive[i]) alist[n++] = (uint32_t)i;
        }
        isl_lo[islands] = N; isl_off[islands] = n;
    }
    int64_t nsims = 0, naccs = 0; int fail = 0;
    memcpy(ntheta, theta, sizeof(double) * (size_t)(N * d));
    memcpy(nlogpi, logpi, sizeof(double) * (size_t)N);
    memcpy(ndelta, delta, sizeof(double) * (size_t)N);
    if (B && blobs && nblobs) memcpy(nblobs, blobs, (size_t)(N * B));
#pragma omp parallel for schedule(static) reduction(+:nsims,naccs) reduction(|:fail)
    for (int64_t i = 0; i < N; ++i) {
        if (flags) flags[i] = 0;
        if (!alive[i]) continue;                                           /* :114 */
        uint32_t pid = (uint32_t)(id0 + i);
        int64_t a, b;
        int isl = 0;
        while (isl + 1 < islands && i >= isl_lo[isl + 1]) isl++;
        const uint32_t* ilist = alist + isl_off[isl];
        const int64_t in_alive = isl_off[isl + 1] - isl_off[isl], ifirst = isl_lo[isl];
        if (inj_a && inj_b) { a = inj_a[i]; b = inj_b[i]; }
        else {
            stream_t ps = mk_stream(seed, pid, epoch, TAG_PARTNER);
            uint32_t att = 0; double u1, u2;
            a = i;
            while (a == i) {                                               /* :119-122 */
                if (att >= ORC_PARTNER_MAX_ATTEMPTS) { fail = 1; break; }
                stream_u2(&ps, att++, &u1, &u2);
                a = faithful_wsample ? wsample_faithful(alive, N, u1) : wsample_list(ilist, in_alive, u1, ifirst);
            }
            att = 0; b = a;
            while (b == a || b == i) {                                     /* :123-126 */
                if (att >= ORC_PARTNER_MAX_ATTEMPTS) { fail = 1; break; }
                stream_u2(&ps, att++, &u1, &u2);
                b = faithful_wsample ? wsample_faithful(alive, N, u2) : wsample_list(ilist, in_alive, u2, ifirst);
            }
            if (fail) continue;
        }
        stream_t ms = mk_stream(seed, pid, epoch, TAG_MOVE);
        double z, z2;
        if (inj_z) z = inj_z[i]; else stream_n2(&ms, 0, &z, &z2);
        double g = gamma0 * (1.0 + z * gsig);                              /* :128 */
        double thp[ORC_MAXD], x[ORC_MAXD];
        for (int k = 0; k < d; ++k) {
            double diff = theta[a * d + k] - theta[b * d + k];
            double sc = diff * g;
            thp[k] = theta[i * d + k] + sc;
        }
        push_p(&pr, thp, x);
        double lp = prior_logpdf(&pr, x);                                  /* :134 */
        if (lp < 0.0 && isinf(lp)) continue;                               /* :135 */
        simrng_t r; r.s = mk_stream(seed, pid, epoch, TAG_MODEL); r.blk = 0;
        uint8_t blob[ORC_MAXBLOB];
        double dp = simulate(model, x, data, &r, blob);                    /* :137 */
        nsims++;                                                           /* :138 */
        if (flags) flags[i] |= FLAG_SIM;
        double w = lp - logpi[i];                                          /* :140-141, left to right */
        w = w + orc_kernel_logpdf(kind, eps, dp);
        w = w - orc_kernel_logpdf(kind, eps, delta[i]);
        int acc = (0.0 <= w);
        if (!acc) {                                                        /* :145, uniform only if w<0 */
            double u, u2;
            if (inj_u) u = inj_u[i]; else stream_u2(&ms, 1, &u, &u2);
            acc = (plog(u) < w);
        }
        if (acc) {                                                         /* :146-150 */
            ndelta[i] = dp;
            for (int k = 0; k < d; ++k) ntheta[i * d + k] = thp[k];
            nlogpi[i] = lp;
            if (B && nblobs) memcpy(nblobs + i * B, blob, (size_t)B);
            naccs++;
            if (flags) flags[i] |= FLAG_ACC;
        }
    }
    free(alist);
    if (nsims_out) *nsims_out = nsims;
    if (naccs_out) *naccs_out = naccs;
    return fail ? 6 : 0;
}

int orc_smc_sweep(int d, const int32_t* family, const double* params, int model, const double* data,
                  int64_t N, const double* theta, const double* logpi, const double* delta,
                  const uint8_t* blobs, const uint8_t* alive,
                  double eps, int kind, double gamma0, double gsig,
                  uint64_t seed, uint32_t epoch, int64_t id0,
                  const int32_t* inj_a, const int32_t* inj_b, const double* inj_z, const double* inj_u,
                  int faithful_wsample,
                  double* ntheta, double* nlogpi, double* ndelta, uint8_t* nblobs, uint8_t* flags,
                  int64_t* nsims_out, int64_t* naccs_out)
{
    return orc_smc_sweep_islands(d, family, params, model, data, N, theta, logpi, delta, blobs, alive, eps, kind,
                                 gamma0, gsig, seed, epoch, id0, inj_a, inj_b, inj_z, inj_u, faithful_wsample, 1,
                                 ntheta, nlogpi, ndelta, nblobs, flags, nsims_out, naccs_out);
}

/* Statistics.quantile(v, p) default (type 7) over the alive distances,
 * src/abcdez_smc.jl:301.  aleph = fma(n, p, 1-p) (Statistics.jl >= 1.9; the
 * unfused n*p+m differs by <= 1 ulp of aleph), j = clamp(trunc(aleph), 1, n-1),
 * g = clamp(aleph - j, 0, 1); a + g*(b-a) if both finite else (1-g)*a + g*b. */
static int cmp_double(const void* a, const void* b)
{
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

int orc_quantile_alive(int64_t N, const double* delta, const uint8_t* alive, double p,
                       double* q_out, double* a_out, double* b_out, int64_t* j_out)
{
    double* v = (double*)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
    int64_t n = 0;
    for (int64_t i = 0; i < N; ++i) if (alive[i]) { if (isnan(delta[i])) { free(v); return 3; } v[n++] = delta[i]; }
    if (n == 0) { free(v); return 4; }
    double m = 1.0 - p;                     /* alpha + p*(1-alpha-beta), alpha=beta=1 */
    double aleph = fma((double)n, p, m);
    int64_t j = (int64_t)trunc(aleph);
    if (j < 1) j = 1;
    if (j > n - 1) j = n - 1;
    double g = aleph - (double)j;
    if (g < 0.0) g = 0.0;
    if (g > 1.0) g = 1.0;
    double a, b;
    if (n == 1) { a = v[0]; b = v[0]; j = 1; }
    else {
        /* Statistics.quantile partial-sorts (sort!(v, 1:n, PartialQuickSort(j:j+1))): v[j] by quickselect, then v[j+1] is
         * the minimum of what the selection left above position j -- the same two order statistics as a full sort */
        int64_t lo = 0, hi = n - 1, k = j - 1;
        while (lo < hi) {
            double piv = v[lo + (hi - lo) / 2];
            int64_t i = lo, t = hi;
            while (i <= t) {
                while (cmp_double(&v[i], &piv) < 0) ++i;
                while (cmp_double(&v[t], &piv) > 0) --t;
                if (i <= t) { double tmp = v[i]; v[i] = v[t]; v[t] = tmp; ++i; --t; }
            }
            if (k <= t) hi = t; else if (k >= i) lo = i; else break;
        }
        a = v[k]; b = v[k + 1];
        for (int64_t i = k + 2; i < n; ++i) if (cmp_double(&v[i], &b) < 0) b = v[i];
    }
    double q = (isfinite(a) && isfinite(b)) ? a + g * (b - a) : (1.0 - g) * a + g * b;
    free(v);
    if (q_out) *q_out = q;
    if (a_out) *a_out = a;
    if (b_out) *b_out = b;
    if (j_out) *j_out = j;
    return 0;
}

/* Base.sum pairwise summation (block 128 here; Julia's is 1024 with a SIMD
 * inner loop whose association order is unspecified, so only <= few-ulp
 * agreement is meaningful). */
static double pairwise_sum(const double* x, int64_t n)
{
    if (n <= 128) { double s = 0.0; for (int64_t i = 0; i < n; ++i) s += x[i]; return s; }
    int64_t h = n / 2;
    return pairwise_sum(x, h) + pairwise_sum(x + h, n - h);
}

/* abcdesmc_update_ws! + the statements of src/abcdez_smc.jl:305-315,323:
 * ws (alive only), wprod, wnorm, Wns, alive, ess.  Dead particles keep a
 * stale ws in the reference; their Wns is 0 so wprod is 0 (ws is finite). */
int orc_reweight(int64_t N, const double* delta, double* W, uint8_t* alive,
                 double eps_old, double eps_new, int kind,
                 double* wnorm_out, double* ess_out, int64_t* nalive_out)
{
    double* wprod = (double*)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
    for (int64_t i = 0; i < N; ++i) {
        double ws = 0.0;
        if (alive[i])                                                      /* :71-76 */
            ws = pexp(orc_kernel_logpdf(kind, eps_new, delta[i]) - orc_kernel_logpdf(kind, eps_old, delta[i]));
        wprod[i] = alive[i] ? W[i] * ws : 0.0;                              /* :308 */
    }
    double wnorm = pairwise_sum(wprod, N);                                  /* :309 */
    int64_t na = 0;
    for (int64_t i = 0; i < N; ++i) {
        W[i] = wprod[i] / wnorm;                                            /* :310 */
        alive[i] = (W[i] > 0.0);                                            /* :311 */
        na += alive[i];
        wprod[i] = W[i] * W[i];
    }
    double ess = 1.0 / pairwise_sum(wprod, N);                              /* :8,:323 */
    free(wprod);
    if (wnorm_out) *wnorm_out = wnorm;
    if (ess_out) *ess_out = ess;
    if (nalive_out) *nalive_out = na;
    return 0;
}

/* wsample_stratified!, src/abcdez_smc.jl:15-56, statement by statement.
 * inds are returned 1-based exactly as the reference produces them, so the
 * quirks are visible: 0 when r == 0, and > N if rounding pushes r above the
 * sequential total (the reference would then throw a BoundsError at :50). */
int orc_wsample_stratified(int64_t N, const double* weights, const double* uniforms, int64_t* inds)
{
    double sval = 1.0 / (double)N;                                         /* :34 */
    double wsum = 0.0; int64_t i = 0;                                      /* :37-38 */
    double unif0 = 0.0, unif1 = 0.0;                                       /* :41-42 */
    for (int64_t si = 0; si < N; ++si) {                                   /* :45 */
        unif1 = unif0 + sval;                                              /* :46 */
        double r = unif0 + (unif1 - unif0) * uniforms[si];                 /* :47 rand(Uniform(a,b)) = a+(b-a)*u */
        while (r > wsum) {                                                 /* :48-51 */
            i += 1;
            if (i > N) break;
            wsum += weights[i - 1];
        }
        unif0 = unif1;                                                     /* :52 */
        inds[si] = i;                                                      /* :53 */
    }
    return 0;
}

/* the library's documented clamp of the two out-of-range quirks */
static inline int64_t clamp_ind(int64_t i, int64_t N) { return i < 1 ? 1 : (i > N ? N : i); }

void orc_resample_uniforms(int64_t N, uint64_t seed, uint32_t epoch, int64_t id0, double* u)
{
    for (int64_t si = 0; si < N; ++si) {
        stream_t s = mk_stream(seed, (uint32_t)(id0 + si), epoch, TAG_RESAMPLE);
        double u2;
        stream_u2(&s, 0, &u[si], &u2);
    }
}

/* ------------------------------------------------------------------ */
/* abcdesmc!, src/abcdez_smc.jl:215-394                                */
/* ------------------------------------------------------------------ */
typedef struct {
    int64_t nparticles;
    double alpha, delta_ess;
    int64_t nsims_max;
    int32_t Kmcmc;
    double Kmcmc_min;
    int32_t kind;
    double facc_stop, facc_min, facc_tune;
    uint64_t seed;
    int32_t faithful_wsample;   /* 1: O(N) StatsBase scan, 0: O(1) alive list */
    int32_t max_iters;          /* safety bound for benchmarks; 0 = unbounded */
    int32_t islands;            /* 0/1: the reference's global partner pool; R > 1: rank-local partners (sharded runs) */
} orc_smc_opts;

typedef struct {
    double eps, logZ;
    int64_t iters, nsims;
    int32_t status;             /* 0 ok, 1 no alive particles */
    int32_t hist_len;           /* entries written to the history arrays */
    double sweep_seconds;       /* wall time spent inside abcdesmc_swarm! */
    int64_t nsweeps;
} orc_smc_result;

static double now_s(void)
{
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return (double)clock() / CLOCKS_PER_SEC;
#endif
}

/* History arrays (length hist_cap each, may be NULL): eps, dmin, dmax, logZ,
 * ess, facc, gamma0, Kmcmc -- entry 0 is the pre-loop record of :284-292. */
int orc_smc_run(int d, const int32_t* family, const double* params, int model, const double* data,
                double eps_target, const orc_smc_opts* o, orc_smc_result* res,
                double* P, double* Wout, double* C, uint8_t* blobs_out,
                int32_t hist_cap, double* h_eps, double* h_dmin, double* h_dmax, double* h_logZ,
                double* h_ess, double* h_facc, double* h_gamma0, int32_t* h_K)
{
    prior_t pr; mk_prior(&pr, d, family, params);
    int64_t N = o->nparticles;
    int B = MODELS[model].blob;
    size_t nb = (size_t)(B ? N * B : 1);
    double* th = (double*)malloc(sizeof(double) * (size_t)(N * d));
    double* nth = (double*)malloc(sizeof(double) * (size_t)(N * d));
    double* lp = (double*)malloc(sizeof(double) * (size_t)N);
    double* nlp = (double*)malloc(sizeof(double) * (size_t)N);
    double* dl = (double*)malloc(sizeof(double) * (size_t)N);
    double* ndl = (double*)malloc(sizeof(double) * (size_t)N);
    uint8_t* bl = (uint8_t*)calloc(nb, 1);
    uint8_t* nbl = (uint8_t*)calloc(nb, 1);
    double* W = (double*)malloc(sizeof(double) * (size_t)N);
    uint8_t* alive = (uint8_t*)malloc((size_t)N);
    double* us = (double*)malloc(sizeof(double) * (size_t)N);
    int64_t* inds = (int64_t*)malloc(sizeof(int64_t) * (size_t)N);
    int rc = orc_init(d, family, params, model, data, N, o->seed, 0, 1, th, lp, dl, bl, NULL);   /* :242-252 */
    if (rc) goto done;

    double eps = INFINITY, eps_k = INFINITY;                               /* :255-256 */
    double ess_min = (double)N * o->delta_ess;                             /* :259 */
    double logZ = 0.0;                                                     /* :263 */
    for (int64_t i = 0; i < N; ++i) { W[i] = 1.0 / (double)N; alive[i] = 1; }   /* :266-271 */
    double ess = 0.0, facc = 1.0;
    int64_t nsims = 0, naccs = 0, n_alive = N;
    int Ki = o->Kmcmc;
    double gamma0 = 2.38 / sqrt(2.0 * (double)d), gsig = 1e-5;             /* :280-281 */
    int32_t hl = 0;
    uint32_t sweep_epoch = 0;
    res->status = 0; res->sweep_seconds = 0.0; res->nsweeps = 0;

#define PUSH_HIST(e_, ess_, facc_, K_) do { if (hl < hist_cap) { \
        double mn = INFINITY, mx = -INFINITY; \
        for (int64_t q = 0; q < N; ++q) { if (dl[q] < mn) mn = dl[q]; if (dl[q] > mx) mx = dl[q]; } \
        if (h_eps) h_eps[hl] = (e_); if (h_dmin) h_dmin[hl] = mn; if (h_dmax) h_dmax[hl] = mx; \
        if (h_logZ) h_logZ[hl] = logZ; if (h_ess) h_ess[hl] = (ess_); if (h_facc) h_facc[hl] = (facc_); \
        if (h_gamma0) h_gamma0[hl] = gamma0; if (h_K) h_K[hl] = (K_); hl++; } } while (0)

    {   /* :284-292, esss[1] = get_ess(Wns) */
        double s2 = 0.0; for (int64_t i = 0; i < N; ++i) s2 += W[i] * W[i];
        PUSH_HIST(eps, 1.0 / s2, facc, Ki);
    }

    int64_t iters = 0;
    for (;;) {                                                             /* :295 */
        iters++;
        double q;
        rc = orc_quantile_alive(N, dl, alive, o->alpha, &q, NULL, NULL, NULL);   /* :301 */
        if (rc) break;
        eps = fmax(fmin(q, eps), eps_target);
        double wnorm;
        orc_reweight(N, dl, W, alive, eps_k, eps, o->kind, &wnorm, &ess, &n_alive);   /* :305-311,323 */
        logZ += plog(wnorm);                                               /* :315 */
        naccs = 0; Ki = o->Kmcmc;                                          /* :318-319 */
        if (facc < o->facc_min) gamma0 *= o->facc_tune;                    /* :320 */
        if (ess < ess_min) {                                               /* :324-326 */
            orc_resample_uniforms(N, o->seed, (uint32_t)iters, 0, us);
            orc_wsample_stratified(N, W, us, inds);
            for (int64_t i = 0; i < N; ++i) {                              /* :96-99 through a temp */
                int64_t s = clamp_ind(inds[i], N) - 1;
                memcpy(nth + i * d, th + s * d, sizeof(double) * (size_t)d);
                nlp[i] = lp[s]; ndl[i] = dl[s];
                if (B) memcpy(nbl + i * B, bl + s * B, (size_t)B);
            }
            { double* t; uint8_t* tb;
              t = th; th = nth; nth = t; t = lp; lp = nlp; nlp = t; t = dl; dl = ndl; ndl = t;
              tb = bl; bl = nbl; nbl = tb; }
            for (int64_t i = 0; i < N; ++i) { W[i] = 1.0 / (double)N; alive[i] = 1; }   /* :102-103 */
            n_alive = N;
            double s2 = 0.0; for (int64_t i = 0; i < N; ++i) s2 += W[i] * W[i];
            ess = 1.0 / s2;
        }
        for (int i = 1; i <= o->Kmcmc; ++i) {                              /* :336-353 */
            int64_t ns = 0, na = 0;
            double t0 = now_s();
            rc = orc_smc_sweep_islands(d, family, params, model, data, N, th, lp, dl, bl, alive, eps, o->kind,
                               gamma0, gsig, o->seed, sweep_epoch++, 0, NULL, NULL, NULL, NULL,
                               o->faithful_wsample, o->islands, nth, nlp, ndl, nbl, NULL, &ns, &na);
            res->sweep_seconds += now_s() - t0; res->nsweeps++;
            if (rc) break;
            nsims += ns; naccs += na;
            { double* t; uint8_t* tb;                                       /* :347-350 */
              t = th; th = nth; nth = t; t = lp; lp = nlp; nlp = t; t = dl; dl = ndl; ndl = t;
              tb = bl; bl = nbl; nbl = tb; }
            if ((double)naccs / (double)n_alive >= o->Kmcmc_min) { Ki = i; break; }   /* :352 */
        }
        if (rc) break;
        facc = (double)naccs / ((double)n_alive * (double)Ki);             /* :357 */
        eps_k = eps;                                                       /* :360 */
        PUSH_HIST(eps, ess, facc, Ki);                                     /* :362-370 */
        if (n_alive == 0) { res->status = 1; break; }                      /* :375 */
        if (eps <= eps_target || nsims >= o->nsims_max || facc < o->facc_stop) break;   /* :376 */
        if (o->max_iters > 0 && iters >= o->max_iters) break;
    }
    res->eps = eps; res->logZ = logZ; res->iters = iters; res->nsims = nsims; res->hist_len = hl;
    for (int64_t i = 0; i < N; ++i) push_p(&pr, th + i * d, P + i * d);   /* :382 */
    memcpy(Wout, W, sizeof(double) * (size_t)N);
    memcpy(C, dl, sizeof(double) * (size_t)N);
    if (B && blobs_out) memcpy(blobs_out, bl, (size_t)(N * B));
done:
    free(th); free(nth); free(lp); free(nlp); free(dl); free(ndl); free(bl); free(nbl);
    free(W); free(alive); free(us); free(inds);
    return rc;
}

/* ------------------------------------------------------------------ */
/* abcdemc_swarm!, src/abcdez_mc.jl:5-61                               */
/* ------------------------------------------------------------------ */
/* order[] = particle indices sorted by (delta, index) ascending; cnt_le[i] =
 * #{j : delta[j] <= delta[i]}.  The reference draws s uniformly from the
 * index-ordered list (1:N)[delta .<= delta[i]] (:23); with the Philox
 * provider the draw is the floor(u*cnt)-th entry of the (delta,index)-sorted
 * order -- the same distribution through a different enumeration, shared with
 * the CUDA library.  inj_s (0-based) reproduces any reference choice exactly. */
typedef struct { double d; int64_t i; } di_t;
static int cmp_di(const void* a, const void* b)
{
    const di_t* x = (const di_t*)a; const di_t* y = (const di_t*)b;
    if (x->d < y->d) return -1;
    if (x->d > y->d) return 1;
    return (x->i > y->i) - (x->i < y->i);
}

/* islands > 1 (sharded runs): base particle and partners are drawn inside the particle's own contiguous
 * block (rank-local subpopulations); eps_pop stays the caller's global value. */
int orc_mc_sweep_islands(int d, const int32_t* family, const double* params, int model, const double* data,
                 int64_t N, const double* theta, const double* logpi, const double* delta,
                 const uint8_t* blobs, double eps_pop, double eps_target, double gamma0, double gsig,
                 uint64_t seed, uint32_t epoch, int64_t id0,
                 const int32_t* inj_s, const int32_t* inj_a, const int32_t* inj_b,
                 const double* inj_z, const double* inj_u, int islands,
                 double* ntheta, double* nlogpi, double* ndelta, uint8_t* nblobs, uint8_t* flags,
                 int64_t* nsims_out)
{
    prior_t pr; mk_prior(&pr, d, family, params);
    int B = MODELS[model].blob;
    if (islands < 1) islands = 1;
    if (islands > ORC_MAX_ISLANDS) return 1;
    int64_t isl_lo[ORC_MAX_ISLANDS + 1];
    for (int r = 0; r <= islands; ++r) isl_lo[r] = (int64_t)(((__int128)N * r) / islands);
    di_t* srt = (di_t*)malloc(sizeof(di_t) * (size_t)N);
    for (int64_t i = 0; i < N; ++i) { srt[i].d = delta[i]; srt[i].i = i; }
    for (int r = 0; r < islands; ++r)           /* each block sorted on its own, in place */
        qsort(srt + isl_lo[r], (size_t)(isl_lo[r + 1] - isl_lo[r]), sizeof(di_t), cmp_di);
    int64_t nsims = 0; int fail = 0;
    memcpy(ntheta, theta, sizeof(double) * (size_t)(N * d));
    memcpy(nlogpi, logpi, sizeof(double) * (size_t)N);
    memcpy(ndelta, delta, sizeof(double) * (size_t)N);
    if (B && blobs && nblobs) memcpy(nblobs, blobs, (size_t)(N * B));
#pragma omp parallel for schedule(static) reduction(+:nsims) reduction(|:fail)
    for (int64_t i = 0; i < N; ++i) {
        if (flags) flags[i] = 0;
        uint32_t pid = (uint32_t)(id0 + i);
        stream_t cs = mk_stream(seed, pid, epoch, TAG_MC);
        int isl = 0;
        while (isl + 1 < islands && i >= isl_lo[isl + 1]) isl++;
        const int64_t L0 = isl_lo[isl], NL = isl_lo[isl + 1] - isl_lo[isl];   /* the particle's block */
        int64_t s = i;                                                     /* :18 */
        double eps = (delta[i] <= eps_target) ? eps_target : eps_pop;      /* :19 */
        if (delta[i] > eps) {                                              /* :20-24 */
            if (inj_s) s = inj_s[i];
            else {
                /* cnt = #{delta <= delta[i]} by upper-bound search on the sorted keys */
                int64_t lo = 0, hi = NL;
                while (lo < hi) { int64_t m = (lo + hi) / 2; if (srt[L0 + m].d <= delta[i]) lo = m + 1; else hi = m; }
                double u1, u2; stream_u2(&cs, 0, &u1, &u2);
                int64_t k = (int64_t)floor(u1 * (double)lo);
                if (k >= lo) k = lo - 1;
                s = srt[L0 + k].i;
            }
        }
        int64_t a, b;
        if (inj_a && inj_b) { a = inj_a[i]; b = inj_b[i]; }
        else {
            stream_t ps = mk_stream(seed, pid, epoch, TAG_PARTNER);
            uint32_t att = 0; double u1, u2;
            a = s;
            while (a == s) {                                               /* :25-28, rand(1:N) -> floor(u*N) */
                if (att >= ORC_PARTNER_MAX_ATTEMPTS) { fail = 1; break; }
                stream_u2(&ps, att++, &u1, &u2);
                a = (int64_t)floor(u1 * (double)NL); if (a >= NL) a = NL - 1;
                a += L0;
            }
            att = 0; b = a;
            while (b == a || b == s) {                                     /* :29-32 */
                if (att >= ORC_PARTNER_MAX_ATTEMPTS) { fail = 1; break; }
                stream_u2(&ps, att++, &u1, &u2);
                b = (int64_t)floor(u2 * (double)NL); if (b >= NL) b = NL - 1;
                b += L0;
            }
            if (fail) continue;
        }
        stream_t ms = mk_stream(seed, pid, epoch, TAG_MOVE);
        double z, z2;
        if (inj_z) z = inj_z[i]; else stream_n2(&ms, 0, &z, &z2);
        double g = gamma0 * (1.0 + z * gsig);                              /* :34 */
        double thp[ORC_MAXD], x[ORC_MAXD];
        for (int k = 0; k < d; ++k) {
            double diff = theta[a * d + k] - theta[b * d + k];
            double sc = diff * g;
            thp[k] = theta[s * d + k] + sc;
        }
        push_p(&pr, thp, x);
        double lp = prior_logpdf(&pr, x);                                  /* :41 */
        double w_prior = lp - logpi[i];                                    /* :42, logpi[i] not [s] */
        double u, u2;
        if (inj_u) u = inj_u[i]; else stream_u2(&ms, 1, &u, &u2);
        if (plog(u) > fmin(0.0, w_prior)) continue;                        /* :43, uniform always drawn */
        nsims++;                                                           /* :44 */
        if (flags) flags[i] |= FLAG_SIM;
        simrng_t r; r.s = mk_stream(seed, pid, epoch, TAG_MODEL); r.blk = 0;
        uint8_t blob[ORC_MAXBLOB];
        double dp = simulate(model, x, data, &r, blob);                    /* :45 */
        if (dp <= fmax(eps, delta[i])) {                                   /* :54-59 */
            ndelta[i] = dp;
            for (int k = 0; k < d; ++k) ntheta[i * d + k] = thp[k];
            nlogpi[i] = lp;
            if (B && nblobs) memcpy(nblobs + i * B, blob, (size_t)B);
            if (flags) flags[i] |= FLAG_ACC;
        }
    }
    free(srt);
    if (nsims_out) *nsims_out = nsims;
    return fail ? 6 : 0;
}

int orc_mc_sweep(int d, const int32_t* family, const double* params, int model, const double* data,
                 int64_t N, const double* theta, const double* logpi, const double* delta,
                 const uint8_t* blobs, double eps_pop, double eps_target, double gamma0, double gsig,
                 uint64_t seed, uint32_t epoch, int64_t id0,
                 const int32_t* inj_s, const int32_t* inj_a, const int32_t* inj_b,
                 const double* inj_z, const double* inj_u,
                 double* ntheta, double* nlogpi, double* ndelta, uint8_t* nblobs, uint8_t* flags,
                 int64_t* nsims_out)
{
    return orc_mc_sweep_islands(d, family, params, model, data, N, theta, logpi, delta, blobs, eps_pop, eps_target,
                                gamma0, gsig, seed, epoch, id0, inj_s, inj_a, inj_b, inj_z, inj_u, 1,
                                ntheta, nlogpi, ndelta, nblobs, flags, nsims_out);
}

typedef struct {
    int64_t nparticles;
    int32_t generations;
    uint64_t seed;
    int32_t islands;            /* 0/1: global pools (the reference); R > 1: rank-local pools (sharded runs) */
} orc_mc_opts;

typedef struct {
    int32_t reached_eps;
    int64_t nsims;
    double dmin, dmax;
    double sweep_seconds;
} orc_mc_result;

/* abcdemc!, src/abcdez_mc.jl:102-172 */
int orc_mc_run(int d, const int32_t* family, const double* params, int model, const double* data,
               double eps_target, const orc_mc_opts* o, orc_mc_result* res,
               double* P, double* C, uint8_t* blobs_out)
{
    prior_t pr; mk_prior(&pr, d, family, params);
    int64_t N = o->nparticles;
    int B = MODELS[model].blob;
    size_t nb = (size_t)(B ? N * B : 1);
    double* th = (double*)malloc(sizeof(double) * (size_t)(N * d));
    double* nth = (double*)malloc(sizeof(double) * (size_t)(N * d));
    double* lp = (double*)malloc(sizeof(double) * (size_t)N);
    double* nlp = (double*)malloc(sizeof(double) * (size_t)N);
    double* dl = (double*)malloc(sizeof(double) * (size_t)N);
    double* ndl = (double*)malloc(sizeof(double) * (size_t)N);
    uint8_t* bl = (uint8_t*)calloc(nb, 1);
    uint8_t* nbl = (uint8_t*)calloc(nb, 1);
    int rc = orc_init(d, family, params, model, data, N, o->seed, 0, 1, th, lp, dl, bl, NULL);   /* :117-125 */
    int64_t nsims = 0;
    double gamma0 = 2.38 / sqrt(2.0 * (double)d), gsig = 1e-5;             /* :130-131 */
    res->sweep_seconds = 0.0;
    for (int it = 0; rc == 0 && it < o->generations; ++it) {               /* :134 */
        double el = INFINITY, eh = -INFINITY;                              /* :146 */
        for (int64_t i = 0; i < N; ++i) { if (dl[i] < el) el = dl[i]; if (dl[i] > eh) eh = dl[i]; }
        double eps_pop = fmax(eps_target, el + 0.0 * (eh - el));           /* :147, alpha = 0 (:107) */
        int64_t ns = 0;
        double t0 = now_s();
        rc = orc_mc_sweep_islands(d, family, params, model, data, N, th, lp, dl, bl, eps_pop, eps_target,
                          gamma0, gsig, o->seed, (uint32_t)it, 0, NULL, NULL, NULL, NULL, NULL, o->islands,
                          nth, nlp, ndl, nbl, NULL, &ns);
        res->sweep_seconds += now_s() - t0;
        nsims += ns;
        { double* t; uint8_t* tb;                                           /* :152-155 */
          t = th; th = nth; nth = t; t = lp; lp = nlp; nlp = t; t = dl; dl = ndl; ndl = t;
          tb = bl; bl = nbl; nbl = tb; }
    }
    double mn = INFINITY, mx = -INFINITY;
    for (int64_t i = 0; i < N; ++i) { if (dl[i] < mn) mn = dl[i]; if (dl[i] > mx) mx = dl[i]; }
    res->reached_eps = (mx <= eps_target);                                 /* :163 */
    res->nsims = nsims; res->dmin = mn; res->dmax = mx;
    for (int64_t i = 0; i < N; ++i) push_p(&pr, th + i * d, P + i * d);   /* :166 */
    memcpy(C, dl, sizeof(double) * (size_t)N);
    if (B && blobs_out) memcpy(blobs_out, bl, (size_t)(N * B));
    free(th); free(nth); free(lp); free(nlp); free(dl); free(ndl); free(bl); free(nbl);
    return rc;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
