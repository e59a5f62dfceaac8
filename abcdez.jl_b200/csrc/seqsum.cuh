// seqsum.cuh -- closed form of a *sequential* FP64 running sum with a constant addend.
//
// The reference's stratified resampling (wsample_stratified!, src/abcdez_smc.jl:15-56) is a
// strictly sequential loop: the stratum edge is the running sum unif0 += 1/N (:46,:52) and
// the cumulative weight is the running sum wsum += weights[i] (:50).  With the indicator
// kernels (the default, src/abcdez_smc.jl:218) every alive weight is the same double c and
// dead weights are 0.0, so wsum after k alive particles is S_c(k) with
//        S(0) = 0,  S(k) = fl(S(k-1) + c)          (round-to-nearest-even)
// A parallel scan rounds differently; this file reproduces S(k) *bit-exactly* in O(log) per
// query instead.  Within one binade [2^e, 2^(e+1)) the ulp u is fixed and fl(s + c) - s is a
// constant multiple of u -- except that under a round-half-even tie (c mod u == u/2) the very
// first step into the binade may differ, after which the running sum is an even multiple of
// u and the increment is constant again.  So each binade is: <= 2 explicitly stepped values,
// then one arithmetic progression.  ~3 segments per binade, ~log2(N) binades.
#pragma once
#include "internal.h"

namespace abcdez {

__host__ __device__ inline int f64_exponent(double x)
{
    // exponent field of a positive normal double
#ifdef __CUDA_ARCH__
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
#else
    unsigned long long b; memcpy(&b, &x, 8);
#endif
    return (int)((b >> 52) & 0x7ff);
}

__host__ __device__ inline double f64_add_rn(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}

// build the table for S(k), k = 0..n
__host__ __device__ inline void seqtab_build(SeqTab* t, double c, unsigned long long n)
{
    int m = 0;
    unsigned long long k = 0;
    double s = 0.0;
    t->kmax = n;
    // segment 0: the single point S(0) = 0
    t->kst[m] = 0; t->sst[m] = 0.0; t->inc[m] = 0.0; m++;
    if (!(c > 0.0) || n == 0) { t->n = m; return; }
    while (k < n && m < SEQTAB_MAX - 3) {
        // explicit step 1
        double s1 = f64_add_rn(s, c);
        if (s1 == s) {                      // stagnation: c < ulp(s)/2, the sum never moves again
            t->kst[m] = k; t->sst[m] = s; t->inc[m] = 0.0; m++;
            k = n; break;
        }
        k += 1; s = s1;
        t->kst[m] = k; t->sst[m] = s; t->inc[m] = 0.0; m++;
        if (k >= n) break;
        // explicit step 2 (settles the parity under a half-even tie)
        double s2 = f64_add_rn(s, c);
        if (s2 == s) continue;
        if (f64_exponent(s2) != f64_exponent(s)) continue;   // crossed a binade: restart there
        k += 1; s = s2;
        t->kst[m] = k; t->sst[m] = s; t->inc[m] = 0.0; m++;
        if (k >= n) break;
        // arithmetic progression from s while the values stay inside this binade
        double s3 = f64_add_rn(s, c);
        if (f64_exponent(s3) != f64_exponent(s) || s3 == s) continue;
        double inc = s3 - s;                                  // exact (same binade)
        // number of further steps j with s + j*inc < 2^(e+1): work on the integer mantissas
#ifdef __CUDA_ARCH__
        unsigned long long sb = (unsigned long long)__double_as_longlong(s);
#else
        unsigned long long sb; memcpy(&sb, &s, 8);
#endif
        unsigned long long mant = (sb & 0x000fffffffffffffull) | 0x0010000000000000ull;   // in [2^52, 2^53)
        // inc / ulp as an integer: ulp = 2^(e-1075+... ) -> compute via scaling by the same exponent
        int e = f64_exponent(s);
        // ulp(s) = 2^(e - 1023 - 52); inc is a multiple of ulp(s)
        double ulp = ldexp(1.0, e - 1023 - 52);
        unsigned long long inc_i = (unsigned long long)(inc / ulp);                          // exact
        unsigned long long room = (0x0020000000000000ull - 1ull - mant) / inc_i;             // steps that stay < 2^53
        unsigned long long steps = room;
        if (steps > n - k) steps = n - k;
        if (steps >= 1) {
            // the progression segment starts at (k, s): overwrite the last explicit entry's inc
            t->inc[m - 1] = inc;
            k += steps;
            s = s + (double)steps * inc;                       // exact
        }
    }
    if (k < n) {
        // table overflow (cannot happen for n <= 2^40 with SEQTAB_MAX = 256): mark by clamping kmax
        t->kmax = k;
    }
    // sentinel
    t->kst[m] = t->kmax + 1; t->sst[m] = INFINITY; t->inc[m] = 0.0;
    t->n = m;
}

// S(k) for 0 <= k <= kmax
__host__ __device__ inline double seqtab_value(const SeqTab* t, unsigned long long k)
{
    if (k > t->kmax) k = t->kmax;
    int lo = 0, hi = t->n - 1;           // last segment with kst <= k
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (t->kst[mid] <= k) lo = mid; else hi = mid - 1; }
    return t->sst[lo] + (double)(k - t->kst[lo]) * t->inc[lo];
}

// min{ k >= 0 : S(k) >= r }, clamped to kmax (the reference would run out of bounds, :48-51)
__host__ __device__ inline unsigned long long seqtab_first_ge(const SeqTab* t, double r)
{
    if (!(r > 0.0)) return 0;
    // last segment whose start value is < r
    int lo = 0, hi = t->n - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (t->sst[mid] < r) lo = mid; else hi = mid - 1; }
    unsigned long long k0 = t->kst[lo], knext = t->kst[lo + 1];   // sentinel at n
    double s0 = t->sst[lo], inc = t->inc[lo];
    unsigned long long len = knext - 1 - k0;                      // progression covers k0 .. k0+len
    if (inc > 0.0 && len > 0) {
        double q = (r - s0) / inc;
        unsigned long long j = (q >= (double)len) ? len : (unsigned long long)q;
        // fix the floating estimate: want the smallest j with s0 + j*inc >= r
        while (j < len && s0 + (double)j * inc < r) j++;
        while (j > 0 && s0 + (double)(j - 1) * inc >= r) j--;
        if (s0 + (double)j * inc >= r) return k0 + j;
    }
    // not reached inside this segment: the next segment starts at a value >= r (or is the sentinel)
    return knext > t->kmax ? t->kmax : knext;
}

}  // namespace abcdez
