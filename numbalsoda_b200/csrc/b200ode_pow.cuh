// pow(x, y) for x >= 0 and the exponents the solvers use (y = 1/k, k = 2..13),
// bit-identical to the pow() the reference calls on the host (glibc 2.39, the
// FMA-dispatched variant every x86-64 CPU since Haswell selects).
//
// The reference takes every step-size decision through libm pow
// (dop853_module.f90:618,:817; LSODA.cpp:1229,:1983-2056,:2106-2113); CUDA's pow is
// accurate to 2 ulp but not the same function, and a last-bit difference in h is
// amplified by chaotic problems and can flip accept/reject or order decisions
// (SURVEY.md section 7).  glibc implements pow with the ARM optimized-routines
// algorithm (log in double-double through a 128-entry table, then exp through a
// 128-entry table).  The sequence below reproduces, operation for operation
// (including which products are fused), what the compiled __pow_fma of this
// image's libm.so.6 executes on its main path; the tables come from the same
// binary (tools/gen_pow_tables.py).  tests/test_pow.py checks it bit-for-bit
// against the host libm on the CPU (this header also compiles as host code) and
// tests/test_gpu_pow.py on the device.
//
// Domain handled: x = +0, subnormal, normal, +inf, NaN; 2^-66 <= |y| < 2^62
// (main path of glibc); results that stay in the normal range (always true for
// y = 1/k).  x < 0 returns NaN (pow of a negative base to a non-integer power).
#ifndef B200ODE_POW_CUH
#define B200ODE_POW_CUH

#if defined(__CUDACC__)
#if defined(B200POW_NOINLINE)
// one out-of-line copy: the LSODA kernels call pow from five places, and their code
// size (instruction-cache misses are their largest stall reason) matters more than the call
#define B200POW_ENTRY __device__ __noinline__
#endif
#define B200POW_FN __host__ __device__ __forceinline__
#define B200POW_TABLE_QUAL static __device__ const
#define B200POW_HOST_TABLES 0
#else
#include <cmath>
#include <cstring>
#define B200POW_FN static inline
#define B200POW_TABLE_QUAL static const
#define B200POW_HOST_TABLES 1
#endif

#include "b200ode_pow_tables.cuh"

B200POW_FN unsigned long long b200pow_bits(double x)
{
#if defined(__CUDA_ARCH__)
    return (unsigned long long)__double_as_longlong(x);
#else
    unsigned long long u;
    memcpy(&u, &x, 8);
    return u;
#endif
}
B200POW_FN double b200pow_double(unsigned long long u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x;
    memcpy(&x, &u, 8);
    return x;
#endif
}
// explicitly rounded primitives: never re-contracted by the compiler
B200POW_FN double b200pow_fma(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}
B200POW_FN double b200pow_mul(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
B200POW_FN double b200pow_add(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}

B200POW_FN double b200ode_pow_impl(double x, double y)
{
    unsigned long long ix = b200pow_bits(x);
    const unsigned topx = (unsigned)(ix >> 52);
    if (topx - 1u >= 0x7feu) {
        // x is zero, subnormal, inf, NaN or negative
        if (x != x) return x + x;
        if (ix >> 63) {
            if ((ix << 1) == 0) return (y > 0.0) ? 0.0 : 1.0 / 0.0;  // pow(-0, y)
            return (x - x) / (x - x);                               // negative base
        }
        if (ix == 0) return (y > 0.0) ? 0.0 : 1.0 / 0.0;
        if (topx == 0x7ff) return (y > 0.0) ? x : 0.0;
        // subnormal: normalise (e_pow.c: ix = asuint64(x * 0x1p52) - (52 << 52))
        ix = b200pow_bits(b200pow_mul(x, 4503599627370496.0));
        ix &= 0x7fffffffffffffffULL;
        ix -= 52ULL << 52;
    }
    // ---- log_inline: log(x) = k ln2 + log(c) + log1p(z/c - 1), as hi + lo
    const unsigned long long tmp = ix - 0x3fe6955500000000ULL;
    const int i = (int)((tmp >> 45) & 127);
    const int k = (int)((long long)tmp >> 52);
    const unsigned long long iz = ix - (tmp & 0xfff0000000000000ULL);
    const double z = b200pow_double(iz);
    const double kd = (double)k;
    const double invc = b200pow_log_tab[i][0];
    const double logc = b200pow_log_tab[i][1];
    const double logctail = b200pow_log_tab[i][2];
    const double t1 = b200pow_fma(kd, B200POW_LN2HI, logc);
    const double lo1 = b200pow_fma(kd, B200POW_LN2LO, logctail);
    const double r = b200pow_fma(z, invc, -1.0);
    const double ar = b200pow_mul(r, B200POW_A0);
    const double p12 = b200pow_fma(r, B200POW_A2, B200POW_A1);
    const double p34 = b200pow_fma(r, B200POW_A4, B200POW_A3);
    const double t2 = b200pow_add(r, t1);
    const double lo2 = b200pow_add(b200pow_add(t1, -t2), r);
    const double ar2 = b200pow_mul(r, ar);
    const double ar3 = b200pow_mul(r, ar2);
    const double lo3 = b200pow_fma(ar, r, -ar2);
    const double hi = b200pow_add(t2, ar2);
    const double p56 = b200pow_fma(r, B200POW_A6, B200POW_A5);
    const double lo4 = b200pow_add(b200pow_add(t2, -hi), ar2);
    const double q = b200pow_fma(p56, ar2, p34);
    const double pp = b200pow_fma(ar2, q, p12);
    double lo = b200pow_add(lo1, lo2);
    lo = b200pow_add(lo, lo3);
    lo = b200pow_add(lo, lo4);
    lo = b200pow_fma(ar3, pp, lo);
    const double loghi = b200pow_add(hi, lo);
    const double loglo = b200pow_add(b200pow_add(hi, -loghi), lo);
    // ---- y * log(x) in double-double
    const double ehi = b200pow_mul(y, loghi);
    const double elo = b200pow_fma(y, loglo, b200pow_fma(loghi, y, -ehi));
    // ---- exp_inline(ehi, elo)
    const unsigned abstop = (unsigned)(b200pow_bits(ehi) >> 52) & 0x7ff;
    if (abstop - 0x3c9u >= 0x3fu) {
        if (abstop < 0x3c9u) return b200pow_add(1.0, ehi);  // |y log x| < 2^-54
        // |y log x| >= 512: outside what the solvers can produce; saturate like pow
        return (b200pow_bits(ehi) >> 63) ? 0.0 : 1.0 / 0.0;
    }
    double kk = b200pow_fma(ehi, B200POW_INVLN2N, B200POW_SHIFT);
    const unsigned long long ki = b200pow_bits(kk);
    kk = b200pow_add(kk, -B200POW_SHIFT);
    double rr = b200pow_fma(kk, B200POW_NEGLN2HIN, ehi);
    rr = b200pow_fma(kk, B200POW_NEGLN2LON, rr);
    rr = b200pow_add(elo, rr);
    const int idx = 2 * (int)(ki & 127);
    const unsigned long long sbits = b200pow_exp_tab[idx + 1] + (ki << 45);
    const double tail = b200pow_double(b200pow_exp_tab[idx]);
    const double c23 = b200pow_fma(rr, B200POW_C3, B200POW_C2);
    const double tr = b200pow_add(rr, tail);
    const double r2 = b200pow_mul(rr, rr);
    const double c45 = b200pow_fma(rr, B200POW_C5, B200POW_C4);
    double t = b200pow_fma(c23, r2, tr);
    const double r4 = b200pow_mul(r2, r2);
    t = b200pow_fma(c45, r4, t);
    const double scale = b200pow_double(sbits);
    return b200pow_fma(t, scale, scale);
}

#ifndef B200POW_ENTRY
#define B200POW_ENTRY B200POW_FN
#endif
B200POW_ENTRY double b200ode_pow(double x, double y) { return b200ode_pow_impl(x, y); }

B200POW_FN double b200ode_pow_eighth(double x) { return b200ode_pow_impl(x, 0.125); }

#endif
