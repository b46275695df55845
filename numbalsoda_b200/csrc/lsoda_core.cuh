// LSODA for one trajectory held by one thread (systems of N <= 8 equations).
//
// Replaces, per trajectory, the reference's lsoda_wrapper (src/wrapper.cpp:10-57) ->
// LSODA::lsoda_update (src/LSODA.cpp:2247-2277) -> LSODA::lsoda (:287-993) ->
// stoda (:995-1366) -> correction / prja / dgefa / dgesl / methodswitch /
// orderswitch / scaleh / intdy.  Every floating-point expression keeps the
// reference's operand order, so that with FMA contraction disabled at link time the
// states and the step / RHS / Jacobian / method-switch counters are those of the CPU
// path bit for bit.  What is different is where things live and how control flows:
//
//  * storage.  The Nordsieck history yh (13 x N) and the iteration matrix wm (N x N)
//    are indexed by the *run-time* order and pivot, which registers cannot do; they
//    live in shared memory, one column per thread ([slot][thread]: consecutive lanes
//    hit consecutive banks, no conflicts), so nothing is ever spilled to local
//    memory.  y, savf, acor, ewt and ~25 scalars stay in registers.  The pivot
//    vector is packed 4 bits per row into one register.
//  * coefficients.  cfode's elco / tesco tables depend only on the method: they are
//    compile-time tables (lsoda_tables.cuh, bit-equal to cfode's arithmetic) instead
//    of being recomputed at every method switch; el[] is read from them in place.
//  * driver.  Instead of re-entering the solver once per output time (wrapper.cpp:32)
//    the lane steps freely and, after each accepted step, interpolates every t_eval
//    point the step has passed -- the same arithmetic.
//  * stoda's many exits funnel through one rescale site and one order-selection
//    site, so a warp whose lanes take different exits re-converges there.
//
// This header also compiles as host C++ (tests/lsoda_core_host.cpp) so that the
// restated control flow can be checked against the CPU oracle without a GPU; the
// product only ever runs it inside lsoda_kernel.cu.
#ifndef B200ODE_LSODA_CORE_CUH
#define B200ODE_LSODA_CORE_CUH

#include "b200ode_pow.cuh"
#include "lsoda_tables.cuh"

#ifndef N
#define N B200_NEQ
#endif

#if defined(__CUDACC__)
#define LS_FN __device__ __forceinline__
#else
#include <cmath>
#define LS_FN static inline
#endif

#ifndef LS_STRIDE
#error "define LS_STRIDE: distance, in doubles, between two slots of one trajectory"
#endif

#define LS_FOR_N _Pragma("unroll") for (int i = 0; i < N; i++)

#define LS_ETA 2.2204460492503131e-16      // LSODA.cpp:33
#define LS_SQRTETA 1.4901161193847656e-08  // sqrt(ETA) = 2^-26, LSODA.cpp:483
#define LS_MXORDN 12                       // LSODA.cpp:38
#define LS_MXORDS 5
#define LS_MAXCOR 3   // LSODA.cpp:564
#define LS_MSBP 20    // LSODA.cpp:565
#define LS_MXHNIL 10  // LSODA.cpp:294
#define LS_CCMAX 0.3  // LSODA.cpp:563

// yh[j][i], j = 1..13 (LSODA.h:114); wm[r][c] = d f_r / d y_c (LSODA.cpp:1680)
#define YH(j, i) yhp[(((j) - 1) * N + (i)) * LS_STRIDE]
#define WM(r, c) wmp[((r) * N + (c)) * LS_STRIDE]
// tables of the method in use (cfode is re-run for `meth` at :1086 before any use)
#define ELCO(q, k) T->elco[s.meth - 1][q][k]
#define TESCO(q, k) T->tesco[s.meth - 1][q][k]

struct LsLane {
    double y[N];  // the caller's vector: last corrector iterate (LSODA.cpp:1809,:1828)
    double h, hu, hold, tn, rc, el0, crate, rmax, pdest, pdlast, pdnorm;
    int nq, meth, mused, ialth, ipup, nslp, icount, irflag, jstart, jcur, kflag;
    int nst, nfe, nje, nqu;
    unsigned ipvt;  // 4 bits per row
};

// std::max / std::min as the reference uses them (they differ from fmax/fmin on NaN)
LS_FN double ls_max(double a, double b) { return (a < b) ? b : a; }
LS_FN double ls_min(double a, double b) { return (b < a) ? b : a; }

// hmin / |h| with hmin = 0 (LSODA.cpp:401, :1155, :1243 ...): +0, or NaN when h is 0 or NaN
LS_FN double ls_hmin_ratio(double h) { return 0.0 / fabs(h); }

// vmnorm, LSODA.cpp:1705-1712
LS_FN double ls_vmnorm(const double (&v)[N], const double (&w)[N])
{
    double vm = 0.0;
    LS_FOR_N vm = ls_max(vm, fabs(v[i]) * w[i]);
    return vm;
}
LS_FN double ls_vmnorm_yh(const double* yhp, int j, const double (&w)[N])
{
    double vm = 0.0;
    LS_FOR_N vm = ls_max(vm, fabs(YH(j, i)) * w[i]);
    return vm;
}

// resetcoeff, LSODA.cpp:2211-2225 (el[] is read from the table where it is used)
LS_FN void ls_resetcoeff(LsLane& s, const B200LsodaTables* T)
{
    const double el1 = ELCO(s.nq, 1);
    s.rc = s.rc * el1 / s.el0;
    s.el0 = el1;
}

// scaleh, LSODA.cpp:1594-1631 (hmxi = 0)
LS_FN void ls_scaleh(LsLane& s, double* yhp, const B200LsodaTables* T, double& rh, double& pdh)
{
    rh = ls_min(rh, s.rmax);
    rh = rh / ls_max(1.0, fabs(s.h) * 0.0 * rh);
    if (s.meth == 1) {
        s.irflag = 0;
        pdh = ls_max(fabs(s.h) * s.pdlast, 0.000001);
        if ((rh * pdh * 1.00001) >= T->sm1[s.nq]) {
            rh = T->sm1[s.nq] / pdh;
            s.irflag = 1;
        }
    }
    double r = 1.0;
    const int l = s.nq + 1;
    for (int j = 2; j <= l; j++) {
        r *= rh;
        LS_FOR_N YH(j, i) = YH(j, i) * r;
    }
    s.h *= rh;
    s.rc *= rh;
    s.ialth = l;
}

// Pascal-triangle sweeps: prediction (LSODA.cpp:1141-1144) and retraction
// (:1290-1295, :1909-1912).  For fixed j the inner loop runs i1 = j..nq upwards, so
// yh[i1+1] is still the value of the previous pass when it is added: it is carried in
// a register and every element costs one shared-memory load and one store.
template <bool FORWARD>
LS_FN void ls_pascal(const LsLane& s, double* yhp)
{
    for (int j = s.nq; j >= 1; j--) {
        double a[N];
        LS_FOR_N a[i] = YH(j, i);
        for (int i1 = j; i1 <= s.nq; i1++) {
            LS_FOR_N {
                const double b = YH(i1 + 1, i);
                YH(i1, i) = FORWARD ? a[i] + b : a[i] - b;
                a[i] = b;
            }
        }
    }
}

// idamax1's integer truncation (LSODA.cpp:59-77) as x86-64 evaluates (size_t)|x|:
// cvttsd2si below 2^63, (x - 2^63) with the top bit set up to 2^64, 0 above and for NaN.
LS_FN unsigned long long ls_trunc_u64(double x)
{
    const double a = fabs(x);
    if (!(a < 18446744073709551616.0)) return 0ULL;
#if defined(__CUDA_ARCH__)
    return __double2ull_rz(a);
#else
    return (unsigned long long)a;
#endif
}

// dgefa, LSODA.cpp:173-231: row-oriented elimination; the pivot is searched along
// row k and columns are swapped.  Returns nonzero when a zero pivot was met.
LS_FN int ls_dgefa(LsLane& s, double* wmp)
{
    int info = 0;
    unsigned ipvt = 0;
#pragma unroll
    for (int k = 0; k < N - 1; k++) {
        int j = k;
        unsigned long long vmax = 0;
#pragma unroll
        for (int c = k; c < N; c++) {
            const unsigned long long v = ls_trunc_u64(WM(k, c));
            if (v > vmax) {
                vmax = v;
                j = c;
            }
        }
        ipvt |= (unsigned)j << (4 * k);
        const double akj = WM(k, j);
        if (akj == 0.0) {
            info = k + 1;
            continue;
        }
        if (j != k) {
            WM(k, j) = WM(k, k);
            WM(k, k) = akj;
        }
        double t = -1.0 / akj;
        double rowk[N];
#pragma unroll
        for (int c = k + 1; c < N; c++) {
            rowk[c] = t * WM(k, c);
            WM(k, c) = rowk[c];
        }
#pragma unroll
        for (int r = k + 1; r < N; r++) {
            t = WM(r, j);
            if (j != k) {
                WM(r, j) = WM(r, k);
                WM(r, k) = t;
            }
#pragma unroll
            for (int c = k + 1; c < N; c++) WM(r, c) = t * rowk[c] + WM(r, c);
        }
    }
    ipvt |= (unsigned)(N - 1) << (4 * (N - 1));
    if (WM(N - 1, N - 1) == 0.0) info = N;
    s.ipvt = ipvt;
    return info;
}

// dgesl with job = 0, LSODA.cpp:119-144
LS_FN void ls_dgesl(unsigned ipvt, const double* wmp, double (&b)[N])
{
#pragma unroll
    for (int k = 0; k < N; k++) {
        double t = 0.0;
#pragma unroll
        for (int c = 0; c < k; c++) t += WM(k, c) * b[c];
        b[k] = (b[k] - t) / WM(k, k);
    }
#pragma unroll
    for (int k = N - 2; k >= 0; k--) {
        double t = 0.0;
#pragma unroll
        for (int c = k + 1; c < N; c++) t += WM(k, c) * b[c];
        b[k] = b[k] + t;
        const int j = (int)((ipvt >> (4 * k)) & 15u);
#pragma unroll
        for (int c = k + 1; c < N; c++) {
            if (c == j) {
                const double tt = b[c];
                b[c] = b[k];
                b[k] = tt;
            }
        }
    }
}

// prja, LSODA.cpp:1633-1696 (miter == 2: dense finite-difference Jacobian), fnorm
// (:1714-1736) and the LU factorisation.  Sets s.jcur; returns ierpj.
LS_FN int ls_prja(LsLane& s, double* wmp, const double (&savf)[N], double (&acor)[N],
                  const double (&ewt)[N], double* pp)
{
    s.nje++;
    s.jcur = 1;
    const double hl0 = s.h * s.el0;
    double fac = ls_vmnorm(savf, ewt);
    double r0 = 1000.0 * fabs(s.h) * LS_ETA * ((double)N) * fac;
    if (r0 == 0.0) r0 = 1.0;
    // one column per pass; a single (not unrolled) loop keeps one inlined copy of the RHS
#pragma unroll 1
    for (int j = 0; j < N; j++) {
        double yj = 0.0, ewtj = 0.0;
        LS_FOR_N {
            if (i == j) {
                yj = s.y[i];
                ewtj = ewt[i];
            }
        }
        const double r = ls_max(LS_SQRTETA * fabs(yj), r0 / ewtj);
        double yp[N];
        LS_FOR_N yp[i] = (i == j) ? s.y[i] + r : s.y[i];
        fac = -hl0 / r;
        b200ode_user_rhs(s.tn, yp, acor, pp);
        LS_FOR_N WM(i, j) = (acor[i] - savf[i]) * fac;
    }
    s.nfe += N;
    double an = 0.0;
#pragma unroll
    for (int r = 0; r < N; r++) {
        double sum = 0.0;
#pragma unroll
        for (int c = 0; c < N; c++) sum += fabs(WM(r, c)) / ewt[c];
        an = ls_max(an, sum * ewt[r]);
    }
    s.pdnorm = an / fabs(hl0);
    LS_FOR_N WM(i, i) = WM(i, i) + 1.0;
    return ls_dgefa(s, wmp) != 0 ? 1 : 0;
}

// corfailure, LSODA.cpp:1904-1922.  The reference increments the pointer `ncf`
// (:1906), so its mxncf limit never fires; with hmin = 0 corflag 2 needs h == 0.
LS_FN int ls_corfailure(LsLane& s, double* yhp, double told, double& rh)
{
    s.rmax = 2.0;
    s.tn = told;
    ls_pascal<false>(s, yhp);
    if (fabs(s.h) <= 0.0 * 1.00001) return 2;
    rh = 0.25;
    s.ipup = (s.meth == 2) ? 2 : 0;
    return 1;
}

// correction, LSODA.cpp:1743-1902.  Returns corflag; m = corrector iterations done.
LS_FN int ls_correction(LsLane& s, double* yhp, double* wmp, const B200LsodaTables* T,
                        double (&savf)[N], double (&acor)[N], const double (&ewt)[N],
                        double* pp, double pnorm, double& del, double told, double& rh, int& m)
{
    const int miter = (s.meth == 2) ? 2 : 0;
    double rate = 0.0, delp = 0.0;
    m = 0;
    del = 0.0;
    LS_FOR_N s.y[i] = YH(1, i);
    for (;;) {
        b200ode_user_rhs(s.tn, s.y, savf, pp);
        s.nfe++;
        if (m == 0) {
            if (s.ipup > 0) {
                const int ierpj = ls_prja(s, wmp, savf, acor, ewt, pp);
                s.ipup = 0;
                s.rc = 1.0;
                s.nslp = s.nst;
                s.crate = 0.7;
                if (ierpj != 0) return ls_corfailure(s, yhp, told, rh);
            }
            LS_FOR_N acor[i] = 0.0;
        }
        const double el1 = ELCO(s.nq, 1);
        if (miter == 0) {
            // functional iteration, LSODA.cpp:1792-1809
            LS_FOR_N {
                savf[i] = s.h * savf[i] - YH(2, i);
                s.y[i] = savf[i] - acor[i];
            }
            del = ls_vmnorm(s.y, ewt);
            LS_FOR_N {
                s.y[i] = YH(1, i) + el1 * savf[i];
                acor[i] = savf[i];
            }
        } else {
            // chord iteration, LSODA.cpp:1816-1829
            LS_FOR_N s.y[i] = s.h * savf[i] - (YH(2, i) + acor[i]);
            ls_dgesl(s.ipvt, wmp, s.y);
            del = ls_vmnorm(s.y, ewt);
            LS_FOR_N {
                acor[i] += s.y[i];
                s.y[i] = YH(1, i) + el1 * acor[i];
            }
        }
        // convergence test, LSODA.cpp:1842-1862
        if (del <= 100.0 * pnorm * LS_ETA) break;
        if (m != 0 || s.meth != 1) {
            if (m != 0) {
                double rm = 1024.0;
                if (del <= (1024.0 * delp)) rm = del / delp;
                rate = ls_max(rate, rm);
                s.crate = ls_max(0.2 * s.crate, rm);
            }
            const double dcon = del * ls_min(1.0, 1.5 * s.crate) / (TESCO(s.nq, 2) * T->conit[s.nq]);
            if (dcon <= 1.0) {
                s.pdest = ls_max(s.pdest, rate / fabs(s.h * el1));
                if (s.pdest != 0.0) s.pdlast = s.pdest;
                break;
            }
        }
        m++;
        if (m == LS_MAXCOR || (m >= 2 && del > 2.0 * delp)) {
            // LSODA.cpp:1871-1891
            if (miter == 0 || s.jcur == 1) return ls_corfailure(s, yhp, told, rh);
            s.ipup = miter;
            m = 0;
            rate = 0.0;
            del = 0.0;
            LS_FOR_N s.y[i] = YH(1, i);
        } else {
            delp = del;
        }
    }
    return 0;
}

// methodswitch, LSODA.cpp:1946-2069
LS_FN void ls_methodswitch(LsLane& s, const double* yhp, const B200LsodaTables* T,
                           const double (&ewt)[N], double dsm, double pnorm, double& pdh, double& rh)
{
    const int l = s.nq + 1;
    if (s.meth == 1) {
        int nqm2;
        double rh2;
        if (s.nq > 5) return;
        if (dsm <= (100.0 * pnorm * LS_ETA) || s.pdest == 0.0) {
            if (s.irflag == 0) return;
            rh2 = 2.0;
            nqm2 = s.nq < LS_MXORDS ? s.nq : LS_MXORDS;
        } else {
            const double exsm = 1.0 / (double)l;
            double rh1 = 1.0 / (1.2 * b200ode_pow(dsm, exsm) + 0.0000012);
            double rh1it = 2.0 * rh1;
            pdh = s.pdlast * fabs(s.h);
            if ((pdh * rh1) > 0.00001) rh1it = T->sm1[s.nq] / pdh;
            rh1 = ls_min(rh1, rh1it);
            // (nq <= 5 = mxords here, so the reference's nq > mxords branch, :1988-1994,
            // cannot be taken)
            const double dm2 = dsm * (T->cm1[s.nq] / T->cm2[s.nq]);
            rh2 = 1.0 / (1.2 * b200ode_pow(dm2, exsm) + 0.0000012);
            nqm2 = s.nq;
            if (rh2 < 5.0 * rh1) return;
        }
        rh = rh2;
        s.icount = 20;
        s.meth = 2;
        s.pdlast = 0.0;
        s.nq = nqm2;
        return;
    }
    // currently BDF, LSODA.cpp:2026-2068 (nq <= mxords <= mxordn: :2030-2036 unreachable)
    const double exsm = 1.0 / (double)l;
    double dm1 = dsm * (T->cm2[s.nq] / T->cm1[s.nq]);
    double rh1 = 1.0 / (1.2 * b200ode_pow(dm1, exsm) + 0.0000012);
    const int nqm1 = s.nq;
    const double exm1 = exsm;
    double rh1it = 2.0 * rh1;
    pdh = s.pdnorm * fabs(s.h);
    if ((pdh * rh1) > 0.00001) rh1it = T->sm1[nqm1] / pdh;
    rh1 = ls_min(rh1, rh1it);
    const double rh2 = 1.0 / (1.2 * b200ode_pow(dsm, exsm) + 0.0000012);
    if ((rh1 * 5.0) < (5.0 * rh2)) return;
    const double alpha = ls_max(0.001, rh1);
    dm1 *= b200ode_pow(alpha, exm1);
    if (dm1 <= 1000.0 * LS_ETA * pnorm) return;
    rh = rh1;
    s.icount = 20;
    s.meth = 1;
    s.pdlast = 0.0;
    s.nq = nqm1;
    (void)yhp;
    (void)ewt;
}

// orderswitch, LSODA.cpp:2097-2209.  Returns orderflag.
LS_FN int ls_orderswitch(LsLane& s, double* yhp, const B200LsodaTables* T, const double (&acor)[N],
                         const double (&ewt)[N], double rhup, double dsm, double& pdh, double& rh)
{
    const int l = s.nq + 1;
    const int lmax = (s.meth == 1 ? LS_MXORDN : LS_MXORDS) + 1;
    int newq;
    const double exsm = 1.0 / (double)l;
    double rhsm = 1.0 / (1.2 * b200ode_pow(dsm, exsm) + 0.0000012);
    double rhdn = 0.0;
    if (s.nq != 1) {
        const double ddn = ls_vmnorm_yh(yhp, l, ewt) / TESCO(s.nq, 1);
        const double exdn = 1.0 / (double)s.nq;
        rhdn = 1.0 / (1.3 * b200ode_pow(ddn, exdn) + 0.0000013);
    }
    if (s.meth == 1) {
        pdh = ls_max(fabs(s.h) * s.pdlast, 0.000001);
        if (l < lmax) rhup = ls_min(rhup, T->sm1[l] / pdh);
        rhsm = ls_min(rhsm, T->sm1[s.nq] / pdh);
        if (s.nq > 1) rhdn = ls_min(rhdn, T->sm1[s.nq - 1] / pdh);
        s.pdest = 0.0;
    }
    if (rhsm >= rhup) {
        if (rhsm >= rhdn) {
            newq = s.nq;
            rh = rhsm;
        } else {
            newq = s.nq - 1;
            rh = rhdn;
            if (s.kflag < 0 && rh > 1.0) rh = 1.0;
        }
    } else {
        if (rhup <= rhdn) {
            newq = s.nq - 1;
            rh = rhdn;
            if (s.kflag < 0 && rh > 1.0) rh = 1.0;
        } else {
            rh = rhup;
            if (rh >= 1.1) {
                const double r = ELCO(s.nq, l) / (double)l;
                s.nq = l;
                LS_FOR_N YH(s.nq + 1, i) = acor[i] * r;
                return 2;
            }
            s.ialth = 3;
            return 0;
        }
    }
    if (s.meth == 1) {
        if ((rh * pdh * 1.00001) < T->sm1[newq])
            if (s.kflag == 0 && rh < 1.1) {
                s.ialth = 3;
                return 0;
            }
    } else {
        if (s.kflag == 0 && rh < 1.1) {
            s.ialth = 3;
            return 0;
        }
    }
    if (s.kflag <= -2) rh = ls_min(rh, 0.2);
    if (newq == s.nq) return 1;
    s.nq = newq;
    return 2;
}

// stoda, LSODA.cpp:995-1366: one accepted step, or a failure reported in s.kflag
// (-1 repeated error-test failures, -2 corrector could not converge).
LS_FN void ls_stoda(LsLane& s, double* yhp, double* wmp, const B200LsodaTables* T,
                    const double (&ewt)[N], double* pp)
{
    double savf[N], acor[N];
    double del = 0.0, dsm = 0.0, rh = 0.0, pdh = 0.0, pnorm = 0.0;
    int m = 0;
    bool rescale = false;   // a step-size change is pending (scaleh)
    bool finished = false;  // ... and it ends an accepted step (rmax = 10, endstoda)
    bool fix_rmax = false;
    s.kflag = 0;
    const double told = s.tn;
    s.jcur = 0;
    if (s.jstart == 0) {
        // first call, LSODA.cpp:1058-1084 (lmax, l follow from meth and nq)
        s.nq = 1;
        s.ialth = 2;
        s.rmax = 10000.0;
        s.rc = 0.0;
        s.el0 = 1.0;
        s.crate = 0.7;
        s.hold = s.h;
        s.nslp = 0;
        s.ipup = 0;  // miter = 0 for meth = 1
        s.icount = 20;
        s.irflag = 0;
        s.pdest = 0.0;
        s.pdlast = 0.0;
        ls_resetcoeff(s, T);
    }
    if (s.jstart == -1) {
        // after a method switch, LSODA.cpp:1096-1124
        s.ipup = (s.meth == 2) ? 2 : 0;
        if (s.ialth == 1) s.ialth = 2;
        if (s.meth != s.mused) {
            s.ialth = s.nq + 1;
            ls_resetcoeff(s, T);
        }
        if (s.h != s.hold) {
            rh = s.h / s.hold;
            s.h = s.hold;
            rescale = true;
        }
    }
    for (;;) {
        if (rescale) {
            ls_scaleh(s, yhp, T, rh, pdh);
            rescale = false;
            if (finished) {
                if (fix_rmax) s.rmax = 10.0;
                break;
            }
        }
        // prediction, LSODA.cpp:1132-1147
        const int miter = (s.meth == 2) ? 2 : 0;
        if (fabs(s.rc - 1.0) > LS_CCMAX) s.ipup = miter;
        if (s.nst >= s.nslp + LS_MSBP) s.ipup = miter;
        s.tn += s.h;
        ls_pascal<true>(s, yhp);
        pnorm = ls_vmnorm_yh(yhp, 1, ewt);
        const int corflag = ls_correction(s, yhp, wmp, T, savf, acor, ewt, pp, pnorm, del, told, rh, m);
        if (corflag == 1) {
            rh = ls_max(rh, ls_hmin_ratio(s.h));
            rescale = true;
            continue;
        }
        if (corflag == 2) {
            s.kflag = -2;
            s.hold = s.h;
            s.jstart = 1;
            return;
        }
        // local error test, LSODA.cpp:1170-1177
        s.jcur = 0;
        if (m == 0) dsm = del / TESCO(s.nq, 2);
        if (m > 0) dsm = ls_vmnorm(acor, ewt) / TESCO(s.nq, 2);
        double rhup = 0.0;
        bool select_order = false;
        const bool accepted = dsm <= 1.0;
        if (accepted) {
            // LSODA.cpp:1191-1277
            const int l = s.nq + 1;
            const int lmax = (s.meth == 1 ? LS_MXORDN : LS_MXORDS) + 1;
            s.kflag = 0;
            s.nst++;
            s.hu = s.h;
            s.nqu = s.nq;
            s.mused = s.meth;
            for (int j = 1; j <= l; j++) {
                const double r = ELCO(s.nq, j);
                LS_FOR_N YH(j, i) = YH(j, i) + r * acor[i];
            }
            s.icount--;
            if (s.icount < 0) {
                ls_methodswitch(s, yhp, T, ewt, dsm, pnorm, pdh, rh);
                if (s.meth != s.mused) {
                    rh = ls_max(rh, ls_hmin_ratio(s.h));
                    rescale = finished = fix_rmax = true;
                    continue;
                }
            }
            s.ialth--;
            if (s.ialth == 0) {
                if (l != lmax) {
                    LS_FOR_N savf[i] = acor[i] - YH(lmax, i);
                    const double dup = ls_vmnorm(savf, ewt) / TESCO(s.nq, 3);
                    const double exup = 1.0 / (double)(l + 1);
                    rhup = 1.0 / (1.4 * b200ode_pow(dup, exup) + 0.0000014);
                }
                select_order = true;
            } else {
                if (!(s.ialth > 1 || l == lmax)) LS_FOR_N YH(lmax, i) = acor[i];
                break;
            }
        } else {
            // error test failed, LSODA.cpp:1286-1363
            s.kflag--;
            s.tn = told;
            ls_pascal<false>(s, yhp);
            s.rmax = 2.0;
            if (fabs(s.h) <= 0.0 * 1.00001) {
                s.kflag = -1;
                s.hold = s.h;
                s.jstart = 1;
                return;
            }
            if (s.kflag > -3) {
                select_order = true;
            } else {
                if (s.kflag == -10) {
                    s.kflag = -1;
                    s.hold = s.h;
                    s.jstart = 1;
                    return;
                }
                // three or more failures: back to order 1 with a fresh derivative
                rh = 0.1;
                rh = ls_max(ls_hmin_ratio(s.h), rh);
                s.h *= rh;
                LS_FOR_N s.y[i] = YH(1, i);
                b200ode_user_rhs(s.tn, s.y, savf, pp);
                s.nfe++;
                LS_FOR_N YH(2, i) = s.h * savf[i];
                s.ipup = (s.meth == 2) ? 2 : 0;
                s.ialth = 5;
                if (s.nq != 1) {
                    s.nq = 1;
                    ls_resetcoeff(s, T);
                }
                continue;
            }
        }
        if (select_order) {
            const int orderflag = ls_orderswitch(s, yhp, T, acor, ewt, rhup, dsm, pdh, rh);
            if (accepted) {
                if (orderflag == 0) break;
                if (orderflag == 2) ls_resetcoeff(s, T);
                rh = ls_max(rh, ls_hmin_ratio(s.h));
                rescale = finished = fix_rmax = true;
            } else {
                if (orderflag == 0) rh = ls_min(rh, 0.2);
                if (orderflag == 2) ls_resetcoeff(s, T);
                rh = ls_max(rh, ls_hmin_ratio(s.h));
                rescale = true;
            }
        }
    }
    // endstoda, LSODA.cpp:2075-2082.  acor is not read again before the next step
    // clears it (:1789), so its scaling by 1/tesco[nqu][2] is dropped.
    s.hold = s.h;
    s.jstart = 1;
}

// intdy with k = 0, LSODA.cpp:1435-1453 (c = 1: the reference's products with it are exact)
LS_FN void ls_intdy(const LsLane& s, const double* yhp, double t, double* out)
{
    const double sfrac = (t - s.tn) / s.h;
    double v[N];
    LS_FOR_N v[i] = 1.0 * YH(s.nq + 1, i);
    for (int j = s.nq - 1; j >= 0; j--) LS_FOR_N v[i] = 1.0 * YH(j + 1, i) + sfrac * v[i];
    LS_FOR_N out[i] = v[i];
}

// ---------------------------------------------------------------------------
// The driver: what lsoda_wrapper's loop over t_eval and LSODA::lsoda's blocks a-e
// do for one trajectory, reorganised as  begin / advance-one-step.

struct LsRun {
    int row;      // next usol row to write
    int nslast;   // nst at the last output (mxstep is per output interval, LSODA.cpp:671,:778)
    int nhnil;    // t + h == t warnings
    int nswitch;  // Adams <-> BDF switches
    int istate;   // 1 running; 2 finished; <= 0 failed (as LSODA.cpp:782-976; 0 = t_eval not monotone)
};

// First call: wrapper.cpp:25-36 and blocks b, c of LSODA::lsoda (LSODA.cpp:341-663).
// Returns true when the trajectory is already finished (R.istate says how).
LS_FN bool ls_begin(LsLane& s, LsRun& R, double* yhp, const double* u0, double* pp, int nt,
                    const double* teval, double* usol, double rtol, double atol)
{
    LS_FOR_N usol[i] = u0[i];
    R.row = 1;
    R.nslast = 0;
    R.nhnil = 0;
    R.nswitch = 0;
    R.istate = 2;
    s.nst = s.nfe = s.nje = s.nqu = 0;
    s.meth = 0;
    s.mused = 0;
    s.nq = 1;
    s.jstart = 0;
    s.jcur = 0;
    s.kflag = 0;
    s.ipvt = 0;
    s.hu = 0.0;
    s.tn = 0.0;
    s.h = 0.0;
    if (nt < 2) return true;
    if (teval[1] < teval[0]) {
        R.istate = 0;
        return true;
    }
    const double t0 = teval[0], tout = teval[1];
    s.tn = t0;
    s.meth = 1;
    if (rtol < 0.0 || atol < 0.0) {  // LSODA.cpp:509-520
        R.istate = -3;
        return true;
    }
    LS_FOR_N s.y[i] = u0[i];
    double f0[N];
    b200ode_user_rhs(t0, s.y, f0, pp);
    s.nfe = 1;
    LS_FOR_N YH(1, i) = s.y[i];
    double ewt[N];
    bool bad_ewt = false;
    LS_FOR_N {
        ewt[i] = rtol * fabs(s.y[i]) + atol;  // ewset with itol = 2
        if (ewt[i] <= 0.0) bad_ewt = true;
        ewt[i] = 1.0 / ewt[i];
    }
    if (bad_ewt) {
        // LSODA.cpp:585-590 leaves through terminate2 with istate still 1, which
        // wrapper.cpp:41 does not treat as a failure: every later call re-initialises,
        // fails the same way and hands u0 back.
        for (int r = 1; r < nt; r++) {
            if (teval[r] < teval[r - 1]) {
                R.istate = 0;
                return true;
            }
            LS_FOR_N usol[(size_t)r * N + i] = u0[i];
            R.row = r + 1;
        }
        R.istate = 2;
        return true;
    }
    const double tdist = fabs(tout - t0);
    const double w0 = ls_max(fabs(t0), fabs(tout));
    if (tdist < 2.0 * LS_ETA * w0) {  // LSODA.cpp:618-623
        R.istate = -3;
        return true;
    }
    double tol = rtol;
    if (tol <= 0.0) {
        LS_FOR_N {
            const double ayi = fabs(s.y[i]);
            if (ayi != 0.0) tol = ls_max(tol, atol / ayi);
        }
    }
    tol = ls_max(tol, 100.0 * LS_ETA);
    tol = ls_min(tol, 0.001);
    double sum = ls_vmnorm(f0, ewt);
    sum = 1.0 / (tol * w0 * w0) + tol * sum * sum;
    double h0 = 1.0 / sqrt(sum);
    h0 = ls_min(h0, tdist);
    h0 = h0 * ((tout - t0 >= 0.0) ? 1.0 : -1.0);
    s.h = h0;
    LS_FOR_N YH(2, i) = f0[i] * h0;
    R.istate = 1;
    return false;
}

// One pass of the step loop of LSODA::lsoda (block e, LSODA.cpp:774-992) followed by
// the output rows the step has reached.  Returns true when the trajectory is finished.
LS_FN bool ls_advance(LsLane& s, LsRun& R, double* yhp, double* wmp, const B200LsodaTables* T,
                      double* pp, int nt, const double* teval, double* usol, double rtol,
                      double atol, unsigned long long mxstep_u, int exit_on_warning)
{
    double ewt[N];
    if (s.nst != 0 && (unsigned long long)(long long)(s.nst - R.nslast) >= mxstep_u) {  // :778
        R.istate = -1;
        return true;
    }
    // ewset (:787-798).  Before the first step yh[1] is u0, whose weights were checked
    // by ls_begin, and the test for non-positive weights is skipped (:776).
    bool bad_ewt = false;
    LS_FOR_N {
        ewt[i] = rtol * fabs(YH(1, i)) + atol;
        if (ewt[i] <= 0.0) bad_ewt = true;
        ewt[i] = 1.0 / ewt[i];
    }
    if (s.nst != 0 && bad_ewt) {
        R.istate = -6;
        return true;
    }
    const double tolsf = LS_ETA * ls_vmnorm_yh(yhp, 1, ewt);
    if (tolsf > 0.01) {  // :801-818
        R.istate = (s.nst == 0) ? -3 : -2;
        return true;
    }
    if ((s.tn + s.h) == s.tn) {  // :820-845
        R.nhnil++;
        if (R.nhnil <= LS_MXHNIL && exit_on_warning) {
            R.istate = -2;
            return true;
        }
    }
    ls_stoda(s, yhp, wmp, T, ewt, pp);
    if (s.kflag != 0) {  // :961-991
        R.istate = (s.kflag == -1) ? -4 : -5;
        return true;
    }
    if (s.meth != s.mused) {  // :861-877
        s.jstart = -1;
        R.nswitch++;
    }
    // Every output time this step has reached: the first is what the reference
    // interpolates at :887, the others are its block-d returns (:675-689) on the
    // following calls, which re-check t_eval's order (wrapper.cpp:33) and intdy's
    // interval (:1426-1434).
    bool first_in_step = true;
    while (R.row < nt) {
        if (teval[R.row] < teval[R.row - 1]) {
            R.istate = 0;
            return true;
        }
        const double tout = teval[R.row];
        if ((s.tn - tout) * s.h < 0.0) break;
        const double tp = s.tn - s.hu - 100.0 * LS_ETA * (s.tn + s.hu);
        if ((tout - tp) * (tout - s.tn) > 0.0) {
            // intdy refuses (:1426-1434): after a step the reference ignores that and
            // returns the last corrector iterate; on a later call it is illegal input
            if (!first_in_step) {
                R.istate = -3;
                return true;
            }
            LS_FOR_N usol[(size_t)R.row * N + i] = s.y[i];
        } else {
            ls_intdy(s, yhp, tout, usol + (size_t)R.row * N);
        }
        first_in_step = false;
        R.nslast = s.nst;
        R.row++;
    }
    if (R.row >= nt) {
        R.istate = 2;
        return true;
    }
    return false;
}

#endif
