// LSODA vector backend: one thread owns one trajectory of N <= 8 equations.
//
// The vector work of the reference's routines -- vmnorm / fnorm (LSODA.cpp:1705-1736),
// ewset (:1368-1390), the Pascal sweeps of stoda (:1141-1144, :1290-1295), scaleh's
// rescale (:1621-1627), prja (:1633-1696), dgefa / dgesl (:110-231), the corrector
// updates (:1792-1829), the Nordsieck update (:1197-1202) and intdy (:1435-1453) -- for
// the thread-per-trajectory kernel (lsoda_kernel.cu).  Control flow: lsoda_core.cuh.
//
// Storage: y, ewt, acor, savf in registers (compile-time N, loops fully unrolled); the
// Nordsieck history yh (13 x N) and the iteration matrix wm (N x N), which are indexed
// by the *run-time* order and pivot, in shared memory as one column per thread
// ([slot][thread]: consecutive lanes hit consecutive banks, no conflicts), so nothing is
// ever spilled to local memory.  The pivot vector is packed 4 bits per row into one
// register.
#ifndef B200ODE_LSODA_VEC_THREAD_CUH
#define B200ODE_LSODA_VEC_THREAD_CUH

#include "lsoda_core.cuh"

#ifndef N
#define N B200_NEQ
#endif
#ifndef LS_STRIDE
#error "define LS_STRIDE: distance, in doubles, between two slots of one trajectory"
#endif

#define LS_FOR_N _Pragma("unroll") for (int i = 0; i < N; i++)
// request fields of the deferred operations (LS_DEFERRED)
#define LS_OP_RETRACT 1u
#define LS_OP_SCALE 2u
#define LS_OP_ROWS 4u
#define LS_OP_JAC 8u
#define LS_RD_RH 0
#define LS_RD_TN 1
#define LS_RD_H 2
#define LS_RD_JTN 3  // in: tn; out: pdnorm
#define LS_RD_JH 4
#define LS_RD_JHL0 5
#define LS_RD_COUNT 6
#define LS_RI_RETRACT_NQ 0
#define LS_RI_SCALE_L 1
#define LS_RI_ROW0 2
#define LS_RI_NROWS 3
#define LS_RI_NQ 4
#define LS_RI_IERPJ 5
#define LS_RI_IPVT 6
#define LS_RI_BLO 7
#define LS_RI_BHI 8
#define LS_RI_COUNT 9
#define LS_JAC_SLOT_Y 8
#define LS_JAC_SLOT_F 9
#define LS_JAC_SLOT_W 10
// yh[j][i], j = 1..13 (LSODA.h:114); wm[r][c] = d f_r / d y_c (LSODA.cpp:1680)
#define YH(j, i) yhp[(((j) - 1) * N + (i)) * LS_STRIDE]
#define WM(r, c) wmp[((r) * N + (c)) * LS_STRIDE]

struct LsThreadVec {
    double y[N];     // the caller's vector: last corrector iterate (LSODA.cpp:1809,:1828)
    double ewt[N];   // inverse error weights of the step (LSODA.cpp:787-798)
    double acor[N];  // accumulated correction (LSODA.cpp:1789,:1827)
    double savf[N];  // f(tn, y) of the current pass
    double* yhp;
    double* wmp;
    unsigned ipvt;  // 4 bits per row

    LS_MFN int size() const { return N; }
    LS_MFN int jac_rhs_calls() const { return N; }
    // which of the lanes in `lanes` satisfy `pred` (all of them take part in the vote)
    LS_MFN unsigned vote(unsigned lanes, bool pred) const
    {
#if defined(__CUDA_ARCH__)
        return __ballot_sync(lanes, pred);
#else
        (void)pred;
        return lanes;
#endif
    }

#if defined(LS_DEFERRED)
    // ---- deferred operations (lsoda_core.cuh: deferred backends; served by ls_serve in
    //      lsoda_kernel.cu).  A request is a few fields in this thread's slots of the block's
    //      request area (LS_REQD / LS_REQI: [field][thread], defined by the kernel file) and a
    //      bit in `ops`; the state the server needs beyond that -- yh, the matrix -- already
    //      lives in shared memory.  The Jacobian's inputs (y, f(tn, y), the weights) travel in
    //      yh slots 8-10: a Jacobian is only ever requested by the BDF method, whose orders
    //      (<= 5) use slots 1-6, and the Adams method initialises every slot it grows into.
    static const bool kDeferred = true;
    unsigned ops;
    LS_MFN void scale_yh(int l, double rh)
    {
        LS_REQI(LS_RI_SCALE_L) = l;
        LS_REQD(LS_RD_RH) = rh;
        ops |= LS_OP_SCALE;
    }
    LS_MFN void request_retract(int nq)
    {
        LS_REQI(LS_RI_RETRACT_NQ) = nq;
        ops |= LS_OP_RETRACT;
    }
    LS_MFN void intdy_at(int nq, double tn, double h, double, int row, double*)
    {
        if (!(ops & LS_OP_ROWS)) {
            LS_REQI(LS_RI_ROW0) = row;
            LS_REQI(LS_RI_NROWS) = 0;
            LS_REQI(LS_RI_NQ) = nq;
            LS_REQD(LS_RD_TN) = tn;
            LS_REQD(LS_RD_H) = h;
            ops |= LS_OP_ROWS;
        }
        LS_REQI(LS_RI_NROWS) = LS_REQI(LS_RI_NROWS) + 1;
    }
    LS_MFN void request_jac(double tn, double h, double hl0)
    {
        LS_REQD(LS_RD_JTN) = tn;
        LS_REQD(LS_RD_JH) = h;
        LS_REQD(LS_RD_JHL0) = hl0;
        LS_FOR_N {
            YH(LS_JAC_SLOT_Y, i) = y[i];
            YH(LS_JAC_SLOT_F, i) = savf[i];
            YH(LS_JAC_SLOT_W, i) = ewt[i];
        }
        ops |= LS_OP_JAC;
    }
    LS_MFN int take_jac(double& pdnorm)
    {
        pdnorm = LS_REQD(LS_RD_JTN);
        ipvt = (unsigned)LS_REQI(LS_RI_IPVT);
        return LS_REQI(LS_RI_IERPJ);
    }
#else
    static const bool kDeferred = false;
    LS_MFN void request_jac(double, double, double) {}
    LS_MFN int take_jac(double&) { return 0; }
    LS_MFN void request_retract(int nq) { pascal<false>(nq); }
    LS_MFN void intdy_at(int nq, double tn, double h, double tout, int row, double* usol) const
    {
        intdy(nq, ls_div(tout - tn, h), usol + (size_t)row * N);
    }
#endif
    LS_MFN void sync(unsigned lanes)
    {
#if defined(__CUDA_ARCH__)
        __syncwarp(lanes);
#else
        (void)lanes;
#endif
    }

    LS_MFN double norm(const double (&v)[N]) const
    {
        double vm = 0.0;
        LS_FOR_N vm = ls_max(vm, fabs(v[i]) * ewt[i]);
        return vm;
    }
    LS_MFN double norm_yh(int j) const
    {
        double vm = 0.0;
        LS_FOR_N vm = ls_max(vm, fabs(YH(j, i)) * ewt[i]);
        return vm;
    }
    LS_MFN double norm_acor() const { return norm(acor); }
    LS_MFN double norm_acor_minus_yh(int j)
    {
        LS_FOR_N savf[i] = acor[i] - YH(j, i);
        return norm(savf);
    }

    // ewset (LSODA.cpp:1368-1390) and inversion (:787-798); true if a weight is <= 0
    LS_MFN bool set_ewt(const LsTol& tol)
    {
        bool bad = false;
        LS_FOR_N {
            ewt[i] = tol.rt(i) * fabs(YH(1, i)) + tol.at(i);
            if (ewt[i] <= 0.0) bad = true;
            ewt[i] = ls_div(1.0, ewt[i]);
        }
        return bad;
    }

    // Pascal-triangle sweeps: prediction (LSODA.cpp:1141-1144) and retraction
    // (:1290-1295, :1909-1912).  For fixed j the inner loop runs i1 = j..nq upwards, so
    // yh[i1+1] is still the value of the previous pass when it is added: it is carried in
    // a register and every element costs one shared-memory load and one store.
    template <bool FORWARD>
    LS_MFN void pascal(int nq)
    {
        for (int j = nq; j >= 1; j--) {
            double a[N];
            LS_FOR_N a[i] = YH(j, i);
            for (int i1 = j; i1 <= nq; i1++) {
                LS_FOR_N {
                    const double b = YH(i1 + 1, i);
                    YH(i1, i) = FORWARD ? a[i] + b : a[i] - b;
                    a[i] = b;
                }
            }
        }
    }

    LS_MFN void scale_yh_now(int l, double rh)
    {
        double r = 1.0;
        for (int j = 2; j <= l; j++) {
            r *= rh;
            LS_FOR_N YH(j, i) = YH(j, i) * r;
        }
    }
#if !defined(LS_DEFERRED)
    LS_MFN void scale_yh(int l, double rh) { scale_yh_now(l, rh); }
#endif

    LS_MFN void y_from_yh1() { LS_FOR_N y[i] = YH(1, i); }
    LS_MFN void rhs_savf(double t, double* pp) { b200ode_user_rhs(t, y, savf, pp); }
    LS_MFN void acor_zero() { LS_FOR_N acor[i] = 0.0; }

    // functional iteration, LSODA.cpp:1792-1809; returns del
    LS_MFN double functional(double h, double el1)
    {
        LS_FOR_N {
            savf[i] = h * savf[i] - YH(2, i);
            y[i] = savf[i] - acor[i];
        }
        const double del = norm(y);
        LS_FOR_N {
            y[i] = YH(1, i) + el1 * savf[i];
            acor[i] = savf[i];
        }
        return del;
    }

    // chord iteration, LSODA.cpp:1816-1829; returns del
    LS_MFN double chord(double h, double el1)
    {
        LS_FOR_N y[i] = h * savf[i] - (YH(2, i) + acor[i]);
        dgesl(y);
        const double del = norm(y);
        LS_FOR_N {
            acor[i] += y[i];
            y[i] = YH(1, i) + el1 * acor[i];
        }
        return del;
    }

    // Nordsieck update of an accepted step, LSODA.cpp:1197-1202; el = elco[meth][nq]
    LS_MFN void accept(int l, const double* el)
    {
        for (int j = 1; j <= l; j++) {
            const double r = el[j];
            LS_FOR_N YH(j, i) = YH(j, i) + r * acor[i];
        }
    }
    LS_MFN void yh_from_acor(int j) { LS_FOR_N YH(j, i) = acor[i]; }
    LS_MFN void yh_from_acor_scaled(int j, double r) { LS_FOR_N YH(j, i) = acor[i] * r; }

    // recovery after three error-test failures, LSODA.cpp:1347-1353
    LS_MFN void refresh_yh2(double t, double h, double* pp)
    {
        LS_FOR_N y[i] = YH(1, i);
        b200ode_user_rhs(t, y, savf, pp);
        LS_FOR_N YH(2, i) = h * savf[i];
    }

    // intdy with k = 0, LSODA.cpp:1435-1453 (c = 1: the reference's products with it are exact)
    LS_MFN void intdy(int nq, double sfrac, double* out) const
    {
        double v[N];
        LS_FOR_N v[i] = 1.0 * YH(nq + 1, i);
        for (int j = nq - 1; j >= 0; j--) LS_FOR_N v[i] = 1.0 * YH(j + 1, i) + sfrac * v[i];
        LS_FOR_N out[i] = v[i];
    }
    LS_MFN void store_y(double* out) const { LS_FOR_N out[i] = y[i]; }

    // ---- first call (ls_begin)
    LS_MFN void begin(const double* u0, double* usol)
    {
        LS_FOR_N {
            y[i] = u0[i];
            usol[i] = y[i];
        }
        ipvt = 0;
    }
    LS_MFN void fill_row(double* out, const double* u0) const { LS_FOR_N out[i] = u0[i]; }
    LS_MFN void rhs_yh2(double t0, double* pp)
    {
        b200ode_user_rhs(t0, y, savf, pp);
        LS_FOR_N {
            YH(1, i) = y[i];
            YH(2, i) = savf[i];
        }
    }
    LS_MFN bool tol_negative(const LsTol& tol) const
    {
        bool neg = false;
        LS_FOR_N if (tol.rt(i) < 0.0 || tol.at(i) < 0.0) neg = true;
        return neg;
    }
    LS_MFN double tol_rtol_max(const LsTol& tol) const
    {
        double t = tol.rt(0);
        LS_FOR_N if (i > 0) t = ls_max(t, tol.rt(i));
        return t;
    }
    LS_MFN double tol_floor(const LsTol& tol, double t) const
    {
        LS_FOR_N {
            const double ayi = fabs(y[i]);
            if (ayi != 0.0) t = ls_max(t, tol.at(i) / ayi);
        }
        return t;
    }
    LS_MFN void scale_yh2(double h0) { LS_FOR_N YH(2, i) = YH(2, i) * h0; }

    // ---- linear algebra
    // dgefa, LSODA.cpp:173-231: row-oriented elimination; the pivot is searched along
    // row k and columns are swapped.  Returns nonzero when a zero pivot was met.
    LS_MFN int dgefa()
    {
        int info = 0;
        unsigned piv = 0;
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
            piv |= (unsigned)j << (4 * k);
            const double akj = WM(k, j);
            if (akj == 0.0) {
                info = k + 1;
                continue;
            }
            if (j != k) {
                WM(k, j) = WM(k, k);
                WM(k, k) = akj;
            }
            double t = ls_div_rare(-1.0, akj);
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
        piv |= (unsigned)(N - 1) << (4 * (N - 1));
        if (WM(N - 1, N - 1) == 0.0) info = N;
        ipvt = piv;
        return info;
    }

    // dgesl with job = 0, LSODA.cpp:119-144
    LS_MFN void dgesl(double (&b)[N]) const
    {
#pragma unroll
        for (int k = 0; k < N; k++) {
            double t = 0.0;
#pragma unroll
            for (int c = 0; c < k; c++) t += WM(k, c) * b[c];
            b[k] = ls_div(b[k] - t, WM(k, k));
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

#if defined(LS_DEFERRED)
    // prja for a requester (request_jac): the same operations as prja() below, with y,
    // savf = f(tn, y) and the weights read from the yh slots the requester left them in, so that
    // the serving thread -- which keeps its own trajectory in registers meanwhile -- needs
    // registers for one perturbed state and one right-hand side only.  acor of the requester is
    // not touched (the caller clears it afterwards, LSODA.cpp:1789).  Returns ierpj; ipvt = pivots.
    LS_MFN int prja_remote(double tn, double h, double hl0, double* pp, double& pdnorm)
    {
        double fac = 0.0;
        LS_FOR_N fac = ls_max(fac, fabs(YH(LS_JAC_SLOT_F, i)) * YH(LS_JAC_SLOT_W, i));
        double r0 = 1000.0 * fabs(h) * LS_ETA * ((double)N) * fac;
        if (r0 == 0.0) r0 = 1.0;
#pragma unroll 1
        for (int j = 0; j < N; j++) {
            const double yj = YH(LS_JAC_SLOT_Y, j), ewtj = YH(LS_JAC_SLOT_W, j);
            const double r = ls_max(LS_SQRTETA * fabs(yj), ls_div_rare(r0, ewtj));
            double yp[N], fp[N];
            LS_FOR_N yp[i] = (i == j) ? YH(LS_JAC_SLOT_Y, i) + r : YH(LS_JAC_SLOT_Y, i);
            fac = ls_div_rare(-hl0, r);
            b200ode_user_rhs(tn, yp, fp, pp);
            LS_FOR_N WM(i, j) = (fp[i] - YH(LS_JAC_SLOT_F, i)) * fac;
        }
        double an = 0.0;
#pragma unroll
        for (int r = 0; r < N; r++) {
            double sum = 0.0;
#pragma unroll
            for (int c = 0; c < N; c++) sum += ls_div_rare(fabs(WM(r, c)), YH(LS_JAC_SLOT_W, c));
            an = ls_max(an, sum * YH(LS_JAC_SLOT_W, r));
        }
        pdnorm = ls_div_rare(an, fabs(hl0));
        LS_FOR_N WM(i, i) = WM(i, i) + 1.0;
        return dgefa() != 0 ? 1 : 0;
    }
#endif

    // prja, LSODA.cpp:1633-1696 (miter == 2: dense finite-difference Jacobian), fnorm
    // (:1714-1736) and the LU factorisation.  Uses y, savf = f(tn, y), ewt; acor is
    // scratch (the caller clears it afterwards, :1789).  Returns ierpj.
    LS_MFN int prja(double tn, double h, double hl0, double* pp, double& pdnorm)
    {
        double fac = norm(savf);
        double r0 = 1000.0 * fabs(h) * LS_ETA * ((double)N) * fac;
        if (r0 == 0.0) r0 = 1.0;
        // one column per pass; a single (not unrolled) loop keeps one inlined copy of the RHS
#pragma unroll 1
        for (int j = 0; j < N; j++) {
            double yj = 0.0, ewtj = 0.0;
            LS_FOR_N {
                if (i == j) {
                    yj = y[i];
                    ewtj = ewt[i];
                }
            }
            const double r = ls_max(LS_SQRTETA * fabs(yj), ls_div_rare(r0, ewtj));
            double yp[N];
            LS_FOR_N yp[i] = (i == j) ? y[i] + r : y[i];
            fac = ls_div_rare(-hl0, r);
            b200ode_user_rhs(tn, yp, acor, pp);
            LS_FOR_N WM(i, j) = (acor[i] - savf[i]) * fac;
        }
        double an = 0.0;
#pragma unroll
        for (int r = 0; r < N; r++) {
            double sum = 0.0;
#pragma unroll
            for (int c = 0; c < N; c++) sum += ls_div_rare(fabs(WM(r, c)), ewt[c]);
            an = ls_max(an, sum * ewt[r]);
        }
        pdnorm = ls_div_rare(an, fabs(hl0));
        LS_FOR_N WM(i, i) = WM(i, i) + 1.0;
        return dgefa() != 0 ? 1 : 0;
    }
};

// One request of a deferred backend, served for the owner whose shared-memory columns `sv`
// views (LS_REQD_OF / LS_REQI_OF(o, f): field f of owner o's request, defined by the includer):
//   kind 0  Jacobian + LU                          (prja, dgefa: LSODA.cpp:1633-1696, :173-231)
//   kind 1  retraction, then the rescale if any    (:1290-1295 / :1909-1912, scaleh :1621-1627)
//   kind 2  rescale, then the output rows if any   (scaleh; intdy :1435-1453)
//   kind 3  output rows only
// pp: the owner's parameters (kind 0); teval / usol: t_eval and the owner's [nt, N] block (kinds 2, 3).
#if defined(LS_DEFERRED)
LS_FN void ls_serve_request(int kind, LsThreadVec& sv, int o, double* pp, const double* teval, double* usol)
{
    if (kind == 0) {
        double pdnorm;
        const int ierpj = sv.prja_remote(LS_REQD_OF(o, LS_RD_JTN), LS_REQD_OF(o, LS_RD_JH),
                                         LS_REQD_OF(o, LS_RD_JHL0), pp, pdnorm);
        LS_REQD_OF(o, LS_RD_JTN) = pdnorm;
        LS_REQI_OF(o, LS_RI_IERPJ) = ierpj;
        LS_REQI_OF(o, LS_RI_IPVT) = (int)sv.ipvt;
        return;
    }
    if (kind == 1) sv.template pascal<false>(LS_REQI_OF(o, LS_RI_RETRACT_NQ));
    const int l = LS_REQI_OF(o, LS_RI_SCALE_L);
    if (kind <= 2 && l > 0) sv.scale_yh_now(l, LS_REQD_OF(o, LS_RD_RH));
    const int nrows = LS_REQI_OF(o, LS_RI_NROWS);
    if (kind >= 2 && nrows > 0) {
        const int row0 = LS_REQI_OF(o, LS_RI_ROW0), nq = LS_REQI_OF(o, LS_RI_NQ);
        const double tn = LS_REQD_OF(o, LS_RD_TN), h = LS_REQD_OF(o, LS_RD_H);
        double* out = usol + (size_t)row0 * N;
        for (int k = 0; k < nrows; k++, out += N) sv.intdy(nq, ls_div(teval[row0 + k] - tn, h), out);
    }
}
// the queue a thread with pending operations `ops` belongs to (-1: none)
LS_FN int ls_request_kind(unsigned ops)
{
    return (ops & LS_OP_JAC) ? 0 : (ops & LS_OP_RETRACT) ? 1 : (ops & LS_OP_SCALE) ? 2 : (ops & LS_OP_ROWS) ? 3 : -1;
}
#endif

#endif
