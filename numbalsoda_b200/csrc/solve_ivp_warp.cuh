// solve_ivp for one trajectory owned by one warp: systems of 9 <= n <= 128 equations.
//
// Same algorithm as solve_ivp_core.cuh (the reference's dop853_solve_ivp_wrapper,
// src/dop853_solve_ivp.f90:211-434: dp86co with iout = 2, solout with the event test and the
// output rows, find_root / zeroin from src/brent_mod.f90:237-373) where the twelve stage
// vectors no longer fit one thread's registers.  Organised like dop853_warp.cuh:
//
//  * y, y1, k1..k10, the 8 n dense-output coefficients, two vectors of norm terms and one
//    interpolated state live in shared memory; lane i owns components i, i + 32, ...;
//  * sums the reference accumulates in index order (error norms, hinit, the stiffness
//    quotient) are taken sequentially by every lane on the lanes' terms: the reference's
//    rounding, no broadcast;
//  * the right-hand side and the events function are scalar user code: lane 0;
//  * step-size control, the event test and Brent's iteration run redundantly on identical
//    scalars in every lane (no divergence); each of Brent's function values is one parallel
//    interpolation (contd8 per owned component) + one events call;
//  * rows are written by the lanes that own the components (coalesced).
//
// Also compiles as host C++ (tests/solve_ivp_warp_host.cpp, one emulated lane).
#ifndef B200ODE_SOLVE_IVP_WARP_CUH
#define B200ODE_SOLVE_IVP_WARP_CUH

#include "b200ode_launch.h"
#include "b200ode_pow.cuh"
#include "b200ode_warp.cuh"
#include "dop853_tableau.cuh"

#if !defined(__CUDACC__)
#include <cmath>
#endif

#define IW_UROUND 2.220446049250313e-16  // epsilon(1.0d0), dop853_constants.f90:11
#define IW_SAFE 0.9                      // dop853_module.f90:55
#define IW_FAC1 0.333                    // :56
#define IW_FAC2 6.0                      // :59
#define IW_NMAX 100000                   // :46
#define IW_NSTIFF 1000                   // :47
#define IW_BRENT_TOL 1.0e-8              // dop853_solve_ivp.f90:96
#define IW_OWN(i) for (int i = WV_LANE; i < n; i += WV_NL)

// gfortran's MAX / MIN ignore a NaN operand, like fmax / fmin
WV_FN double iw_max(double a, double b) { return fmax(a, b); }
WV_FN double iw_min(double a, double b) { return fmin(a, b); }

// The user's events function, called from one place (see wv_user_rhs, b200ode_warp.cuh).
#if defined(__CUDA_ARCH__)
extern "C" __device__ long long b200ode_user_events(double t, double* u, double* ev, double* p);
__device__ __noinline__ void wv_user_events(double t, double* u, double* ev, double* p)
{
    b200ode_user_events(t, u, ev, p);
}
#else
#define wv_user_events(t, u, ev, p) b200ode_user_events(t, u, ev, p)
#endif

WV_FN double iw_ordered_sum(const double* t, int n)
{
    double s = 0.0;
    for (int i = 0; i < n; i++) s = s + t[i];
    return s;
}

struct IvpWarp {
    int n;
    double *cont, *ytmp, *ev_tmp;  // shared memory
    double xold, hout;
    double* pp;
};

// contd8 (module :832-872) of the owned components into out[]
WV_FN void iw_contd8(const IvpWarp& W, double x, double* out)
{
    const int n = W.n;
    const double* cont = W.cont;
    const double s = (x - W.xold) / W.hout;
    const double s1 = 1.0 - s;
    IW_OWN(i) {
        const double conpar = cont[4 * n + i] + s * (cont[5 * n + i] + s1 * (cont[6 * n + i] + s * cont[7 * n + i]));
        out[i] = cont[i] + s * (cont[n + i] + s1 * (cont[2 * n + i] + s * (cont[3 * n + i] + s1 * conpar)));
    }
}

// find_root's fcn, dop853_solve_ivp.f90:116-128
WV_FN double iw_event_at(const IvpWarp& W, int ind, double x)
{
    WV_SYNC();
    iw_contd8(W, x, W.ytmp);
    WV_SYNC();
    if (WV_LANE == 0) wv_user_events(x, W.ytmp, W.ev_tmp, W.pp);
    WV_SYNC();
    return W.ev_tmp[ind];
}

// zeroin, brent_mod.f90:237-373; every lane runs the same scalar iteration.  Returns iflag.
WV_FN int iw_zeroin(const IvpWarp& W, int ind, double ax, double bx, double* xzero)
{
    const double eps = IW_UROUND, tol = IW_BRENT_TOL;
    double a = ax, b = bx, c, d, e, fa, fb, fc, tol1, xm, p, q, r, s;
    fa = iw_event_at(W, ind, a);
    fb = iw_event_at(W, ind, b);
    if (fa == 0.0) {
        *xzero = a;
        return 0;
    }
    if (fb == 0.0) {
        *xzero = b;
        return 0;
    }
    if (!(fa * fb / fabs(fb) < 0.0)) return -1;
    c = a;
    fc = fa;
    d = b - a;
    e = d;
    for (;;) {
        if (fabs(fc) < fabs(fb)) {
            a = b;
            b = c;
            c = a;
            fa = fb;
            fb = fc;
            fc = fa;
        }
        tol1 = 2.0 * eps * fabs(b) + 0.5 * tol;
        xm = 0.5 * (c - b);
        if (fabs(xm) <= tol1 || fb == 0.0) break;
        if (fabs(e) >= tol1 && fabs(fa) > fabs(fb)) {
            s = fb / fa;
            if (a != c) {
                q = fa / fc;
                r = fb / fc;
                p = s * (2.0 * xm * q * (q - r) - (b - a) * (r - 1.0));
                q = (q - 1.0) * (r - 1.0) * (s - 1.0);
            } else {
                p = 2.0 * xm * s;
                q = 1.0 - s;
            }
            if (p <= 0.0) p = -p;
            else q = -q;
            s = e;
            e = d;
            if (2.0 * p >= 3.0 * xm * q - fabs(tol1 * q) || p >= fabs(0.5 * s * q)) {
                d = xm;
                e = d;
            } else {
                d = p / q;
            }
        } else {
            d = xm;
            e = d;
        }
        a = b;
        fa = fb;
        if (fabs(d) <= tol1) {
            if (xm <= 0.0) b = b - tol1;
            else b = b + tol1;
        } else {
            b = b + d;
        }
        fb = iw_event_at(W, ind, b);
        if (fb * fc / fabs(fc) > 0.0) {
            c = a;
            fc = fa;
            d = b - a;
            e = d;
        }
    }
    *xzero = b;
    return 0;
}

struct IvpWarpResult {
    int success, message, nt_out, rows_needed, event_found, ind_event, idid, nstep, naccpt, nrejct;
    double t_event;
};

// One trajectory.  `w`: B200ODE_IVPW_DOUBLES(n) doubles of (shared) memory.  Rows go to
// t_out[0..cap) / y_out[0..cap) (this trajectory's block); y_event: its [n] block.
WV_FN void ivp_warp_solve(int n, double* w, const B200IvpLaunch& L, const double* ts, const double* y0,
                          double* pp, double rtol, double atol, int te_mono, int cap, double* t_out,
                          double* y_out, double* y_event, IvpWarpResult& R)
{
    double *y = w, *y1 = y + n, *k1 = y1 + n, *k2 = k1 + n, *k3 = k2 + n, *k4 = k3 + n;
    double *k5 = k4 + n, *k6 = k5 + n, *k7 = k6 + n, *k8 = k7 + n, *k9 = k8 + n, *k10 = k9 + n;
    double* cont = k10 + n;     // 8 n
    double* t1 = cont + 8 * n;  // norm terms
    double* t2 = t1 + n;
    double* ytmp = t2 + n;
    double* ev_new = ytmp + n;  // B200ODE_IVP_NEVMAX each
    double* ev_old = ev_new + B200ODE_IVP_NEVMAX;
    double* ev_tmp = ev_old + B200ODE_IVP_NEVMAX;
    const int nt = L.nt, n_events = L.n_events;
    const bool t_eval_given = nt > 0;
    R.success = 1;
    R.message = B200ODE_IVP_MSG_END;
    R.nt_out = R.rows_needed = R.event_found = R.ind_event = R.idid = 0;
    R.nstep = R.naccpt = R.nrejct = 0;
    R.t_event = 0.0;
#define IW_RHS(t, u, du) wv_user_rhs_all(t, u, du, pp)  // (barriers inside)
#define IW_EVENTS(t, u, ev)                              \
    do {                                                 \
        WV_SYNC();                                       \
        if (WV_LANE == 0) wv_user_events(t, u, ev, pp);  \
        WV_SYNC();                                       \
    } while (0)
    double x = ts[0];
    const double xend = ts[1];
    const bool forward = xend >= x;
    IW_OWN(i) y_event[i] = 0.0;  // ODEResult.y_event without an event, driver_solve_ivp.py:195
    if (t_eval_given && !(te_mono & (forward ? 1 : 2))) {
        // '"t_eval" does not always increase or decrease', dop853_solve_ivp.f90:278-316
        R.success = 0;
        R.message = B200ODE_IVP_MSG_TEVAL;
        return;
    }
    IW_OWN(i) y[i] = y0[i];
    IW_OWN(i) for (int v = 0; v < 8; v++) cont[v * n + i] = 0.0;  // set_parameters, module :284
    double hmax = (L.max_step == 0.0) ? xend - x : L.max_step;    // module :431-435
    const double facc1 = 1.0 / IW_FAC1, facc2 = 1.0 / IW_FAC2;
    const double posneg = copysign(1.0, xend - x);
    IW_RHS(x, y, k1);
    hmax = fabs(hmax);
    double h = L.first_step;
    if (h == 0.0) {
        // hinit, module :745-823 (itol = 0, iord = 8)
        IW_OWN(i) {
            const double sk = atol + rtol * fabs(y[i]);
            double q = k1[i] / sk;
            t1[i] = q * q;
            q = y[i] / sk;
            t2[i] = q * q;
        }
        WV_SYNC();
        const double dnf = iw_ordered_sum(t1, n), dny = iw_ordered_sum(t2, n);
        double hh;
        if (dnf <= 1.0e-10 || dny <= 1.0e-10) hh = 1.0e-6;
        else hh = sqrt(dny / dnf) * 0.01;
        hh = iw_min(hh, hmax);
        hh = copysign(fabs(hh), posneg);
        WV_SYNC();
        IW_OWN(i) y1[i] = y[i] + hh * k1[i];
        IW_RHS(x + hh, y1, k2);
        IW_OWN(i) {
            const double sk = atol + rtol * fabs(y[i]);
            const double q = (k2[i] - k1[i]) / sk;
            t1[i] = q * q;
        }
        WV_SYNC();
        double der2 = iw_ordered_sum(t1, n);
        der2 = sqrt(der2) / hh;
        const double der12 = iw_max(fabs(der2), sqrt(dnf));
        double h1;
        if (der12 <= 1.0e-15) h1 = iw_max(1.0e-6, fabs(hh) * 1.0e-3);
        else h1 = b200ode_pow_eighth(0.01 / der12);
        hh = iw_min(iw_min(100.0 * fabs(hh), h1), hmax);
        h = copysign(fabs(hh), posneg);
        WV_SYNC();
    }
    IvpWarp W;
    W.n = n;
    W.cont = cont;
    W.ytmp = ytmp;
    W.ev_tmp = ev_tmp;
    W.xold = x;
    W.hout = 1.0;
    W.pp = pp;
    int row = 0;  // t_eval: rows written; otherwise rows saved
    // solout, nr = 1 (dop853_solve_ivp.f90:150-157, :204-208)
    if (n_events > 0) {
        IW_EVENTS(x, y, ev_new);
        if (WV_LANE == 0)
            for (int j = 0; j < n_events; j++) ev_old[j] = ev_new[j];
        WV_SYNC();
    }
    if (!t_eval_given) {
        if (row < cap) {
            if (WV_LANE == 0) t_out[row] = x;
            IW_OWN(i) y_out[(size_t)row * n + i] = y[i];
        }
        row++;
    }
    bool last = false, reject = false;
    int iasti = 0, nonsti = 0, idid = 0, aborted = 0;
    double hlamb = 0.0;
    for (;;) {
        if (R.nstep > IW_NMAX) {  // module :539
            idid = -2;
            break;
        }
        if (0.1 * fabs(h) <= fabs(x) * IW_UROUND) {  // module :546
            idid = -3;
            break;
        }
        if ((x + 1.01 * h - xend) * posneg > 0.0) {  // module :554-557
            h = xend - x;
            last = true;
        }
        R.nstep++;
        // the twelve stages, module :561-587
        IW_OWN(i) y1[i] = y[i] + h * DP_A2_1 * k1[i];
        IW_RHS(x + DP_C2 * h, y1, k2);
        IW_OWN(i) y1[i] = y[i] + h * (DP_A3_1 * k1[i] + DP_A3_2 * k2[i]);
        IW_RHS(x + DP_C3 * h, y1, k3);
        IW_OWN(i) y1[i] = y[i] + h * (DP_A4_1 * k1[i] + DP_A4_3 * k3[i]);
        IW_RHS(x + DP_C4 * h, y1, k4);
        IW_OWN(i) y1[i] = y[i] + h * (DP_A5_1 * k1[i] + DP_A5_3 * k3[i] + DP_A5_4 * k4[i]);
        IW_RHS(x + DP_C5 * h, y1, k5);
        IW_OWN(i) y1[i] = y[i] + h * (DP_A6_1 * k1[i] + DP_A6_4 * k4[i] + DP_A6_5 * k5[i]);
        IW_RHS(x + DP_C6 * h, y1, k6);
        IW_OWN(i) y1[i] = y[i] + h * (DP_A7_1 * k1[i] + DP_A7_4 * k4[i] + DP_A7_5 * k5[i] + DP_A7_6 * k6[i]);
        IW_RHS(x + DP_C7 * h, y1, k7);
        IW_OWN(i) y1[i] = y[i] + h * (DP_A8_1 * k1[i] + DP_A8_4 * k4[i] + DP_A8_5 * k5[i] + DP_A8_6 * k6[i] +
                                      DP_A8_7 * k7[i]);
        IW_RHS(x + DP_C8 * h, y1, k8);
        IW_OWN(i) y1[i] = y[i] + h * (DP_A9_1 * k1[i] + DP_A9_4 * k4[i] + DP_A9_5 * k5[i] + DP_A9_6 * k6[i] +
                                      DP_A9_7 * k7[i] + DP_A9_8 * k8[i]);
        IW_RHS(x + DP_C9 * h, y1, k9);
        IW_OWN(i) y1[i] = y[i] + h * (DP_A10_1 * k1[i] + DP_A10_4 * k4[i] + DP_A10_5 * k5[i] +
                                      DP_A10_6 * k6[i] + DP_A10_7 * k7[i] + DP_A10_8 * k8[i] +
                                      DP_A10_9 * k9[i]);
        IW_RHS(x + DP_C10 * h, y1, k10);
        IW_OWN(i) y1[i] = y[i] + h * (DP_A11_1 * k1[i] + DP_A11_4 * k4[i] + DP_A11_5 * k5[i] +
                                      DP_A11_6 * k6[i] + DP_A11_7 * k7[i] + DP_A11_8 * k8[i] +
                                      DP_A11_9 * k9[i] + DP_A11_10 * k10[i]);
        IW_RHS(x + DP_C11 * h, y1, k2);
        const double xph = x + h;
        IW_OWN(i) y1[i] = y[i] + h * (DP_A12_1 * k1[i] + DP_A12_4 * k4[i] + DP_A12_5 * k5[i] +
                                      DP_A12_6 * k6[i] + DP_A12_7 * k7[i] + DP_A12_8 * k8[i] +
                                      DP_A12_9 * k9[i] + DP_A12_10 * k10[i] + DP_A12_11 * k2[i]);
        IW_RHS(xph, y1, k3);
        // 8th-order solution and the terms of the error norms, module :588-612
        IW_OWN(i) {
            k4[i] = DP_B1 * k1[i] + DP_B6 * k6[i] + DP_B7 * k7[i] + DP_B8 * k8[i] + DP_B9 * k9[i] +
                    DP_B10 * k10[i] + DP_B11 * k2[i] + DP_B12 * k3[i];
            k5[i] = y[i] + h * k4[i];
            const double sk = atol + rtol * iw_max(fabs(y[i]), fabs(k5[i]));
            double erri = k4[i] - DP_BHH1 * k1[i] - DP_BHH2 * k9[i] - DP_BHH3 * k3[i];
            double q = erri / sk;
            t2[i] = q * q;
            erri = DP_ER1 * k1[i] + DP_ER6 * k6[i] + DP_ER7 * k7[i] + DP_ER8 * k8[i] + DP_ER9 * k9[i] +
                   DP_ER10 * k10[i] + DP_ER11 * k2[i] + DP_ER12 * k3[i];
            q = erri / sk;
            t1[i] = q * q;
        }
        WV_SYNC();
        double err = iw_ordered_sum(t1, n);
        const double err2 = iw_ordered_sum(t2, n);
        double deno = err + 0.01 * err2;
        if (deno <= 0.0) deno = 1.0;
        err = fabs(h) * err * sqrt(1.0 / ((double)n * deno));
        // step size controller, module :618-623 (beta = 0: facold**beta = 1)
        const double fac11 = b200ode_pow_eighth(err);
        double fac = fac11 / 1.0;
        fac = iw_max(facc2, iw_min(facc1, fac / IW_SAFE));
        double hnew = h / fac;
        WV_SYNC();
        if (err <= 1.0) {
            // ---- accepted, module :625-722
            R.naccpt++;
            IW_RHS(xph, k5, k4);
            // stiffness detection, module :631-655; iprint = 0: exit with idid = -4
            if (R.naccpt % IW_NSTIFF == 0 || iasti > 0) {
                IW_OWN(i) {
                    t1[i] = (k4[i] - k3[i]) * (k4[i] - k3[i]);
                    t2[i] = (k5[i] - y1[i]) * (k5[i] - y1[i]);
                }
                WV_SYNC();
                const double stnum = iw_ordered_sum(t1, n), stden = iw_ordered_sum(t2, n);
                WV_SYNC();
                if (stden > 0.0) hlamb = fabs(h) * sqrt(stnum / stden);
                if (hlamb > 6.1) {
                    nonsti = 0;
                    iasti++;
                    if (iasti == 15) {
                        idid = -4;
                        break;
                    }
                } else {
                    nonsti++;
                    if (nonsti == 6) iasti = 0;
                }
            }
            // dense output of every accepted step (iout = 2), module :657-705
            double *c0 = cont, *c1 = cont + n, *c2 = cont + 2 * n, *c3 = cont + 3 * n;
            double *c4 = cont + 4 * n, *c5 = cont + 5 * n, *c6 = cont + 6 * n, *c7 = cont + 7 * n;
            IW_OWN(i) {
                c0[i] = y[i];
                const double ydiff = k5[i] - y[i];
                c1[i] = ydiff;
                const double bspl = h * k1[i] - ydiff;
                c2[i] = bspl;
                c3[i] = ydiff - h * k4[i] - bspl;
                c4[i] = DP_D4_1 * k1[i] + DP_D4_6 * k6[i] + DP_D4_7 * k7[i] + DP_D4_8 * k8[i] +
                        DP_D4_9 * k9[i] + DP_D4_10 * k10[i] + DP_D4_11 * k2[i] + DP_D4_12 * k3[i];
                c5[i] = DP_D5_1 * k1[i] + DP_D5_6 * k6[i] + DP_D5_7 * k7[i] + DP_D5_8 * k8[i] +
                        DP_D5_9 * k9[i] + DP_D5_10 * k10[i] + DP_D5_11 * k2[i] + DP_D5_12 * k3[i];
                c6[i] = DP_D6_1 * k1[i] + DP_D6_6 * k6[i] + DP_D6_7 * k7[i] + DP_D6_8 * k8[i] +
                        DP_D6_9 * k9[i] + DP_D6_10 * k10[i] + DP_D6_11 * k2[i] + DP_D6_12 * k3[i];
                c7[i] = DP_D7_1 * k1[i] + DP_D7_6 * k6[i] + DP_D7_7 * k7[i] + DP_D7_8 * k8[i] +
                        DP_D7_9 * k9[i] + DP_D7_10 * k10[i] + DP_D7_11 * k2[i] + DP_D7_12 * k3[i];
            }
            // the next three function evaluations, module :682-691
            IW_OWN(i) y1[i] = y[i] + h * (DP_A14_1 * k1[i] + DP_A14_7 * k7[i] + DP_A14_8 * k8[i] +
                                          DP_A14_9 * k9[i] + DP_A14_10 * k10[i] + DP_A14_11 * k2[i] +
                                          DP_A14_12 * k3[i] + DP_A14_13 * k4[i]);
            IW_RHS(x + DP_C14 * h, y1, k10);
            IW_OWN(i) y1[i] = y[i] + h * (DP_A15_1 * k1[i] + DP_A15_6 * k6[i] + DP_A15_7 * k7[i] +
                                          DP_A15_8 * k8[i] + DP_A15_11 * k2[i] + DP_A15_12 * k3[i] +
                                          DP_A15_13 * k4[i] + DP_A15_14 * k10[i]);
            IW_RHS(x + DP_C15 * h, y1, k2);
            IW_OWN(i) y1[i] = y[i] + h * (DP_A16_1 * k1[i] + DP_A16_6 * k6[i] + DP_A16_7 * k7[i] +
                                          DP_A16_8 * k8[i] + DP_A16_9 * k9[i] + DP_A16_13 * k4[i] +
                                          DP_A16_14 * k10[i] + DP_A16_15 * k2[i]);
            IW_RHS(x + DP_C16 * h, y1, k3);
            IW_OWN(i) {
                c4[i] = h * (c4[i] + DP_D4_13 * k4[i] + DP_D4_14 * k10[i] + DP_D4_15 * k2[i] + DP_D4_16 * k3[i]);
                c5[i] = h * (c5[i] + DP_D5_13 * k4[i] + DP_D5_14 * k10[i] + DP_D5_15 * k2[i] + DP_D5_16 * k3[i]);
                c6[i] = h * (c6[i] + DP_D6_13 * k4[i] + DP_D6_14 * k10[i] + DP_D6_15 * k2[i] + DP_D6_16 * k3[i]);
                c7[i] = h * (c7[i] + DP_D7_13 * k4[i] + DP_D7_14 * k10[i] + DP_D7_15 * k2[i] + DP_D7_16 * k3[i]);
            }
            W.hout = h;
            IW_OWN(i) {
                k1[i] = k4[i];
                y[i] = k5[i];
            }
            W.xold = x;
            x = xph;
            WV_SYNC();

            // ---- solout, nr > 1 (dop853_solve_ivp.f90:149-207)
            if (n_events > 0) {
                IW_EVENTS(x, y, ev_new);
                for (int j = 0; j < n_events; j++) {
                    const double eo = ev_old[j], en = ev_new[j];
                    if ((eo < 0 && en > 0) || (eo > 0 && en < 0)) {
                        double xzero = 0.0;
                        if (iw_zeroin(W, j, W.xold, x, &xzero) != 0) {
                            // the reference prints 'brent failed' and stops the process
                            aborted = 2;
                            break;
                        }
                        R.event_found = 1;
                        R.ind_event = j;  // 0-based (dop853_solve_ivp.f90:432)
                        R.t_event = xzero;
                        WV_SYNC();
                        iw_contd8(W, xzero, y_event);
                        aborted = 1;
                        break;
                    }
                }
                WV_SYNC();
                if (WV_LANE == 0)
                    for (int j = 0; j < n_events; j++) ev_old[j] = ev_new[j];
                WV_SYNC();
            }
            if (aborted == 2) break;
            if (t_eval_given) {
                const double xx = R.event_found ? R.t_event : x;
                while (row < nt && (forward ? L.t_eval[row] <= xx : L.t_eval[row] >= xx)) {
                    iw_contd8(W, L.t_eval[row], y_out + (size_t)row * n);
                    row++;
                }
            } else if (!R.event_found) {
                if (row < cap) {
                    if (WV_LANE == 0) t_out[row] = x;
                    IW_OWN(i) y_out[(size_t)row * n + i] = y[i];
                }
                row++;
            }
            if (aborted) {  // irtrn < 0, module :711-714
                idid = 2;
                break;
            }
            if (last) {  // module :715-719
                idid = 1;
                break;
            }
            if (fabs(hnew) > hmax) hnew = posneg * hmax;
            if (reject) hnew = posneg * iw_min(fabs(hnew), fabs(h));
            reject = false;
        } else {
            // ---- rejected, module :724-728
            hnew = h / iw_min(facc1, fac11 / IW_SAFE);
            reject = true;
            if (R.naccpt >= 1) R.nrejct++;
            last = false;
        }
        h = hnew;
    }
    // the wrapper's epilogue, dop853_solve_ivp.f90:337-395
    R.idid = idid;
    R.rows_needed = t_eval_given ? nt : row;
    if (aborted == 2) {
        R.success = 0;
        R.message = B200ODE_IVP_MSG_BRENT;
        R.nt_out = row;
        R.event_found = 0;
    } else if (idid < 0) {
        R.success = 0;
        R.message = B200ODE_IVP_MSG_FAILED;
        R.nt_out = R.rows_needed;
    } else if (R.event_found) {
        R.message = B200ODE_IVP_MSG_EVENT;
        R.nt_out = row;  // rows are trimmed to d%nt - 1 (:356-373)
    } else {
        R.message = B200ODE_IVP_MSG_END;
        R.nt_out = R.rows_needed;
    }
    if (!t_eval_given && R.rows_needed > cap) {
        // not in the reference, whose arrays grow (:204-207): the caller's rows ran out
        R.success = 0;
        R.message = B200ODE_IVP_MSG_CAPACITY;
        R.nt_out = cap;
    }
#undef IW_RHS
#undef IW_EVENTS
}

#endif
