/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's third entry point, solve_ivp (DOP853 with a
 * solver-chosen or user-given output grid, backward integration, first_step / max_step
 * and terminal events located by Brent's method on the dense interpolant):
 *
 *   numbalsoda/driver_solve_ivp.py:86-218      the Python driver and the ODEResult fields
 *   src/dop853_solve_ivp.f90:131-209           solout: event test, output rows
 *   src/dop853_solve_ivp.f90:78-129            find_root (tol = 1e-8 on the interpolant)
 *   src/dop853_solve_ivp.f90:211-395           dop853_solve_ivp_wrapper
 *   src/brent_mod.f90:237-373                  zeroin
 *   src/dop853_module.f90:464-738              dp86co with iout = 2 (dense output after
 *                                              every accepted step), iprint = 0 (the
 *                                              stiffness detector aborts: idid = -4),
 *                                              nmax = 100000, nstiff = 1000 (:46-47)
 *
 * The reference is Fortran and cannot be built in this image; this file is pinned by the
 * outputs the reference itself printed in /root/reference/solve_ivp.ipynb (cells 11, 26,
 * 27: twelve event times, one event state) and its tests (tests/test_numbalsoda.py:33-56):
 * tests/test_oracle_solve_ivp.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "dop853_tableau.h"
#include "oracle.h"

#define UROUND 2.220446049250313e-16 /* epsilon(1.0d0) */

/* test-only: scipy's (Hairer's original) rejection rule hnew = h/facc1 instead of the
 * reference's h/min(facc1, fac11/safe) (dop853_module.f90:725) -- used to show that the event
 * times the reference printed distinguish the two (tests/test_oracle_solve_ivp.py) */
static int g_ivp_scipy_reject_rule = 0;
void oracle_ivp_set_scipy_reject_rule(int on) { g_ivp_scipy_reject_rule = on; }

static inline double dmax(double a, double b) { return (b > a || isnan(a)) ? b : a; }
static inline double dmin(double a, double b) { return (b < a || isnan(a)) ? b : a; }
static inline double dsign(double a, double b) { return signbit(b) ? -fabs(a) : fabs(a); }

/* hinit, dop853_module.f90:745-823 (itol == 0, iord == 8) */
static double ivp_hinit(oracle_rhs_fn f, void *data, int n, double x, const double *y,
                        double posneg, const double *f0, double hmax, double atol, double rtol,
                        double *f1, double *y1)
{
    double dnf = 0.0, dny = 0.0, sk, h, h1, der2, der12, q;
    int i;
    for (i = 0; i < n; i++) {
        sk = atol + rtol * fabs(y[i]);
        q = f0[i] / sk;
        dnf = dnf + q * q;
        q = y[i] / sk;
        dny = dny + q * q;
    }
    if (dnf <= 1.0e-10 || dny <= 1.0e-10)
        h = 1.0e-6;
    else
        h = sqrt(dny / dnf) * 0.01;
    h = dmin(h, hmax);
    h = dsign(h, posneg);
    for (i = 0; i < n; i++)
        y1[i] = y[i] + h * f0[i];
    f(x + h, y1, f1, data);
    der2 = 0.0;
    for (i = 0; i < n; i++) {
        sk = atol + rtol * fabs(y[i]);
        q = (f1[i] - f0[i]) / sk;
        der2 = der2 + q * q;
    }
    der2 = sqrt(der2) / h;
    der12 = dmax(fabs(der2), sqrt(dnf));
    if (der12 <= 1.0e-15)
        h1 = dmax(1.0e-6, fabs(h) * 1.0e-3);
    else
        h1 = pow(0.01 / der12, 1.0 / 8.0);
    h = dmin(dmin(100.0 * fabs(h), h1), hmax);
    return dsign(h, posneg);
}

typedef struct {
    int n;
    const double *cont; /* 8 n */
    double xold, hout;
    oracle_rhs_fn events;
    void *data;
    int ind;
    double *ytmp, *etmp;
} ivp_dense;

/* contd8, dop853_module.f90:832-872, for every component */
static void ivp_contd8(const ivp_dense *d, double x, double *out)
{
    const int n = d->n;
    const double s = (x - d->xold) / d->hout;
    const double s1 = 1.0 - s;
    int i;
    for (i = 0; i < n; i++) {
        const double conpar = d->cont[4 * n + i] +
                              s * (d->cont[5 * n + i] + s1 * (d->cont[6 * n + i] + s * d->cont[7 * n + i]));
        out[i] = d->cont[i] +
                 s * (d->cont[n + i] + s1 * (d->cont[2 * n + i] + s * (d->cont[3 * n + i] + s1 * conpar)));
    }
}

/* find_root's fcn, dop853_solve_ivp.f90:116-128 */
static double ivp_event_at(ivp_dense *d, double x)
{
    ivp_contd8(d, x, d->ytmp);
    d->events(x, d->ytmp, d->etmp, d->data);
    return d->etmp[d->ind];
}

/* zeroin, brent_mod.f90:237-373.  Returns iflag. */
static int ivp_zeroin(ivp_dense *dd, double ax, double bx, double tol, double *xzero)
{
    const double eps = UROUND;
    double a = ax, b = bx, c, d, e, fa, fb, fc, tol1, xm, p, q, r, s;
    fa = ivp_event_at(dd, a);
    fb = ivp_event_at(dd, b);
    if (fa == 0.0) {
        *xzero = a;
        return 0;
    }
    if (fb == 0.0) {
        *xzero = b;
        return 0;
    }
    if (!(fa * fb / fabs(fb) < 0.0))
        return -1;
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
        if (fabs(xm) <= tol1 || fb == 0.0)
            break;
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
            if (p <= 0.0)
                p = -p;
            else
                q = -q;
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
            if (xm <= 0.0)
                b = b - tol1;
            else
                b = b + tol1;
        } else {
            b = b + d;
        }
        fb = ivp_event_at(dd, b);
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

void oracle_solve_ivp(oracle_rhs_fn f, const double *t_span, int neq, const double *y0, int nt,
                      const double *teval, int n_events, oracle_rhs_fn events, void *data,
                      double first_step, double max_step, double rtol, double atol, int cap,
                      double *t_out, double *y_out, double *y_event, oracle_ivp_result *res)
{
    const int n = neq;
    const int t_eval_given = nt > 0;
    const int events_given = n_events > 0 && events != NULL;
    const int forward = t_span[1] >= t_span[0];
    const double safe = 0.9, fac1 = 0.333, fac2 = 6.0;
    const int nmax = 100000, nstiff = 1000;
    double *w, *y, *y1, *k1, *k2, *k3, *k4, *k5, *k6, *k7, *k8, *k9, *k10, *cont, *ev_old, *ev_new;
    double x, xend, h, hmax, hnew, posneg, xph, xold, facc1, facc2, err, err2, deno, fac, fac11;
    double hlamb = 0.0;
    int i, nfcn = 0, nstep = 0, naccpt = 0, nrejct = 0, last = 0, reject = 0, iasti = 0, nonsti = 0;
    int idid = 0, row = 0 /* d%nt - 1 */, saved = 0 /* size(d%t) when no t_eval */, aborted = 0;
    ivp_dense dd;

    memset(res, 0, sizeof *res);
    res->success = 1;
    if (t_eval_given) {
        for (i = 1; i < nt; i++)
            if (forward ? teval[i] < teval[i - 1] : teval[i] > teval[i - 1]) {
                res->success = 0; /* dop853_solve_ivp.f90:278-303 */
                res->message = ORACLE_IVP_MSG_TEVAL;
                return;
            }
    }
    w = (double *)calloc((size_t)n * 22 + 2 * (size_t)(n_events > 0 ? n_events : 1), sizeof(double));
    y = w;
    y1 = y + n;
    k1 = y1 + n;
    k2 = k1 + n;
    k3 = k2 + n;
    k4 = k3 + n;
    k5 = k4 + n;
    k6 = k5 + n;
    k7 = k6 + n;
    k8 = k7 + n;
    k9 = k8 + n;
    k10 = k9 + n;
    cont = k10 + n; /* 8 n, zero: set_parameters module :284 */
    dd.ytmp = cont + 8 * n;
    dd.etmp = dd.ytmp + n;
    ev_old = dd.etmp + n; /* (n doubles of slack before the event arrays) */
    ev_new = ev_old + (n_events > 0 ? n_events : 1);
    /* etmp must hold n_events values: use a separate block when n_events > n */
    if (n_events > n)
        dd.etmp = (double *)calloc((size_t)n_events, sizeof(double));
    dd.n = n;
    dd.cont = cont;
    dd.events = events;
    dd.data = data;

    for (i = 0; i < n; i++)
        y[i] = y0[i];
    x = t_span[0];
    xend = t_span[1];
    hmax = (max_step == 0.0) ? xend - x : max_step; /* module :431-435 */
    h = first_step;
    facc1 = 1.0 / fac1;
    facc2 = 1.0 / fac2;
    posneg = dsign(1.0, xend - x);
    f(x, y, k1, data);
    hmax = fabs(hmax);
    if (h == 0.0)
        h = ivp_hinit(f, data, n, x, y, posneg, k1, hmax, atol, rtol, k2, y1);
    nfcn += 2;
    xold = x;
    dd.xold = xold;
    dd.hout = 1.0;

    /* ---- solout, nr == 1 (dop853_solve_ivp.f90:150-157, :204-208) */
    if (events_given) {
        events(x, y, ev_new, data);
        for (i = 0; i < n_events; i++)
            ev_old[i] = ev_new[i];
    }
    if (!t_eval_given) {
        if (saved < cap) {
            t_out[saved] = x;
            for (i = 0; i < n; i++)
                y_out[(size_t)saved * n + i] = y[i];
        }
        saved++;
    }

    for (;;) {
        if (nstep > nmax) {
            idid = -2;
            break;
        }
        if (0.1 * fabs(h) <= fabs(x) * UROUND) {
            idid = -3;
            break;
        }
        if ((x + 1.01 * h - xend) * posneg > 0.0) {
            h = xend - x;
            last = 1;
        }
        nstep++;
        for (i = 0; i < n; i++)
            y1[i] = y[i] + h * DP_A2_1 * k1[i];
        f(x + DP_C2 * h, y1, k2, data);
        for (i = 0; i < n; i++)
            y1[i] = y[i] + h * (DP_A3_1 * k1[i] + DP_A3_2 * k2[i]);
        f(x + DP_C3 * h, y1, k3, data);
        for (i = 0; i < n; i++)
            y1[i] = y[i] + h * (DP_A4_1 * k1[i] + DP_A4_3 * k3[i]);
        f(x + DP_C4 * h, y1, k4, data);
        for (i = 0; i < n; i++)
            y1[i] = y[i] + h * (DP_A5_1 * k1[i] + DP_A5_3 * k3[i] + DP_A5_4 * k4[i]);
        f(x + DP_C5 * h, y1, k5, data);
        for (i = 0; i < n; i++)
            y1[i] = y[i] + h * (DP_A6_1 * k1[i] + DP_A6_4 * k4[i] + DP_A6_5 * k5[i]);
        f(x + DP_C6 * h, y1, k6, data);
        for (i = 0; i < n; i++)
            y1[i] = y[i] + h * (DP_A7_1 * k1[i] + DP_A7_4 * k4[i] + DP_A7_5 * k5[i] + DP_A7_6 * k6[i]);
        f(x + DP_C7 * h, y1, k7, data);
        for (i = 0; i < n; i++)
            y1[i] = y[i] + h * (DP_A8_1 * k1[i] + DP_A8_4 * k4[i] + DP_A8_5 * k5[i] + DP_A8_6 * k6[i] +
                                DP_A8_7 * k7[i]);
        f(x + DP_C8 * h, y1, k8, data);
        for (i = 0; i < n; i++)
            y1[i] = y[i] + h * (DP_A9_1 * k1[i] + DP_A9_4 * k4[i] + DP_A9_5 * k5[i] + DP_A9_6 * k6[i] +
                                DP_A9_7 * k7[i] + DP_A9_8 * k8[i]);
        f(x + DP_C9 * h, y1, k9, data);
        for (i = 0; i < n; i++)
            y1[i] = y[i] + h * (DP_A10_1 * k1[i] + DP_A10_4 * k4[i] + DP_A10_5 * k5[i] +
                                DP_A10_6 * k6[i] + DP_A10_7 * k7[i] + DP_A10_8 * k8[i] +
                                DP_A10_9 * k9[i]);
        f(x + DP_C10 * h, y1, k10, data);
        for (i = 0; i < n; i++)
            y1[i] = y[i] + h * (DP_A11_1 * k1[i] + DP_A11_4 * k4[i] + DP_A11_5 * k5[i] +
                                DP_A11_6 * k6[i] + DP_A11_7 * k7[i] + DP_A11_8 * k8[i] +
                                DP_A11_9 * k9[i] + DP_A11_10 * k10[i]);
        f(x + DP_C11 * h, y1, k2, data);
        xph = x + h;
        for (i = 0; i < n; i++)
            y1[i] = y[i] + h * (DP_A12_1 * k1[i] + DP_A12_4 * k4[i] + DP_A12_5 * k5[i] +
                                DP_A12_6 * k6[i] + DP_A12_7 * k7[i] + DP_A12_8 * k8[i] +
                                DP_A12_9 * k9[i] + DP_A12_10 * k10[i] + DP_A12_11 * k2[i]);
        f(xph, y1, k3, data);
        nfcn += 11;
        for (i = 0; i < n; i++) {
            k4[i] = DP_B1 * k1[i] + DP_B6 * k6[i] + DP_B7 * k7[i] + DP_B8 * k8[i] + DP_B9 * k9[i] +
                    DP_B10 * k10[i] + DP_B11 * k2[i] + DP_B12 * k3[i];
            k5[i] = y[i] + h * k4[i];
        }
        err = 0.0;
        err2 = 0.0;
        for (i = 0; i < n; i++) {
            const double sk = atol + rtol * dmax(fabs(y[i]), fabs(k5[i]));
            double erri = k4[i] - DP_BHH1 * k1[i] - DP_BHH2 * k9[i] - DP_BHH3 * k3[i];
            double q = erri / sk;
            err2 = err2 + q * q;
            erri = DP_ER1 * k1[i] + DP_ER6 * k6[i] + DP_ER7 * k7[i] + DP_ER8 * k8[i] + DP_ER9 * k9[i] +
                   DP_ER10 * k10[i] + DP_ER11 * k2[i] + DP_ER12 * k3[i];
            q = erri / sk;
            err = err + q * q;
        }
        deno = err + 0.01 * err2;
        if (deno <= 0.0)
            deno = 1.0;
        err = fabs(h) * err * sqrt(1.0 / ((double)n * deno));
        fac11 = pow(err, 1.0 / 8.0);
        fac = fac11 / 1.0; /* facold**beta, beta = 0 */
        fac = dmax(facc2, dmin(facc1, fac / safe));
        hnew = h / fac;
        if (err <= 1.0) {
            double *c0 = cont, *c1 = cont + n, *c2 = cont + 2 * n, *c3 = cont + 3 * n;
            double *c4 = cont + 4 * n, *c5 = cont + 5 * n, *c6 = cont + 6 * n, *c7 = cont + 7 * n;
            naccpt++;
            f(xph, k5, k4, data);
            nfcn++;
            /* stiffness detection, module :631-655; iprint == 0: fail exit */
            if (naccpt % nstiff == 0 || iasti > 0) {
                double stnum = 0.0, stden = 0.0;
                for (i = 0; i < n; i++) {
                    stnum = stnum + (k4[i] - k3[i]) * (k4[i] - k3[i]);
                    stden = stden + (k5[i] - y1[i]) * (k5[i] - y1[i]);
                }
                if (stden > 0.0)
                    hlamb = fabs(h) * sqrt(stnum / stden);
                if (hlamb > 6.1) {
                    nonsti = 0;
                    iasti++;
                    if (iasti == 15) {
                        idid = -4;
                        break;
                    }
                } else {
                    nonsti++;
                    if (nonsti == 6)
                        iasti = 0;
                }
            }
            /* dense output of every accepted step (iout == 2), module :657-705 */
            for (i = 0; i < n; i++) {
                double ydiff, bspl;
                c0[i] = y[i];
                ydiff = k5[i] - y[i];
                c1[i] = ydiff;
                bspl = h * k1[i] - ydiff;
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
            for (i = 0; i < n; i++)
                y1[i] = y[i] + h * (DP_A14_1 * k1[i] + DP_A14_7 * k7[i] + DP_A14_8 * k8[i] +
                                    DP_A14_9 * k9[i] + DP_A14_10 * k10[i] + DP_A14_11 * k2[i] +
                                    DP_A14_12 * k3[i] + DP_A14_13 * k4[i]);
            f(x + DP_C14 * h, y1, k10, data);
            for (i = 0; i < n; i++)
                y1[i] = y[i] + h * (DP_A15_1 * k1[i] + DP_A15_6 * k6[i] + DP_A15_7 * k7[i] +
                                    DP_A15_8 * k8[i] + DP_A15_11 * k2[i] + DP_A15_12 * k3[i] +
                                    DP_A15_13 * k4[i] + DP_A15_14 * k10[i]);
            f(x + DP_C15 * h, y1, k2, data);
            for (i = 0; i < n; i++)
                y1[i] = y[i] + h * (DP_A16_1 * k1[i] + DP_A16_6 * k6[i] + DP_A16_7 * k7[i] +
                                    DP_A16_8 * k8[i] + DP_A16_9 * k9[i] + DP_A16_13 * k4[i] +
                                    DP_A16_14 * k10[i] + DP_A16_15 * k2[i]);
            f(x + DP_C16 * h, y1, k3, data);
            nfcn += 3;
            for (i = 0; i < n; i++) {
                c4[i] = h * (c4[i] + DP_D4_13 * k4[i] + DP_D4_14 * k10[i] + DP_D4_15 * k2[i] +
                             DP_D4_16 * k3[i]);
                c5[i] = h * (c5[i] + DP_D5_13 * k4[i] + DP_D5_14 * k10[i] + DP_D5_15 * k2[i] +
                             DP_D5_16 * k3[i]);
                c6[i] = h * (c6[i] + DP_D6_13 * k4[i] + DP_D6_14 * k10[i] + DP_D6_15 * k2[i] +
                             DP_D6_16 * k3[i]);
                c7[i] = h * (c7[i] + DP_D7_13 * k4[i] + DP_D7_14 * k10[i] + DP_D7_15 * k2[i] +
                             DP_D7_16 * k3[i]);
            }
            dd.hout = h;
            for (i = 0; i < n; i++) {
                k1[i] = k4[i];
                y[i] = k5[i];
            }
            xold = x;
            dd.xold = xold;
            x = xph;

            /* ---- solout, nr > 1 (dop853_solve_ivp.f90:149-207) */
            if (events_given) {
                events(x, y, ev_new, data);
                for (i = 0; i < n_events; i++) {
                    if ((ev_old[i] < 0 && ev_new[i] > 0) || (ev_old[i] > 0 && ev_new[i] < 0)) {
                        double xzero = 0.0;
                        dd.ind = i;
                        if (ivp_zeroin(&dd, xold, x, 1.0e-8, &xzero) != 0) {
                            /* the reference prints 'brent failed' and stops the process */
                            res->message = ORACLE_IVP_MSG_BRENT;
                            res->success = 0;
                            aborted = 2;
                            break;
                        }
                        res->event_found = 1;
                        res->t_event = xzero;
                        res->ind_event = i; /* already 0-based (:432) */
                        ivp_contd8(&dd, xzero, y_event);
                        aborted = 1;
                        break;
                    }
                }
                for (i = 0; i < n_events; i++)
                    ev_old[i] = ev_new[i];
            }
            if (aborted == 2)
                break;
            if (t_eval_given) {
                const double xx = res->event_found ? res->t_event : x;
                while (row < nt && (forward ? teval[row] <= xx : teval[row] >= xx)) {
                    ivp_contd8(&dd, teval[row], y_out + (size_t)row * n);
                    row++;
                }
            } else if (!res->event_found) {
                if (saved < cap) {
                    t_out[saved] = x;
                    for (i = 0; i < n; i++)
                        y_out[(size_t)saved * n + i] = y[i];
                }
                saved++;
            }
            if (aborted) {
                idid = 2;
                break;
            }
            if (last) {
                idid = 1;
                break;
            }
            if (fabs(hnew) > hmax)
                hnew = posneg * hmax;
            if (reject)
                hnew = posneg * dmin(fabs(hnew), fabs(h));
            reject = 0;
        } else {
            hnew = g_ivp_scipy_reject_rule ? h / facc1 : h / dmin(facc1, fac11 / safe);
            reject = 1;
            if (naccpt >= 1)
                nrejct++;
            last = 0;
        }
        h = hnew;
    }

    res->idid = idid;
    res->nstep = nstep;
    res->naccpt = naccpt;
    res->nrejct = nrejct;
    res->rows_needed = t_eval_given ? nt : saved;
    if (aborted == 2) {
        res->nt_out = t_eval_given ? row : (saved < cap ? saved : cap);
    } else if (idid < 0) {
        /* dop853_solve_ivp.f90:340-347: nfev stays 0, nt_out = size(d%t) */
        res->success = 0;
        res->message = ORACLE_IVP_MSG_FAILED;
        res->nt_out = t_eval_given ? nt : saved;
    } else {
        res->nfev = nfcn;
        if (res->event_found) {
            res->message = ORACLE_IVP_MSG_EVENT;
            res->nt_out = t_eval_given ? row : saved; /* trimmed to d%nt - 1, :356-373 */
        } else {
            res->message = ORACLE_IVP_MSG_END;
            res->nt_out = t_eval_given ? nt : saved;
        }
    }
    if (!t_eval_given && saved > cap) {
        res->success = 0;
        res->message = ORACLE_IVP_MSG_CAPACITY;
        res->nt_out = cap;
    }
    if (n_events > n)
        free(dd.etmp);
    free(w);
}
