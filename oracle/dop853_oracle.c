/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * CPU oracle for the DOP853 half of the hot path: a plain-C restatement of what
 * the reference computes when Python calls numbalsoda.dop853():
 *
 *   dop853_wrapper + solout   /root/reference/src/dop853_c_interface.f90:40-130
 *   dop853 (argument set-up)  /root/reference/src/dop853_module.f90:316-456
 *   dp86co (the step loop)    /root/reference/src/dop853_module.f90:464-738
 *   hinit                     /root/reference/src/dop853_module.f90:745-823
 *   contd8                    /root/reference/src/dop853_module.f90:832-872
 *
 * The reference is Fortran and there is no Fortran compiler in this image, so it
 * cannot be built here (see oracle/README.md).  This file therefore restates the
 * arithmetic in the exact evaluation order gfortran uses for the source
 * expressions (left to right, parentheses honoured, no FMA contraction: build with
 * -ffp-contract=off) and is pinned against
 *   - the reference's own known-answer test (tests/test_numbalsoda.py:22-31),
 *   - Hairer's original Fortran DOP853 as shipped inside scipy
 *     (scipy.integrate.ode('dop853')) with the reference's controller constants,
 * by tests/test_oracle_dop853.py, and
 *   - the event times the reference itself printed in solve_ivp.ipynb, through
 *     solve_ivp_oracle.c (an independent restatement of the same dp86co that takes the same
 *     steps and produces the same rows bit for bit, rejected steps included):
 *     tests/test_oracle_solve_ivp.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this file's shared object.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "dop853_tableau.h"
#include "oracle.h"

#define UROUND 2.220446049250313e-16 /* epsilon(1.0d0), dop853_constants.f90:11 */

/* Test-only knob.  scipy.integrate.ode('dop853') (scipy >= 1.15 ships a C
 * translation of Hairer's code) shrinks a rejected step by the full factor fac1
 * instead of h/min(1/fac1, fac11/safe) (reference: module :725).  With the knob on,
 * this file reproduces scipy's accepted-step sequence exactly, which pins every
 * other part of the restatement (stages, error norms, controller, hinit, dense
 * output) against an independent implementation; see tests/test_oracle_dop853.py.
 * It is never set outside that test. */
static int g_scipy_reject_rule = 0;
void oracle_dop853_set_scipy_reject_rule(int on) { g_scipy_reject_rule = on; }

/* Test-only knob: emulate the controller arithmetic of the product's default
 * (throughput) link mode -- numbalsoda_b200/csrc/b200ode_fastmath.cuh, dop853_kernel.cu:
 * the squared error and a single-precision err**(-1/8) (libm's log2f / exp2f stand in for
 * the GPU's MUFU.LG2 / EX2: same accuracy class, not the same bits).  With it
 * tests/test_oracle_dop853.py measures, without a GPU, how far that mode moves the solution
 * from the reference arithmetic.  Off (the reference's arithmetic) everywhere else. */
static int g_fast_ctrl = 0;
void oracle_dop853_set_fast_ctrl(int on) { g_fast_ctrl = on; }

/* gfortran's MAX/MIN for reals (trans-intrinsic.c, gfc_conv_intrinsic_minmax):
 * m = a; if (b > m || isnan(m)) m = b;   i.e. a NaN operand is ignored.  This is
 * what lets the reference recover from an overflowing trial step (err = NaN ->
 * rejected with the maximal shrink factor). */
static inline double dmax(double a, double b) { return (b > a || isnan(a)) ? b : a; }
static inline double dmin(double a, double b) { return (b < a || isnan(a)) ? b : a; }
static inline double dsign(double a, double b) { return signbit(b) ? -fabs(a) : fabs(a); }

/* rtol_v / atol_v: per-component tolerances (itol = 1: module :603-612, :779-784, :799-803), or
 * NULL: the scalars apply (itol = 0, what the reference's wrapper always passes) */
#define RT(i) (rtol_v ? rtol_v[i] : rtol)
#define AT(i) (atol_v ? atol_v[i] : atol)

/* hinit, dop853_module.f90:745-823 (iord == 8). */
static double first_step(oracle_rhs_fn f, void *data, int n, double x, const double *y,
                         double posneg, const double *f0, double hmax, double atol,
                         double rtol, const double *rtol_v, const double *atol_v, double *f1, double *y1)
{
    double dnf = 0.0, dny = 0.0, sk, h, h1, der2, der12, q;
    int i;
    for (i = 0; i < n; i++) {
        sk = AT(i) + RT(i) * fabs(y[i]);
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
        sk = AT(i) + RT(i) * fabs(y[i]);
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

void oracle_dop853_v(oracle_rhs_fn f, int neq, const double *u0, void *data, int nt,
                     const double *teval, double *usol, double rtol, double atol,
                     const double *rtol_v, const double *atol_v, int mxstep,
                     int *success, oracle_dop853_stats *st)
{
    const int n = neq;
    oracle_dop853_stats local;
    double *w, *y, *y1, *k1, *k2, *k3, *k4, *k5, *k6, *k7, *k8, *k9, *k10, *cont;
    double x, xend, h, hmax, hnew, posneg, xph, xold = 0.0, hout = 1.0, xout;
    double facold, expo1, facc1, facc2, err, err2, erri, sk, deno, fac, fac11, ifac = 0.0;
    const double safe = 0.9, fac1 = 0.333, fac2 = 6.0, beta = 0.0; /* module :55-62 */
    int i, row, last = 0, reject = 0, event;

    if (!st)
        st = &local;
    memset(st, 0, sizeof(*st));
    *success = 1; /* c_interface :98 */

    /* set_parameters: nmax <= 0 is rejected, module :243-246 -> c_interface :106-109 */
    if (mxstep <= 0) {
        *success = 0;
        st->idid = -1;
        return;
    }

    w = (double *)malloc(sizeof(double) * (size_t)n * 20);
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
    cont = k10 + n; /* 8*n */

    for (i = 0; i < n; i++)
        y[i] = u0[i];
    x = teval[0];       /* c_interface :122 */
    xend = teval[nt - 1]; /* c_interface :123 */

    /* dop853(): hmax = xend - x, h = hinitial = 0  (module :431-438) */
    hmax = xend - x;
    h = 0.0;

    /* dp86co initialisations, module :502-534 */
    facold = 1.0e-4;
    expo1 = 1.0 / 8.0 - beta * 0.2;
    facc1 = 1.0 / fac1;
    facc2 = 1.0 / fac2;
    posneg = dsign(1.0, xend - x);
    f(x, y, k1, data);
    hmax = fabs(hmax);
    if (h == 0.0)
        h = first_step(f, data, n, x, y, posneg, k1, hmax, atol, rtol, rtol_v, atol_v, k2, y1);
    st->nfcn += 2;
    xold = x;
    /* first solout call (nr == 1), c_interface :56-58 */
    xout = teval[0];
    row = 0;

    for (;;) {
        if (st->nstep > mxstep) { /* module :539 */
            st->idid = -2;
            break;
        }
        if (0.1 * fabs(h) <= fabs(x) * UROUND) { /* module :546 */
            st->idid = -3;
            break;
        }
        if ((x + 1.01 * h - xend) * posneg > 0.0) { /* module :554-557 */
            h = xend - x;
            last = 1;
        }
        st->nstep++;
        /* the twelve stages, module :561-587 */
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
            y1[i] = y[i] + h * (DP_A7_1 * k1[i] + DP_A7_4 * k4[i] + DP_A7_5 * k5[i] +
                                DP_A7_6 * k6[i]);
        f(x + DP_C7 * h, y1, k7, data);
        for (i = 0; i < n; i++)
            y1[i] = y[i] + h * (DP_A8_1 * k1[i] + DP_A8_4 * k4[i] + DP_A8_5 * k5[i] +
                                DP_A8_6 * k6[i] + DP_A8_7 * k7[i]);
        f(x + DP_C8 * h, y1, k8, data);
        for (i = 0; i < n; i++)
            y1[i] = y[i] + h * (DP_A9_1 * k1[i] + DP_A9_4 * k4[i] + DP_A9_5 * k5[i] +
                                DP_A9_6 * k6[i] + DP_A9_7 * k7[i] + DP_A9_8 * k8[i]);
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
        st->nfcn += 11;
        for (i = 0; i < n; i++) {
            k4[i] = DP_B1 * k1[i] + DP_B6 * k6[i] + DP_B7 * k7[i] + DP_B8 * k8[i] +
                    DP_B9 * k9[i] + DP_B10 * k10[i] + DP_B11 * k2[i] + DP_B12 * k3[i];
            k5[i] = y[i] + h * k4[i];
        }
        /* error estimation, module :591-616 */
        err = 0.0;
        err2 = 0.0;
        for (i = 0; i < n; i++) {
            double q;
            sk = AT(i) + RT(i) * dmax(fabs(y[i]), fabs(k5[i]));
            erri = k4[i] - DP_BHH1 * k1[i] - DP_BHH2 * k9[i] - DP_BHH3 * k3[i];
            q = erri / sk;
            err2 = err2 + q * q;
            erri = DP_ER1 * k1[i] + DP_ER6 * k6[i] + DP_ER7 * k7[i] + DP_ER8 * k8[i] +
                   DP_ER9 * k9[i] + DP_ER10 * k10[i] + DP_ER11 * k2[i] + DP_ER12 * k3[i];
            q = erri / sk;
            err = err + q * q;
        }
        deno = err + 0.01 * err2;
        if (deno <= 0.0)
            deno = 1.0;
        if (g_fast_ctrl) {
            /* throughput-mode emulation: err holds the square of the reference's err */
            double e2 = (h * h) * (err * err) * (1.0 / ((double)n * deno)), m;
            int ex;
            err = e2;
            if (e2 != e2) {
                ifac = e2;
            } else {
                if (e2 > 1.0e300)
                    e2 = 1.0e300;
                if (e2 < 2.2250738585072014e-308) {
                    ifac = safe * (double)exp2f(1023.0f * 0.0625f);
                } else {
                    m = 2.0 * frexp(e2, &ex);
                    ifac = safe * (double)exp2f((log2f((float)m) + (float)(ex - 1)) * -0.0625f);
                }
            }
            fac11 = 0.0;
            hnew = h * dmin(fac2, dmax(fac1, ifac));
        } else {
        err = fabs(h) * err * sqrt(1.0 / ((double)n * deno));
        /* step size controller, module :618-623 */
        fac11 = pow(err, expo1);
        fac = fac11 / pow(facold, beta);
        fac = dmax(facc2, dmin(facc1, fac / safe));
        hnew = h / fac;
        }
        if (err <= 1.0) {
            /* accepted, module :625-722 */
            facold = dmax(err, 1.0e-4);
            st->naccpt++;
            f(xph, k5, k4, data);
            st->nfcn++;
            /* The stiffness detector (module :631-655) only prints: iprint is
             * output_unit, so idid = -4 is unreachable.  Nothing to restate. */
            event = (xout <= xph); /* iout == 3, module :657 */
            if (event) {
                double *c0 = cont, *c1 = cont + n, *c2 = cont + 2 * n, *c3 = cont + 3 * n;
                double *c4 = cont + 4 * n, *c5 = cont + 5 * n, *c6 = cont + 6 * n,
                       *c7 = cont + 7 * n;
                st->nevent++;
                for (i = 0; i < n; i++) {
                    double ydiff, bspl;
                    c0[i] = y[i];
                    ydiff = k5[i] - y[i];
                    c1[i] = ydiff;
                    bspl = h * k1[i] - ydiff;
                    c2[i] = bspl;
                    c3[i] = ydiff - h * k4[i] - bspl;
                    c4[i] = DP_D4_1 * k1[i] + DP_D4_6 * k6[i] + DP_D4_7 * k7[i] +
                            DP_D4_8 * k8[i] + DP_D4_9 * k9[i] + DP_D4_10 * k10[i] +
                            DP_D4_11 * k2[i] + DP_D4_12 * k3[i];
                    c5[i] = DP_D5_1 * k1[i] + DP_D5_6 * k6[i] + DP_D5_7 * k7[i] +
                            DP_D5_8 * k8[i] + DP_D5_9 * k9[i] + DP_D5_10 * k10[i] +
                            DP_D5_11 * k2[i] + DP_D5_12 * k3[i];
                    c6[i] = DP_D6_1 * k1[i] + DP_D6_6 * k6[i] + DP_D6_7 * k7[i] +
                            DP_D6_8 * k8[i] + DP_D6_9 * k9[i] + DP_D6_10 * k10[i] +
                            DP_D6_11 * k2[i] + DP_D6_12 * k3[i];
                    c7[i] = DP_D7_1 * k1[i] + DP_D7_6 * k6[i] + DP_D7_7 * k7[i] +
                            DP_D7_8 * k8[i] + DP_D7_9 * k9[i] + DP_D7_10 * k10[i] +
                            DP_D7_11 * k2[i] + DP_D7_12 * k3[i];
                }
                /* three more stages, module :682-691 */
                for (i = 0; i < n; i++)
                    y1[i] = y[i] + h * (DP_A14_1 * k1[i] + DP_A14_7 * k7[i] + DP_A14_8 * k8[i] +
                                        DP_A14_9 * k9[i] + DP_A14_10 * k10[i] +
                                        DP_A14_11 * k2[i] + DP_A14_12 * k3[i] +
                                        DP_A14_13 * k4[i]);
                f(x + DP_C14 * h, y1, k10, data);
                for (i = 0; i < n; i++)
                    y1[i] = y[i] + h * (DP_A15_1 * k1[i] + DP_A15_6 * k6[i] + DP_A15_7 * k7[i] +
                                        DP_A15_8 * k8[i] + DP_A15_11 * k2[i] +
                                        DP_A15_12 * k3[i] + DP_A15_13 * k4[i] +
                                        DP_A15_14 * k10[i]);
                f(x + DP_C15 * h, y1, k2, data);
                for (i = 0; i < n; i++)
                    y1[i] = y[i] + h * (DP_A16_1 * k1[i] + DP_A16_6 * k6[i] + DP_A16_7 * k7[i] +
                                        DP_A16_8 * k8[i] + DP_A16_9 * k9[i] +
                                        DP_A16_13 * k4[i] + DP_A16_14 * k10[i] +
                                        DP_A16_15 * k2[i]);
                f(x + DP_C16 * h, y1, k3, data);
                st->nfcn += 3;
                for (i = 0; i < n; i++) {
                    c4[i] = h * (c4[i] + DP_D4_13 * k4[i] + DP_D4_14 * k10[i] +
                                 DP_D4_15 * k2[i] + DP_D4_16 * k3[i]);
                    c5[i] = h * (c5[i] + DP_D5_13 * k4[i] + DP_D5_14 * k10[i] +
                                 DP_D5_15 * k2[i] + DP_D5_16 * k3[i]);
                    c6[i] = h * (c6[i] + DP_D6_13 * k4[i] + DP_D6_14 * k10[i] +
                                 DP_D6_15 * k2[i] + DP_D6_16 * k3[i]);
                    c7[i] = h * (c7[i] + DP_D7_13 * k4[i] + DP_D7_14 * k10[i] +
                                 DP_D7_15 * k2[i] + DP_D7_16 * k3[i]);
                }
                hout = h;
            }
            for (i = 0; i < n; i++) {
                k1[i] = k4[i];
                y[i] = k5[i];
            }
            xold = x;
            x = xph;
            if (event) {
                /* solout for nr > 1, c_interface :60-68, with contd8 (module :861-869) */
                while (row < nt && teval[row] <= x) {
                    const double s = (teval[row] - xold) / hout;
                    const double s1 = 1.0 - s;
                    for (i = 0; i < n; i++) {
                        double conpar = cont[4 * n + i] +
                                        s * (cont[5 * n + i] +
                                             s1 * (cont[6 * n + i] + s * cont[7 * n + i]));
                        usol[(size_t)row * n + i] =
                            cont[i] +
                            s * (cont[n + i] +
                                 s1 * (cont[2 * n + i] + s * (cont[3 * n + i] + s1 * conpar)));
                    }
                    row++;
                }
                if (row < nt)
                    xout = teval[row];
            }
            if (last) { /* module :715-719 */
                st->idid = 1;
                break;
            }
            if (fabs(hnew) > hmax)
                hnew = posneg * hmax;
            if (reject)
                hnew = posneg * dmin(fabs(hnew), fabs(h));
            reject = 0;
        } else {
            /* rejected, module :724-728 */
            if (g_fast_ctrl)
                hnew = h * dmax(fac1, ifac);
            else
                hnew = g_scipy_reject_rule ? h / facc1 : h / dmin(facc1, fac11 / safe);
            reject = 1;
            if (st->naccpt >= 1)
                st->nrejct++;
            last = 0;
        }
        h = hnew;
    }
    st->rows = row;
    st->hlast = h;
    st->xlast = x;
    if (st->idid < 0)
        *success = 0; /* c_interface :125-128 */
    free(w);
}
#undef RT
#undef AT

void oracle_dop853(oracle_rhs_fn f, int neq, const double *u0, void *data, int nt,
                   const double *teval, double *usol, double rtol, double atol, int mxstep,
                   int *success, oracle_dop853_stats *st)
{
    oracle_dop853_v(f, neq, u0, data, nt, teval, usol, rtol, atol, NULL, NULL, mxstep, success, st);
}
