// DOP853 for one trajectory owned by one warp: systems of 9 <= n <= 128 equations.
//
// Replaces, per trajectory, the reference's dop853_wrapper + solout
// (src/dop853_c_interface.f90:40-130) and dop853 / dp86co / hinit / contd8
// (src/dop853_module.f90:316-872) where the twelve stage vectors no longer fit one
// thread's registers (dop853_kernel.cu covers n <= 8).  Same arithmetic, in the
// reference's evaluation order; what changes is who does it:
//
//  * y, y1, k1..k10 and the 8 n dense-output coefficients live in shared memory;
//    lane i owns components i, i + 32, ...: the stage combinations, the new solution
//    and the interpolation are elementwise;
//  * the error norms are sums over components that the reference accumulates in index
//    order (module :594-602, hinit :768-775): every lane writes its components' terms,
//    then the sum is taken sequentially -- by all lanes redundantly, so that no
//    broadcast is needed -- which keeps the rounding of the reference;
//  * the right-hand side is one call of the user's scalar function: lane 0;
//  * all lanes take the step-size decisions redundantly on identical scalars: no
//    divergence; with every lane on the same trajectory, dense output is done in line
//    (the per-warp event queue of dop853_kernel.cu has nothing to batch here).
//
// Also compiles as host C++ (tests/dop853_warp_host.cpp, one emulated lane) so that the
// restated control flow is checked against the CPU oracle without a GPU.
#ifndef B200ODE_DOP853_WARP_CUH
#define B200ODE_DOP853_WARP_CUH

#include "b200ode_pow.cuh"
#include "b200ode_warp.cuh"
#include "dop853_tableau.cuh"

#if !defined(__CUDACC__)
#include <cmath>
#endif

#define DW_UROUND 2.220446049250313e-16  // epsilon(1.0d0), dop853_constants.f90:11
#define DW_SAFE 0.9                      // dop853_module.f90:55
#define DW_FAC1 0.333                    // :56
#define DW_FAC2 6.0                      // :59
#define DW_OWN(i) for (int i = WV_LANE; i < n; i += WV_NL)

// shared memory of one trajectory (B200ODE_DOP853W_DOUBLES(n)): 12 stage vectors, 8 n
// dense-output coefficients, two vectors of norm terms
#include "b200ode_launch.h"

// gfortran's MAX / MIN ignore a NaN operand, like fmax / fmin
WV_FN double dw_max(double a, double b) { return fmax(a, b); }
WV_FN double dw_min(double a, double b) { return fmin(a, b); }

struct Dop853WarpStats {
    int nstep, naccpt, nrejct, nevent, idid, rows;
};

// sum of t[0..n-1] in index order (every lane computes the same value)
WV_FN double dw_ordered_sum(const double* t, int n)
{
    double s = 0.0;
    for (int i = 0; i < n; i++) s = s + t[i];
    return s;
}

// dop853_wrapper for one trajectory.  `w`: B200ODE_DOP853W_DOUBLES(n) doubles of
// (shared) memory; usol: this trajectory's [nt, n] block.
// rtol_v / atol_v: per-component tolerances (the reference's itol = 1 branches, module
// :603-612, :779-803), or null: the scalars apply to every component
WV_FN void dop853_warp_solve(int n, double* w, const double* u0, double* pp, int nt,
                             const double* te, double* usol, double rtol, double atol, int mxstep,
                             Dop853WarpStats& st, const double* rtol_v = nullptr,
                             const double* atol_v = nullptr)
{
#define DW_RT(i) (rtol_v ? rtol_v[i] : rtol)
#define DW_AT(i) (atol_v ? atol_v[i] : atol)
    double *y = w, *y1 = y + n, *k1 = y1 + n, *k2 = k1 + n, *k3 = k2 + n, *k4 = k3 + n;
    double *k5 = k4 + n, *k6 = k5 + n, *k7 = k6 + n, *k8 = k7 + n, *k9 = k8 + n, *k10 = k9 + n;
    double* cont = k10 + n;     // 8 n
    double* t1 = cont + 8 * n;  // norm terms
    double* t2 = t1 + n;
    st.nstep = st.naccpt = st.nrejct = st.nevent = st.idid = st.rows = 0;
#define DW_RHS(t, u, du) wv_user_rhs_all(t, u, du, pp)  // (barriers inside)
    // set_parameters: nmax <= 0 is rejected, module :243-246 -> c_interface :106-109
    if (mxstep <= 0) {
        st.idid = -1;
        return;
    }
    DW_OWN(i) y[i] = u0[i];
    double x = te[0];
    const double xend = te[nt - 1];
    const double hmax = fabs(xend - x);  // module :431-432, :519
    const double facc1 = 1.0 / DW_FAC1, facc2 = 1.0 / DW_FAC2;
    const double posneg = copysign(1.0, xend - x);
    DW_RHS(x, y, k1);
    double h;
    {
        // hinit, module :745-823 (itol = 0, iord = 8)
        DW_OWN(i) {
            const double sk = DW_AT(i) + DW_RT(i) * fabs(y[i]);
            double q = k1[i] / sk;
            t1[i] = q * q;
            q = y[i] / sk;
            t2[i] = q * q;
        }
        WV_SYNC();
        const double dnf = dw_ordered_sum(t1, n), dny = dw_ordered_sum(t2, n);
        double hh;
        if (dnf <= 1.0e-10 || dny <= 1.0e-10) hh = 1.0e-6;
        else hh = sqrt(dny / dnf) * 0.01;
        hh = dw_min(hh, hmax);
        hh = copysign(fabs(hh), posneg);
        WV_SYNC();
        DW_OWN(i) y1[i] = y[i] + hh * k1[i];
        DW_RHS(x + hh, y1, k2);
        DW_OWN(i) {
            const double sk = DW_AT(i) + DW_RT(i) * fabs(y[i]);
            const double q = (k2[i] - k1[i]) / sk;
            t1[i] = q * q;
        }
        WV_SYNC();
        double der2 = dw_ordered_sum(t1, n);
        der2 = sqrt(der2) / hh;
        const double der12 = dw_max(fabs(der2), sqrt(dnf));
        double h1;
        if (der12 <= 1.0e-15) h1 = dw_max(1.0e-6, fabs(hh) * 1.0e-3);
        else h1 = b200ode_pow_eighth(0.01 / der12);
        hh = dw_min(dw_min(100.0 * fabs(hh), h1), hmax);
        h = copysign(fabs(hh), posneg);
        WV_SYNC();
    }
    // first solout call (c_interface :56-58): xout = teval(1), nt = 1
    double xout = te[0], xold = x, hout = 1.0;
    int row = 0;
    bool last = false, reject = false;
    for (;;) {
        if (st.nstep > mxstep) {  // module :539
            st.idid = -2;
            break;
        }
        if (0.1 * fabs(h) <= fabs(x) * DW_UROUND) {  // module :546
            st.idid = -3;
            break;
        }
        if ((x + 1.01 * h - xend) * posneg > 0.0) {  // module :554-557
            h = xend - x;
            last = true;
        }
        st.nstep++;
        // the twelve stages, module :561-587
        DW_OWN(i) y1[i] = y[i] + h * DP_A2_1 * k1[i];
        DW_RHS(x + DP_C2 * h, y1, k2);
        DW_OWN(i) y1[i] = y[i] + h * (DP_A3_1 * k1[i] + DP_A3_2 * k2[i]);
        DW_RHS(x + DP_C3 * h, y1, k3);
        DW_OWN(i) y1[i] = y[i] + h * (DP_A4_1 * k1[i] + DP_A4_3 * k3[i]);
        DW_RHS(x + DP_C4 * h, y1, k4);
        DW_OWN(i) y1[i] = y[i] + h * (DP_A5_1 * k1[i] + DP_A5_3 * k3[i] + DP_A5_4 * k4[i]);
        DW_RHS(x + DP_C5 * h, y1, k5);
        DW_OWN(i) y1[i] = y[i] + h * (DP_A6_1 * k1[i] + DP_A6_4 * k4[i] + DP_A6_5 * k5[i]);
        DW_RHS(x + DP_C6 * h, y1, k6);
        DW_OWN(i) y1[i] = y[i] + h * (DP_A7_1 * k1[i] + DP_A7_4 * k4[i] + DP_A7_5 * k5[i] +
                                      DP_A7_6 * k6[i]);
        DW_RHS(x + DP_C7 * h, y1, k7);
        DW_OWN(i) y1[i] = y[i] + h * (DP_A8_1 * k1[i] + DP_A8_4 * k4[i] + DP_A8_5 * k5[i] +
                                      DP_A8_6 * k6[i] + DP_A8_7 * k7[i]);
        DW_RHS(x + DP_C8 * h, y1, k8);
        DW_OWN(i) y1[i] = y[i] + h * (DP_A9_1 * k1[i] + DP_A9_4 * k4[i] + DP_A9_5 * k5[i] +
                                      DP_A9_6 * k6[i] + DP_A9_7 * k7[i] + DP_A9_8 * k8[i]);
        DW_RHS(x + DP_C9 * h, y1, k9);
        DW_OWN(i) y1[i] = y[i] + h * (DP_A10_1 * k1[i] + DP_A10_4 * k4[i] + DP_A10_5 * k5[i] +
                                      DP_A10_6 * k6[i] + DP_A10_7 * k7[i] + DP_A10_8 * k8[i] +
                                      DP_A10_9 * k9[i]);
        DW_RHS(x + DP_C10 * h, y1, k10);
        DW_OWN(i) y1[i] = y[i] + h * (DP_A11_1 * k1[i] + DP_A11_4 * k4[i] + DP_A11_5 * k5[i] +
                                      DP_A11_6 * k6[i] + DP_A11_7 * k7[i] + DP_A11_8 * k8[i] +
                                      DP_A11_9 * k9[i] + DP_A11_10 * k10[i]);
        DW_RHS(x + DP_C11 * h, y1, k2);
        const double xph = x + h;
        DW_OWN(i) y1[i] = y[i] + h * (DP_A12_1 * k1[i] + DP_A12_4 * k4[i] + DP_A12_5 * k5[i] +
                                      DP_A12_6 * k6[i] + DP_A12_7 * k7[i] + DP_A12_8 * k8[i] +
                                      DP_A12_9 * k9[i] + DP_A12_10 * k10[i] + DP_A12_11 * k2[i]);
        DW_RHS(xph, y1, k3);
        // 8th-order solution and the terms of the error norms, module :588-612
        DW_OWN(i) {
            k4[i] = DP_B1 * k1[i] + DP_B6 * k6[i] + DP_B7 * k7[i] + DP_B8 * k8[i] +
                    DP_B9 * k9[i] + DP_B10 * k10[i] + DP_B11 * k2[i] + DP_B12 * k3[i];
            k5[i] = y[i] + h * k4[i];
            const double sk = DW_AT(i) + DW_RT(i) * dw_max(fabs(y[i]), fabs(k5[i]));
            double erri = k4[i] - DP_BHH1 * k1[i] - DP_BHH2 * k9[i] - DP_BHH3 * k3[i];
            double q = erri / sk;
            t2[i] = q * q;
            erri = DP_ER1 * k1[i] + DP_ER6 * k6[i] + DP_ER7 * k7[i] + DP_ER8 * k8[i] +
                   DP_ER9 * k9[i] + DP_ER10 * k10[i] + DP_ER11 * k2[i] + DP_ER12 * k3[i];
            q = erri / sk;
            t1[i] = q * q;
        }
        WV_SYNC();
        double err = dw_ordered_sum(t1, n);
        const double err2 = dw_ordered_sum(t2, n);
        double deno = err + 0.01 * err2;
        if (deno <= 0.0) deno = 1.0;
        err = fabs(h) * err * sqrt(1.0 / ((double)n * deno));
        // step size controller, module :618-623 (beta = 0: facold**beta = 1)
        const double fac11 = b200ode_pow_eighth(err);
        double fac = fac11 / 1.0;
        fac = dw_max(facc2, dw_min(facc1, fac / DW_SAFE));
        double hnew = h / fac;
        WV_SYNC();
        if (err <= 1.0) {
            // accepted, module :625-722
            st.naccpt++;
            DW_RHS(xph, k5, k4);
            const bool event = xout <= xph;  // iout = 3, module :657
            if (event) {
                st.nevent++;
                double *c0 = cont, *c1 = cont + n, *c2 = cont + 2 * n, *c3 = cont + 3 * n;
                double *c4 = cont + 4 * n, *c5 = cont + 5 * n, *c6 = cont + 6 * n, *c7 = cont + 7 * n;
                DW_OWN(i) {
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
                DW_OWN(i) y1[i] = y[i] + h * (DP_A14_1 * k1[i] + DP_A14_7 * k7[i] + DP_A14_8 * k8[i] +
                                              DP_A14_9 * k9[i] + DP_A14_10 * k10[i] + DP_A14_11 * k2[i] +
                                              DP_A14_12 * k3[i] + DP_A14_13 * k4[i]);
                DW_RHS(x + DP_C14 * h, y1, k10);
                DW_OWN(i) y1[i] = y[i] + h * (DP_A15_1 * k1[i] + DP_A15_6 * k6[i] + DP_A15_7 * k7[i] +
                                              DP_A15_8 * k8[i] + DP_A15_11 * k2[i] + DP_A15_12 * k3[i] +
                                              DP_A15_13 * k4[i] + DP_A15_14 * k10[i]);
                DW_RHS(x + DP_C15 * h, y1, k2);
                DW_OWN(i) y1[i] = y[i] + h * (DP_A16_1 * k1[i] + DP_A16_6 * k6[i] + DP_A16_7 * k7[i] +
                                              DP_A16_8 * k8[i] + DP_A16_9 * k9[i] + DP_A16_13 * k4[i] +
                                              DP_A16_14 * k10[i] + DP_A16_15 * k2[i]);
                DW_RHS(x + DP_C16 * h, y1, k3);
                DW_OWN(i) {
                    c4[i] = h * (c4[i] + DP_D4_13 * k4[i] + DP_D4_14 * k10[i] + DP_D4_15 * k2[i] +
                                 DP_D4_16 * k3[i]);
                    c5[i] = h * (c5[i] + DP_D5_13 * k4[i] + DP_D5_14 * k10[i] + DP_D5_15 * k2[i] +
                                 DP_D5_16 * k3[i]);
                    c6[i] = h * (c6[i] + DP_D6_13 * k4[i] + DP_D6_14 * k10[i] + DP_D6_15 * k2[i] +
                                 DP_D6_16 * k3[i]);
                    c7[i] = h * (c7[i] + DP_D7_13 * k4[i] + DP_D7_14 * k10[i] + DP_D7_15 * k2[i] +
                                 DP_D7_16 * k3[i]);
                }
                hout = h;
            }
            DW_OWN(i) {
                k1[i] = k4[i];
                y[i] = k5[i];
            }
            xold = x;
            x = xph;
            if (event) {
                // solout for nr > 1 (c_interface :60-68) with contd8 (module :861-869)
                while (row < nt && te[row] <= x) {
                    const double s = (te[row] - xold) / hout;
                    const double s1 = 1.0 - s;
                    double* out = usol + (size_t)row * n;
                    DW_OWN(i) {
                        const double conpar = cont[4 * n + i] +
                                              s * (cont[5 * n + i] + s1 * (cont[6 * n + i] + s * cont[7 * n + i]));
                        out[i] = cont[i] + s * (cont[n + i] + s1 * (cont[2 * n + i] +
                                                                   s * (cont[3 * n + i] + s1 * conpar)));
                    }
                    row++;
                }
                if (row < nt) xout = te[row];
            }
            WV_SYNC();
            if (last) {  // module :715-719
                st.idid = 1;
                break;
            }
            if (fabs(hnew) > hmax) hnew = posneg * hmax;
            if (reject) hnew = posneg * dw_min(fabs(hnew), fabs(h));
            reject = false;
        } else {
            // rejected, module :724-728
            hnew = h / dw_min(facc1, fac11 / DW_SAFE);
            reject = true;
            if (st.naccpt >= 1) st.nrejct++;
            last = false;
        }
        h = hnew;
    }
    st.rows = row;
#undef DW_RHS
}

#endif
