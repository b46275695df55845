// Batched DOP853 for sm_100a: one trajectory per thread, everything in registers.
//
// Replaces, for ensembles, the reference's dop853_wrapper + solout
// (src/dop853_c_interface.f90:40-130) and dop853_class: dop853 / dp86co / hinit /
// contd8 (src/dop853_module.f90:316-872).  The arithmetic of every formula is kept
// in the reference's evaluation order, so that with FMA contraction disabled at
// link time (strict mode) the results are bit-identical to the CPU path; what is
// new is the organisation:
//
//  * thread-per-trajectory, state (y, k1..k10, y1 = 12n doubles) and the dense
//    output coefficients (8n) live in registers; the user's right-hand side is
//    inlined by link-time optimisation (see b200ode_device.cuh);
//  * a persistent grid pulls trajectories from a global work counter: a lane that
//    finishes its trajectory immediately starts the next one inside the same
//    step loop, so the adaptive-step imbalance between trajectories costs idle
//    lanes only at the very end of the launch;
//  * the t_eval rows of a trajectory are consecutive in memory ([B,nt,n] layout of
//    the reference's usol), so each lane streams its rows sequentially and the
//    L2 merges them into full-sector HBM writes.
#include "b200ode_device.cuh"
#include "b200ode_pow.cuh"
#include "dop853_tableau.cuh"

#define N B200_NEQ
#define FOR_N _Pragma("unroll") for (int i = 0; i < N; i++)

#define DP_UROUND 2.220446049250313e-16  // epsilon(1.0d0), dop853_constants.f90:11
#define DP_SAFE 0.9                      // dop853_module.f90:55
#define DP_FAC1 0.333                    // :56
#define DP_FAC2 6.0                      // :59

// gfortran's MAX/MIN ignore a NaN operand, exactly like fmax/fmin (DMNMX).
__device__ __forceinline__ double fmax_(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ double fmin_(double a, double b) { return fmin(a, b); }

struct Dop853Rhs {
    double* p;
    __device__ __forceinline__ void operator()(double t, double* u, double* du) const {
        b200ode_user_rhs(t, u, du, p);
    }
};

// STAGED: the first L.np doubles of the trajectory's data record are copied into
// registers once and the right-hand side reads them there; otherwise it receives
// the record's global address unchanged (arbitrary record layout).
template <bool STAGED>
__device__ __forceinline__ void dop853_ensemble(const B200OdeLaunch& L)
{
    double y[N], y1[N], k1[N], k2[N], k3[N], k4[N], k5[N], k6[N], k7[N], k8[N], k9[N], k10[N];
    double pl[STAGED ? B200ODE_NPMAX : 1];
    Dop853Rhs f;
    double x = 0.0, xend = 0.0, h = 0.0, hmax = 0.0, posneg = 1.0, xout = 0.0;
    int nstep = 0, nfcn = 0, naccpt = 0, nrejct = 0, nevent = 0, row = 0;
    bool last = false, reject = false;
    long long b = -1;
    double* out = nullptr;

    const double rtol = L.rtol, atol = L.atol;
    const double facc1 = 1.0 / DP_FAC1;     // module :505
    const double facc2 = 1.0 / DP_FAC2;     // module :506
    const int nt = L.nt;
    const double* __restrict__ te = L.t_eval;

    for (;;) {
        if (b < 0) {
            // ---- take the next trajectory: dop853_wrapper (c_interface :98-124),
            //      dop853 (module :396-438), dp86co prologue (module :502-534)
            b = b200ode_fetch(L.next);
            if (b >= L.B) break;
            const double* u0 = L.u0 + b * L.u0_stride;
            FOR_N y[i] = u0[i];
            const double* pg = reinterpret_cast<const double*>(L.data + b * L.data_stride);
            if (STAGED) {
#pragma unroll
                for (int j = 0; j < B200ODE_NPMAX; j++) pl[j] = (j < L.np) ? pg[j] : 0.0;
                f.p = pl;
            } else {
                f.p = const_cast<double*>(pg);
            }
            out = L.usol + (size_t)b * (size_t)nt * N;
            x = te[0];
            xend = te[nt - 1];
            hmax = xend - x;
            posneg = copysign(1.0, xend - x);
            last = false;
            reject = false;
            nstep = naccpt = nrejct = nevent = 0;
            f(x, y, k1);
            hmax = fabs(hmax);
            {
                // hinit, module :745-823 (itol = 0, iord = 8)
                double dnf = 0.0, dny = 0.0;
                FOR_N {
                    double sk = atol + rtol * fabs(y[i]);
                    double q = k1[i] / sk;
                    dnf = dnf + q * q;
                    q = y[i] / sk;
                    dny = dny + q * q;
                }
                double hh;
                if (dnf <= 1.0e-10 || dny <= 1.0e-10) hh = 1.0e-6;
                else hh = sqrt(dny / dnf) * 0.01;
                hh = fmin_(hh, hmax);
                hh = copysign(fabs(hh), posneg);
                FOR_N y1[i] = y[i] + hh * k1[i];
                f(x + hh, y1, k2);
                double der2 = 0.0;
                FOR_N {
                    double sk = atol + rtol * fabs(y[i]);
                    double q = (k2[i] - k1[i]) / sk;
                    der2 = der2 + q * q;
                }
                der2 = sqrt(der2) / hh;
                double der12 = fmax_(fabs(der2), sqrt(dnf));
                double h1;
                if (der12 <= 1.0e-15) h1 = fmax_(1.0e-6, fabs(hh) * 1.0e-3);
                else h1 = b200ode_pow_eighth(0.01 / der12);
                hh = fmin_(fmin_(100.0 * fabs(hh), h1), hmax);
                h = copysign(fabs(hh), posneg);
            }
            nfcn = 2;
            xout = te[0];  // first solout call, c_interface :56-58
            row = 0;
        }

        // ---- one attempted step, dp86co module :537-731
        int idid = 0;
        if (nstep > L.mxstep) {
            idid = -2;  // module :539
        } else if (0.1 * fabs(h) <= fabs(x) * DP_UROUND) {
            idid = -3;  // module :546
        } else {
            if ((x + 1.01 * h - xend) * posneg > 0.0) {  // module :554-557
                h = xend - x;
                last = true;
            }
            nstep++;
            // the twelve stages, module :561-587
            FOR_N y1[i] = y[i] + h * DP_A2_1 * k1[i];
            f(x + DP_C2 * h, y1, k2);
            FOR_N y1[i] = y[i] + h * (DP_A3_1 * k1[i] + DP_A3_2 * k2[i]);
            f(x + DP_C3 * h, y1, k3);
            FOR_N y1[i] = y[i] + h * (DP_A4_1 * k1[i] + DP_A4_3 * k3[i]);
            f(x + DP_C4 * h, y1, k4);
            FOR_N y1[i] = y[i] + h * (DP_A5_1 * k1[i] + DP_A5_3 * k3[i] + DP_A5_4 * k4[i]);
            f(x + DP_C5 * h, y1, k5);
            FOR_N y1[i] = y[i] + h * (DP_A6_1 * k1[i] + DP_A6_4 * k4[i] + DP_A6_5 * k5[i]);
            f(x + DP_C6 * h, y1, k6);
            FOR_N y1[i] = y[i] + h * (DP_A7_1 * k1[i] + DP_A7_4 * k4[i] + DP_A7_5 * k5[i] +
                                      DP_A7_6 * k6[i]);
            f(x + DP_C7 * h, y1, k7);
            FOR_N y1[i] = y[i] + h * (DP_A8_1 * k1[i] + DP_A8_4 * k4[i] + DP_A8_5 * k5[i] +
                                      DP_A8_6 * k6[i] + DP_A8_7 * k7[i]);
            f(x + DP_C8 * h, y1, k8);
            FOR_N y1[i] = y[i] + h * (DP_A9_1 * k1[i] + DP_A9_4 * k4[i] + DP_A9_5 * k5[i] +
                                      DP_A9_6 * k6[i] + DP_A9_7 * k7[i] + DP_A9_8 * k8[i]);
            f(x + DP_C9 * h, y1, k9);
            FOR_N y1[i] = y[i] + h * (DP_A10_1 * k1[i] + DP_A10_4 * k4[i] + DP_A10_5 * k5[i] +
                                      DP_A10_6 * k6[i] + DP_A10_7 * k7[i] + DP_A10_8 * k8[i] +
                                      DP_A10_9 * k9[i]);
            f(x + DP_C10 * h, y1, k10);
            FOR_N y1[i] = y[i] + h * (DP_A11_1 * k1[i] + DP_A11_4 * k4[i] + DP_A11_5 * k5[i] +
                                      DP_A11_6 * k6[i] + DP_A11_7 * k7[i] + DP_A11_8 * k8[i] +
                                      DP_A11_9 * k9[i] + DP_A11_10 * k10[i]);
            f(x + DP_C11 * h, y1, k2);
            const double xph = x + h;
            FOR_N y1[i] = y[i] + h * (DP_A12_1 * k1[i] + DP_A12_4 * k4[i] + DP_A12_5 * k5[i] +
                                      DP_A12_6 * k6[i] + DP_A12_7 * k7[i] + DP_A12_8 * k8[i] +
                                      DP_A12_9 * k9[i] + DP_A12_10 * k10[i] +
                                      DP_A12_11 * k2[i]);
            f(xph, y1, k3);
            nfcn += 11;
            FOR_N {
                k4[i] = DP_B1 * k1[i] + DP_B6 * k6[i] + DP_B7 * k7[i] + DP_B8 * k8[i] +
                        DP_B9 * k9[i] + DP_B10 * k10[i] + DP_B11 * k2[i] + DP_B12 * k3[i];
                k5[i] = y[i] + h * k4[i];
            }
            // error estimation, module :591-616
            double err = 0.0, err2 = 0.0;
            FOR_N {
                double sk = atol + rtol * fmax_(fabs(y[i]), fabs(k5[i]));
                double erri = k4[i] - DP_BHH1 * k1[i] - DP_BHH2 * k9[i] - DP_BHH3 * k3[i];
                double q = erri / sk;
                err2 = err2 + q * q;
                erri = DP_ER1 * k1[i] + DP_ER6 * k6[i] + DP_ER7 * k7[i] + DP_ER8 * k8[i] +
                       DP_ER9 * k9[i] + DP_ER10 * k10[i] + DP_ER11 * k2[i] + DP_ER12 * k3[i];
                q = erri / sk;
                err = err + q * q;
            }
            double deno = err + 0.01 * err2;
            if (deno <= 0.0) deno = 1.0;
            err = fabs(h) * err * sqrt(1.0 / ((double)N * deno));
            // step size controller, module :618-623 (facold**beta == 1 for beta = 0)
            const double fac11 = b200ode_pow_eighth(err);
            double fac = fac11 / 1.0;
            fac = fmax_(facc2, fmin_(facc1, fac / DP_SAFE));
            double hnew = h / fac;
            if (err <= 1.0) {
                // ---- accepted, module :625-722
                // (facold = max(err, 1e-4), module :626, is dead: beta = 0)
                naccpt++;
                f(xph, k5, k4);
                nfcn++;
                // (the stiffness detector, module :631-655, only prints: iprint /= 0)
                const bool event = (xout <= xph);  // iout == 3, module :657
                if (event) {
                    double c0[N], c1[N], c2[N], c3[N], c4[N], c5[N], c6[N], c7[N];
                    nevent++;
                    FOR_N {
                        c0[i] = y[i];
                        double ydiff = k5[i] - y[i];
                        c1[i] = ydiff;
                        double bspl = h * k1[i] - ydiff;
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
                    // the next three function evaluations, module :682-691
                    FOR_N y1[i] = y[i] + h * (DP_A14_1 * k1[i] + DP_A14_7 * k7[i] +
                                              DP_A14_8 * k8[i] + DP_A14_9 * k9[i] +
                                              DP_A14_10 * k10[i] + DP_A14_11 * k2[i] +
                                              DP_A14_12 * k3[i] + DP_A14_13 * k4[i]);
                    f(x + DP_C14 * h, y1, k10);
                    FOR_N y1[i] = y[i] + h * (DP_A15_1 * k1[i] + DP_A15_6 * k6[i] +
                                              DP_A15_7 * k7[i] + DP_A15_8 * k8[i] +
                                              DP_A15_11 * k2[i] + DP_A15_12 * k3[i] +
                                              DP_A15_13 * k4[i] + DP_A15_14 * k10[i]);
                    f(x + DP_C15 * h, y1, k2);
                    FOR_N y1[i] = y[i] + h * (DP_A16_1 * k1[i] + DP_A16_6 * k6[i] +
                                              DP_A16_7 * k7[i] + DP_A16_8 * k8[i] +
                                              DP_A16_9 * k9[i] + DP_A16_13 * k4[i] +
                                              DP_A16_14 * k10[i] + DP_A16_15 * k2[i]);
                    f(x + DP_C16 * h, y1, k3);
                    nfcn += 3;
                    FOR_N {
                        c4[i] = h * (c4[i] + DP_D4_13 * k4[i] + DP_D4_14 * k10[i] +
                                     DP_D4_15 * k2[i] + DP_D4_16 * k3[i]);
                        c5[i] = h * (c5[i] + DP_D5_13 * k4[i] + DP_D5_14 * k10[i] +
                                     DP_D5_15 * k2[i] + DP_D5_16 * k3[i]);
                        c6[i] = h * (c6[i] + DP_D6_13 * k4[i] + DP_D6_14 * k10[i] +
                                     DP_D6_15 * k2[i] + DP_D6_16 * k3[i]);
                        c7[i] = h * (c7[i] + DP_D7_13 * k4[i] + DP_D7_14 * k10[i] +
                                     DP_D7_15 * k2[i] + DP_D7_16 * k3[i]);
                    }
                    // solout (c_interface :60-68) with contd8 (module :861-869):
                    // xold = x, hout = h, new x = xph
                    while (row < nt) {
                        const double tq = te[row];
                        if (!(tq <= xph)) break;
                        const double s = (tq - x) / h;
                        const double s1 = 1.0 - s;
                        double* o = out + (size_t)row * N;
                        FOR_N {
                            double conpar = c4[i] + s * (c5[i] + s1 * (c6[i] + s * c7[i]));
                            o[i] = c0[i] + s * (c1[i] + s1 * (c2[i] + s * (c3[i] + s1 * conpar)));
                        }
                        row++;
                    }
                    if (row < nt) xout = te[row];
                }
                FOR_N {
                    k1[i] = k4[i];
                    y[i] = k5[i];
                }
                x = xph;
                if (last) {
                    idid = 1;  // module :715-719
                } else {
                    if (fabs(hnew) > hmax) hnew = posneg * hmax;
                    if (reject) hnew = posneg * fmin_(fabs(hnew), fabs(h));
                    reject = false;
                }
            } else {
                // ---- rejected, module :724-728
                hnew = h / fmin_(facc1, fac11 / DP_SAFE);
                reject = true;
                if (naccpt >= 1) nrejct++;
                last = false;
            }
            h = hnew;
        }

        if (idid != 0) {
            // dop853_wrapper epilogue, c_interface :125-128
            L.success[b] = (idid < 0) ? 0 : 1;
            if (L.counters) {
                int* c = L.counters + b * B200ODE_NCOUNTERS;
                c[0] = nstep;
                c[1] = nfcn;
                c[2] = naccpt;
                c[3] = nrejct;
                c[4] = nevent;
                c[5] = idid;
                c[6] = row;
                c[7] = 0;
            }
            b = -1;
        }
    }
}

extern "C" __global__ void __launch_bounds__(128) b200ode_dop853_kernel(const B200OdeLaunch L)
{
    dop853_ensemble<true>(L);
}

extern "C" __global__ void __launch_bounds__(128) b200ode_dop853_kernel_ptr(const B200OdeLaunch L)
{
    dop853_ensemble<false>(L);
}
