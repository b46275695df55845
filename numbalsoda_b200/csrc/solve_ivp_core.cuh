// solve_ivp for one trajectory owned by one thread: systems of n <= 8 equations.
//
// Replaces, per trajectory, the reference's third native entry point
//   dop853_solve_ivp_wrapper + dop853_solve_ivp_result   src/dop853_solve_ivp.f90:211-434
//   solout (event test, output rows)                      src/dop853_solve_ivp.f90:131-209
//   find_root / zeroin (Brent on the dense interpolant)   src/dop853_solve_ivp.f90:78-129,
//                                                         src/brent_mod.f90:237-373
// on top of dop853 / dp86co / hinit / contd8 (src/dop853_module.f90:316-872) run with
// iout = 2 (dense output prepared after *every* accepted step), iprint = 0 (the stiffness
// detector aborts with idid = -4), nmax = 100000, nstiff = 1000 (module :46-47).
// Python side: numbalsoda/driver_solve_ivp.py:86-218.
//
// Organisation (what differs from the CPU code, the arithmetic does not):
//  * a trajectory is a small state machine -- ivp_begin() once, ivp_step() per attempted
//    step -- so that a persistent grid can give a lane its next trajectory inside the same
//    loop (solve_ivp_kernel.cu);
//  * y, k1 and the scalars of the step loop stay in registers; the 8 n dense-output
//    coefficients and the previous event values live in shared memory, strided by the
//    block size (conflict-free), because Brent's iteration and the row loop index them
//    at run time;
//  * every accepted step needs the three extra stages here (iout = 2), so unlike dop853()
//    there is nothing to queue: dense output is in line and converged;
//  * event location (rare: once per trajectory, events are terminal) is out of line.
//
// Also compiles as host C++ (tests/solve_ivp_core_host.cpp) so that the restated control
// flow is compared with the CPU oracle bit for bit in a container without a GPU.
#ifndef B200ODE_SOLVE_IVP_CORE_CUH
#define B200ODE_SOLVE_IVP_CORE_CUH

#include "b200ode_launch.h"
#include "b200ode_pow.cuh"
#include "b200ode_fastmath.cuh"
#include "dop853_tableau.cuh"

#if defined(__CUDACC__)
#define IVP_FN __device__ __forceinline__
#define IVP_SLOW __device__ __noinline__
// throughput mode (handle linked without B200ODE_STRICT_FP / B200ODE_EXACT_CTRL): the error
// norm and the step-size controller through reciprocals and a single-precision power, as in
// dop853_kernel.cu (b200ode_fastmath.cuh); a link-time constant
#define IVP_FAST_CTRL() B200ODE_FAST_CTRL()
#else
#include <cmath>
#define IVP_FN static inline
#define IVP_SLOW static
#ifndef IVP_HOST_FAST
#define IVP_HOST_FAST 0  // tests: the throughput mode's arithmetic with libm stand-ins
#endif
#define IVP_FAST_CTRL() (IVP_HOST_FAST != 0)
#endif

#define IVP_UROUND 2.220446049250313e-16  // epsilon(1.0d0), dop853_constants.f90:11
#define IVP_SAFE 0.9                      // dop853_module.f90:55
#define IVP_FAC1 0.333                    // :56
#define IVP_FAC2 6.0                      // :59
#define IVP_NMAX 100000                   // :46
#define IVP_NSTIFF 1000                   // :47
#define IVP_BRENT_TOL 1.0e-8              // dop853_solve_ivp.f90:96

#define IVP_FOR_N _Pragma("unroll") for (int i = 0; i < N; i++)
#define IVP_FOR_EV _Pragma("unroll") for (int j = 0; j < B200ODE_IVP_NEVMAX; j++)

// gfortran's MAX / MIN ignore a NaN operand, like fmax / fmin
IVP_FN double ivp_max(double a, double b) { return fmax(a, b); }
IVP_FN double ivp_min(double a, double b) { return fmin(a, b); }
IVP_FN double ivp_sign(double a, double b) { return copysign(fabs(a), b); }

// State of one trajectory between attempted steps.
template <int N>
struct IvpLane {
    double y[N], k1[N];
    double x, h, xend, hmax, posneg, rtol, atol, hlamb;
    double* pp;  // parameters the right-hand side and the events function read: registers
                 // (staged) or the record's address
    const double* pg;  // the record's address in global memory
    int nstep, naccpt, nrejct, iasti, nonsti;
    int row;     // t_eval: rows written (d%nt - 1); otherwise rows saved (size(d%t))
    int cap;     // rows this trajectory owns in t_out / y_out
    long long base;  // its first row
    unsigned flags;
};
enum : unsigned { IVP_LAST = 1u, IVP_REJECT = 2u, IVP_FORWARD = 4u };

// Storage a lane indexes at run time: cont[(v*N + i) * cs], v = 0..7 and ev_old[j * cs].
struct IvpShared {
    double* cont;
    double* ev_old;
    int cs;
};
#define IVP_CONT(v, i) sh.cont[((v) * N + (i)) * sh.cs]
#define IVP_EVOLD(j) sh.ev_old[(j) * sh.cs]

// contd8, dop853_module.f90:832-872, for every component
template <int N>
IVP_FN void ivp_contd8(const IvpShared& sh, double xold, double hout, double x, double* out)
{
    const double s = IVP_FAST_CTRL() ? b200_div(x - xold, hout) : (x - xold) / hout;
    const double s1 = 1.0 - s;
    IVP_FOR_N {
        const double conpar = IVP_CONT(4, i) + s * (IVP_CONT(5, i) + s1 * (IVP_CONT(6, i) + s * IVP_CONT(7, i)));
        out[i] = IVP_CONT(0, i) + s * (IVP_CONT(1, i) + s1 * (IVP_CONT(2, i) + s * (IVP_CONT(3, i) + s1 * conpar)));
    }
}

// find_root's fcn, dop853_solve_ivp.f90:116-128: event `ind` on the interpolant at x
template <int N>
IVP_FN double ivp_event_at(const IvpShared& sh, double xold, double hout, double* pp, int ind, double x)
{
    double yt[N], et[B200ODE_IVP_NEVMAX];
    IVP_FOR_EV et[j] = 0.0;
    ivp_contd8<N>(sh, xold, hout, x, yt);
    b200ode_user_events(x, yt, et, pp);
    double v = 0.0;
    IVP_FOR_EV if (j == ind) v = et[j];
    return v;
}

// zeroin, brent_mod.f90:237-373, on [ax, bx] = the step just accepted.
struct IvpRoot {
    double xzero;
    int iflag;
};
// Out of line (once per trajectory at most).  STAGED: the parameters are staged again from
// the record, so that the caller's copy is never addressed and stays in registers.
template <int N, bool STAGED>
IVP_SLOW IvpRoot ivp_zeroin(double* cont, double* ev_old, int cs, double xold, double hout,
                            const double* pg, int np, int ind, double ax, double bx)
{
    double pl[STAGED ? B200ODE_NPMAX : 1];
    double* pp;
    if (STAGED) {
#pragma unroll
        for (int j = 0; j < B200ODE_NPMAX; j++) pl[j] = (j < np) ? pg[j] : 0.0;
        pp = pl;
    } else {
        pp = const_cast<double*>(pg);
    }
    IvpShared sh;
    sh.cont = cont;
    sh.ev_old = ev_old;
    sh.cs = cs;
    IvpRoot root;
    root.xzero = 0.0;
    root.iflag = 0;
    const double eps = IVP_UROUND, tol = IVP_BRENT_TOL;
    double a = ax, b = bx, c, d, e, fa, fb, fc, tol1, xm, p, q, r, s;
    fa = ivp_event_at<N>(sh, xold, hout, pp, ind, a);
    fb = ivp_event_at<N>(sh, xold, hout, pp, ind, b);
    if (fa == 0.0) {
        root.xzero = a;
        return root;
    }
    if (fb == 0.0) {
        root.xzero = b;
        return root;
    }
    if (!(fa * fb / fabs(fb) < 0.0)) {
        root.iflag = -1;
        return root;
    }
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
            if (a != c) {  // inverse quadratic interpolation
                q = fa / fc;
                r = fb / fc;
                p = s * (2.0 * xm * q * (q - r) - (b - a) * (r - 1.0));
                q = (q - 1.0) * (r - 1.0) * (s - 1.0);
            } else {  // linear interpolation
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
        fb = ivp_event_at<N>(sh, xold, hout, pp, ind, b);
        if (fb * fc / fabs(fc) > 0.0) {
            c = a;
            fc = fa;
            d = b - a;
            e = d;
        }
    }
    root.xzero = b;
    return root;
}

// What a finished trajectory reports: result[b*12 + k] (include/b200ode.h) and the event.
template <int N>
IVP_FN void ivp_finish(const B200IvpLaunch& L, long long b, const IvpLane<N>& S, int idid, int success,
                       int message, int nt_out, int rows_needed, int event_found, int ind_event,
                       double t_event)
{
    int* r = L.result + b * B200ODE_IVP_NRESULT;
    if (L.nt == 0 && rows_needed > S.cap) {
        // not in the reference, whose arrays grow (dop853_solve_ivp.f90:204-207): the
        // caller's rows ran out; rows_needed says how many a complete result has
        success = 0;
        message = B200ODE_IVP_MSG_CAPACITY;
        nt_out = S.cap;
    }
    r[0] = success;
    r[1] = message;
    // nfev stays 0 unless the integration ended normally (dop853_solve_ivp.f90:340-347);
    // nfcn, module :524,:587,:629,:691: 2 + 11 per attempt + 1 + 3 per accepted step
    r[2] = (idid > 0) ? 2 + 11 * S.nstep + 4 * S.naccpt : 0;
    r[3] = nt_out;
    r[4] = rows_needed;
    r[5] = event_found;
    r[6] = ind_event;
    r[7] = idid;
    r[8] = S.nstep;
    r[9] = S.naccpt;
    r[10] = S.nrejct;
    r[11] = 0;
    if (!event_found) {
        // ODEResult.t_event / y_event are zeros without an event (driver_solve_ivp.py:194-195)
        L.t_event[b] = 0.0;
        IVP_FOR_N L.y_event[b * N + i] = 0.0;
    } else {
        L.t_event[b] = t_event;
    }
}

// Start trajectory b: dop853_solve_ivp_wrapper (dop853_solve_ivp.f90:257-335), dop853
// (module :396-438), the prologue of dp86co (module :502-534) and solout's first call
// (nr = 1, dop853_solve_ivp.f90:150-157, :204-208).  Returns true if the trajectory is
// already finished (t_eval not monotone in the direction of integration).
template <int N>
IVP_FN bool ivp_begin(IvpLane<N>& S, const B200IvpLaunch& L, long long b, const IvpShared& sh, int te_mono)
{
    const double* ts = L.t_span + b * L.t_span_stride;
    const double* y0 = L.y0 + b * L.y0_stride;
    S.x = ts[0];
    S.xend = ts[1];
    S.rtol = L.rtol_b ? L.rtol_b[b] : L.rtol;
    S.atol = L.atol_b ? L.atol_b[b] : L.atol;
    S.nstep = S.naccpt = S.nrejct = S.iasti = S.nonsti = 0;
    S.row = 0;
    if (L.nt > 0) {
        S.base = b * L.nt;
        S.cap = L.nt;
    } else if (L.row_offset) {
        // ragged rows: trajectory b owns rows [row_offset[b], row_offset[b+1])
        S.base = L.row_offset[b];
        S.cap = (int)(L.row_offset[b + 1] - S.base);
    } else {
        S.base = b * L.cap;
        S.cap = (int)L.cap;
    }
    S.hlamb = 0.0;
    S.flags = (S.xend >= S.x) ? IVP_FORWARD : 0u;
    if (L.nt > 0 && !(te_mono & ((S.flags & IVP_FORWARD) ? 1 : 2))) {
        // '"t_eval" does not always increase or decrease', dop853_solve_ivp.f90:278-316
        ivp_finish<N>(L, b, S, 0, 0, B200ODE_IVP_MSG_TEVAL, 0, 0, 0, 0, 0.0);
        return true;
    }
    IVP_FOR_N S.y[i] = y0[i];
    // set_parameters zeroes cont (module :284)
    IVP_FOR_N {
#pragma unroll
        for (int v = 0; v < 8; v++) IVP_CONT(v, i) = 0.0;
    }
    double hmax = (L.max_step == 0.0) ? S.xend - S.x : L.max_step;  // module :431-435
    S.posneg = ivp_sign(1.0, S.xend - S.x);                         // module :507
    b200ode_user_rhs(S.x, S.y, S.k1, S.pp);
    hmax = fabs(hmax);
    S.hmax = hmax;
    double h = L.first_step;
    if (h == 0.0) {
        // hinit, module :745-823 (itol = 0, iord = 8)
        const double rtol = S.rtol, atol = S.atol, posneg = S.posneg;
        double y1[N], k2[N];
        double dnf = 0.0, dny = 0.0;
        IVP_FOR_N {
            const double sk = atol + rtol * fabs(S.y[i]);
            double qq = S.k1[i] / sk;
            dnf = dnf + qq * qq;
            qq = S.y[i] / sk;
            dny = dny + qq * qq;
        }
        if (dnf <= 1.0e-10 || dny <= 1.0e-10) h = 1.0e-6;
        else h = sqrt(dny / dnf) * 0.01;
        h = ivp_min(h, hmax);
        h = ivp_sign(h, posneg);
        IVP_FOR_N y1[i] = S.y[i] + h * S.k1[i];
        b200ode_user_rhs(S.x + h, y1, k2, S.pp);
        double der2 = 0.0;
        IVP_FOR_N {
            const double sk = atol + rtol * fabs(S.y[i]);
            const double qq = (k2[i] - S.k1[i]) / sk;
            der2 = der2 + qq * qq;
        }
        der2 = sqrt(der2) / h;
        const double der12 = ivp_max(fabs(der2), sqrt(dnf));
        double h1;
        if (der12 <= 1.0e-15) h1 = ivp_max(1.0e-6, fabs(h) * 1.0e-3);
        else h1 = b200ode_pow_eighth(0.01 / der12);
        h = ivp_min(ivp_min(100.0 * fabs(h), h1), hmax);
        h = ivp_sign(h, posneg);
    }
    S.h = h;
    // solout, nr = 1
    if (L.n_events > 0) {
        double ev[B200ODE_IVP_NEVMAX];
        IVP_FOR_EV ev[j] = 0.0;
        b200ode_user_events(S.x, S.y, ev, S.pp);
        IVP_FOR_EV if (j < L.n_events) IVP_EVOLD(j) = ev[j];
    }
    if (L.nt == 0) {
        if (S.row < S.cap) {
            L.t_out[S.base + S.row] = S.x;
            double* o = L.y_out + (S.base + S.row) * N;
            IVP_FOR_N o[i] = S.y[i];
        }
        S.row++;
    }
    return false;
}

// One attempted step of trajectory b: dp86co's loop body (module :537-731) with solout
// (dop853_solve_ivp.f90:131-209) and the wrapper's epilogue (:337-395) when the
// integration ends.  Returns true when the trajectory is finished.
template <int N, bool STAGED>
IVP_FN bool ivp_step(IvpLane<N>& S, const B200IvpLaunch& L, long long b, const IvpShared& sh)
{
    const double facc1 = 1.0 / IVP_FAC1, facc2 = 1.0 / IVP_FAC2;  // module :505-506
    const int t_eval_given = L.nt > 0;
    int idid = 0;
    if (S.nstep > IVP_NMAX) idid = -2;                                 // module :539
    else if (0.1 * fabs(S.h) <= fabs(S.x) * IVP_UROUND) idid = -3;     // module :546
    if (idid < 0) {
        // 'Failed to complete integration.', dop853_solve_ivp.f90:340-353: nt_out = size(d%t)
        const int rows = t_eval_given ? L.nt : S.row;
        ivp_finish<N>(L, b, S, idid, 0, B200ODE_IVP_MSG_FAILED, rows, rows, 0, 0, 0.0);
        return true;
    }
    double* const y = S.y;
    double* const k1 = S.k1;
    double* const pp = S.pp;
    const double x = S.x, posneg = S.posneg;
    if ((x + 1.01 * S.h - S.xend) * posneg > 0.0) {  // module :554-557
        S.h = S.xend - x;
        S.flags |= IVP_LAST;
    }
    const double h = S.h;
    S.nstep++;
    double y1[N], k2[N], k3[N], k4[N], k5[N], k6[N], k7[N], k8[N], k9[N], k10[N];
    // the twelve stages, module :561-587
    IVP_FOR_N y1[i] = y[i] + h * DP_A2_1 * k1[i];
    b200ode_user_rhs(x + DP_C2 * h, y1, k2, pp);
    IVP_FOR_N y1[i] = y[i] + h * (DP_A3_1 * k1[i] + DP_A3_2 * k2[i]);
    b200ode_user_rhs(x + DP_C3 * h, y1, k3, pp);
    IVP_FOR_N y1[i] = y[i] + h * (DP_A4_1 * k1[i] + DP_A4_3 * k3[i]);
    b200ode_user_rhs(x + DP_C4 * h, y1, k4, pp);
    IVP_FOR_N y1[i] = y[i] + h * (DP_A5_1 * k1[i] + DP_A5_3 * k3[i] + DP_A5_4 * k4[i]);
    b200ode_user_rhs(x + DP_C5 * h, y1, k5, pp);
    IVP_FOR_N y1[i] = y[i] + h * (DP_A6_1 * k1[i] + DP_A6_4 * k4[i] + DP_A6_5 * k5[i]);
    b200ode_user_rhs(x + DP_C6 * h, y1, k6, pp);
    IVP_FOR_N y1[i] = y[i] + h * (DP_A7_1 * k1[i] + DP_A7_4 * k4[i] + DP_A7_5 * k5[i] + DP_A7_6 * k6[i]);
    b200ode_user_rhs(x + DP_C7 * h, y1, k7, pp);
    IVP_FOR_N y1[i] = y[i] + h * (DP_A8_1 * k1[i] + DP_A8_4 * k4[i] + DP_A8_5 * k5[i] + DP_A8_6 * k6[i] +
                                  DP_A8_7 * k7[i]);
    b200ode_user_rhs(x + DP_C8 * h, y1, k8, pp);
    IVP_FOR_N y1[i] = y[i] + h * (DP_A9_1 * k1[i] + DP_A9_4 * k4[i] + DP_A9_5 * k5[i] + DP_A9_6 * k6[i] +
                                  DP_A9_7 * k7[i] + DP_A9_8 * k8[i]);
    b200ode_user_rhs(x + DP_C9 * h, y1, k9, pp);
    IVP_FOR_N y1[i] = y[i] + h * (DP_A10_1 * k1[i] + DP_A10_4 * k4[i] + DP_A10_5 * k5[i] + DP_A10_6 * k6[i] +
                                  DP_A10_7 * k7[i] + DP_A10_8 * k8[i] + DP_A10_9 * k9[i]);
    b200ode_user_rhs(x + DP_C10 * h, y1, k10, pp);
    IVP_FOR_N y1[i] = y[i] + h * (DP_A11_1 * k1[i] + DP_A11_4 * k4[i] + DP_A11_5 * k5[i] + DP_A11_6 * k6[i] +
                                  DP_A11_7 * k7[i] + DP_A11_8 * k8[i] + DP_A11_9 * k9[i] +
                                  DP_A11_10 * k10[i]);
    b200ode_user_rhs(x + DP_C11 * h, y1, k2, pp);
    const double xph = x + h;
    IVP_FOR_N y1[i] = y[i] + h * (DP_A12_1 * k1[i] + DP_A12_4 * k4[i] + DP_A12_5 * k5[i] + DP_A12_6 * k6[i] +
                                  DP_A12_7 * k7[i] + DP_A12_8 * k8[i] + DP_A12_9 * k9[i] +
                                  DP_A12_10 * k10[i] + DP_A12_11 * k2[i]);
    b200ode_user_rhs(xph, y1, k3, pp);
    IVP_FOR_N {
        k4[i] = DP_B1 * k1[i] + DP_B6 * k6[i] + DP_B7 * k7[i] + DP_B8 * k8[i] + DP_B9 * k9[i] +
                DP_B10 * k10[i] + DP_B11 * k2[i] + DP_B12 * k3[i];
        k5[i] = y[i] + h * k4[i];
    }
    // error estimation, module :591-616
    double err = 0.0, err2 = 0.0, fac11 = 0.0, ifac = 0.0, hnew;
    if (IVP_FAST_CTRL()) {
        // the same quantities through reciprocals; err holds the *square* of the reference's
        // err (the test err <= 1 is unchanged by that), ifac = safe / err**(1/8)
        IVP_FOR_N {
            const double isk = b200_rcp(S.atol + S.rtol * ivp_max(fabs(y[i]), fabs(k5[i])));
            double erri = k4[i] - DP_BHH1 * k1[i] - DP_BHH2 * k9[i] - DP_BHH3 * k3[i];
            double qq = erri * isk;
            err2 = err2 + qq * qq;
            erri = DP_ER1 * k1[i] + DP_ER6 * k6[i] + DP_ER7 * k7[i] + DP_ER8 * k8[i] + DP_ER9 * k9[i] +
                   DP_ER10 * k10[i] + DP_ER11 * k2[i] + DP_ER12 * k3[i];
            qq = erri * isk;
            err = err + qq * qq;
        }
        double deno = err + 0.01 * err2;
        if (deno <= 0.0) deno = 1.0;
        err = (h * h) * (err * err) * b200_rcp((double)N * deno);
        ifac = IVP_SAFE * b200_inv_root16(err);
        // fac = max(facc2, min(facc1, fac11/safe)); hnew = h/fac   (module :620-623)
        hnew = h * ivp_min(IVP_FAC2, ivp_max(IVP_FAC1, ifac));
    } else {
        IVP_FOR_N {
            const double sk = S.atol + S.rtol * ivp_max(fabs(y[i]), fabs(k5[i]));
            double erri = k4[i] - DP_BHH1 * k1[i] - DP_BHH2 * k9[i] - DP_BHH3 * k3[i];
            double qq = erri / sk;
            err2 = err2 + qq * qq;
            erri = DP_ER1 * k1[i] + DP_ER6 * k6[i] + DP_ER7 * k7[i] + DP_ER8 * k8[i] + DP_ER9 * k9[i] +
                   DP_ER10 * k10[i] + DP_ER11 * k2[i] + DP_ER12 * k3[i];
            qq = erri / sk;
            err = err + qq * qq;
        }
        double deno = err + 0.01 * err2;
        if (deno <= 0.0) deno = 1.0;
        err = fabs(h) * err * sqrt(1.0 / ((double)N * deno));
        // step size controller, module :618-623 (facold**beta = 1: beta = 0)
        fac11 = b200ode_pow_eighth(err);
        double fac = fac11 / 1.0;
        fac = ivp_max(facc2, ivp_min(facc1, fac / IVP_SAFE));
        hnew = h / fac;
    }
    if (!(err <= 1.0)) {
        // ---- rejected, module :724-728
        if (IVP_FAST_CTRL()) hnew = h * ivp_max(IVP_FAC1, ifac);
        else hnew = h / ivp_min(facc1, fac11 / IVP_SAFE);
        if (S.naccpt >= 1) S.nrejct++;
        S.flags = (S.flags & IVP_FORWARD) | IVP_REJECT;  // and last = .false.
        S.h = hnew;
        return false;
    }
    // ---- accepted, module :625-722
    S.naccpt++;
    b200ode_user_rhs(xph, k5, k4, pp);
    // stiffness detection, module :631-655; iprint = 0: exit with idid = -4
    if (S.naccpt % IVP_NSTIFF == 0 || S.iasti > 0) {
        double stnum = 0.0, stden = 0.0;
        IVP_FOR_N {
            stnum = stnum + (k4[i] - k3[i]) * (k4[i] - k3[i]);
            stden = stden + (k5[i] - y1[i]) * (k5[i] - y1[i]);
        }
        if (stden > 0.0) S.hlamb = fabs(h) * sqrt(stnum / stden);
        if (S.hlamb > 6.1) {
            S.nonsti = 0;
            S.iasti++;
            if (S.iasti == 15) {
                const int rows = t_eval_given ? L.nt : S.row;
                ivp_finish<N>(L, b, S, -4, 0, B200ODE_IVP_MSG_FAILED, rows, rows, 0, 0, 0.0);
                return true;
            }
        } else {
            S.nonsti++;
            if (S.nonsti == 6) S.iasti = 0;
        }
    }
    // dense output of every accepted step (iout = 2), module :657-705
    IVP_FOR_N {
        IVP_CONT(0, i) = y[i];
        const double ydiff = k5[i] - y[i];
        IVP_CONT(1, i) = ydiff;
        const double bspl = h * k1[i] - ydiff;
        IVP_CONT(2, i) = bspl;
        IVP_CONT(3, i) = ydiff - h * k4[i] - bspl;
    }
    double c4[N], c5[N], c6[N], c7[N];
    IVP_FOR_N {
        c4[i] = DP_D4_1 * k1[i] + DP_D4_6 * k6[i] + DP_D4_7 * k7[i] + DP_D4_8 * k8[i] + DP_D4_9 * k9[i] +
                DP_D4_10 * k10[i] + DP_D4_11 * k2[i] + DP_D4_12 * k3[i];
        c5[i] = DP_D5_1 * k1[i] + DP_D5_6 * k6[i] + DP_D5_7 * k7[i] + DP_D5_8 * k8[i] + DP_D5_9 * k9[i] +
                DP_D5_10 * k10[i] + DP_D5_11 * k2[i] + DP_D5_12 * k3[i];
        c6[i] = DP_D6_1 * k1[i] + DP_D6_6 * k6[i] + DP_D6_7 * k7[i] + DP_D6_8 * k8[i] + DP_D6_9 * k9[i] +
                DP_D6_10 * k10[i] + DP_D6_11 * k2[i] + DP_D6_12 * k3[i];
        c7[i] = DP_D7_1 * k1[i] + DP_D7_6 * k6[i] + DP_D7_7 * k7[i] + DP_D7_8 * k8[i] + DP_D7_9 * k9[i] +
                DP_D7_10 * k10[i] + DP_D7_11 * k2[i] + DP_D7_12 * k3[i];
    }
    // the next three function evaluations, module :682-691
    IVP_FOR_N y1[i] = y[i] + h * (DP_A14_1 * k1[i] + DP_A14_7 * k7[i] + DP_A14_8 * k8[i] + DP_A14_9 * k9[i] +
                                  DP_A14_10 * k10[i] + DP_A14_11 * k2[i] + DP_A14_12 * k3[i] +
                                  DP_A14_13 * k4[i]);
    b200ode_user_rhs(x + DP_C14 * h, y1, k10, pp);
    IVP_FOR_N y1[i] = y[i] + h * (DP_A15_1 * k1[i] + DP_A15_6 * k6[i] + DP_A15_7 * k7[i] + DP_A15_8 * k8[i] +
                                  DP_A15_11 * k2[i] + DP_A15_12 * k3[i] + DP_A15_13 * k4[i] +
                                  DP_A15_14 * k10[i]);
    b200ode_user_rhs(x + DP_C15 * h, y1, k2, pp);
    IVP_FOR_N y1[i] = y[i] + h * (DP_A16_1 * k1[i] + DP_A16_6 * k6[i] + DP_A16_7 * k7[i] + DP_A16_8 * k8[i] +
                                  DP_A16_9 * k9[i] + DP_A16_13 * k4[i] + DP_A16_14 * k10[i] +
                                  DP_A16_15 * k2[i]);
    b200ode_user_rhs(x + DP_C16 * h, y1, k3, pp);
    // final preparation, module :693-703
    IVP_FOR_N {
        IVP_CONT(4, i) = h * (c4[i] + DP_D4_13 * k4[i] + DP_D4_14 * k10[i] + DP_D4_15 * k2[i] + DP_D4_16 * k3[i]);
        IVP_CONT(5, i) = h * (c5[i] + DP_D5_13 * k4[i] + DP_D5_14 * k10[i] + DP_D5_15 * k2[i] + DP_D5_16 * k3[i]);
        IVP_CONT(6, i) = h * (c6[i] + DP_D6_13 * k4[i] + DP_D6_14 * k10[i] + DP_D6_15 * k2[i] + DP_D6_16 * k3[i]);
        IVP_CONT(7, i) = h * (c7[i] + DP_D7_13 * k4[i] + DP_D7_14 * k10[i] + DP_D7_15 * k2[i] + DP_D7_16 * k3[i]);
    }
    IVP_FOR_N {
        k1[i] = k4[i];
        y[i] = k5[i];
    }
    const double xold = x;  // me%xold, me%hout = h (module :704,:708)
    S.x = xph;

    // ---- solout, nr > 1 (dop853_solve_ivp.f90:149-207)
    int event_found = 0, ind_event = 0, brent_failed = 0;
    double t_event = 0.0;
    if (L.n_events > 0) {
        double ev[B200ODE_IVP_NEVMAX];
        IVP_FOR_EV ev[j] = 0.0;
        b200ode_user_events(xph, y, ev, pp);
        int hit = -1;
        IVP_FOR_EV {
            if (j < L.n_events && hit < 0) {
                const double eo = IVP_EVOLD(j);
                if ((eo < 0 && ev[j] > 0) || (eo > 0 && ev[j] < 0)) hit = j;
            }
        }
        if (hit >= 0) {
            const IvpRoot root = ivp_zeroin<N, STAGED>(sh.cont, sh.ev_old, sh.cs, xold, h, S.pg, L.np, hit, xold, xph);
            if (root.iflag != 0) {
                brent_failed = 1;  // the reference prints 'brent failed' and stops the process
            } else {
                event_found = 1;
                ind_event = hit;  // 0-based (dop853_solve_ivp.f90:432)
                t_event = root.xzero;
                double ye[N];
                ivp_contd8<N>(sh, xold, h, t_event, ye);
                IVP_FOR_N L.y_event[b * N + i] = ye[i];
            }
        }
        IVP_FOR_EV if (j < L.n_events) IVP_EVOLD(j) = ev[j];
    }
    if (brent_failed) {
        ivp_finish<N>(L, b, S, 0, 0, B200ODE_IVP_MSG_BRENT, S.row, t_eval_given ? L.nt : S.row, 0, 0, 0.0);
        return true;
    }
    if (t_eval_given) {
        const double xx = event_found ? t_event : xph;
        const bool fwd = (S.flags & IVP_FORWARD) != 0;
        int row = S.row;
        double* o = L.y_out + (S.base + row) * N;
        while (row < L.nt) {
            const double tq = L.t_eval[row];
            if (!(fwd ? tq <= xx : tq >= xx)) break;
            double yr[N];
            ivp_contd8<N>(sh, xold, h, tq, yr);
            IVP_FOR_N o[i] = yr[i];
            o += N;
            row++;
        }
        S.row = row;
    } else if (!event_found) {
        if (S.row < S.cap) {
            L.t_out[S.base + S.row] = xph;
            double* o = L.y_out + (S.base + S.row) * N;
            IVP_FOR_N o[i] = y[i];
        }
        S.row++;
    }
    if (event_found || (S.flags & IVP_LAST)) {
        // irtrn < 0: idid = 2 (module :711-714); last step: idid = 1 (module :715-719);
        // the wrapper's epilogue, dop853_solve_ivp.f90:355-393
        int nt_out, message;
        const int rows_needed = t_eval_given ? L.nt : S.row;
        if (event_found) {
            message = B200ODE_IVP_MSG_EVENT;
            nt_out = S.row;  // rows are trimmed to d%nt - 1 (:356-373)
        } else {
            message = B200ODE_IVP_MSG_END;
            nt_out = rows_needed;
        }
        ivp_finish<N>(L, b, S, event_found ? 2 : 1, 1, message, nt_out, rows_needed, event_found,
                      ind_event, t_event);
        return true;
    }
    if (fabs(hnew) > S.hmax) hnew = posneg * S.hmax;                             // module :720
    if (S.flags & IVP_REJECT) hnew = posneg * ivp_min(fabs(hnew), fabs(h));      // module :721
    S.flags &= ~IVP_REJECT;
    S.h = hnew;
    return false;
}

#endif
