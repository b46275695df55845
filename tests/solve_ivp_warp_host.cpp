// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
// Compiles the warp-per-trajectory solve_ivp core (numbalsoda_b200/csrc/solve_ivp_warp.cuh) as
// host C++ with one emulated lane, so that its control flow and arithmetic order can be compared
// with the CPU oracle bit for bit without a GPU (tests/test_solve_ivp_core_host.py).
// Build: g++ -O2 -ffp-contract=off -mfma.
#include <cstddef>
#include <cstring>
#include <vector>

typedef void (*rhs_fn)(double, double*, double*, void*);
static thread_local rhs_fn g_rhs, g_events;
static inline long long b200ode_user_rhs(double t, double* u, double* du, double* p)
{
    g_rhs(t, u, du, p);
    return 0;
}
static inline long long b200ode_user_events(double t, double* u, double* ev, double* p)
{
    if (g_events) g_events(t, u, ev, p);
    return 0;
}

#include "../numbalsoda_b200/csrc/solve_ivp_warp.cuh"

extern "C" int solve_ivp_warp_host(rhs_fn f, rhs_fn events, int n, B200IvpLaunch* Lp)
{
    const B200IvpLaunch& L = *Lp;
    g_rhs = f;
    g_events = events;
    int inc = 1, dec = 1;
    for (int i = 1; i < L.nt; i++) {
        if (L.t_eval[i] < L.t_eval[i - 1]) inc = 0;
        if (L.t_eval[i] > L.t_eval[i - 1]) dec = 0;
    }
    const int te_mono = inc | (dec << 1);
    std::vector<double> w(B200ODE_IVPW_DOUBLES(n), 0.0);
    for (long long b = 0; b < L.B; b++) {
        double* pp = reinterpret_cast<double*>(const_cast<char*>(L.data + b * L.data_stride));
        long long base;
        int cap;
        if (L.nt > 0) base = b * L.nt, cap = L.nt;
        else if (L.row_offset) base = L.row_offset[b], cap = (int)(L.row_offset[b + 1] - base);
        else base = b * L.cap, cap = (int)L.cap;
        IvpWarpResult R;
        ivp_warp_solve(n, w.data(), L, L.t_span + b * L.t_span_stride, L.y0 + b * L.y0_stride, pp,
                       L.rtol_b ? L.rtol_b[b] : L.rtol, L.atol_b ? L.atol_b[b] : L.atol, te_mono, cap,
                       L.t_out ? L.t_out + base : nullptr, L.y_out + base * n, L.y_event + b * n, R);
        int* r = L.result + b * B200ODE_IVP_NRESULT;
        r[0] = R.success, r[1] = R.message;
        r[2] = R.idid > 0 ? 2 + 11 * R.nstep + 4 * R.naccpt : 0;
        r[3] = R.nt_out, r[4] = R.rows_needed, r[5] = R.event_found, r[6] = R.ind_event, r[7] = R.idid;
        r[8] = R.nstep, r[9] = R.naccpt, r[10] = R.nrejct, r[11] = 0;
        L.t_event[b] = R.event_found ? R.t_event : 0.0;
    }
    return 0;
}
