// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
// Compiles the thread-per-trajectory solve_ivp core (numbalsoda_b200/csrc/solve_ivp_core.cuh)
// as host C++, so that its control flow and arithmetic order can be compared with the CPU
// oracle bit for bit in a container without a GPU (tests/test_solve_ivp_core_host.py).
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

#include "../numbalsoda_b200/csrc/solve_ivp_core.cuh"

template <int N>
static void run(B200IvpLaunch& L)
{
    std::vector<double> w(B200ODE_IVP_DOUBLES(N), 0.0);
    IvpShared sh;
    sh.cont = w.data();
    sh.ev_old = w.data() + 8 * N;
    sh.cs = 1;
    int inc = 1, dec = 1;
    for (int i = 1; i < L.nt; i++) {
        if (L.t_eval[i] < L.t_eval[i - 1]) inc = 0;
        if (L.t_eval[i] > L.t_eval[i - 1]) dec = 0;
    }
    const int te_mono = inc | (dec << 1);
    for (long long b = 0; b < L.B; b++) {
        IvpLane<N> S;
        S.pp = reinterpret_cast<double*>(const_cast<char*>(L.data + b * L.data_stride));
        S.pg = S.pp;
        if (ivp_begin<N>(S, L, b, sh, te_mono)) continue;
        while (!ivp_step<N, false>(S, L, b, sh)) {
        }
    }
}

extern "C" int solve_ivp_core_host(rhs_fn f, rhs_fn events, int neq, B200IvpLaunch* L)
{
    g_rhs = f;
    g_events = events;
    switch (neq) {
    case 1: run<1>(*L); break;
    case 2: run<2>(*L); break;
    case 3: run<3>(*L); break;
    case 4: run<4>(*L); break;
    case 5: run<5>(*L); break;
    case 6: run<6>(*L); break;
    case 7: run<7>(*L); break;
    case 8: run<8>(*L); break;
    default: return -1;
    }
    return 0;
}
