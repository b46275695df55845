// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
// Compiles the LSODA device core (numbalsoda_b200/csrc/lsoda_core.cuh) as host C++ for
// one system size (-DB200_NEQ=n, thread backend) or any size (-DB200_WARP_BACKEND) so that its restated control flow can be compared
// with the CPU oracle bit for bit in a container without a GPU
// (tests/test_lsoda_core_host.py).  Build: g++ -O2 -ffp-contract=off -mfma.
#include <cstddef>
#include <cstring>

typedef void (*rhs_fn)(double, double*, double*, void*);
static thread_local rhs_fn g_rhs;
static inline long long b200ode_user_rhs(double t, double* u, double* du, double* p)
{
    g_rhs(t, u, du, p);
    return 0;
}

#ifdef B200_WARP_BACKEND
// warp-per-trajectory backend, one emulated lane, any neq <= 128
#include <vector>
#include "../numbalsoda_b200/csrc/lsoda_vec_warp.cuh"
#else
#define LS_STRIDE 1
#include "../numbalsoda_b200/csrc/lsoda_vec_thread.cuh"
#endif

static const B200LsodaTables g_tables = B200_LSODA_TABLES_INIT;

extern "C" const B200LsodaTables* lsoda_core_tables(void) { return &g_tables; }

static void core_run(rhs_fn f, int neq, const double* u0, void* data, int nt, const double* teval,
                     double* usol, const LsTol& tol, int mxstep, int* success, int exit_on_warning,
                     int* counters, int ml = -1, int mu = -1);

// banded finite-difference Jacobian (warp backend only)
extern "C" void lsoda_core_host_band(rhs_fn f, int neq, const double* u0, void* data, int nt,
                                     const double* teval, double* usol, double rtol, double atol,
                                     int ml, int mu, int mxstep, int* success, int exit_on_warning,
                                     int* counters)
{
    LsTol tol = {rtol, atol, nullptr, nullptr};
    core_run(f, neq, u0, data, nt, teval, usol, tol, mxstep, success, exit_on_warning, counters, ml, mu);
}

extern "C" void lsoda_core_host(rhs_fn f, int neq, const double* u0, void* data, int nt,
                                const double* teval, double* usol, double rtol, double atol,
                                int mxstep, int* success, int exit_on_warning, int* counters)
{
    LsTol tol = {rtol, atol, nullptr, nullptr};
    core_run(f, neq, u0, data, nt, teval, usol, tol, mxstep, success, exit_on_warning, counters);
}

// per-component tolerance vectors (null: the scalar applies)
extern "C" void lsoda_core_host_vtol(rhs_fn f, int neq, const double* u0, void* data, int nt,
                                     const double* teval, double* usol, double rtol, double atol,
                                     const double* rtol_v, const double* atol_v, int mxstep,
                                     int* success, int exit_on_warning, int* counters)
{
    LsTol tol = {rtol, atol, rtol_v, atol_v};
    core_run(f, neq, u0, data, nt, teval, usol, tol, mxstep, success, exit_on_warning, counters);
}

static void core_run(rhs_fn f, int neq, const double* u0, void* data, int nt, const double* teval,
                     double* usol, const LsTol& tol, int mxstep, int* success, int exit_on_warning,
                     int* counters, int ml, int mu)
{
    g_rhs = f;
    LsLane s;
    LsRun R;
#ifdef B200_WARP_BACKEND
    if (neq < 1 || neq > WV_NMAX) {
        *success = -99;
        return;
    }
    std::vector<double> store(ml >= 0 ? B200ODE_LSODAW_BAND_DOUBLES(neq, ml, mu) : B200ODE_LSODAW_DOUBLES(neq), 0.0);
    LsWarpVec v;
    v.attach(store.data(), neq, (exit_on_warning & 2) != 0, nullptr, ml, mu);  // bit 1 of the flag: fast solves
    exit_on_warning &= 1;
#else
    if (neq != N) {
        *success = -99;
        return;
    }
    (void)ml, (void)mu;
    double yh[13 * N], wm[N * N];
    memset(yh, 0, sizeof yh);
    memset(wm, 0, sizeof wm);
    LsThreadVec v;
    v.yhp = yh;
    v.wmp = wm;
#endif
    double* pp = static_cast<double*>(data);
    bool done = ls_begin(s, R, v, u0, pp, nt, teval, usol, tol);
    const unsigned long long mxstep_u = (unsigned long long)(long long)mxstep;
    while (!done)
        done = ls_advance(s, R, v, &g_tables, pp, nt, teval, usol, tol, mxstep_u,
                          exit_on_warning, 0u);
    *success = R.istate > 0 ? 1 : 0;
    if (counters) {
        counters[0] = s.nst;
        counters[1] = s.nfe;
        counters[2] = s.nje;
        counters[3] = R.nswitch;
        counters[4] = s.nqu;
        counters[5] = R.istate;
        counters[6] = R.row;
        counters[7] = s.meth;
    }
}
