// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
// Compiles the warp-per-trajectory DOP853 core (numbalsoda_b200/csrc/dop853_warp.cuh) as
// host C++ with one emulated lane, so that its control flow and arithmetic order can be
// compared with the CPU oracle bit for bit in a container without a GPU
// (tests/test_dop853_warp_host.py).  Build: g++ -O2 -ffp-contract=off -mfma.
#include <cstddef>
#include <cstring>
#include <vector>

typedef void (*rhs_fn)(double, double*, double*, void*);
static thread_local rhs_fn g_rhs;
static inline long long b200ode_user_rhs(double t, double* u, double* du, double* p)
{
    g_rhs(t, u, du, p);
    return 0;
}

#include "../numbalsoda_b200/csrc/dop853_warp.cuh"

extern "C" void dop853_warp_host_vtol(rhs_fn f, int neq, const double* u0, void* data, int nt,
                                      const double* teval, double* usol, double rtol, double atol,
                                      const double* rtol_v, const double* atol_v, int mxstep,
                                      int* success, int* counters);

extern "C" void dop853_warp_host(rhs_fn f, int neq, const double* u0, void* data, int nt,
                                 const double* teval, double* usol, double rtol, double atol,
                                 int mxstep, int* success, int* counters)
{
    dop853_warp_host_vtol(f, neq, u0, data, nt, teval, usol, rtol, atol, nullptr, nullptr, mxstep,
                          success, counters);
}

extern "C" void dop853_warp_host_vtol(rhs_fn f, int neq, const double* u0, void* data, int nt,
                                      const double* teval, double* usol, double rtol, double atol,
                                      const double* rtol_v, const double* atol_v, int mxstep,
                                      int* success, int* counters)
{
    g_rhs = f;
    std::vector<double> w(B200ODE_DOP853W_DOUBLES(neq), 0.0);
    Dop853WarpStats st;
    dop853_warp_solve(neq, w.data(), u0, static_cast<double*>(data), nt, teval, usol, rtol, atol,
                      mxstep, st, rtol_v, atol_v);
    *success = st.idid < 0 ? 0 : 1;  // c_interface :125-128
    if (counters) {
        counters[0] = st.nstep;
        counters[1] = st.idid == -1 ? 0 : 2 + 11 * st.nstep + st.naccpt + 3 * st.nevent;
        counters[2] = st.naccpt;
        counters[3] = st.nrejct;
        counters[4] = st.nevent;
        counters[5] = st.idid;
        counters[6] = st.rows;
        counters[7] = 0;
    }
}
