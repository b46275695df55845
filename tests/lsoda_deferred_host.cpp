// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
// The deferred organisation of the thread-per-trajectory LSODA kernel (lsoda_kernel.cu:
// requests of the rare blocks served block-wide between the two halves of a pass), emulated
// on the host: T virtual threads run an ensemble in lockstep exactly like a block does --
// fetch, first half of the pass, queue building, service of every request by kind, second
// half -- with the same headers (lsoda_core.cuh, lsoda_vec_thread.cuh with LS_DEFERRED).  So
// the restated control flow (ls_advance_a / ls_advance_b, request_jac / take_jac, the
// retraction / rescale / output-row requests, the Jacobian's inputs travelling in yh slots
// 8-10) is compared with the CPU oracle bit for bit without a GPU
// (tests/test_lsoda_deferred_host.py).  Build: g++ -O2 -ffp-contract=off -mfma -DB200_NEQ=n.
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

#ifndef LS_T
#define LS_T 8  // virtual threads of the emulated block
#endif
#define LS_STRIDE LS_T
#define LS_DEFERRED 1
static thread_local int g_tid;  // the virtual thread that is running
static thread_local double* g_reqd;
static thread_local int* g_reqi;
#define LS_REQD(f) g_reqd[(f) * LS_T + g_tid]
#define LS_REQI(f) g_reqi[(f) * LS_T + g_tid]
#define LS_REQD_OF(o, f) g_reqd[(f) * LS_T + (o)]
#define LS_REQI_OF(o, f) g_reqi[(f) * LS_T + (o)]
#include "../numbalsoda_b200/csrc/lsoda_vec_thread.cuh"

static const B200LsodaTables g_tables = B200_LSODA_TABLES_INIT;

// B trajectories (u0 [B,N] or shared, data [B] records of data_stride bytes) on LS_T virtual threads
extern "C" void lsoda_deferred_host(rhs_fn f, long B, int neq, const double* u0, long u0_stride,
                                    const char* data, long data_stride, int nt, const double* teval,
                                    double* usol, double rtol, double atol, int mxstep, int* success,
                                    int exit_on_warning, int* counters)
{
    if (neq != N) {
        for (long b = 0; b < B; b++) success[b] = -99;
        return;
    }
    g_rhs = f;
    std::vector<double> yh(13 * N * LS_T, 0.0), wm(N * N * LS_T, 0.0), reqd(LS_RD_COUNT * LS_T, 0.0);
    std::vector<int> reqi(LS_RI_COUNT * LS_T, 0);
    g_reqd = reqd.data();
    g_reqi = reqi.data();
    LsLane s[LS_T];
    LsRun R[LS_T];
    LsThreadVec v[LS_T];
    long b[LS_T];
    bool done[LS_T];
    for (int t = 0; t < LS_T; t++) {
        v[t].yhp = yh.data() + t;
        v[t].wmp = wm.data() + t;
        v[t].ops = 0;
        b[t] = -1;
    }
    long next = 0;
    const unsigned long long mxstep_u = (unsigned long long)(long long)mxstep;
    const LsTol tol = {rtol, atol, nullptr, nullptr};
    for (;;) {
        bool any = false;
        for (int t = 0; t < LS_T; t++) {
            g_tid = t;
            done[t] = false;
            if (b[t] < 0 && next < B) {
                b[t] = next++;
                done[t] = ls_begin(s[t], R[t], v[t], u0 + b[t] * u0_stride,
                                   (double*)(data + b[t] * data_stride), nt, teval,
                                   usol + (size_t)b[t] * nt * N, tol);
            }
            if (b[t] >= 0) any = true;
            if (b[t] >= 0 && !done[t]) {
                LS_REQI(LS_RI_SCALE_L) = 0;
                LS_REQI(LS_RI_NROWS) = 0;
                done[t] = ls_advance_a(s[t], R[t], v[t], &g_tables, nt, teval, usol + (size_t)b[t] * nt * N, 0u);
            }
        }
        if (!any) break;
        // the service, kind by kind as the block's warps do
        for (int kind = 0; kind < 4; kind++)
            for (int o = 0; o < LS_T; o++)
                if (ls_request_kind(v[o].ops) == kind) {
                    LsThreadVec sv;
                    sv.yhp = yh.data() + o;
                    sv.wmp = wm.data() + o;
                    ls_serve_request(kind, sv, o, (double*)(data + b[o] * data_stride), teval,
                                     usol + (size_t)b[o] * nt * N);
                }
        for (int t = 0; t < LS_T; t++) v[t].ops = 0;
        for (int t = 0; t < LS_T; t++) {
            g_tid = t;
            if (b[t] >= 0 && !done[t])
                done[t] = ls_advance_b(s[t], R[t], v[t], &g_tables, (double*)(data + b[t] * data_stride), tol,
                                       mxstep_u, exit_on_warning, 0u);
            if (b[t] >= 0 && done[t]) {
                success[b[t]] = R[t].istate > 0 ? 1 : 0;
                if (counters) {
                    int* c = counters + b[t] * 8;
                    c[0] = s[t].nst, c[1] = s[t].nfe, c[2] = s[t].nje, c[3] = R[t].nswitch;
                    c[4] = s[t].nqu, c[5] = R[t].istate, c[6] = R[t].row, c[7] = s[t].meth;
                }
                b[t] = -1;
                v[t].ops = 0;
            }
        }
    }
}
