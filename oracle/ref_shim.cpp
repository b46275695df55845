// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// Thin shim around the UNMODIFIED reference LSODA sources.  oracle/Makefile
// compiles it together with /root/reference/src/LSODA.cpp (never copied into this
// repository) into oracle/_ref/libref_probe.so.  It drives the reference exactly
// the way /root/reference/src/wrapper.cpp:10-57 does -- one LSODA object, one
// lsoda_update() call per output time -- and additionally reads the solver's
// private counters so that the restatement in lsoda_oracle.c and the CUDA kernels
// can be checked on step / RHS / Jacobian / method-switch counts.
#define private public
#include "LSODA.h"
#undef private

#include <array>
#include <atomic>
#include <thread>
#include <vector>

extern "C" {

typedef void (*rhs_t)(double, double *, double *, void *);

// stats: [nst, nfe, nje, nswitch(seen at output granularity), nqu, istate, rows, meth]
void ref_lsoda_stats(rhs_t rhs, int neq, const double *u0, void *data, int nt,
                     const double *teval, double *usol, double rtol, double atol, int mxstep,
                     int *success, int exit_on_warning, int *stats)
{
    LSODA solver;
    solver.mxstep = mxstep;
    std::vector<double> y(neq), yout(neq);
    int istate = 1, rows = 1, nswitch = 0;
    size_t meth_prev = 1;
    *success = 1;
    for (int i = 0; i < neq; i++) {
        y[i] = u0[i];
        usol[i] = u0[i];
    }
    double t = teval[0];
    for (int i = 1; i < nt; i++) {
        if (teval[i] < teval[i - 1]) {
            *success = 0;
            istate = 0;
            break;
        }
        solver.lsoda_update(rhs, neq, y, yout, &t, teval[i], &istate, data, rtol, atol,
                            exit_on_warning != 0);
        if (solver.meth_ != meth_prev) {
            nswitch++;
            meth_prev = solver.meth_;
        }
        if (istate <= 0) {
            *success = 0;
            break;
        }
        for (int j = 0; j < neq; j++) {
            y[j] = yout[j + 1];
            usol[j + (size_t)neq * i] = yout[j + 1];
        }
        rows = i + 1;
    }
    if (stats) {
        stats[0] = (int)solver.nst;
        stats[1] = (int)solver.nfe;
        stats[2] = (int)solver.nje;
        stats[3] = nswitch;
        stats[4] = (int)solver.nqu;
        stats[5] = istate;
        stats[6] = rows;
        stats[7] = (int)solver.meth_;
    }
}

// The same loop with per-component tolerances: what lsoda_update (LSODA.cpp:2247-2277) does,
// but with the private tolerance vectors filled per component and itol_ = 4, then the public
// LSODA::lsoda with lsoda_update's fixed options (itask = 1, iopt = 0, jt = 2).
void ref_lsoda_stats_vtol(rhs_t rhs, int neq, const double *u0, void *data, int nt,
                          const double *teval, double *usol, const double *rtol_v,
                          const double *atol_v, int mxstep, int *success, int exit_on_warning,
                          int *stats)
{
    LSODA solver;
    solver.mxstep = mxstep;
    solver.itol_ = 4;
    std::vector<double> y(neq), yout(neq + 1);
    int istate = 1, rows = 1, nswitch = 0;
    size_t meth_prev = 1;
    *success = 1;
    for (int i = 0; i < neq; i++) {
        y[i] = u0[i];
        usol[i] = u0[i];
    }
    double t = teval[0];
    for (int i = 1; i < nt; i++) {
        if (teval[i] < teval[i - 1]) {
            *success = 0;
            istate = 0;
            break;
        }
        std::array<int, 7> iworks = {{0}};
        std::array<double, 4> rworks = {{0.0}};
        solver.rtol_.assign(neq + 1, 0.0);
        solver.atol_.assign(neq + 1, 0.0);
        for (int j = 0; j < neq; j++) {
            solver.rtol_[j + 1] = rtol_v[j];
            solver.atol_[j + 1] = atol_v[j];
            yout[j + 1] = y[j];
        }
        solver.lsoda(rhs, neq, yout, &t, teval[i], 1, &istate, 0, 2, iworks, rworks, data,
                     exit_on_warning != 0);
        if (solver.meth_ != meth_prev) {
            nswitch++;
            meth_prev = solver.meth_;
        }
        if (istate <= 0) {
            *success = 0;
            break;
        }
        for (int j = 0; j < neq; j++) {
            y[j] = yout[j + 1];
            usol[j + (size_t)neq * i] = yout[j + 1];
        }
        rows = i + 1;
    }
    if (stats) {
        stats[0] = (int)solver.nst;
        stats[1] = (int)solver.nfe;
        stats[2] = (int)solver.nje;
        stats[3] = nswitch;
        stats[4] = (int)solver.nqu;
        stats[5] = istate;
        stats[6] = rows;
        stats[7] = (int)solver.meth_;
    }
}

// The reference's own entry point (compiled from wrapper.cpp, same library).
void lsoda_wrapper(rhs_t rhs, int neq, double *u0, void *data, int nt, double *teval,
                   double *usol, double rtol, double atol, int mxstep, int *success,
                   bool exit_on_warning);

// Ensemble loop over the reference entry point: what a numba prange loop around
// numbalsoda.lsoda() executes (README.md:60-67).  Returns the threads used.
int ref_lsoda_batch(rhs_t rhs, long B, int neq, const double *u0, long u0_stride,
                    const void *data, long data_stride, int nt, const double *teval,
                    double *usol, double rtol, double atol, int mxstep, int exit_on_warning,
                    int *success, int nthreads)
{
    if (nthreads <= 0)
        nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1)
        nthreads = 1;
    std::atomic<long> next(0);
    auto work = [&]() {
        for (;;) {
            long b0 = next.fetch_add(8), b1 = b0 + 8;
            if (b0 >= B)
                break;
            if (b1 > B)
                b1 = B;
            for (long b = b0; b < b1; b++) {
                int ok = 0;
                lsoda_wrapper(rhs, neq, const_cast<double *>(u0 + b * u0_stride),
                              (void *)((const char *)data + b * data_stride), nt,
                              const_cast<double *>(teval), usol + (size_t)b * nt * neq, rtol,
                              atol, mxstep, &ok, exit_on_warning != 0);
                success[b] = ok;
            }
        }
    };
    std::vector<std::thread> pool;
    for (int i = 1; i < nthreads; i++)
        pool.emplace_back(work);
    work();
    for (auto &t : pool)
        t.join();
    return nthreads;
}
}
