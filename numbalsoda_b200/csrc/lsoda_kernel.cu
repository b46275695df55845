// Batched LSODA for sm_100a: one trajectory per thread, systems of N <= 8 equations.
//
// Replaces, for ensembles, the reference's lsoda_wrapper (src/wrapper.cpp:10-57) and
// class LSODA (src/LSODA.cpp); the solver is lsoda_core.cuh (control flow) +
// lsoda_vec_thread.cuh (vector work of one thread's trajectory).  This file is
// the ensemble organisation:
//
//  * a persistent grid (blocks = SMs x resident blocks) pulls trajectories from a
//    global work counter.  LSODA's cost per trajectory varies several-fold over a
//    parameter sweep (stiffness sets the number of steps, Jacobians and method
//    switches), so a lane that finishes starts its next trajectory inside the same
//    step loop rather than idling until the slowest lane of its warp is done;
//  * per-thread solver storage that must be indexed at run time (Nordsieck history,
//    iteration matrix / LU factors) sits in dynamic shared memory as [slot][thread]
//    columns: (13 N + N^2) doubles per thread, bank-conflict free;
//  * the method-coefficient tables are staged once per block into shared memory
//    (lanes of a warp sit at different orders: divergent indices cost bank conflicts
//    there instead of serialised constant-cache replays);
//  * output rows of one trajectory are consecutive in usol[B,nt,N]; each lane streams
//    its rows as they are passed and the L2 merges them into full sectors.
#include "b200ode_device.cuh"

#ifndef B200_MINBLOCKS
#define B200_MINBLOCKS 1
#endif
#ifndef B200_THREADS
#define B200_THREADS 128
#endif
#define LS_STRIDE B200_THREADS
#include "lsoda_vec_thread.cuh"

#define FULL 0xffffffffu
#define LS_TABLE_DOUBLES ((int)(sizeof(B200LsodaTables) / sizeof(double)))

__constant__ B200LsodaTables b200_lsoda_tables = B200_LSODA_TABLES_INIT;

// The tolerances of trajectory b.  Per-component vectors (rtol_vec / atol_vec) are read from
// global memory where they are used (once per step, in ewset); the pointers are rebuilt from
// the launch block and b instead of being carried in registers.
__device__ __forceinline__ LsTol lane_tol(const B200OdeLaunch& L, long long b, double rtol, double atol)
{
    LsTol t;
    t.rtol = rtol;
    t.atol = atol;
    t.rv = L.rtol_vec ? L.rtol_b + b * L.rtol_sb : nullptr;
    t.av = L.atol_vec ? L.atol_b + b * L.atol_sb : nullptr;
    return t;
}

template <bool STAGED>
__device__ __forceinline__ void lsoda_ensemble(const B200OdeLaunch& L)
{
    extern __shared__ double b200_smem[];
    {
        const double* src = reinterpret_cast<const double*>(&b200_lsoda_tables);
        for (int k = threadIdx.x; k < LS_TABLE_DOUBLES; k += B200_THREADS) b200_smem[k] = src[k];
    }
    __syncthreads();
    const B200LsodaTables* T = reinterpret_cast<const B200LsodaTables*>(b200_smem);
    LsLane s;
    LsRun R;
    LsThreadVec v;
    v.yhp = b200_smem + LS_TABLE_DOUBLES + threadIdx.x;
    v.wmp = v.yhp + 13 * N * B200_THREADS;
    double pl[STAGED ? B200ODE_NPMAX : 1];
    double* pp = pl;
    long long b = -1;  // trajectory of this lane, -1 = idle
    bool exhausted = false;
    const int nt = L.nt;
    const double* __restrict__ te = L.t_eval;
    double rtol = L.rtol, atol = L.atol;  // of this lane's trajectory
    // wrapper.cpp:16 stores the int in a size_t member (LSODA.h:130)
    const unsigned long long mxstep_u = (unsigned long long)(long long)L.mxstep;

    for (;;) {
        bool done = false;
        if (b < 0 && !exhausted) {
            const long long bn = b200ode_fetch(L.next);
            if (bn >= L.B) {
                exhausted = true;
            } else {
                b = bn;
                if (L.rtol_b && !L.rtol_vec) rtol = L.rtol_b[b * L.rtol_sb];
                if (L.atol_b && !L.atol_vec) atol = L.atol_b[b * L.atol_sb];
                const double* pg = reinterpret_cast<const double*>(L.data + b * L.data_stride);
                if (STAGED) {
#pragma unroll
                    for (int j = 0; j < B200ODE_NPMAX; j++) pl[j] = (j < L.np) ? pg[j] : 0.0;
                    pp = pl;
                } else {
                    pp = const_cast<double*>(pg);
                }
                done = ls_begin(s, R, v, L.u0 + b * L.u0_stride, pp, nt, te,
                                L.usol + (size_t)b * (size_t)nt * N, lane_tol(L, b, rtol, atol));
            }
        }
        // lanes with a running trajectory take one pass of the state machine together
        const unsigned lanes = __ballot_sync(FULL, b >= 0 && !done);
        if (lanes == 0 && !__any_sync(FULL, b >= 0)) break;
        if (b >= 0 && !done)
            done = ls_advance(s, R, v, T, pp, nt, te, L.usol + (size_t)b * (size_t)nt * N,
                              lane_tol(L, b, rtol, atol), mxstep_u, L.exit_on_warning, lanes);
        if (b >= 0 && done) {
            // wrapper.cpp:41-45: istate <= 0 -> success = 0
            L.success[b] = (R.istate > 0) ? 1 : 0;
            if (L.counters) {
                int* c = L.counters + b * B200ODE_NCOUNTERS;
                c[0] = s.nst;
                c[1] = s.nfe;
                c[2] = s.nje;
                c[3] = R.nswitch;
                c[4] = s.nqu;
                c[5] = R.istate;
                c[6] = R.row;
                c[7] = s.meth;
            }
            b = -1;
        }
    }
}

extern "C" __global__ void __launch_bounds__(B200_THREADS, B200_MINBLOCKS)
b200ode_lsoda_kernel(const B200OdeLaunch L)
{
    lsoda_ensemble<true>(L);
}

extern "C" __global__ void __launch_bounds__(B200_THREADS, B200_MINBLOCKS)
b200ode_lsoda_kernel_ptr(const B200OdeLaunch L)
{
    lsoda_ensemble<false>(L);
}

// Build-time geometry of this image, read once by the host library after loading:
// {doubles of dynamic shared memory per block, 0, min blocks, N, threads}.
extern "C" __global__ void b200ode_lsoda_config(int* out)
{
    out[0] = LS_TABLE_DOUBLES + (13 * N + N * N) * B200_THREADS;
    out[1] = 0;
    out[2] = B200_MINBLOCKS;
    out[3] = N;
    out[4] = B200_THREADS;
}
