// Batched solve_ivp for sm_100a: one trajectory per thread (n <= 8), persistent grid.
//
// GPU counterpart of the reference's dop853_solve_ivp_wrapper / _result
// (src/dop853_solve_ivp.f90:211-434; Python side numbalsoda/driver_solve_ivp.py:86-218):
// DOP853 with a user-given or solver-chosen output grid, forward or backward, first_step /
// max_step, terminal events located by Brent's method on the dense interpolant.  The
// per-trajectory algorithm is solve_ivp_core.cuh; this file is the ensemble around it:
//
//  * a persistent grid pulls trajectories from a global work counter; a lane that finishes
//    (end of span, event, failure) starts its next trajectory inside the same step loop, so
//    trajectories cut short by an event cost no idle lanes;
//  * the user's right-hand side and events function are inlined by link-time optimisation
//    (b200ode_user_rhs, b200ode_user_events: LTO-IR from numba.cuda or nvcc);
//  * rows of a trajectory are consecutive in memory ([B,cap,n]: each y_out[b] is the
//    reference's ODEResult.y), so they stream out sequentially per lane and L2 merges them
//    into full sectors.
#include "b200ode_device.cuh"

// The reference's events contract  void events(double t, double* y, double* out, void* data)
// (dop853_solve_ivp.f90:9-17), compiled like the right-hand side.  Handles linked without
// events get an empty one (events_builtin.cu: none) that LTO removes.
extern "C" __device__ long long b200ode_user_events(double t, double* u, double* ev, double* p);

#include "solve_ivp_core.cuh"

#define N B200_NEQ
#ifndef B200_THREADS
#define B200_THREADS 128
#endif
#ifndef B200_MINBLOCKS
#define B200_MINBLOCKS 1
#endif

template <bool STAGED>
__device__ __forceinline__ void ivp_ensemble(const B200IvpLaunch& L)
{
    extern __shared__ double b200_smem[];
    IvpShared sh;
    sh.cont = b200_smem + threadIdx.x;
    sh.ev_old = b200_smem + 8 * N * B200_THREADS + threadIdx.x;
    sh.cs = B200_THREADS;
    // '"t_eval" does not always increase or decrease' (dop853_solve_ivp.f90:278-316): t_eval is
    // shared by the ensemble, the direction of integration is per trajectory
    int inc = 1, dec = 1;
    for (int i = 1 + threadIdx.x; i < L.nt; i += B200_THREADS) {
        if (L.t_eval[i] < L.t_eval[i - 1]) inc = 0;
        if (L.t_eval[i] > L.t_eval[i - 1]) dec = 0;
    }
    inc = __syncthreads_and(inc);
    dec = __syncthreads_and(dec);
    const int te_mono = inc | (dec << 1);
    IvpLane<N> S;
    double pl[STAGED ? B200ODE_NPMAX : 1];
    long long b = -1;
    for (;;) {
        if (b < 0) {
            b = b200ode_fetch(L.next);
            if (b >= L.B) break;
            const double* pg = reinterpret_cast<const double*>(L.data + b * L.data_stride);
            S.pg = pg;
            if (STAGED) {
#pragma unroll
                for (int j = 0; j < B200ODE_NPMAX; j++) pl[j] = (j < L.np) ? pg[j] : 0.0;
                S.pp = pl;
            } else {
                S.pp = const_cast<double*>(pg);
            }
            if (ivp_begin<N>(S, L, b, sh, te_mono)) {
                b = -1;
                continue;
            }
        }
        if (ivp_step<N, STAGED>(S, L, b, sh)) b = -1;
    }
}

extern "C" __global__ void __launch_bounds__(B200_THREADS, B200_MINBLOCKS)
b200ode_ivp_kernel(const B200IvpLaunch L)
{
    ivp_ensemble<true>(L);
}

extern "C" __global__ void __launch_bounds__(B200_THREADS, B200_MINBLOCKS)
b200ode_ivp_kernel_ptr(const B200IvpLaunch L)
{
    ivp_ensemble<false>(L);
}

// Build-time geometry of this image, read once by the host library after loading.
extern "C" __global__ void b200ode_ivp_config(int* out)
{
    out[0] = B200ODE_IVP_DOUBLES(N) * B200_THREADS;  // doubles of dynamic shared memory
    out[1] = B200ODE_IVP_NEVMAX;
    out[2] = B200_MINBLOCKS;
    out[3] = N;
    out[4] = B200_THREADS;
    out[5] = 0;
}
