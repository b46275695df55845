// Batched DOP853 for sm_100a: one WARP per trajectory, systems of 9 <= n <= 128
// equations (dop853_kernel.cu keeps n <= 8 in one thread's registers).
//
// Replaces, for ensembles, the reference's dop853_wrapper (src/dop853_c_interface.f90:75-130)
// and dop853_class (src/dop853_module.f90); the solver is dop853_warp.cuh.  One block =
// one warp; a persistent grid pulls trajectories from a global counter; the trajectory's
// 22 n doubles of stage vectors, dense-output coefficients and norm terms live in dynamic
// shared memory; `data` is passed to the RHS by address; n is a run-time argument (one
// image serves every size).
#include "b200ode_device_nosize.cuh"
#include "dop853_warp.cuh"

extern "C" __global__ void __launch_bounds__(32)
b200ode_dop853w_kernel(const B200OdeLaunch L)
{
    extern __shared__ double b200_smem[];
    const int n = L.neq;
    const int nt = L.nt;
    for (;;) {
        long long b = 0;
        if (WV_LANE == 0) b = b200ode_fetch(L.next);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= L.B) break;
        double* pp = reinterpret_cast<double*>(const_cast<char*>(L.data + b * L.data_stride));
        Dop853WarpStats st;
        dop853_warp_solve(n, b200_smem, L.u0 + b * L.u0_stride, pp, nt, L.t_eval,
                          L.usol + (size_t)b * (size_t)nt * (size_t)n,
                          (L.rtol_b && !L.rtol_vec) ? L.rtol_b[b * L.rtol_sb] : L.rtol,
                          (L.atol_b && !L.atol_vec) ? L.atol_b[b * L.atol_sb] : L.atol, L.mxstep, st,
                          L.rtol_vec ? L.rtol_b + b * L.rtol_sb : nullptr,
                          L.atol_vec ? L.atol_b + b * L.atol_sb : nullptr);
        if (WV_LANE == 0) {
            // c_interface :125-128
            L.success[b] = (st.idid < 0) ? 0 : 1;
            if (L.counters) {
                int* c = L.counters + b * B200ODE_NCOUNTERS_DEV;
                c[0] = st.nstep;
                c[1] = st.idid == -1 ? 0 : 2 + 11 * st.nstep + st.naccpt + 3 * st.nevent;
                c[2] = st.naccpt;
                c[3] = st.nrejct;
                c[4] = st.nevent;
                c[5] = st.idid;
                c[6] = st.rows;
                c[7] = 0;
            }
        }
        __syncwarp();
    }
}
