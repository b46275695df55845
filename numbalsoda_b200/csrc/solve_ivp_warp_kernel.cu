// Batched solve_ivp for sm_100a: one WARP per trajectory, systems of 9 <= n <= 128 equations
// (solve_ivp_kernel.cu keeps n <= 8 in one thread's registers).
//
// GPU counterpart of the reference's dop853_solve_ivp_wrapper / _result
// (src/dop853_solve_ivp.f90:211-434); the solver is solve_ivp_warp.cuh.  One block = one warp;
// a persistent grid pulls trajectories from a global counter; the trajectory's 23 n + 24 doubles
// (stage vectors, dense-output coefficients, norm terms, interpolated state, event values) live
// in dynamic shared memory; `data` is passed to the right-hand side and the events function by
// address; n is a run-time argument (one image serves every size).
#include "b200ode_device_nosize.cuh"
#include "solve_ivp_warp.cuh"

extern "C" __global__ void __launch_bounds__(32)
b200ode_ivpw_kernel(const B200IvpLaunch L, const int n)
{
    extern __shared__ double b200_smem[];
    // '"t_eval" does not always increase or decrease' (dop853_solve_ivp.f90:278-316): t_eval is
    // shared by the ensemble, the direction of integration is per trajectory
    int inc = 1, dec = 1;
    for (int i = 1 + WV_LANE; i < L.nt; i += 32) {
        if (L.t_eval[i] < L.t_eval[i - 1]) inc = 0;
        if (L.t_eval[i] > L.t_eval[i - 1]) dec = 0;
    }
    inc = __all_sync(0xffffffffu, inc);
    dec = __all_sync(0xffffffffu, dec);
    const int te_mono = (inc ? 1 : 0) | (dec ? 2 : 0);
    for (;;) {
        long long b = 0;
        if (WV_LANE == 0) b = b200ode_fetch(L.next);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= L.B) break;
        double* pp = reinterpret_cast<double*>(const_cast<char*>(L.data + b * L.data_stride));
        long long base;
        int cap;
        if (L.nt > 0) {
            base = b * L.nt;
            cap = L.nt;
        } else if (L.row_offset) {
            base = L.row_offset[b];
            cap = (int)(L.row_offset[b + 1] - base);
        } else {
            base = b * L.cap;
            cap = (int)L.cap;
        }
        IvpWarpResult R;
        ivp_warp_solve(n, b200_smem, L, L.t_span + b * L.t_span_stride, L.y0 + b * L.y0_stride, pp,
                       L.rtol_b ? L.rtol_b[b] : L.rtol, L.atol_b ? L.atol_b[b] : L.atol, te_mono, cap,
                       L.t_out + base, L.y_out + base * n, L.y_event + b * n, R);
        if (WV_LANE == 0) {
            int* r = L.result + b * B200ODE_IVP_NRESULT;
            r[0] = R.success;
            r[1] = R.message;
            // nfev stays 0 unless the integration ended normally (dop853_solve_ivp.f90:340-347)
            r[2] = R.idid > 0 ? 2 + 11 * R.nstep + 4 * R.naccpt : 0;
            r[3] = R.nt_out;
            r[4] = R.rows_needed;
            r[5] = R.event_found;
            r[6] = R.ind_event;
            r[7] = R.idid;
            r[8] = R.nstep;
            r[9] = R.naccpt;
            r[10] = R.nrejct;
            r[11] = 0;
            L.t_event[b] = R.event_found ? R.t_event : 0.0;
        }
        __syncwarp();
    }
}
