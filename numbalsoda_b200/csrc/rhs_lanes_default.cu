// Default of the optional lane-parallel right-hand side (b200ode_warp.cuh): lane 0 evaluates the
// user's scalar function.  Linked into warp-per-trajectory handles whose right-hand side does not
// provide b200ode_user_rhs_lanes itself.
extern "C" __device__ long long b200ode_user_rhs(double t, double* u, double* du, double* p);
extern "C" __device__ long long b200ode_user_rhs_lanes(double t, double* u, double* du, double* p, int lane,
                                                       int nlanes)
{
    (void)nlanes;
    if (lane == 0) b200ode_user_rhs(t, u, du, p);
    return 0;
}
