// Built-in events functions of solve_ivp (device code), one per translation: compile with
// -DB200_EVENTS_<NAME>.  Same name and signature as a user's events function after
// numba.cuda compiled it (solve_ivp_kernel.cu); the reference's contract is
// void events(double t, double* y, double* out, void* data), src/dop853_solve_ivp.f90:9-17.
//   none      a handle linked without events (n_events = 0): never called, removed by LTO
//   lv_prey   /root/reference/solve_ivp.ipynb cell 9: prey population u[0] crosses 1
//   lorenz_z  z crosses the level in the 4th parameter: data = [sigma, rho, beta, level]
extern "C" __device__ long long b200ode_user_events(double t, double* u, double* out, double* p)
{
#if defined(B200_EVENTS_NONE)
    (void)u, (void)out, (void)p;
#elif defined(B200_EVENTS_LV_PREY)
    out[0] = u[0] - 1;
    (void)p;
#elif defined(B200_EVENTS_LORENZ_Z)
    out[0] = u[2] - p[3];
#else
#error "define one of B200_EVENTS_*"
#endif
    (void)t;
    return 0;
}
