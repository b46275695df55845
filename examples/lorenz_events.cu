// Right-hand side and events function of examples/lorenz_events.c as device functions with the
// contracts of include/b200ode.h (the reference's: benchmark/benchmark_lorenz.py:9-16 and
// src/dop853_solve_ivp.f90:9-17).  Build each to an LTO fatbin:
//   nvcc -gencode arch=compute_100a,code=lto_100a -rdc=true -fatbin -DLORENZ_RHS    -o lorenz_rhs.fatbin    lorenz_events.cu
//   nvcc -gencode arch=compute_100a,code=lto_100a -rdc=true -fatbin -DLORENZ_EVENTS -o lorenz_events.fatbin lorenz_events.cu
#if defined(LORENZ_RHS)
extern "C" __device__ long long b200ode_user_rhs(double t, double* u, double* du, double* p)
{
    const double sigma = p[0], rho = p[1], beta = p[2];
    const double x = u[0], y = u[1], z = u[2];
    du[0] = sigma * (y - x);
    du[1] = x * (rho - z) - y;
    du[2] = x * y - beta * z;
    (void)t;
    return 0;
}
#elif defined(LORENZ_EVENTS)
// two terminal events: z reaches the level in p[3]; x changes sign
extern "C" __device__ long long b200ode_user_events(double t, double* u, double* out, double* p)
{
    out[0] = u[2] - p[3];
    out[1] = u[0];
    (void)t;
    return 0;
}
#else
#error "define LORENZ_RHS or LORENZ_EVENTS"
#endif
