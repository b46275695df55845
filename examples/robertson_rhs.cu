// Right-hand side of the Robertson problem (reference benchmark/benchmark_robber.py:8-14)
// as a device function with the contract of include/b200ode.h.  Build to an LTO fatbin:
//   nvcc -gencode arch=compute_100a,code=lto_100a -rdc=true -fatbin -o robertson_rhs.fatbin robertson_rhs.cu
extern "C" __device__ long long b200ode_user_rhs(double t, double* u, double* du, double* p)
{
    const double k1 = p[0], k2 = p[1], k3 = p[2];
    du[0] = -k1 * u[0] + k3 * u[1] * u[2];
    du[1] = k1 * u[0] - k2 * (u[1] * u[1]) - k3 * u[1] * u[2];
    du[2] = k2 * (u[1] * u[1]);
    (void)t;
    return 0;
}
