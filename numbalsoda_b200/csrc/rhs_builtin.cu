// Built-in right-hand sides (device code), one per translation: compile with
// -DB200_RHS_<NAME>.  They carry the name and signature a user RHS has after
// numba.cuda compiled it (b200ode_device.cuh) and perform the same IEEE operations,
// in the same order, as the reference's Python definitions:
//   Lotka-Volterra  /root/reference/README.md:26-29
//   Lorenz          /root/reference/benchmark/benchmark_lorenz.py:9-16
//   Robertson       /root/reference/benchmark/benchmark_robber.py:8-14
// brusselator1d is this repository's n = 64 method-of-lines problem (config C4).
extern "C" __device__ long long b200ode_user_rhs(double t, double* u, double* du, double* p)
{
#if defined(B200_RHS_LOTKA_VOLTERRA)
    du[0] = u[0] - u[0] * u[1];
    du[1] = u[0] * u[1] - u[1] * p[0];
#elif defined(B200_RHS_LOTKA_VOLTERRA2)
    // /root/reference/solve_ivp.ipynb cell 21 with the two sums in p[0], p[1]
    du[0] = u[0] - u[0] * u[1] * p[0];
    du[1] = u[0] * u[1] - u[1] * p[1];
#elif defined(B200_RHS_LORENZ)
    const double sigma = p[0], rho = p[1], beta = p[2];
    const double x = u[0], y = u[1], z = u[2];
    du[0] = sigma * (y - x);
    du[1] = x * (rho - z) - y;
    du[2] = x * y - beta * z;
#elif defined(B200_RHS_ROBERTSON)
    const double k1 = p[0], k2 = p[1], k3 = p[2];
    du[0] = -k1 * u[0] + k3 * u[1] * u[2];
    du[1] = k1 * u[0] - k2 * (u[1] * u[1]) - k3 * u[1] * u[2];
    du[2] = k2 * (u[1] * u[1]);
#elif defined(B200_RHS_BRUSSELATOR1D)
    const double A = p[0], B = p[1], D = p[2];
    const int m = 32;
    // (not unrolled: in the warp-per-trajectory kernels one lane evaluates this, and code
    // size -- instruction fetch -- is what binds them; numba emits a loop as well)
#pragma unroll 1
    for (int i = 0; i < m; i++) {
        const double uu = u[i], v = u[m + i];
        const double ul = (i == 0) ? A : u[i - 1];
        const double ur = (i == m - 1) ? A : u[i + 1];
        const double vl = (i == 0) ? B : u[m + i - 1];
        const double vr = (i == m - 1) ? B : u[m + i + 1];
        const double uuv = uu * uu * v;
        du[i] = A + uuv - (B + 1.0) * uu + D * (ul - 2.0 * uu + ur);
        du[m + i] = B * uu - uuv + D * (vl - 2.0 * v + vr);
    }
#elif defined(B200_RHS_BRUSSELATOR1D_IL)
    // the same problem with the species interleaved, y = (u_0, v_0, u_1, v_1, ...): the
    // Jacobian is banded with ml = mu = 2 (the layout for lsoda(..., ml=2, mu=2))
    const double A = p[0], B = p[1], D = p[2];
    const int m = 32;
#pragma unroll 1
    for (int i = 0; i < m; i++) {
        const double uu = u[2 * i], v = u[2 * i + 1];
        const double ul = (i == 0) ? A : u[2 * i - 2];
        const double ur = (i == m - 1) ? A : u[2 * i + 2];
        const double vl = (i == 0) ? B : u[2 * i - 1];
        const double vr = (i == m - 1) ? B : u[2 * i + 3];
        const double uuv = uu * uu * v;
        du[2 * i] = A + uuv - (B + 1.0) * uu + D * (ul - 2.0 * uu + ur);
        du[2 * i + 1] = B * uu - uuv + D * (vl - 2.0 * v + vr);
    }
#else
#error "define one of B200_RHS_*"
#endif
    (void)t;
    return 0;
}

// The lane-parallel form (b200ode_warp.cuh): the method-of-lines problems split their loop over
// cells, cell i on lane i % nlanes -- the same operations per component as above, so strict mode
// stays bit-exact; the small systems never run on the warp kernels and let lane 0 do the work.
extern "C" __device__ long long b200ode_user_rhs_lanes(double t, double* u, double* du, double* p, int lane,
                                                       int nlanes)
{
#if defined(B200_RHS_BRUSSELATOR1D)
    const double A = p[0], B = p[1], D = p[2];
    const int m = 32;
#pragma unroll 1
    for (int i = lane; i < m; i += nlanes) {
        const double uu = u[i], v = u[m + i];
        const double ul = (i == 0) ? A : u[i - 1];
        const double ur = (i == m - 1) ? A : u[i + 1];
        const double vl = (i == 0) ? B : u[m + i - 1];
        const double vr = (i == m - 1) ? B : u[m + i + 1];
        const double uuv = uu * uu * v;
        du[i] = A + uuv - (B + 1.0) * uu + D * (ul - 2.0 * uu + ur);
        du[m + i] = B * uu - uuv + D * (vl - 2.0 * v + vr);
    }
#elif defined(B200_RHS_BRUSSELATOR1D_IL)
    const double A = p[0], B = p[1], D = p[2];
    const int m = 32;
#pragma unroll 1
    for (int i = lane; i < m; i += nlanes) {
        const double uu = u[2 * i], v = u[2 * i + 1];
        const double ul = (i == 0) ? A : u[2 * i - 2];
        const double ur = (i == m - 1) ? A : u[2 * i + 2];
        const double vl = (i == 0) ? B : u[2 * i - 1];
        const double vr = (i == m - 1) ? B : u[2 * i + 3];
        const double uuv = uu * uu * v;
        du[2 * i] = A + uuv - (B + 1.0) * uu + D * (ul - 2.0 * uu + ur);
        du[2 * i + 1] = B * uu - uuv + D * (vl - 2.0 * v + vr);
    }
#else
    (void)nlanes;
    if (lane == 0) b200ode_user_rhs(t, u, du, p);
#endif
    (void)t;
    return 0;
}
