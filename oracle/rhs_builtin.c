/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * Right-hand sides of the benchmark problems, written so that every floating
 * point operation happens in the order numba emits for the reference's Python
 * definitions (numba/LLVM neither re-associates nor contracts by default):
 *   Lotka-Volterra  /root/reference/README.md:26-29
 *   Lorenz          /root/reference/benchmark/benchmark_lorenz.py:9-16
 *   Robertson       /root/reference/benchmark/benchmark_robber.py:8-14
 * plus the n = 64 method-of-lines problem this repository defines for config C4
 * (no reference shape exists for it).
 */
#include <string.h>

#include "oracle.h"

void oracle_rhs_lotka_volterra(double t, double *u, double *du, void *pv)
{
    const double *p = (const double *)pv;
    (void)t;
    du[0] = u[0] - u[0] * u[1];
    du[1] = u[0] * u[1] - u[1] * p[0];
}

/* /root/reference/solve_ivp.ipynb cell 21 with np.sum(p.arr1), np.sum(p.arr2) in p[0], p[1] */
void oracle_rhs_lotka_volterra2(double t, double *u, double *du, void *pv)
{
    const double *p = (const double *)pv;
    (void)t;
    du[0] = u[0] - u[0] * u[1] * p[0];
    du[1] = u[0] * u[1] - u[1] * p[1];
}

void oracle_rhs_lorenz(double t, double *u, double *du, void *pv)
{
    const double *p = (const double *)pv;
    const double sigma = p[0], rho = p[1], beta = p[2];
    const double x = u[0], y = u[1], z = u[2];
    (void)t;
    du[0] = sigma * (y - x);
    du[1] = x * (rho - z) - y;
    du[2] = x * y - beta * z;
}

void oracle_rhs_robertson(double t, double *u, double *du, void *pv)
{
    const double *p = (const double *)pv;
    const double k1 = p[0], k2 = p[1], k3 = p[2];
    (void)t;
    du[0] = -k1 * u[0] + k3 * u[1] * u[2];
    du[1] = k1 * u[0] - k2 * (u[1] * u[1]) - k3 * u[1] * u[2];
    du[2] = k2 * (u[1] * u[1]);
}

/* 1-D Brusselator, 32 cells, 2 species, Dirichlet walls u = A, v = B.
 * State layout: u_0..u_31, v_0..v_31.  p = [A, B, D] with D = alpha/dx^2. */
void oracle_rhs_brusselator1d(double t, double *y, double *dy, void *pv)
{
    const double *p = (const double *)pv;
    const double A = p[0], B = p[1], D = p[2];
    const int m = 32;
    (void)t;
    for (int i = 0; i < m; i++) {
        const double u = y[i], v = y[m + i];
        const double ul = (i == 0) ? A : y[i - 1];
        const double ur = (i == m - 1) ? A : y[i + 1];
        const double vl = (i == 0) ? B : y[m + i - 1];
        const double vr = (i == m - 1) ? B : y[m + i + 1];
        const double uuv = u * u * v;
        dy[i] = A + uuv - (B + 1.0) * u + D * (ul - 2.0 * u + ur);
        dy[m + i] = B * u - uuv + D * (vl - 2.0 * v + vr);
    }
}

/* The same problem with the two species interleaved, y = (u_0, v_0, u_1, v_1, ...): the
 * Jacobian is then banded with ml = mu = 2 (the layout a method-of-lines user picks for a
 * banded solver); same operations per cell as above. */
void oracle_rhs_brusselator1d_il(double t, double *y, double *dy, void *pv)
{
    const double *p = (const double *)pv;
    const double A = p[0], B = p[1], D = p[2];
    const int m = 32;
    (void)t;
    for (int i = 0; i < m; i++) {
        const double u = y[2 * i], v = y[2 * i + 1];
        const double ul = (i == 0) ? A : y[2 * i - 2];
        const double ur = (i == m - 1) ? A : y[2 * i + 2];
        const double vl = (i == 0) ? B : y[2 * i - 1];
        const double vr = (i == m - 1) ? B : y[2 * i + 3];
        const double uuv = u * u * v;
        dy[2 * i] = A + uuv - (B + 1.0) * u + D * (ul - 2.0 * u + ur);
        dy[2 * i + 1] = B * u - uuv + D * (vl - 2.0 * v + vr);
    }
}

/* Events functions of solve_ivp (numbalsoda_b200/csrc/events_builtin.cu):
 *   lv_prey   /root/reference/solve_ivp.ipynb cell 9: prey population u[0] crosses 1
 *   lorenz_z  z crosses the level in the 4th parameter */
void oracle_events_lv_prey(double t, double *u, double *out, void *pv)
{
    (void)t, (void)pv;
    out[0] = u[0] - 1;
}

void oracle_events_lorenz_z(double t, double *u, double *out, void *pv)
{
    const double *p = (const double *)pv;
    (void)t;
    out[0] = u[2] - p[3];
}

oracle_rhs_fn oracle_builtin_rhs(const char *name)
{
    if (!strcmp(name, "events_lv_prey"))
        return oracle_events_lv_prey;
    if (!strcmp(name, "events_lorenz_z"))
        return oracle_events_lorenz_z;
    if (!strcmp(name, "lotka_volterra"))
        return oracle_rhs_lotka_volterra;
    if (!strcmp(name, "lotka_volterra2"))
        return oracle_rhs_lotka_volterra2;
    if (!strcmp(name, "lorenz"))
        return oracle_rhs_lorenz;
    if (!strcmp(name, "robertson"))
        return oracle_rhs_robertson;
    if (!strcmp(name, "brusselator1d"))
        return oracle_rhs_brusselator1d;
    if (!strcmp(name, "brusselator1d_il"))
        return oracle_rhs_brusselator1d_il;
    return 0;
}
