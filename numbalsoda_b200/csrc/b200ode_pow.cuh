// pow() for the handful of exponents the solvers use (1/k, k = 2..13).
// PLACEHOLDER: CUDA's pow (<= 2 ulp) -- replaced by the glibc-exact version.
#ifndef B200ODE_POW_CUH
#define B200ODE_POW_CUH
__device__ __forceinline__ double b200ode_pow(double x, double y) { return pow(x, y); }
__device__ __forceinline__ double b200ode_pow_eighth(double x) { return pow(x, 0.125); }
#endif
