// Device-side common definitions for the batched ODE kernels (sm_100a).
//
// Every solver kernel is compiled by nvcc to LTO-IR once per system size
// (-DB200_NEQ=n) and linked at run time, by nvJitLink, with the user's
// right-hand side, which numba.cuda (or nvcc, for the built-in problems) compiled
// to LTO-IR as well.  Link-time optimisation inlines the right-hand side into the
// stage code, so the state vectors the reference passes by pointer
// (numbalsoda/driver.py:8-11, LSODA.h:32) never leave registers.
#ifndef B200ODE_DEVICE_CUH
#define B200ODE_DEVICE_CUH

#ifndef B200_NEQ
#error "compile with -DB200_NEQ=<system size>"
#endif

// The reference's RHS contract  void f(double t, double* u, double* du, void* data)
// (src/LSODA.h:32).  numba's C-ABI device functions return a dummy 64-bit value
// even for a void Python function, hence the return type.
extern "C" __device__ long long b200ode_user_rhs(double t, double* u, double* du, double* p);

#include "b200ode_launch.h"
#define B200ODE_NCOUNTERS B200ODE_NCOUNTERS_DEV

__device__ __forceinline__ long long b200ode_fetch(unsigned long long* next) {
    return (long long)atomicAdd(next, 1ULL);
}

#endif
