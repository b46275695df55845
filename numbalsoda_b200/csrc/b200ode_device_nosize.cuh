// Device-side common definitions that do not depend on a compile-time system size
// (see b200ode_device.cuh for the per-size kernels).
#ifndef B200ODE_DEVICE_NOSIZE_CUH
#define B200ODE_DEVICE_NOSIZE_CUH

// The reference's RHS contract  void f(double t, double* u, double* du, void* data)
// (src/LSODA.h:32).  numba's C-ABI device functions return a dummy 64-bit value
// even for a void Python function, hence the return type.
extern "C" __device__ long long b200ode_user_rhs(double t, double* u, double* du, double* p);

#include "b200ode_launch.h"

__device__ __forceinline__ long long b200ode_fetch(unsigned long long* next) {
    return (long long)atomicAdd(next, 1ULL);
}

#endif
