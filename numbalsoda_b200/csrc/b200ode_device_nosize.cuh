// Device-side common definitions that do not depend on a compile-time system size
// (see b200ode_device.cuh for the per-size kernels).
#ifndef B200ODE_DEVICE_NOSIZE_CUH
#define B200ODE_DEVICE_NOSIZE_CUH

// The reference's RHS contract  void f(double t, double* u, double* du, void* data)
// (src/LSODA.h:32).  numba's C-ABI device functions return a dummy 64-bit value
// even for a void Python function, hence the return type.
extern "C" __device__ long long b200ode_user_rhs(double t, double* u, double* du, double* p);

// The flags the handle was linked with (include/b200ode.h: B200ODE_STRICT_FP = 1,
// B200ODE_EXACT_CTRL = 2), a link-time constant: link_flags.cu.
extern "C" __device__ unsigned b200ode_link_flags();
#define B200ODE_FAST_CTRL() ((b200ode_link_flags() & 3u) == 0u)

#include "b200ode_launch.h"

__device__ __forceinline__ long long b200ode_fetch(unsigned long long* next) {
    return (long long)atomicAdd(next, 1ULL);
}

#endif
