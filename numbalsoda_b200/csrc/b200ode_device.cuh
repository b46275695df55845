// Device-side common definitions for the batched ODE kernels (sm_100a).
//
// Every thread-per-trajectory solver kernel is compiled by nvcc to LTO-IR once per
// system size (-DB200_NEQ=n) and linked at run time, by nvJitLink, with the user's
// right-hand side, which numba.cuda (or nvcc, for the built-in problems) compiled
// to LTO-IR as well.  Link-time optimisation inlines the right-hand side into the
// stage code, so the state vectors the reference passes by pointer
// (numbalsoda/driver.py:8-11, LSODA.h:32) never leave registers.
#ifndef B200ODE_DEVICE_CUH
#define B200ODE_DEVICE_CUH

#ifndef B200_NEQ
#error "compile with -DB200_NEQ=<system size>"
#endif

#include "b200ode_device_nosize.cuh"
#define B200ODE_NCOUNTERS B200ODE_NCOUNTERS_DEV

#endif
