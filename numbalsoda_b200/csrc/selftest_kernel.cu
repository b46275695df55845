// Device self-test kernels (not on the solver path): evaluate the device pow on an
// array so that tests can compare it bit-for-bit with the host libm.
#include "b200ode_pow.cuh"

extern "C" __global__ void b200ode_selftest_pow_kernel(const double* x, double y, double* out,
                                                       long long n)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = b200ode_pow(x[i], y);
}
