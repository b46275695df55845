// Test helper: the device pow of numbalsoda_b200/csrc/b200ode_pow.cuh compiled as
// host code, so that it can be compared with the host libm without a GPU.
#include "../numbalsoda_b200/csrc/b200ode_pow.cuh"

extern "C" void b200pow_host(const double* x, double y, double* out, long n)
{
    for (long i = 0; i < n; i++) out[i] = b200ode_pow(x[i], y);
}
extern "C" void libm_pow(const double* x, double y, double* out, long n)
{
    for (long i = 0; i < n; i++) out[i] = pow(x[i], y);
}
