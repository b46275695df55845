// Warp-level primitives shared by the warp-per-trajectory kernels (lsoda_vec_warp.cuh,
// dop853_warp.cuh): one warp owns one trajectory, lane i owns components i, i + 32, ...
// The same headers compile as host C++ for the CPU parity harnesses (tests/*_host.cpp):
// there a single "lane" walks every component in order.
#ifndef B200ODE_WARP_CUH
#define B200ODE_WARP_CUH

#if defined(__CUDACC__)
#define WV_FN __device__ __forceinline__
#else
#define WV_FN static inline
#endif

// std::max semantics (a NaN second operand is ignored): the reduction of per-lane
// partial maxima that started from 0 and ignored NaNs the same way is order independent
WV_FN double wv_stdmax(double a, double b) { return (a < b) ? b : a; }

// The user's right-hand side, called from one place: with link-time optimisation every
// call site would get its own inlined copy of the RHS (a method-of-lines RHS is thousands
// of instructions), and the warp kernels -- a few warps per SM, each in a different part
// of a large program -- are bound by instruction fetch before anything else.
#if defined(__CUDA_ARCH__)
extern "C" __device__ long long b200ode_user_rhs(double t, double* u, double* du, double* p);
__device__ __noinline__ void wv_user_rhs(double t, double* u, double* du, double* p)
{
    b200ode_user_rhs(t, u, du, p);
}
// The same right-hand side evaluated by the whole warp: every lane computes the components it
// owns (`lane`, `nlanes` are a partition hint; all nlanes lanes call, and together they must
// write every du[i] exactly as the scalar function does).  A method-of-lines right-hand side is a
// loop over cells; with one trajectory per warp the scalar function kept 1 of 32 lanes busy for
// 26 % of the warp instructions of the banded n = 64 workload (profiles/r2_lsodaw_band_v0.txt).
// Optional: right-hand sides that do not provide it are linked with a default that lets lane 0
// call the scalar function (rhs_lanes_default.cu).
extern "C" __device__ long long b200ode_user_rhs_lanes(double t, double* u, double* du, double* p, int lane,
                                                       int nlanes);
__device__ __noinline__ void wv_user_rhs_all(double t, double* u, double* du, double* p)
{
    __syncwarp();
    b200ode_user_rhs_lanes(t, u, du, p, (int)(threadIdx.x & 31u), 32);
    __syncwarp();
}
#else
#define wv_user_rhs(t, u, du, p) b200ode_user_rhs(t, u, du, p)
#define wv_user_rhs_all(t, u, du, p) b200ode_user_rhs(t, u, du, p)
#endif

#if defined(__CUDA_ARCH__)
#define WV_LANE ((int)(threadIdx.x & 31u))
#define WV_NL 32
#define WV_SYNC() __syncwarp()
WV_FN double wv_bcast(double x, int src) { return __shfl_sync(0xffffffffu, x, src); }
#if defined(WV_MAX_NOINLINE)
__device__ __noinline__
#else
WV_FN
#endif
double wv_max(double x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = wv_stdmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}
WV_FN bool wv_any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
// first position of the largest value: larger v wins, ties go to the smaller index
WV_FN void wv_argmax(unsigned long long& v, int& j)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oj = __shfl_xor_sync(0xffffffffu, j, o);
        if (ov > v || (ov == v && oj < j)) {
            v = ov;
            j = oj;
        }
    }
}
#else
// host build (tests/lsoda_core_host.cpp): one "lane" walks every component in order
#define WV_LANE 0
#define WV_NL 1
#define WV_SYNC() (void)0
WV_FN double wv_bcast(double x, int) { return x; }
WV_FN double wv_max(double x) { return x; }
WV_FN bool wv_any(bool p) { return p; }
WV_FN void wv_argmax(unsigned long long&, int&) {}
#endif

#endif
