// Batched LSODA for sm_100a: one WARP per trajectory, systems of 9 <= n <= 128
// equations (BASELINE config 4: stiff method-of-lines problems, n = 64).
//
// Replaces, for ensembles, the reference's lsoda_wrapper (src/wrapper.cpp:10-57) and
// class LSODA (src/LSODA.cpp) where the system is too large for one thread: the dense
// Jacobian (n RHS evaluations, LSODA.cpp:1663-1691), its LU factorisation (dgefa,
// :173-231) and the triangular solves (dgesl, :110-170) dominate, and they are what the
// 32 lanes share (lsoda_vec_warp.cuh).  The solver's control flow is the same code as
// the thread-per-trajectory kernel's (lsoda_core.cuh), executed redundantly by every
// lane on identical scalars: no divergence, no inter-lane communication for decisions.
//
//  * one block = one warp; a persistent grid pulls trajectories from a global counter;
//  * all solver storage of the trajectory in dynamic shared memory
//    (B200ODE_LSODAW_DOUBLES(n) doubles: 47 KB at n = 64 -> 4 warps per SM), or -- when the
//    launch provides `scratch` -- the iteration matrix (70 % of it) in a per-block slice of
//    global memory that stays L2 resident, leaving 14 KB per warp -> 16 warps per SM;
//    nothing is ever re-read from HBM;
//  * method-coefficient tables in constant memory (every lane reads the same entry);
//  * `data` is passed to the RHS by address (the record stays in global memory / L1);
//  * n is a run-time argument: one image serves every size.
#include "b200ode_device_nosize.cuh"
#include "lsoda_vec_warp.cuh"

__constant__ B200LsodaTables b200_lsoda_tables = B200_LSODA_TABLES_INIT;

template <bool TM, int BAND>
__device__ __forceinline__ void lsodaw_ensemble(const B200OdeLaunch& L, unsigned tm = 0)
{
    extern __shared__ double b200_smem[];
    const B200LsodaTables* T = &b200_lsoda_tables;
    const int n = L.neq;
    // parallel triangular solves unless the handle was linked for bit parity: a link-time constant
    // (link_flags.cu), so each linked image carries one of the two sets of sweeps only
    const bool kFastSolves = (b200ode_link_flags() & 1u) == 0u;
    LsWarpVecT<TM, BAND> v;
    if (TM) {
        // a block of independent warps: each its own slice of shared and of tensor memory
        v.tm = tm;
        v.attach(b200_smem + (size_t)(threadIdx.x >> 5) * B200ODE_LSODAW_TMEM_DOUBLES(n), n, kFastSolves);
    } else {
        v.attach(b200_smem, n, kFastSolves,
                 L.scratch ? L.scratch + (size_t)blockIdx.x * B200ODE_LSODAW_MATRIX(n) : nullptr, L.ml, L.mu);
    }
    const int nt = L.nt;
    const double* __restrict__ te = L.t_eval;
    // wrapper.cpp:16 stores the int in a size_t member (LSODA.h:130)
    const unsigned long long mxstep_u = (unsigned long long)(long long)L.mxstep;
    for (;;) {
        long long b = 0;
        if (WV_LANE == 0) b = b200ode_fetch(L.next);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= L.B) break;
        double* pp = reinterpret_cast<double*>(const_cast<char*>(L.data + b * L.data_stride));
        double* usol = L.usol + (size_t)b * (size_t)nt * (size_t)n;
        LsTol tol;
        tol.rtol = (L.rtol_b && !L.rtol_vec) ? L.rtol_b[b * L.rtol_sb] : L.rtol;
        tol.atol = (L.atol_b && !L.atol_vec) ? L.atol_b[b * L.atol_sb] : L.atol;
        tol.rv = L.rtol_vec ? L.rtol_b + b * L.rtol_sb : nullptr;
        tol.av = L.atol_vec ? L.atol_b + b * L.atol_sb : nullptr;
        LsLane s;
        LsRun R;
        bool done = ls_begin(s, R, v, L.u0 + b * L.u0_stride, pp, nt, te, usol, tol);
        while (!done)
            done = ls_advance(s, R, v, T, pp, nt, te, usol, tol, mxstep_u,
                              L.exit_on_warning, 0xffffffffu);
        if (WV_LANE == 0) {
            // wrapper.cpp:41-45: istate <= 0 -> success = 0
            L.success[b] = (R.istate > 0) ? 1 : 0;
            if (L.counters) {
                int* c = L.counters + b * B200ODE_NCOUNTERS_DEV;
                c[0] = s.nst;
                c[1] = s.nfe;
                c[2] = s.nje;
                c[3] = R.nswitch;
                c[4] = s.nqu;
                c[5] = R.istate;
                c[6] = R.row;
                c[7] = s.meth;
            }
        }
        __syncwarp();
    }
}

// Two entry points, one body: the dense iteration matrix limits an SM to 4 warps by shared memory
// at n = 64, so that kernel takes the registers it likes (252); the banded one (ml, mu >= 0) fits
// 12 warps and is compiled for that (168 registers).  (The dense bound is explicit -- 8 blocks of
// one warp = 255 registers: left to itself ptxas once settled on 128 registers and 144 bytes of
// spills for this kernel.)
extern "C" __global__ void __launch_bounds__(32, 8)
b200ode_lsodaw_kernel(const B200OdeLaunch L)
{
    lsodaw_ensemble<false, 0>(L);
}

extern "C" __global__ void __launch_bounds__(32, 12)
b200ode_lsodaw_band_kernel(const B200OdeLaunch L)
{
    lsodaw_ensemble<false, 1>(L);
}

// Dense matrix in tensor memory (lsoda_vec_warp.cuh), 32 < n <= 64: one block of eight warps per
// SM, each warp a trajectory of its own.  Warp 0 allocates all 512 columns; warp w works in the
// lanes of its quadrant (w mod 4) and in the column half w / 4.  The only block-wide
// synchronisation is around the allocation and the release.
extern "C" __global__ void __launch_bounds__(32 * B200ODE_LSODAW_TMEM_WARPS, 1)
b200ode_lsodaw_tmem_kernel(const B200OdeLaunch L)
{
    __shared__ unsigned tm_base;
    const unsigned warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(
            (unsigned)__cvta_generic_to_shared(&tm_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    lsodaw_ensemble<true, 0>(L, tm_base + (((warp & 3u) * 32u) << 16) + (warp >> 2) * 256u);
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tm_base));
}
