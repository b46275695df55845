// Batched LSODA for sm_100a: one trajectory per thread, systems of N <= 8 equations.
//
// Replaces, for ensembles, the reference's lsoda_wrapper (src/wrapper.cpp:10-57) and
// class LSODA (src/LSODA.cpp); the solver is lsoda_core.cuh (control flow) +
// lsoda_vec_thread.cuh (vector work of one thread's trajectory).  This file is
// the ensemble organisation:
//
//  * a persistent grid (blocks = SMs x resident blocks) pulls trajectories from a
//    global work counter.  LSODA's cost per trajectory varies several-fold over a
//    parameter sweep (stiffness sets the number of steps, Jacobians and method
//    switches), so a lane that finishes starts its next trajectory inside the same
//    step loop rather than idling until the slowest lane of its warp is done;
//  * per-thread solver storage that must be indexed at run time (Nordsieck history,
//    iteration matrix / LU factors) sits in dynamic shared memory as [slot][thread]
//    columns: (13 N + N^2) doubles per thread, bank-conflict free;
//  * the method-coefficient tables are staged once per block into shared memory
//    (lanes of a warp sit at different orders: divergent indices cost bank conflicts
//    there instead of serialised constant-cache replays);
//  * block-level compaction of the rare blocks (north_star: "sorting and compacting lanes by
//    status between step bursts") is built and MEASURED SLOWER, so it is a build option
//    (-DLS_COMPACT), not the default.  The blocks of a pass that only a few lanes of a warp need
//    at a time -- Jacobian + LU (9 % of the lanes per pass), retraction after a failed attempt
//    (10 %), rescale after a step-size change (25 %), output rows (18 % of the steps) -- run
//    with 2-4 of 32 lanes and are 28 % of the warp instructions
//    (profiles/r2_lsoda_v0_baseline.txt).  With LS_COMPACT they are *requested*
//    (lsoda_vec_thread.cuh, LS_DEFERRED) and served once per pass for the whole block, between
//    the two halves of the pass: every thread with a request enters the queue of its kind, and
//    each warp of the block serves one queue, request r on lane r.  The server needs nothing
//    from the requester's registers -- yh and the matrix are shared-memory columns it addresses
//    by the owner's thread index, the few scalars travel in the request -- and performs the same
//    operations in the same order, so strict mode stays bit-exact (all 150 GPU tests pass with
//    it, and tests/test_lsoda_deferred_host.py emulates it on the CPU).  What it buys and costs
//    (profiles/r2_lsoda_v1_deferred.txt, Robertson): the served blocks shrink from 4.4 G to 2.0 G
//    warp instructions and SIMT efficiency rises from 11.8 to 13.9 lanes, but a block of 128
//    trajectories only has 5-12 requests of a kind per pass (served at 5-13 lanes, not 32), the
//    queues cost 0.6 G, the lanes that wait for a matrix re-enter the corrector, and the two
//    block barriers per pass show up as the third-largest stall (0.85 warps per issue cycle):
//    16.5 G instructions against 16.1 G, 8.02 M trajectories/s against 8.66 M.  Larger blocks
//    (more requests per queue) lose more to the barriers: 192 threads 8.14 M, 384 threads
//    7.64 M (profiles/r2_sweep_lsoda.log).  The kernel is bound by the latency of dependent
//    instructions (stall_wait 2.5 warps per issue cycle at 3 warps per scheduler), not by issue
//    slots, so removing low-lane instructions does not shorten a pass while barriers lengthen it;
//  * output rows of one trajectory are consecutive in usol[B,nt,N]; a trajectory's rows
//    stream out as they are passed and the L2 merges them into full sectors.
#include "b200ode_device.cuh"

#ifndef B200_MINBLOCKS
#define B200_MINBLOCKS 1
#endif
#ifndef B200_THREADS
#define B200_THREADS 128
#endif
#define LS_STRIDE B200_THREADS
#ifdef LS_COMPACT
#define LS_DEFERRED 1
#endif

#include "lsoda_tables.cuh"
#define LS_TABLE_DOUBLES ((int)(sizeof(B200LsodaTables) / sizeof(double)))
// dynamic shared memory of a block, in doubles:
//   tables | yh [13 N][T] | wm [N N][T] | request doubles [6][T] | request ints [9][T],
//   queues [4][T], queue lengths [2][4]
#define N_ B200_NEQ
#define LS_OFF_YH LS_TABLE_DOUBLES
#define LS_OFF_WM (LS_OFF_YH + 13 * N_ * B200_THREADS)
#define LS_OFF_REQD (LS_OFF_WM + N_ * N_ * B200_THREADS)
#define LS_OFF_REQI (LS_OFF_REQD + 6 * B200_THREADS)
#define LS_REQI_INTS (9 * B200_THREADS + 4 * B200_THREADS + 8)
#ifdef LS_COMPACT
#define LS_SMEM_DOUBLES (LS_OFF_REQI + (LS_REQI_INTS + 1) / 2)
#else
#define LS_SMEM_DOUBLES LS_OFF_REQD  // (the request area and the queues exist with LS_COMPACT only)
#endif
extern __shared__ double b200_smem[];
// this thread's request slots (lsoda_vec_thread.cuh)
#define LS_REQD(f) b200_smem[LS_OFF_REQD + (f) * B200_THREADS + threadIdx.x]
#define LS_REQI(f) reinterpret_cast<int*>(b200_smem + LS_OFF_REQI)[(f) * B200_THREADS + threadIdx.x]
// ... and those of the thread `o` a server works for
#define LS_REQD_OF(o, f) b200_smem[LS_OFF_REQD + (f) * B200_THREADS + (o)]
#define LS_REQI_OF(o, f) reinterpret_cast<int*>(b200_smem + LS_OFF_REQI)[(f) * B200_THREADS + (o)]
#include "lsoda_vec_thread.cuh"

#define FULL 0xffffffffu

__constant__ B200LsodaTables b200_lsoda_tables = B200_LSODA_TABLES_INIT;

// The tolerances of trajectory b.  Per-component vectors (rtol_vec / atol_vec) are read from
// global memory where they are used (once per step, in ewset); the pointers are rebuilt from
// the launch block and b instead of being carried in registers.
__device__ __forceinline__ LsTol lane_tol(const B200OdeLaunch& L, long long b)
{
    LsTol t;
    // (scalars from the launch block -- the constant bank -- or this trajectory's entry, read
    // where they are used, once per step: two registers less than carrying them)
    t.rtol = (L.rtol_b && !L.rtol_vec) ? L.rtol_b[b * L.rtol_sb] : L.rtol;
    t.atol = (L.atol_b && !L.atol_vec) ? L.atol_b[b * L.atol_sb] : L.atol;
    t.rv = L.rtol_vec ? L.rtol_b + b * L.rtol_sb : nullptr;
    t.av = L.atol_vec ? L.atol_b + b * L.atol_sb : nullptr;
    return t;
}

#if defined(LS_DEFERRED)
#define LS_NKINDS 4
// ---- the block's server: queue `kind` is served by one warp, request r on lane r
//   0  Jacobian + LU                          (prja, dgefa: LSODA.cpp:1633-1696, :173-231)
//   1  retraction, then the rescale if any    (:1290-1295 / :1909-1912, scaleh :1621-1627)
//   2  rescale, then the output rows if any   (scaleh; intdy :1435-1453)
//   3  output rows only
template <bool STAGED>
__device__ __forceinline__ void ls_serve(const B200OdeLaunch& L, const int* q, const int* qn)
{
    const int lane = threadIdx.x & 31;
    // blocks of at least four warps: warp w serves kind w & 3, every (w >> 2)-th batch of 32
    // requests of it; smaller blocks: warp w serves kinds w, w + warps, ...
    constexpr int kWarps = B200_THREADS / 32;
    constexpr int kGroups = kWarps >= LS_NKINDS ? kWarps / LS_NKINDS : 1;
    const int w = threadIdx.x >> 5;
    if (kWarps >= LS_NKINDS && w >= kGroups * LS_NKINDS) return;
#pragma unroll 1
    for (int kind = kWarps >= LS_NKINDS ? (w & (LS_NKINDS - 1)) : w; kind < LS_NKINDS;
         kind += kWarps >= LS_NKINDS ? LS_NKINDS : kWarps) {
        const int cnt = qn[kind];
        const int r0 = kWarps >= LS_NKINDS ? lane + 32 * (w / LS_NKINDS) : lane;
#pragma unroll 1
        for (int r = r0; r < cnt; r += 32 * kGroups) {
            const int o = q[kind * B200_THREADS + r];  // the owner's thread index
            LsThreadVec sv;  // a view of the owner's shared-memory columns (its registers stay unused)
            sv.yhp = b200_smem + LS_OFF_YH + o;
            sv.wmp = b200_smem + LS_OFF_WM + o;
            const long long bo = ((long long)LS_REQI_OF(o, LS_RI_BHI) << 32) | (unsigned)LS_REQI_OF(o, LS_RI_BLO);
            double pl2[STAGED ? B200ODE_NPMAX : 1];
            double* pp2 = pl2;
            if (kind == 0) {
                const double* pg = reinterpret_cast<const double*>(L.data + bo * L.data_stride);
                if (STAGED) {
#pragma unroll
                    for (int j = 0; j < B200ODE_NPMAX; j++) pl2[j] = (j < L.np) ? pg[j] : 0.0;
                } else {
                    pp2 = const_cast<double*>(pg);
                }
            }
            ls_serve_request(kind, sv, o, pp2, L.t_eval, L.usol + (size_t)bo * (size_t)L.nt * N);
        }
    }
}
#endif

template <bool STAGED>
__device__ __forceinline__ void lsoda_ensemble(const B200OdeLaunch& L)
{
    {
        const double* src = reinterpret_cast<const double*>(&b200_lsoda_tables);
        for (int k = threadIdx.x; k < LS_TABLE_DOUBLES; k += B200_THREADS) b200_smem[k] = src[k];
    }
#if defined(LS_DEFERRED)
    int* const q = reinterpret_cast<int*>(b200_smem + LS_OFF_REQI) + 9 * B200_THREADS;
    int* const qn = q + LS_NKINDS * B200_THREADS;  // [2][LS_NKINDS]: lengths, double-buffered
    if (threadIdx.x < 2 * LS_NKINDS) qn[threadIdx.x] = 0;
#endif
    __syncthreads();
    const B200LsodaTables* T = reinterpret_cast<const B200LsodaTables*>(b200_smem);
    LsLane s;
    LsRun R;
    LsThreadVec v;
    v.yhp = b200_smem + LS_OFF_YH + threadIdx.x;
    v.wmp = b200_smem + LS_OFF_WM + threadIdx.x;
#if defined(LS_DEFERRED)
    v.ops = 0;
#endif
    double pl[STAGED ? B200ODE_NPMAX : 1];
    double* pp = pl;
    long long b = -1;  // trajectory of this lane; -1 = idle, -2 = idle and the work counter has run out
    const int nt = L.nt;
    const double* __restrict__ te = L.t_eval;
    // wrapper.cpp:16 stores the int in a size_t member (LSODA.h:130)
#define MXSTEP_U ((unsigned long long)(long long)L.mxstep)

    for (unsigned pass = 0;; pass++) {
        bool done = false;
        if (b == -1) {
            const long long bn = b200ode_fetch(L.next);
            if (bn >= L.B) {
                b = -2;
            } else {
                b = bn;
                const double* pg = reinterpret_cast<const double*>(L.data + b * L.data_stride);
                if (STAGED) {
#pragma unroll
                    for (int j = 0; j < B200ODE_NPMAX; j++) pl[j] = (j < L.np) ? pg[j] : 0.0;
                    pp = pl;
                } else {
                    pp = const_cast<double*>(pg);
                }
#if defined(LS_DEFERRED)
                LS_REQI(LS_RI_BLO) = (int)(unsigned)(b & 0xffffffffLL);
                LS_REQI(LS_RI_BHI) = (int)(b >> 32);
#endif
                done = ls_begin(s, R, v, L.u0 + b * L.u0_stride, pp, nt, te,
                                L.usol + (size_t)b * (size_t)nt * N, lane_tol(L, b));
            }
        }
#if defined(LS_DEFERRED)
        // ---- first half of the pass: end of the previous step (rescale, output rows) -- requests
        bool active = b >= 0 && !done;
        unsigned lanes = __ballot_sync(FULL, active);
        if (active) {
            LS_REQI(LS_RI_SCALE_L) = 0;
            LS_REQI(LS_RI_NROWS) = 0;
            done = ls_advance_a(s, R, v, T, nt, te, L.usol + (size_t)b * (size_t)nt * N, lanes);
        }
        // ---- every thread with a request enters the queue of its kind
        int* const qc = qn + (pass & 1u) * LS_NKINDS;
        {
            const int kind = ls_request_kind(v.ops);
            const int lane = threadIdx.x & 31;
#pragma unroll
            for (int k = 0; k < LS_NKINDS; k++) {
                const unsigned m = __ballot_sync(FULL, kind == k);
                if (m) {
                    const int leader = __ffs(m) - 1;
                    int base = 0;
                    if (lane == leader) base = atomicAdd(&qc[k], __popc(m));
                    base = __shfl_sync(FULL, base, leader);
                    if (kind == k) q[k * B200_THREADS + base + __popc(m & ((1u << lane) - 1u))] = threadIdx.x;
                }
            }
            v.ops = 0;
        }
        // (the barrier also decides whether the block is finished: nobody holds or can fetch work)
        if (!__syncthreads_or((b >= -1) ? 1 : 0)) break;
        if (threadIdx.x < LS_NKINDS) qn[((pass + 1u) & 1u) * LS_NKINDS + threadIdx.x] = 0;
        ls_serve<STAGED>(L, q, qc);
        __syncthreads();
        // ---- second half: start of the step, prediction, corrector, error test, selection
        active = b >= 0 && !done;
        lanes = __ballot_sync(FULL, active);
        if (active)
            done = ls_advance_b(s, R, v, T, pp, lane_tol(L, b), MXSTEP_U, L.exit_on_warning, lanes);
#else
        // lanes with a running trajectory take one pass of the state machine together
        const unsigned lanes = __ballot_sync(FULL, b >= 0 && !done);
        if (lanes == 0 && !__any_sync(FULL, b >= 0)) break;
        if (b >= 0 && !done)
            done = ls_advance(s, R, v, T, pp, nt, te, L.usol + (size_t)b * (size_t)nt * N,
                              lane_tol(L, b), MXSTEP_U, L.exit_on_warning, lanes);
#endif
        if (b >= 0 && done) {
            // wrapper.cpp:41-45: istate <= 0 -> success = 0
            L.success[b] = (R.istate > 0) ? 1 : 0;
            if (L.counters) {
                int* c = L.counters + b * B200ODE_NCOUNTERS;
                c[0] = s.nst;
                c[1] = s.nfe;
                c[2] = s.nje;
                c[3] = R.nswitch;
                c[4] = s.nqu;
                c[5] = R.istate;
                c[6] = R.row;
                c[7] = s.meth;
            }
            b = -1;
#if defined(LS_DEFERRED)
            v.ops = 0;  // (a request of the attempt that ended the trajectory is void)
#endif
        }
    }
}

// (B200_MAXNREG: an explicit register budget instead of the one ptxas derives from the launch
// bounds, for geometry experiments: tools/sweep_lsoda.sh)
#ifdef B200_MAXNREG
#define LS_BOUNDS __maxnreg__(B200_MAXNREG)
#else
#define LS_BOUNDS __launch_bounds__(B200_THREADS, B200_MINBLOCKS)
#endif
extern "C" __global__ void LS_BOUNDS
b200ode_lsoda_kernel(const B200OdeLaunch L)
{
    lsoda_ensemble<true>(L);
}

extern "C" __global__ void LS_BOUNDS
b200ode_lsoda_kernel_ptr(const B200OdeLaunch L)
{
    lsoda_ensemble<false>(L);
}

// Build-time geometry of this image, read once by the host library after loading:
// {doubles of dynamic shared memory per block, 0, min blocks, N, threads}.
extern "C" __global__ void b200ode_lsoda_config(int* out)
{
    out[0] = LS_SMEM_DOUBLES;
    out[1] = 0;
    out[2] = B200_MINBLOCKS;
    out[3] = N;
    out[4] = B200_THREADS;
}
