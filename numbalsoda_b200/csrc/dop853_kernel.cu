// Batched DOP853 for sm_100a: one trajectory per thread, everything in registers.
//
// Replaces, for ensembles, the reference's dop853_wrapper + solout
// (src/dop853_c_interface.f90:40-130) and dop853_class: dop853 / dp86co / hinit /
// contd8 (src/dop853_module.f90:316-872).  The arithmetic of every formula is kept
// in the reference's evaluation order, so that with FMA contraction disabled at
// link time (strict mode) the results are bit-identical to the CPU path; what is
// new is the organisation:
//
//  * thread-per-trajectory: the state (y, k1..k10, y1 = 12n doubles) lives in
//    registers and the user's right-hand side is inlined by link-time
//    optimisation (b200ode_device.cuh);
//  * a persistent grid pulls trajectories from a global work counter: a lane that
//    finishes its trajectory starts the next one inside the same step loop, so
//    the adaptive-step imbalance between trajectories costs idle lanes only at
//    the very end of the launch;
//  * dense output is decoupled from stepping.  Only ~1/3 of the accepted steps
//    contain a t_eval point, but in a warp of 32 trajectories almost every
//    iteration has *some* lane in that case, so doing the 3 extra stages and the
//    interpolation in line ran that code on every iteration at ~9 active lanes
//    (ncu, profiles/r1_dop853_v0_inline_events.txt).  Instead a lane whose step passed output
//    times pushes the step's stage values (4+11n doubles) into a per-warp queue in
//    shared memory and goes on stepping; when 32 records are queued the whole
//    warp pops one record per lane and does the dense-output work fully converged;
//  * the t_eval rows of a trajectory are consecutive in memory ([B,nt,n] layout of
//    the reference's usol), so rows stream out sequentially per trajectory and the
//    L2 merges them into full-sector HBM writes.
#include "b200ode_device.cuh"
#include "b200ode_pow.cuh"
#include "b200ode_fastmath.cuh"
#include "dop853_tableau.cuh"

#define N B200_NEQ
#define FOR_N _Pragma("unroll") for (int i = 0; i < N; i++)

#define DP_UROUND 2.220446049250313e-16  // epsilon(1.0d0), dop853_constants.f90:11
#define DP_SAFE 0.9                      // dop853_module.f90:55
#define DP_FAC1 0.333                    // :56
#define DP_FAC2 6.0                      // :59

#ifndef B200_QCAP
#define B200_QCAP 64
#endif
#ifndef B200_ST_HINT
#define B200_ST_HINT 0  // cache hint of the output-row stores: 0 plain, 1 L2 evict_last, 2 .cs, 3 .cg
#endif
#ifndef B200_MINBLOCKS
#define B200_MINBLOCKS 1
#endif
#ifndef B200_THREADS
#define B200_THREADS 128
#endif
#define QCAP B200_QCAP                    // records per warp queue (>= 32)
#define QF B200ODE_DOP853_QFIELDS(N)      // doubles per record
#define FULL 0xffffffffu
// The queue is drained at the top of an iteration as soon as fewer than 32 slots are
// free (one iteration can push at most 32 records), 32 records at a time.  With
// QCAP = 63 every drain is a full warp; smaller queues trade drain efficiency
// (QCAP - 31 .. 32 lanes busy) for shared memory.
#define QDRAIN_AT (QCAP - 31 < 32 ? QCAP - 31 : 32)
#ifndef B200_DRAIN_NOINLINE
#define B200_DRAIN_NOINLINE 0
#endif
#if B200_DRAIN_NOINLINE
#define DRAIN_INLINE __noinline__
#else
#define DRAIN_INLINE __forceinline__
#endif

// gfortran's MAX/MIN ignore a NaN operand, exactly like fmax/fmin (DMNMX).
__device__ __forceinline__ double fmax_(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ double fmin_(double a, double b) { return fmin(a, b); }

// Record layout in the queue, field-major: q[field * QCAP + slot].
//   0: x (xold)   1: h (hout)   2: trajectory index   3: row range (lo | hi << 32)
//   4 + v*N + i : vector v, component i; v = 0 y, 1 ynew(k5), 2 k1, 3 f(xph,ynew),
//                 4..8 k6..k10, 9 k11 (stored in k2), 10 k12 (stored in k3)
#define QV(v, i) (4 + (v) * N + (i))

// Dense output for `m` queued steps, one per lane: dp86co module :657-705 (the 8n
// continuous-output coefficients and the three extra stages) followed by solout
// (c_interface :60-68) with contd8 (module :861-869).
template <bool STAGED>
__device__ DRAIN_INLINE void dop853_drain(const char* data, long long data_stride, int np_,
                                          const double* __restrict__ te_g,
                                          const double* te_s, double* usol, int nt,
                                          const double* q, int qhead, int m)
{
    const int lane = threadIdx.x & 31;
    if (lane >= m) return;
    int slot = qhead + lane;
    if (slot >= QCAP) slot -= QCAP;
    const double* r = q + slot;
    const double x = r[0 * QCAP], h = r[1 * QCAP];
    const long long b = __double_as_longlong(r[2 * QCAP]);
    const unsigned long long rows = (unsigned long long)__double_as_longlong(r[3 * QCAP]);
    int row = (int)(rows & 0xffffffffu);
    const int row_end = (int)(rows >> 32);
    // t_eval from the block's copy in shared memory if there is one (te_s), else from global
#define DRAIN_TE(i) (te_s ? te_s[(i)] : te_g[(i)])

    // Parameters of the record's trajectory (the queue holds steps of other lanes).
    double pl[STAGED ? B200ODE_NPMAX : 1];
    double* pp;
    const double* pg = reinterpret_cast<const double*>(data + b * data_stride);
    if (STAGED) {
#pragma unroll
        for (int j = 0; j < B200ODE_NPMAX; j++) pl[j] = (j < np_) ? pg[j] : 0.0;
        pp = pl;
    } else {
        pp = const_cast<double*>(pg);
    }
    // The stage vectors are consumed one at a time, in the order in which the
    // reference's sums run over them (k1, k6..k10, k11, k12, then f(xph,ynew)), so
    // every partial sum below is a prefix of the corresponding source expression:
    //   c4..c7 : cont(4n..8n) = d_1*k1 + d_6*k6 + ... + d_12*k3          module :668-679
    //   s14    : a141*k1 + a147*k7 + ... + a1412*k3 + a1413*k4          module :682-683
    //   s15    : a151*k1 + a156*k6 + ... + a1513*k4 (+ a1514*k14)       module :685-686
    //   s16    : a161*k1 + a166*k6 + ... + a1613*k4 (+ ...)             module :688-689
    // and only 13 vectors are live at once instead of 20.
    double c0[N], c1[N], c2[N], c3[N], c4[N], c5[N], c6[N], c7[N], s14[N], s15[N], s16[N];
    double kv[N], k4[N], y1[N];
    FOR_N {
        c0[i] = r[QV(0, i) * QCAP];                    // y
        const double ydiff = r[QV(1, i) * QCAP] - c0[i];  // ynew - y
        c1[i] = ydiff;
        kv[i] = r[QV(2, i) * QCAP];                    // k1
        k4[i] = r[QV(3, i) * QCAP];                    // f(xph, ynew)
        const double bspl = h * kv[i] - ydiff;
        c2[i] = bspl;
        c3[i] = ydiff - h * k4[i] - bspl;
        c4[i] = DP_D4_1 * kv[i];
        c5[i] = DP_D5_1 * kv[i];
        c6[i] = DP_D6_1 * kv[i];
        c7[i] = DP_D7_1 * kv[i];
        s14[i] = DP_A14_1 * kv[i];
        s15[i] = DP_A15_1 * kv[i];
        s16[i] = DP_A16_1 * kv[i];
    }
    FOR_N {
        kv[i] = r[QV(4, i) * QCAP];  // k6
        c4[i] = c4[i] + DP_D4_6 * kv[i];
        c5[i] = c5[i] + DP_D5_6 * kv[i];
        c6[i] = c6[i] + DP_D6_6 * kv[i];
        c7[i] = c7[i] + DP_D7_6 * kv[i];
        s15[i] = s15[i] + DP_A15_6 * kv[i];
        s16[i] = s16[i] + DP_A16_6 * kv[i];
    }
    FOR_N {
        kv[i] = r[QV(5, i) * QCAP];  // k7
        c4[i] = c4[i] + DP_D4_7 * kv[i];
        c5[i] = c5[i] + DP_D5_7 * kv[i];
        c6[i] = c6[i] + DP_D6_7 * kv[i];
        c7[i] = c7[i] + DP_D7_7 * kv[i];
        s14[i] = s14[i] + DP_A14_7 * kv[i];
        s15[i] = s15[i] + DP_A15_7 * kv[i];
        s16[i] = s16[i] + DP_A16_7 * kv[i];
    }
    FOR_N {
        kv[i] = r[QV(6, i) * QCAP];  // k8
        c4[i] = c4[i] + DP_D4_8 * kv[i];
        c5[i] = c5[i] + DP_D5_8 * kv[i];
        c6[i] = c6[i] + DP_D6_8 * kv[i];
        c7[i] = c7[i] + DP_D7_8 * kv[i];
        s14[i] = s14[i] + DP_A14_8 * kv[i];
        s15[i] = s15[i] + DP_A15_8 * kv[i];
        s16[i] = s16[i] + DP_A16_8 * kv[i];
    }
    FOR_N {
        kv[i] = r[QV(7, i) * QCAP];  // k9
        c4[i] = c4[i] + DP_D4_9 * kv[i];
        c5[i] = c5[i] + DP_D5_9 * kv[i];
        c6[i] = c6[i] + DP_D6_9 * kv[i];
        c7[i] = c7[i] + DP_D7_9 * kv[i];
        s14[i] = s14[i] + DP_A14_9 * kv[i];
        s16[i] = s16[i] + DP_A16_9 * kv[i];
    }
    FOR_N {
        kv[i] = r[QV(8, i) * QCAP];  // k10
        c4[i] = c4[i] + DP_D4_10 * kv[i];
        c5[i] = c5[i] + DP_D5_10 * kv[i];
        c6[i] = c6[i] + DP_D6_10 * kv[i];
        c7[i] = c7[i] + DP_D7_10 * kv[i];
        s14[i] = s14[i] + DP_A14_10 * kv[i];
    }
    FOR_N {
        kv[i] = r[QV(9, i) * QCAP];  // k11
        c4[i] = c4[i] + DP_D4_11 * kv[i];
        c5[i] = c5[i] + DP_D5_11 * kv[i];
        c6[i] = c6[i] + DP_D6_11 * kv[i];
        c7[i] = c7[i] + DP_D7_11 * kv[i];
        s14[i] = s14[i] + DP_A14_11 * kv[i];
        s15[i] = s15[i] + DP_A15_11 * kv[i];
    }
    FOR_N {
        kv[i] = r[QV(10, i) * QCAP];  // k12
        c4[i] = c4[i] + DP_D4_12 * kv[i];
        c5[i] = c5[i] + DP_D5_12 * kv[i];
        c6[i] = c6[i] + DP_D6_12 * kv[i];
        c7[i] = c7[i] + DP_D7_12 * kv[i];
        s14[i] = s14[i] + DP_A14_12 * kv[i];
        s15[i] = s15[i] + DP_A15_12 * kv[i];
    }
    // the next three function evaluations, module :682-691, and the final
    // preparation, module :693-703: cont = h*(cont + d_13*k4 + d_14*k14 + d_15*k15 + d_16*k16)
    FOR_N {
        y1[i] = c0[i] + h * (s14[i] + DP_A14_13 * k4[i]);
        s15[i] = s15[i] + DP_A15_13 * k4[i];
        s16[i] = s16[i] + DP_A16_13 * k4[i];
        c4[i] = c4[i] + DP_D4_13 * k4[i];
        c5[i] = c5[i] + DP_D5_13 * k4[i];
        c6[i] = c6[i] + DP_D6_13 * k4[i];
        c7[i] = c7[i] + DP_D7_13 * k4[i];
    }
    b200ode_user_rhs(x + DP_C14 * h, y1, kv, pp);  // k14
    FOR_N {
        y1[i] = c0[i] + h * (s15[i] + DP_A15_14 * kv[i]);
        s16[i] = s16[i] + DP_A16_14 * kv[i];
        c4[i] = c4[i] + DP_D4_14 * kv[i];
        c5[i] = c5[i] + DP_D5_14 * kv[i];
        c6[i] = c6[i] + DP_D6_14 * kv[i];
        c7[i] = c7[i] + DP_D7_14 * kv[i];
    }
    b200ode_user_rhs(x + DP_C15 * h, y1, kv, pp);  // k15
    FOR_N {
        y1[i] = c0[i] + h * (s16[i] + DP_A16_15 * kv[i]);
        c4[i] = c4[i] + DP_D4_15 * kv[i];
        c5[i] = c5[i] + DP_D5_15 * kv[i];
        c6[i] = c6[i] + DP_D6_15 * kv[i];
        c7[i] = c7[i] + DP_D7_15 * kv[i];
    }
    b200ode_user_rhs(x + DP_C16 * h, y1, kv, pp);  // k16
    FOR_N {
        c4[i] = h * (c4[i] + DP_D4_16 * kv[i]);
        c5[i] = h * (c5[i] + DP_D5_16 * kv[i]);
        c6[i] = h * (c6[i] + DP_D6_16 * kv[i]);
        c7[i] = h * (c7[i] + DP_D7_16 * kv[i]);
    }
    // solout rows [row, row_end): xold = x, hout = h
    double* out = usol + ((size_t)b * (size_t)nt + (size_t)row) * N;
#if B200_ST_HINT == 1
    // experiment: keep the partially written lines of the output stream in
    // L2 until their trajectory has completed them
    unsigned long long st_policy;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(st_policy));
#endif
    for (; row < row_end; row++, out += N) {
        const double tq = DRAIN_TE(row);
        const double s = B200ODE_FAST_CTRL() ? b200_div(tq - x, h) : (tq - x) / h;
        const double s1 = 1.0 - s;
        FOR_N {
            double conpar = c4[i] + s * (c5[i] + s1 * (c6[i] + s * c7[i]));
            const double val = c0[i] + s * (c1[i] + s1 * (c2[i] + s * (c3[i] + s1 * conpar)));
#if B200_ST_HINT == 1
            asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(out + i), "d"(val), "l"(st_policy) : "memory");
#elif B200_ST_HINT == 2
            __stcs(out + i, val);
#elif B200_ST_HINT == 3
            __stcg(out + i, val);
#else
            out[i] = val;
#endif
        }
    }
#undef DRAIN_TE
}

// STAGED: the first L.np doubles of the trajectory's data record are copied into
// registers once and the right-hand side reads them there; otherwise it receives
// the record's global address unchanged (arbitrary record layout).
template <bool STAGED, bool VTOL>
__device__ __forceinline__ void dop853_ensemble(const B200OdeLaunch& L)
{
    extern __shared__ double b200_smem[];
    // (lane and the warp's queue base are recomputed where needed: two registers less)
#define LANE ((int)(threadIdx.x & 31u))
#define WARP_Q (b200_smem + (size_t)(threadIdx.x >> 5) * (QF * QCAP))
    int qhead = 0, qcount = 0;  // warp-uniform
    // t_eval is shared by the whole ensemble, so the integration span is the same for
    // every trajectory: dop853_wrapper's t = teval(1), tout = teval(nt) (c_interface
    // :122-123), dop853's hmax = xend - x (module :431-432, :519) and dp86co's posneg
    // (module :507).  Kept in shared memory (4 doubles per block) to save registers.
    double* span = b200_smem + (size_t)(B200_THREADS / 32) * (QF * QCAP);
    if (threadIdx.x == 0) {
        const double x0 = L.t_eval[0], xe = L.t_eval[L.nt - 1];
        span[0] = x0;
        span[1] = xe;
        span[2] = fabs(xe - x0);
        span[3] = copysign(1.0, xe - x0);
    }
    // t_eval is read once per output row by the event test below and once more by the dense
    // output; from global memory those loads were the two largest stall sites of the kernel
    // (ncu source page, profiles/r1_dop853_v3: 7.4 % + 2.0 % of the samples on the compare /
    // subtract that consume them).  When nt doubles more fit the block's shared memory
    // without costing a resident block (the host decides: b200ode.cpp launch()), every block
    // reads its own copy.
    // (no pointer is kept: both candidates are rematerialised from the constant bank)
#define TE_S (span + 4 + 3 * B200_THREADS)
#define TE(i) (L.te_shared ? TE_S[(i)] : L.t_eval[(i)])
    if (L.te_shared)
        for (int i = threadIdx.x; i < L.nt; i += B200_THREADS) TE_S[i] = L.t_eval[i];
    __syncthreads();
#define SPAN_X0 span[0]
#define SPAN_XEND span[1]
#define SPAN_HMAX span[2]
#define SPAN_POSNEG span[3]
    // solout's xout = t_eval[min(row, nt-1)] of this lane's trajectory, cached in shared
    // memory: t_eval is read once per output row instead of once per step (ncu,
    // profiles/r1_dop853_v2: the two t_eval loads were 12 % of the stall samples), at no
    // cost in registers (the kernel sits at the 168-register bound of 3 blocks per SM)
#define XOUT span[4 + threadIdx.x]
    // the tolerances of this lane's trajectory (scalars, or per trajectory: rtol_b / atol_b),
    // also in shared memory for the same reason
#define RTOL span[4 + B200_THREADS + threadIdx.x]
#define ATOL span[4 + 2 * B200_THREADS + threadIdx.x]
    // Per-component tolerances (the reference's itol = 1 branches,
    // dop853_module.f90:603-612, :779-803): read from global memory where they are used -- twice
    // per trajectory in hinit, once per step in the error norm -- so that the scalar case keeps
    // its registers and shared memory.  rtol_vec / atol_vec are uniform over the launch.
    // VTOL is a template parameter: the scalar kernels carry none of it.
#define RT_I(i, scalar) ((VTOL && L.rtol_vec) ? __ldg(L.rtol_b + (long long)b * L.rtol_sb + (i)) : (scalar))
#define AT_I(i, scalar) ((VTOL && L.atol_vec) ? __ldg(L.atol_b + (long long)b * L.atol_sb + (i)) : (scalar))
    // (Measured with one kernel for both cases: the loads in line, predicated off in the scalar
    // case, cost 16 issue slots per step; out of line behind a launch-uniform branch they cost
    // 5 % -- 8.02 M against 8.47 M Lorenz trajectories/s: the call site splits the scheduling
    // region of the error norm.  Hence two instantiations.)
#define SK_I(i, ymax) (AT_I(i, atol) + RT_I(i, rtol) * (ymax))

    double y[N], y1[N], k1[N], k2[N], k3[N], k4[N], k5[N], k6[N], k7[N], k8[N], k9[N], k10[N];
    double pl[STAGED ? B200ODE_NPMAX : 1];
    double* pp = pl;
    double x = 0.0, h = 0.0;
    int nstep = 0, naccpt = 0, nrejct = 0, nevent = 0, row = 0;
    // lane state bits (one register instead of three predicates kept across the loop)
    enum : unsigned { F_LAST = 1u, F_REJECT = 2u, F_EXHAUSTED = 4u };
    unsigned flags = 0;
    int b = -1;  // trajectory index of this lane, -1 = idle

    const double facc1 = 1.0 / DP_FAC1;  // module :505
    const double facc2 = 1.0 / DP_FAC2;  // module :506
    const int nt = L.nt;

    for (;;) {
        if (b < 0 && !(flags & F_EXHAUSTED)) {
            // ---- take the next trajectory: dop853_wrapper (c_interface :98-124),
            //      dop853 (module :396-438), dp86co prologue (module :502-534)
            const long long bn = b200ode_fetch(L.next);
            if (bn >= L.B) {
                flags |= F_EXHAUSTED;
            } else {
                b = (int)bn;
                const double* u0 = L.u0 + (long long)b * L.u0_stride;
                FOR_N y[i] = u0[i];
                const double* pg = reinterpret_cast<const double*>(L.data + (long long)b * L.data_stride);
                if (STAGED) {
#pragma unroll
                    for (int j = 0; j < B200ODE_NPMAX; j++) pl[j] = (j < L.np) ? pg[j] : 0.0;
                    pp = pl;
                } else {
                    pp = const_cast<double*>(pg);
                }
                const double rtol = (L.rtol_b && !L.rtol_vec) ? L.rtol_b[(long long)b * L.rtol_sb] : L.rtol;
                const double atol = (L.atol_b && !L.atol_vec) ? L.atol_b[(long long)b * L.atol_sb] : L.atol;
                RTOL = rtol;
                ATOL = atol;
                x = SPAN_X0;
                const double hmax = SPAN_HMAX, posneg = SPAN_POSNEG;
                flags = 0;
                nstep = naccpt = nrejct = nevent = 0;
                b200ode_user_rhs(x, y, k1, pp);
                {
                    // hinit, module :745-823 (itol = 0, iord = 8)
                    double dnf = 0.0, dny = 0.0;
                    FOR_N {
                        double sk = AT_I(i, atol) + RT_I(i, rtol) * fabs(y[i]);
                        double qq = k1[i] / sk;
                        dnf = dnf + qq * qq;
                        qq = y[i] / sk;
                        dny = dny + qq * qq;
                    }
                    double hh;
                    if (dnf <= 1.0e-10 || dny <= 1.0e-10) hh = 1.0e-6;
                    else hh = sqrt(dny / dnf) * 0.01;
                    hh = fmin_(hh, hmax);
                    hh = copysign(fabs(hh), posneg);
                    FOR_N y1[i] = y[i] + hh * k1[i];
                    b200ode_user_rhs(x + hh, y1, k2, pp);
                    double der2 = 0.0;
                    FOR_N {
                        double sk = AT_I(i, atol) + RT_I(i, rtol) * fabs(y[i]);
                        double qq = (k2[i] - k1[i]) / sk;
                        der2 = der2 + qq * qq;
                    }
                    der2 = sqrt(der2) / hh;
                    double der12 = fmax_(fabs(der2), sqrt(dnf));
                    double h1;
                    if (der12 <= 1.0e-15) h1 = fmax_(1.0e-6, fabs(hh) * 1.0e-3);
                    else h1 = b200ode_pow_eighth(0.01 / der12);
                    hh = fmin_(fmin_(100.0 * fabs(hh), h1), hmax);
                    h = copysign(fabs(hh), posneg);
                }
                // first solout call (c_interface :56-58): xout = teval(1), nt = 1.  From here
                // on xout is always t_eval[min(row, nt-1)] (it keeps its last value once
                // every row is written, c_interface :66-68).
                row = 0;
                XOUT = x;
            }
        }

        // ---- dense output for queued steps, converged across the warp
        if (qcount >= QDRAIN_AT || (qcount > 0 && !__any_sync(FULL, b >= 0))) {
            const int m = qcount < 32 ? qcount : 32;
            dop853_drain<STAGED>(L.data, L.data_stride, L.np, L.t_eval,
                                 L.te_shared ? TE_S : nullptr, L.usol, nt, WARP_Q, qhead, m);
            qhead += m;
            if (qhead >= QCAP) qhead -= QCAP;
            qcount -= m;
            __syncwarp();
        }
        if (!__any_sync(FULL, b >= 0)) {  // (voted again: cheaper than a register held across the drain)
            if (qcount == 0) break;
            continue;
        }

        // ---- one attempted step, dp86co module :537-731
        int idid = 0;
        bool accepted = false, push = false;
        int row_lo = 0;
        double hnew = 0.0, xph = 0.0;
        if (b >= 0) {
            if (nstep > L.mxstep) {
                idid = -2;  // module :539
            } else if (0.1 * fabs(h) <= fabs(x) * DP_UROUND) {
                idid = -3;  // module :546
            } else {
                const double xend = SPAN_XEND, posneg = SPAN_POSNEG;
                if ((x + 1.01 * h - xend) * posneg > 0.0) {  // module :554-557
                    h = xend - x;
                    flags |= F_LAST;
                }
                nstep++;
                // the twelve stages, module :561-587
                FOR_N y1[i] = y[i] + h * DP_A2_1 * k1[i];
                b200ode_user_rhs(x + DP_C2 * h, y1, k2, pp);
                FOR_N y1[i] = y[i] + h * (DP_A3_1 * k1[i] + DP_A3_2 * k2[i]);
                b200ode_user_rhs(x + DP_C3 * h, y1, k3, pp);
                FOR_N y1[i] = y[i] + h * (DP_A4_1 * k1[i] + DP_A4_3 * k3[i]);
                b200ode_user_rhs(x + DP_C4 * h, y1, k4, pp);
                FOR_N y1[i] = y[i] + h * (DP_A5_1 * k1[i] + DP_A5_3 * k3[i] + DP_A5_4 * k4[i]);
                b200ode_user_rhs(x + DP_C5 * h, y1, k5, pp);
                FOR_N y1[i] = y[i] + h * (DP_A6_1 * k1[i] + DP_A6_4 * k4[i] + DP_A6_5 * k5[i]);
                b200ode_user_rhs(x + DP_C6 * h, y1, k6, pp);
                FOR_N y1[i] = y[i] + h * (DP_A7_1 * k1[i] + DP_A7_4 * k4[i] + DP_A7_5 * k5[i] +
                                          DP_A7_6 * k6[i]);
                b200ode_user_rhs(x + DP_C7 * h, y1, k7, pp);
                FOR_N y1[i] = y[i] + h * (DP_A8_1 * k1[i] + DP_A8_4 * k4[i] + DP_A8_5 * k5[i] +
                                          DP_A8_6 * k6[i] + DP_A8_7 * k7[i]);
                b200ode_user_rhs(x + DP_C8 * h, y1, k8, pp);
                FOR_N y1[i] = y[i] + h * (DP_A9_1 * k1[i] + DP_A9_4 * k4[i] + DP_A9_5 * k5[i] +
                                          DP_A9_6 * k6[i] + DP_A9_7 * k7[i] + DP_A9_8 * k8[i]);
                b200ode_user_rhs(x + DP_C9 * h, y1, k9, pp);
                FOR_N y1[i] = y[i] + h * (DP_A10_1 * k1[i] + DP_A10_4 * k4[i] + DP_A10_5 * k5[i] +
                                          DP_A10_6 * k6[i] + DP_A10_7 * k7[i] + DP_A10_8 * k8[i] +
                                          DP_A10_9 * k9[i]);
                b200ode_user_rhs(x + DP_C10 * h, y1, k10, pp);
                FOR_N y1[i] = y[i] + h * (DP_A11_1 * k1[i] + DP_A11_4 * k4[i] + DP_A11_5 * k5[i] +
                                          DP_A11_6 * k6[i] + DP_A11_7 * k7[i] + DP_A11_8 * k8[i] +
                                          DP_A11_9 * k9[i] + DP_A11_10 * k10[i]);
                b200ode_user_rhs(x + DP_C11 * h, y1, k2, pp);
                xph = x + h;
                FOR_N y1[i] = y[i] + h * (DP_A12_1 * k1[i] + DP_A12_4 * k4[i] + DP_A12_5 * k5[i] +
                                          DP_A12_6 * k6[i] + DP_A12_7 * k7[i] + DP_A12_8 * k8[i] +
                                          DP_A12_9 * k9[i] + DP_A12_10 * k10[i] +
                                          DP_A12_11 * k2[i]);
                b200ode_user_rhs(xph, y1, k3, pp);
                FOR_N {
                    k4[i] = DP_B1 * k1[i] + DP_B6 * k6[i] + DP_B7 * k7[i] + DP_B8 * k8[i] +
                            DP_B9 * k9[i] + DP_B10 * k10[i] + DP_B11 * k2[i] + DP_B12 * k3[i];
                    k5[i] = y[i] + h * k4[i];
                }
                // error estimation, module :591-616
                const double rtol = RTOL, atol = ATOL;
                double err = 0.0, err2 = 0.0;
                double fac11 = 0.0, ifac = 0.0;
                if (B200ODE_FAST_CTRL()) {
                    // throughput mode (b200ode_fastmath.cuh): the same quantities through
                    // reciprocals; err holds the *square* of the reference's err (the test
                    // err <= 1 is unchanged by that), ifac = safe / err**(1/8)
                    FOR_N {
                        const double isk = b200_rcp(SK_I(i, fmax_(fabs(y[i]), fabs(k5[i]))));
                        double erri = k4[i] - DP_BHH1 * k1[i] - DP_BHH2 * k9[i] - DP_BHH3 * k3[i];
                        double qq = erri * isk;
                        err2 = err2 + qq * qq;
                        erri = DP_ER1 * k1[i] + DP_ER6 * k6[i] + DP_ER7 * k7[i] + DP_ER8 * k8[i] +
                               DP_ER9 * k9[i] + DP_ER10 * k10[i] + DP_ER11 * k2[i] + DP_ER12 * k3[i];
                        qq = erri * isk;
                        err = err + qq * qq;
                    }
                    double deno = err + 0.01 * err2;
                    if (deno <= 0.0) deno = 1.0;
                    err = (h * h) * (err * err) * b200_rcp((double)N * deno);
                    ifac = DP_SAFE * b200_inv_root16(err);
                    // fac = max(facc2, min(facc1, fac11/safe)); hnew = h/fac   (module :620-623)
                    hnew = h * fmin_(DP_FAC2, fmax_(DP_FAC1, ifac));
                } else {
                    FOR_N {
                        double sk = SK_I(i, fmax_(fabs(y[i]), fabs(k5[i])));
                        double erri = k4[i] - DP_BHH1 * k1[i] - DP_BHH2 * k9[i] - DP_BHH3 * k3[i];
                        double qq = erri / sk;
                        err2 = err2 + qq * qq;
                        erri = DP_ER1 * k1[i] + DP_ER6 * k6[i] + DP_ER7 * k7[i] + DP_ER8 * k8[i] +
                               DP_ER9 * k9[i] + DP_ER10 * k10[i] + DP_ER11 * k2[i] + DP_ER12 * k3[i];
                        qq = erri / sk;
                        err = err + qq * qq;
                    }
                    double deno = err + 0.01 * err2;
                    if (deno <= 0.0) deno = 1.0;
                    err = fabs(h) * err * sqrt(1.0 / ((double)N * deno));
                    // step size controller, module :618-623 (facold**beta == 1 for beta = 0,
                    // which also makes facold, module :626, dead)
                    fac11 = b200ode_pow_eighth(err);
                    double fac = fac11 / 1.0;
                    fac = fmax_(facc2, fmin_(facc1, fac / DP_SAFE));
                    hnew = h / fac;
                }
                if (err <= 1.0) {
                    // ---- accepted, module :625-722
                    accepted = true;
                    naccpt++;
                    b200ode_user_rhs(xph, k5, k4, pp);
                    // (the stiffness detector, module :631-655, only prints: iprint /= 0)
                    if (XOUT <= xph) {  // event: xout <= xph, module :657
                        // dense output is prepared for this step: module :658-705 costs
                        // 3 RHS evaluations; solout writes every row with t_eval <= xph
                        nevent++;
                        row_lo = row;
                        double xo;
                        do {
                            row++;
                            xo = TE(row < nt ? row : nt - 1);
                        } while (row < nt && xo <= xph);
                        XOUT = xo;
                        push = row > row_lo;
                    }
                    if (flags & F_LAST) {
                        idid = 1;  // module :715-719
                    } else {
                        const double hmax = SPAN_HMAX;
                        if (fabs(hnew) > hmax) hnew = posneg * hmax;
                        if (flags & F_REJECT) hnew = posneg * fmin_(fabs(hnew), fabs(h));
                        flags &= ~F_REJECT;
                    }
                } else {
                    // ---- rejected, module :724-728
                    if (B200ODE_FAST_CTRL()) hnew = h * fmax_(DP_FAC1, ifac);
                    else hnew = h / fmin_(facc1, fac11 / DP_SAFE);
                    if (naccpt >= 1) nrejct++;
                    flags = F_REJECT;  // and last = .false.
                }
            }
        }

        // ---- queue the steps that passed output times
        const unsigned pm = __ballot_sync(FULL, push);
        if (pm) {
            const int np = __popc(pm);
            if (push) {
                int slot = qhead + qcount + __popc(pm & ((1u << LANE) - 1u));
                if (slot >= QCAP) slot -= QCAP;
                double* r = WARP_Q + slot;
                r[0 * QCAP] = x;
                r[1 * QCAP] = h;
                r[2 * QCAP] = __longlong_as_double((long long)b);
                r[3 * QCAP] = __longlong_as_double(
                    (long long)((unsigned long long)(unsigned)row_lo |
                                ((unsigned long long)(unsigned)row << 32)));
                FOR_N {
                    r[QV(0, i) * QCAP] = y[i];
                    r[QV(1, i) * QCAP] = k5[i];
                    r[QV(2, i) * QCAP] = k1[i];
                    r[QV(3, i) * QCAP] = k4[i];
                    r[QV(4, i) * QCAP] = k6[i];
                    r[QV(5, i) * QCAP] = k7[i];
                    r[QV(6, i) * QCAP] = k8[i];
                    r[QV(7, i) * QCAP] = k9[i];
                    r[QV(8, i) * QCAP] = k10[i];
                    r[QV(9, i) * QCAP] = k2[i];
                    r[QV(10, i) * QCAP] = k3[i];
                }
            }
            qcount += np;
            __syncwarp();
        }

        if (b >= 0) {
            if (accepted) {
                FOR_N {
                    k1[i] = k4[i];
                    y[i] = k5[i];
                }
                x = xph;
            }
            if (idid == 0) {
                h = hnew;
            } else {
                // dop853_wrapper epilogue, c_interface :125-128
                L.success[b] = (idid < 0) ? 0 : 1;
                if (L.counters) {
                    int* c = L.counters + (long long)b * B200ODE_NCOUNTERS;
                    c[0] = nstep;
                    // nfcn (module :524,:587,:629,:691): 2 at start-up, 11 per attempted step,
                    // 1 per accepted step, 3 per step that prepared dense output
                    c[1] = 2 + 11 * nstep + naccpt + 3 * nevent;
                    c[2] = naccpt;
                    c[3] = nrejct;
                    c[4] = nevent;
                    c[5] = idid;
                    c[6] = row;
                    c[7] = 0;
                }
                b = -1;
            }
        }
    }
}

// (B200_MAXNREG: an explicit register budget instead of the one ptxas derives from the launch
// bounds -- it rounds 13 or 14 warps per SM down to 128 registers -- for geometry experiments)
#ifdef B200_MAXNREG
#define DP_BOUNDS __maxnreg__(B200_MAXNREG)
#else
#define DP_BOUNDS __launch_bounds__(B200_THREADS, B200_MINBLOCKS)
#endif
extern "C" __global__ void DP_BOUNDS
b200ode_dop853_kernel(const B200OdeLaunch L)
{
    dop853_ensemble<true, false>(L);
}

extern "C" __global__ void DP_BOUNDS
b200ode_dop853_kernel_ptr(const B200OdeLaunch L)
{
    dop853_ensemble<false, false>(L);
}

// the same two kernels for per-component tolerances (launches with rtol_vec / atol_vec)
extern "C" __global__ void DP_BOUNDS
b200ode_dop853_kernel_vtol(const B200OdeLaunch L)
{
    dop853_ensemble<true, true>(L);
}

extern "C" __global__ void DP_BOUNDS
b200ode_dop853_kernel_ptr_vtol(const B200OdeLaunch L)
{
    dop853_ensemble<false, true>(L);
}

// Build-time geometry of this image, read once by the host library after loading.
extern "C" __global__ void b200ode_dop853_config(int* out)
{
    out[0] = QCAP;
    out[1] = QF;
    out[2] = B200_MINBLOCKS;
    out[3] = N;
    out[4] = B200_THREADS;
    out[5] = (B200_THREADS / 32) * QF * QCAP + 4 + 3 * B200_THREADS;  // doubles of dynamic shared memory
                                                                      // (+ nt with te_shared)
}
