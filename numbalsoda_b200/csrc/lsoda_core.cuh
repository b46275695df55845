// LSODA: the solver's control flow for ONE trajectory, independent of how the
// trajectory's vectors are stored and who works on them.
//
// Replaces, per trajectory, the reference's lsoda_wrapper (src/wrapper.cpp:10-57) ->
// LSODA::lsoda_update (src/LSODA.cpp:2247-2277) -> LSODA::lsoda (:287-993) ->
// stoda (:995-1366) -> correction (:1743-1902) / methodswitch (:1946-2069) /
// orderswitch (:2097-2209) / scaleh (:1594-1631).  Every floating-point expression
// keeps the reference's operand order, so that with FMA contraction disabled at link
// time the states and the step / RHS / Jacobian / method-switch counters are those of
// the CPU path bit for bit.  What is different:
//
//  * vectors.  Everything of length n (Nordsieck history yh, y, savf, acor, ewt, the
//    iteration matrix and its LU factors) belongs to a *vector backend* V:
//      lsoda_vec_thread.cuh  one thread per trajectory, n <= 8: y / ewt / acor in
//                            registers, yh and the matrix in shared memory as
//                            [slot][thread] columns (run-time order and pivot
//                            indices without local memory);
//      lsoda_vec_warp.cuh    one warp per trajectory, n <= 128: all vectors in shared
//                            memory, lanes own components, norms and pivot searches
//                            by shuffles, Jacobian columns evaluated in parallel.
//    This file holds the scalars (LsLane) and the decisions.
//  * coefficients.  cfode's elco / tesco tables depend only on the method: they are
//    compile-time tables (lsoda_tables.cuh, bit-equal to cfode's arithmetic) instead
//    of being recomputed at every method switch; el[] is read from them in place.
//  * driver.  Instead of re-entering the solver once per output time (wrapper.cpp:32)
//    a trajectory steps freely and, after each accepted step, interpolates every t_eval
//    point the step has passed -- the same arithmetic.
//  * control.  The reference's loop nest (steps / retries / corrector iterations) is
//    flattened into a state machine whose pass is one corrector iteration
//    (ls_advance), so that lanes of a warp at different depths of the nest still
//    execute the same code together.
//
// The headers also compile as host C++ (tests/lsoda_core_host.cpp) so that the
// restated control flow can be checked against the CPU oracle without a GPU; the
// product only ever runs them inside lsoda_kernel.cu / lsoda_warp_kernel.cu.
#ifndef B200ODE_LSODA_CORE_CUH
#define B200ODE_LSODA_CORE_CUH

#include "b200ode_pow.cuh"
#include "b200ode_fastmath.cuh"
#include "lsoda_tables.cuh"

#if defined(__CUDACC__)
#define LS_FN __device__ __forceinline__
#define LS_MFN __device__ __forceinline__  // member functions
#else
#include <cmath>
#define LS_FN static inline
#define LS_MFN inline
#endif

#define LS_ETA 2.2204460492503131e-16      // LSODA.cpp:33
#define LS_SQRTETA 1.4901161193847656e-08  // sqrt(ETA) = 2^-26, LSODA.cpp:483
#define LS_MXORDN 12                       // LSODA.cpp:38
#define LS_MXORDS 5
#define LS_MAXCOR 3   // LSODA.cpp:564
#define LS_MSBP 20    // LSODA.cpp:565
#define LS_MXHNIL 10  // LSODA.cpp:294
#define LS_CCMAX 0.3  // LSODA.cpp:563
#ifndef LS_ITERS_PER_PASS
// corrector iterations one pass of ls_advance may take.  3 = maxcor: a pass is then a whole
// attempt unless the corrector restarts with a fresh Jacobian.  Measured on Robertson
// (M traj/s): 1 -> 4.79, 2 -> 4.58, 3 -> 5.35, 4 -> 4.30, 6 -> 4.47.
#define LS_ITERS_PER_PASS 3
#endif

// tables of the method in use (cfode is re-run for `meth` at :1086 before any use)
#define ELCO(q, k) T->elco[s.meth - 1][q][k]
#define TESCO(q, k) T->tesco[s.meth - 1][q][k]

// where a trajectory is in its step (ls_advance)
#define LS_P_STEP 0     // about to start a step
#define LS_P_ATTEMPT 1  // about to predict (first attempt of a step, or a retry)
#define LS_P_ITER 2     // inside the corrector iteration
#define LS_P_JACWAIT 3  // deferred backends: the iteration matrix was requested; the corrector
                        // resumes on the next pass with the right-hand side it already has
// work a pass leaves for the trajectory's next pass
#define LS_F_RESCALE 1u  // apply scaleh with the ratio s.rh
#define LS_F_RMAX10 2u   // ... then rmax = 10 (the rescale ends an accepted step)
#define LS_F_FINISH 4u   // an accepted step is complete: endstoda, output rows
#define LS_F_REFRESH 8u  // deferred backends: yh[2] = h f(tn, yh[1]) is due once the retraction is done

// The scalars of class LSODA (LSODA.h:105-143) plus the locals of stoda / correction
// that outlive a pass of the state machine.
struct LsLane {
    double h, hu, hold, tn, rc, el0, crate, rmax, pdest, pdlast, pdnorm;
    double told, pnorm, del, delp, rate, rh;
    int nq, nslp;
    int nst, nfe, nje;
    int nqu : 16, icount : 16;  // (order of the last step; 20 down to -1, where it stays)
    // The small state variables share one register (the kernel carries ~75 values per thread
    // across a pass; at 128 registers -- 16 warps per SM -- every word counts).  Ranges: meth 1-2,
    // mused 0-2, jcur 0-1, irflag 0-2, ipup 0 / 2 / 5 (miter), m 0-3 (LS_MAXCOR), phase LS_P_*,
    // flags LS_F_*, ialth 0-13, jstart -1..1, kflag -10..0.
    unsigned meth : 2, mused : 2, jcur : 1, irflag : 2, ipup : 3, m : 3, phase : 3, flags : 4;
    int jstart : 2, kflag : 5, ialth : 5;
};

struct LsRun {
    int row;      // next usol row to write
    int nslast;   // nst at the last output (mxstep is per output interval, LSODA.cpp:671,:778)
    int nswitch;  // Adams <-> BDF switches
    unsigned nhnil : 8;  // t + h == t warnings, counted up to LS_MXHNIL + 1 (only "<= LS_MXHNIL" is ever asked)
    int istate : 8;      // 1 running; 2 finished; <= 0 failed (as LSODA.cpp:782-976; 0 = t_eval not monotone)
};

// Tolerances of one trajectory: the two scalars lsoda_update spreads over its private vectors
// (LSODA.cpp:2267-2270), or -- SURVEY.md section 8f rank 1 -- one value per component, the
// reference's itol_ = 3 / 4 cases (ewset, LSODA.cpp:1368-1390; checks :500-520; h0's tol
// :624-640).  A null vector means the scalar applies to every component.
struct LsTol {
    double rtol, atol;
    const double* rv;  // [n] or null
    const double* av;  // [n] or null
    LS_MFN double rt(int i) const { return rv ? rv[i] : rtol; }
    LS_MFN double at(int i) const { return av ? av[i] : atol; }
};

// std::max / std::min as the reference uses them (they differ from fmax/fmin on NaN)
LS_FN double ls_max(double a, double b) { return (a < b) ? b : a; }
LS_FN double ls_min(double a, double b) { return (b < a) ? b : a; }

// Throughput mode (default link flags, include/b200ode.h): divisions through a reciprocal and
// one residual correction (within an ulp, 6 instructions without a slow path instead of ~23),
// the step / order / method ratios through a single-precision pow: b200ode_fastmath.cuh.
// A link-time constant on the device; on the host (tests/lsoda_core_host.cpp) a build flag.
#if defined(__CUDA_ARCH__)
#define LS_FAST() B200ODE_FAST_CTRL()
#elif defined(B200_HOST_FAST)
#define LS_FAST() true
#else
#define LS_FAST() false
#endif

// IEEE division.  ls_div is inlined (the compiler's Newton sequence, ~23 instructions).
// ls_div_rare marks the divisions of blocks that few lanes of a warp execute (Jacobian /
// LU, rescale, coefficient reload).  With -DB200_DIV_NOINLINE it becomes one out-of-line
// routine -- an experiment against the kernel's instruction-cache misses (ncu
// stall_no_instruction, profiles/r1_lsoda_v3) that measured SLOWER on Robertson (4.40 M
// vs 4.66 M traj/s; every division out of line: 4.28 M), so it is off by default.
#if defined(__CUDA_ARCH__) && defined(B200_DIV_NOINLINE_ALL)
__device__ __noinline__ double ls_div_ieee(double a, double b) { return a / b; }
#else
LS_FN double ls_div_ieee(double a, double b) { return a / b; }
#endif
#if defined(__CUDA_ARCH__) && (defined(B200_DIV_NOINLINE) || defined(B200_DIV_NOINLINE_ALL))
__device__ __noinline__ double ls_div_rare_ieee(double a, double b) { return a / b; }
#else
LS_FN double ls_div_rare_ieee(double a, double b) { return a / b; }
#endif
LS_FN double ls_div(double a, double b) { return LS_FAST() ? b200_div(a, b) : ls_div_ieee(a, b); }
LS_FN double ls_div_rare(double a, double b)
{
    return LS_FAST() ? b200_div(a, b) : ls_div_rare_ieee(a, b);
}

// The step-size ratio formula of stoda / methodswitch / orderswitch,
// 1 / (c * x^e + c * 1e-6)  (LSODA.cpp:1229, :1983, :1994, :2037, :2051, :2106, :2111),
// six call sites: out of line together with pow in the thread-per-trajectory kernels.
#if defined(__CUDA_ARCH__) && defined(B200POW_NOINLINE)
__device__ __noinline__ double ls_rhfun_exact(double c1, double x, double e, double c2)
#else
LS_FN double ls_rhfun_exact(double c1, double x, double e, double c2)
#endif
{
    return 1.0 / (c1 * b200ode_pow_impl(x, e) + c2);
}
LS_FN double ls_rhfun(double c1, double x, double e, double c2)
{
    if (LS_FAST()) {
        if (!(x <= 1.7976931348623157e308)) return 0.0 / x;  // +inf -> 0, NaN -> NaN as 1/(c*pow+c)
        return b200_rcp(c1 * b200_pow_fast(x, e) + c2);
    }
    return ls_rhfun_exact(c1, x, e, c2);
}
LS_FN double ls_pow(double x, double e)
{
    if (LS_FAST()) return (x <= 1.7976931348623157e308) ? b200_pow_fast(x, e) : x;
    return b200ode_pow(x, e);
}

// hmin / |h| with hmin = 0 (LSODA.cpp:401, :1155, :1243 ...): +0, or NaN when h is 0 or NaN
LS_FN double ls_hmin_ratio(double h) { return 0.0 / fabs(h); }

// idamax1's integer truncation (LSODA.cpp:59-77) as x86-64 evaluates (size_t)|x|:
// cvttsd2si below 2^63, (x - 2^63) with the top bit set up to 2^64, 0 above and for NaN.
LS_FN unsigned long long ls_trunc_u64(double x)
{
    const double a = fabs(x);
    if (!(a < 18446744073709551616.0)) return 0ULL;
#if defined(__CUDA_ARCH__)
    return __double2ull_rz(a);
#else
    return (unsigned long long)a;
#endif
}

// resetcoeff, LSODA.cpp:2211-2225 (el[] is read from the table where it is used)
LS_FN void ls_resetcoeff(LsLane& s, const B200LsodaTables* T)
{
    const double el1 = ELCO(s.nq, 1);
    s.rc = ls_div_rare(s.rc * el1, s.el0);
    s.el0 = el1;
}

// scaleh, LSODA.cpp:1594-1631 (hmxi = 0)
template <class V>
LS_FN void ls_scaleh(LsLane& s, V& v, const B200LsodaTables* T, double& rh, double& pdh)
{
    rh = ls_min(rh, s.rmax);
    rh = ls_div_rare(rh, ls_max(1.0, fabs(s.h) * 0.0 * rh));
    if (s.meth == 1) {
        s.irflag = 0;
        pdh = ls_max(fabs(s.h) * s.pdlast, 0.000001);
        if ((rh * pdh * 1.00001) >= T->sm1[s.nq]) {
            rh = ls_div_rare(T->sm1[s.nq], pdh);
            s.irflag = 1;
        }
    }
    const int l = s.nq + 1;
    v.scale_yh(l, rh);  // r = 1; for j = 2..l: r *= rh; yh[j] *= r
    s.h *= rh;
    s.rc *= rh;
    s.ialth = l;
}

// methodswitch, LSODA.cpp:1946-2069
LS_FN void ls_methodswitch(LsLane& s, const B200LsodaTables* T, double dsm, double pnorm,
                           double& pdh, double& rh)
{
    const int l = s.nq + 1;
    if (s.meth == 1) {
        int nqm2;
        double rh2;
        if (s.nq > 5) return;
        if (dsm <= (100.0 * pnorm * LS_ETA) || s.pdest == 0.0) {
            if (s.irflag == 0) return;
            rh2 = 2.0;
            nqm2 = s.nq < LS_MXORDS ? s.nq : LS_MXORDS;
        } else {
            const double exsm = T->rcp[l];
            double rh1 = ls_rhfun(1.2, dsm, exsm, 0.0000012);
            double rh1it = 2.0 * rh1;
            pdh = s.pdlast * fabs(s.h);
            if ((pdh * rh1) > 0.00001) rh1it = ls_div(T->sm1[s.nq], pdh);
            rh1 = ls_min(rh1, rh1it);
            // (nq <= 5 = mxords here, so the reference's nq > mxords branch, :1988-1994,
            // cannot be taken)
            const double dm2 = dsm * T->cm12[s.nq];
            rh2 = ls_rhfun(1.2, dm2, exsm, 0.0000012);
            nqm2 = s.nq;
            if (rh2 < 5.0 * rh1) return;
        }
        rh = rh2;
        s.icount = 20;
        s.meth = 2;
        s.pdlast = 0.0;
        s.nq = nqm2;
        return;
    }
    // currently BDF, LSODA.cpp:2026-2068 (nq <= mxords <= mxordn: :2030-2036 unreachable)
    const double exsm = T->rcp[l];
    double dm1 = dsm * T->cm21[s.nq];
    double rh1 = ls_rhfun(1.2, dm1, exsm, 0.0000012);
    const int nqm1 = s.nq;
    const double exm1 = exsm;
    double rh1it = 2.0 * rh1;
    pdh = s.pdnorm * fabs(s.h);
    if ((pdh * rh1) > 0.00001) rh1it = ls_div(T->sm1[nqm1], pdh);
    rh1 = ls_min(rh1, rh1it);
    const double rh2 = ls_rhfun(1.2, dsm, exsm, 0.0000012);
    if ((rh1 * 5.0) < (5.0 * rh2)) return;
    const double alpha = ls_max(0.001, rh1);
    dm1 *= ls_pow(alpha, exm1);
    if (dm1 <= 1000.0 * LS_ETA * pnorm) return;
    rh = rh1;
    s.icount = 20;
    s.meth = 1;
    s.pdlast = 0.0;
    s.nq = nqm1;
}

// orderswitch, LSODA.cpp:2097-2209.  Returns orderflag.
template <class V>
LS_FN int ls_orderswitch(LsLane& s, V& v, const B200LsodaTables* T, double rhup, double dsm,
                         double& pdh, double& rh)
{
    const int l = s.nq + 1;
    const int lmax = (s.meth == 1 ? LS_MXORDN : LS_MXORDS) + 1;
    int newq;
    const double exsm = T->rcp[l];
    double rhsm = ls_rhfun(1.2, dsm, exsm, 0.0000012);
    double rhdn = 0.0;
    if (s.nq != 1) {
        const double ddn = ls_div(v.norm_yh(l), TESCO(s.nq, 1));
        const double exdn = T->rcp[s.nq];
        rhdn = ls_rhfun(1.3, ddn, exdn, 0.0000013);
    }
    if (s.meth == 1) {
        pdh = ls_max(fabs(s.h) * s.pdlast, 0.000001);
        if (l < lmax) rhup = ls_min(rhup, ls_div(T->sm1[l], pdh));
        rhsm = ls_min(rhsm, ls_div(T->sm1[s.nq], pdh));
        if (s.nq > 1) rhdn = ls_min(rhdn, ls_div(T->sm1[s.nq - 1], pdh));
        s.pdest = 0.0;
    }
    if (rhsm >= rhup) {
        if (rhsm >= rhdn) {
            newq = s.nq;
            rh = rhsm;
        } else {
            newq = s.nq - 1;
            rh = rhdn;
            if (s.kflag < 0 && rh > 1.0) rh = 1.0;
        }
    } else {
        if (rhup <= rhdn) {
            newq = s.nq - 1;
            rh = rhdn;
            if (s.kflag < 0 && rh > 1.0) rh = 1.0;
        } else {
            rh = rhup;
            if (rh >= 1.1) {
                const double r = ls_div_rare(ELCO(s.nq, l), (double)l);
                s.nq = l;
                v.yh_from_acor_scaled(s.nq + 1, r);
                return 2;
            }
            s.ialth = 3;
            return 0;
        }
    }
    if (s.meth == 1) {
        if ((rh * pdh * 1.00001) < T->sm1[newq])
            if (s.kflag == 0 && rh < 1.1) {
                s.ialth = 3;
                return 0;
            }
    } else {
        if (s.kflag == 0 && rh < 1.1) {
            s.ialth = 3;
            return 0;
        }
    }
    if (s.kflag <= -2) rh = ls_min(rh, 0.2);
    if (newq == s.nq) return 1;
    s.nq = newq;
    return 2;
}

// ---------------------------------------------------------------------------
// The driver: what lsoda_wrapper's loop over t_eval and LSODA::lsoda's blocks a-e
// do for one trajectory, reorganised as  begin / advance-by-one-pass.

// First call: wrapper.cpp:25-36 and blocks b, c of LSODA::lsoda (LSODA.cpp:341-663).
// Returns true when the trajectory is already finished (R.istate says how).
template <class V>
LS_FN bool ls_begin(LsLane& s, LsRun& R, V& v, const double* u0, double* pp, int nt,
                    const double* teval, double* usol, const LsTol& tol)
{
    const int n = v.size();
    v.begin(u0, usol);  // usol row 0 = u0 (wrapper.cpp:25-28), y = u0
    R.row = 1;
    R.nslast = 0;
    R.nhnil = 0;
    R.nswitch = 0;
    R.istate = 2;
    s.nst = s.nfe = s.nje = s.nqu = 0;
    s.meth = 0;
    s.mused = 0;
    s.nq = 1;
    s.jstart = 0;
    s.jcur = 0;
    s.kflag = 0;
    s.m = 0;
    s.phase = LS_P_STEP;
    s.flags = 0;
    s.told = s.pnorm = s.del = s.delp = s.rate = s.rh = 0.0;
    s.hu = 0.0;
    s.tn = 0.0;
    s.h = 0.0;
    if (nt < 2) return true;
    if (teval[1] < teval[0]) {
        R.istate = 0;
        return true;
    }
    const double t0 = teval[0], tout = teval[1];
    s.tn = t0;
    s.meth = 1;
    if (v.tol_negative(tol)) {  // LSODA.cpp:500-520
        R.istate = -3;
        return true;
    }
    v.rhs_yh2(t0, pp);  // yh[2] = f(t0, y), yh[1] = y (:572-578)
    s.nfe = 1;
    if (v.set_ewt(tol)) {
        // LSODA.cpp:585-590 leaves through terminate2 with istate still 1, which
        // wrapper.cpp:41 does not treat as a failure: every later call re-initialises,
        // fails the same way and hands u0 back.
        for (int r = 1; r < nt; r++) {
            if (teval[r] < teval[r - 1]) {
                R.istate = 0;
                return true;
            }
            v.fill_row(usol + (size_t)r * n, u0);
            R.row = r + 1;
        }
        R.istate = 2;
        return true;
    }
    const double tdist = fabs(tout - t0);
    const double w0 = ls_max(fabs(t0), fabs(tout));
    if (tdist < 2.0 * LS_ETA * w0) {  // LSODA.cpp:618-623
        R.istate = -3;
        return true;
    }
    double tl = v.tol_rtol_max(tol);             // :624-628
    if (tl <= 0.0) tl = v.tol_floor(tol, tl);    // :629-640
    tl = ls_max(tl, 100.0 * LS_ETA);
    tl = ls_min(tl, 0.001);
    double sum = v.norm_yh(2);
    sum = 1.0 / (tl * w0 * w0) + tl * sum * sum;
    double h0 = 1.0 / sqrt(sum);
    h0 = ls_min(h0, tdist);
    h0 = h0 * ((tout - t0 >= 0.0) ? 1.0 : -1.0);
    s.h = h0;
    v.scale_yh2(h0);
    R.istate = 1;
    return false;
}

// One pass of the solver's state machine for one trajectory; returns true when the
// trajectory is finished (R.istate says how).
//
// The reference nests three loops -- lsoda's step loop (LSODA.cpp:774), stoda's retry
// loop (:1127) and the corrector iteration (:1763) -- and a warp that runs them as
// written spends most of its time with a few lanes in an inner loop while the others
// wait (ncu, profiles/r1_lsoda_v0_thread_per_traj.txt: 5.1 of 32 lanes active).  Here the nest is
// flattened: every pass performs exactly ONE corrector iteration (one RHS evaluation),
// preceded, for the lanes that need it, by the end of the previous step (rescale,
// output rows), the start of a new step (weights, stop tests) and the prediction, and
// followed, for the lanes whose corrector converged, by the error test and the step /
// order / method selection.  A retry after a failed error or convergence test is just
// the lane's next pass, taken together with the other lanes' next iterations.  The
// arithmetic and its order per trajectory are unchanged.
//
// The function is written without early exits: a `return` inside a divergent branch
// moves the branch's re-convergence point to the end of the function, and the warp
// then runs everything after it once per group of lanes (ncu, profiles/r1_lsoda_v1_flat_unsynced.txt:
// the blocks below were executed 2-3 times per pass, 7 lanes at a time).  `fin` carries
// the exits instead, and v.sync(lanes) -- __syncwarp over the lanes that entered; a
// no-op in the host build and for the warp backend, whose 32 lanes run one trajectory
// in lock step -- pins the lanes together between blocks.
// Deferred backends (V::kDeferred; the thread-per-trajectory kernel, lsoda_kernel.cu).  The
// blocks of a pass that few lanes of a warp need at a time -- the Jacobian and its LU, the
// retraction after a failed attempt, the rescale after a step-size change, the output rows --
// are what kept 12 of 32 lanes busy (ncu, profiles/r2_lsoda_v0_baseline.txt: 41 % of the warp
// instructions ran with 2-7 lanes).  Their vector work touches only what already lives in
// shared memory (yh, the matrix) plus a few values that are cheap to hand over, so a lane that
// needs one of them *requests* it (v.scale_yh, v.pascal<false>, v.intdy_at, v.request_jac)
// and the block serves all requests of a pass together, compacted by kind, between the two
// halves of the pass (ls_advance_a: blocks 0-1; ls_advance_b: blocks 2-5).  Per trajectory the
// operations and their order are unchanged, so strict mode stays bit-exact.  For the other
// backends the requests execute on the spot and ls_advance is the two halves back to back.
template <class V>
LS_FN bool ls_advance_a(LsLane& s, LsRun& R, V& v, const B200LsodaTables* T, int nt,
                        const double* teval, double* usol, unsigned lanes)
{
#define LS_SYNC() v.sync(lanes)
    const int n = v.size();
    double pdh = 0.0;
    bool fin = false;
    // ---- (0) pending step-size change (scaleh), requested by the previous pass
    if (s.flags & LS_F_RESCALE) {
        ls_scaleh(s, v, T, s.rh, pdh);
        if (s.flags & LS_F_RMAX10) s.rmax = 10.0;  // LSODA.cpp:1211,:1258
        s.flags &= ~(LS_F_RESCALE | LS_F_RMAX10);
    }
    LS_SYNC();
    // ---- (1) end of an accepted step: endstoda (LSODA.cpp:2075-2082; acor is not read
    //      again before :1789 clears it, so its scaling is dropped), the method-switch
    //      bookkeeping of lsoda (:861-877) and every output row the step has reached
    if (s.flags & LS_F_FINISH) {
        s.flags &= ~LS_F_FINISH;
        s.hold = s.h;
        s.jstart = 1;
        if (s.meth != s.mused) {
            s.jstart = -1;
            R.nswitch++;
        }
        // The first row is what the reference interpolates at :887, the others are its
        // block-d returns (:675-689) on the following calls, which re-check t_eval's order
        // (wrapper.cpp:33) and intdy's interval (:1426-1434).
        bool first_in_step = true, more = R.row < nt;
        while (more) {
            const double tout = teval[R.row];
            if (tout < teval[R.row - 1]) {
                R.istate = 0;
                fin = true;
                more = false;
            } else if ((s.tn - tout) * s.h < 0.0) {
                more = false;
            } else {
                const double tp = s.tn - s.hu - 100.0 * LS_ETA * (s.tn + s.hu);
                const bool refused = (tout - tp) * (tout - s.tn) > 0.0;
                if (refused && !first_in_step) {
                    // intdy refuses: on a later call that is illegal input
                    R.istate = -3;
                    fin = true;
                    more = false;
                } else {
                    if (refused) {
                        // ... right after a step the reference ignores intdy's refusal and
                        // returns the last corrector iterate
                        v.store_y(usol + (size_t)R.row * n);
                    } else {
                        // intdy with k = 0, LSODA.cpp:1435-1453
                        v.intdy_at(s.nq, s.tn, s.h, tout, R.row, usol);
                    }
                    first_in_step = false;
                    R.nslast = s.nst;
                    R.row++;
                    more = R.row < nt;
                }
            }
        }
        if (!fin) {
            if (R.row >= nt) {
                R.istate = 2;
                fin = true;
            } else {
                s.phase = LS_P_STEP;
            }
        }
    }
    LS_SYNC();
    return fin;
#undef LS_SYNC
}

template <class V>
LS_FN bool ls_advance_b(LsLane& s, LsRun& R, V& v, const B200LsodaTables* T, double* pp,
                        const LsTol& tol, unsigned long long mxstep_u, int exit_on_warning,
                        unsigned lanes)
{
#define LS_SYNC() v.sync(lanes)
    const int n = v.size();
    double pdh = 0.0;
    bool fin = false;
    if (V::kDeferred && (s.flags & LS_F_REFRESH)) {
        // (the tail of the three-failures recovery, LSODA.cpp:1347-1353, after the retraction)
        s.flags &= ~LS_F_REFRESH;
        v.refresh_yh2(s.tn, s.h, pp);
        s.nfe++;
    }
    // ---- (2) start of a step: lsoda's stop tests (:774-845) and stoda's prologue (:1040-1124)
    if (!fin && s.phase == LS_P_STEP) {
        // ewset (:787-798).  Before the first step yh[1] is u0, whose weights ls_begin
        // checked, and the test for non-positive weights is skipped (:776).
        const bool bad_ewt = v.set_ewt(tol);
        const double tolsf = LS_ETA * v.norm_yh(1);
        if (s.nst != 0 && (unsigned long long)(long long)(s.nst - R.nslast) >= mxstep_u) {  // :778
            R.istate = -1;
            fin = true;
        } else if (s.nst != 0 && bad_ewt) {  // :793
            R.istate = -6;
            fin = true;
        } else if (tolsf > 0.01) {  // :801-818
            R.istate = (s.nst == 0) ? -3 : -2;
            fin = true;
        } else {
            if ((s.tn + s.h) == s.tn) {  // :820-845
                if (R.nhnil <= LS_MXHNIL) R.nhnil++;
                if (R.nhnil <= LS_MXHNIL && exit_on_warning) {
                    R.istate = -2;
                    fin = true;
                }
            }
            if (!fin) {
                s.kflag = 0;
                s.told = s.tn;
                s.jcur = 0;
                s.phase = LS_P_ATTEMPT;
                if (s.jstart == 0) {
                    // first call, LSODA.cpp:1058-1084 (l and lmax follow from nq and meth)
                    s.nq = 1;
                    s.ialth = 2;
                    s.rmax = 10000.0;
                    s.rc = 0.0;
                    s.el0 = 1.0;
                    s.crate = 0.7;
                    s.hold = s.h;
                    s.nslp = 0;
                    s.ipup = 0;  // miter = 0 for meth = 1
                    s.icount = 20;
                    s.irflag = 0;
                    s.pdest = 0.0;
                    s.pdlast = 0.0;
                    ls_resetcoeff(s, T);
                }
                if (s.jstart == -1) {
                    // after a method switch, LSODA.cpp:1096-1124
                    s.ipup = (s.meth == 2) ? 2 : 0;
                    if (s.ialth == 1) s.ialth = 2;
                    if (s.meth != s.mused) {
                        s.ialth = s.nq + 1;
                        ls_resetcoeff(s, T);
                    }
                    if (s.h != s.hold) {
                        // (cannot happen in this driver -- nothing changes h between steps --
                        // but kept: the rescale is done at the top of the lane's next pass,
                        // and this pass ends here)
                        s.rh = ls_div_rare(s.h, s.hold);
                        s.h = s.hold;
                        s.flags |= LS_F_RESCALE;
                    }
                }
            }
        }
    }
    LS_SYNC();
    const int miter = (s.meth == 2) ? 2 : 0;
    const bool go = !fin && !(s.flags & LS_F_RESCALE);
    // ---- (3) start of an attempt: prediction, LSODA.cpp:1132-1147, and the corrector's
    //      initialisation, :1748-1762
    if (go && s.phase == LS_P_ATTEMPT) {
        if (fabs(s.rc - 1.0) > LS_CCMAX) s.ipup = miter;
        if (s.nst >= s.nslp + LS_MSBP) s.ipup = miter;
        s.tn += s.h;
        v.template pascal<true>(s.nq);
        s.pnorm = v.norm_yh(1);
        s.m = 0;
        s.rate = 0.0;
        s.del = 0.0;
        s.delp = 0.0;
        v.y_from_yh1();
        s.phase = LS_P_ITER;
    }
    LS_SYNC();
    // ---- (4) corrector iterations, LSODA.cpp:1763-1900: up to LS_ITERS_PER_PASS of them, the
    //      later ones only for the lanes that have not converged yet.  (With one iteration
    //      per pass, a third of the lanes was in its second iteration at any time and every
    //      other block of the pass -- start of step, prediction, error test, step / order /
    //      method selection -- ran with two thirds of its lanes: ncu, profiles/r1_lsoda_v3.
    //      The loop is one code site: the kernel does not grow.)
    bool corfail = false, converged = false;
    bool iterate = go;
    // deferred backends: the matrix requested on the previous pass is there; take it and go on
    // with the iteration whose right-hand side (savf) this lane still holds
    bool resumed = false;
    if (V::kDeferred && go && s.phase == LS_P_JACWAIT) {
        const int ierpj = v.take_jac(s.pdnorm);
        s.nfe += v.jac_rhs_calls();
        s.ipup = 0;
        s.rc = 1.0;
        s.nslp = s.nst;
        s.crate = 0.7;
        if (ierpj != 0) corfail = true;
        s.phase = LS_P_ITER;
        resumed = true;
    }
#pragma unroll 1
    for (int it = 0; it < LS_ITERS_PER_PASS; it++) {
        if (iterate && !resumed) {
            v.rhs_savf(s.tn, pp);
            s.nfe++;
            if (s.m == 0 && s.ipup > 0) {
                // prja (LSODA.cpp:1633-1696): finite-difference Jacobian, P = I - h el0 J, LU
                s.nje++;
                s.jcur = 1;
                if (V::kDeferred) {
                    v.request_jac(s.tn, s.h, s.h * s.el0);
                    s.phase = LS_P_JACWAIT;
                    iterate = false;  // this lane's pass ends here
                } else {
                    const int ierpj = v.prja(s.tn, s.h, s.h * s.el0, pp, s.pdnorm);
                    s.nfe += v.jac_rhs_calls();  // n (LSODA.cpp:1678), or the column groups of a banded one
                    s.ipup = 0;
                    s.rc = 1.0;
                    s.nslp = s.nst;
                    s.crate = 0.7;
                    if (ierpj != 0) corfail = true;
                }
            }
        }
        resumed = false;
        LS_SYNC();
        if (iterate && !corfail) {
            if (s.m == 0) v.acor_zero();
            const double el1 = ELCO(s.nq, 1);
            if (miter == 0)
                s.del = v.functional(s.h, el1);  // functional iteration, :1792-1809
            else
                s.del = v.chord(s.h, el1);  // chord iteration, :1816-1829
            // convergence test, :1842-1862
            if (s.del <= 100.0 * s.pnorm * LS_ETA) {
                converged = true;
            } else if (s.m != 0 || s.meth != 1) {
                if (s.m != 0) {
                    double rm = 1024.0;
                    if (s.del <= (1024.0 * s.delp)) rm = ls_div(s.del, s.delp);
                    s.rate = ls_max(s.rate, rm);
                    s.crate = ls_max(0.2 * s.crate, rm);
                }
                const double dcon = ls_div(s.del * ls_min(1.0, 1.5 * s.crate), TESCO(s.nq, 2) * T->conit[s.nq]);
                if (dcon <= 1.0) {
                    s.pdest = ls_max(s.pdest, ls_div(s.rate, fabs(s.h * el1)));
                    if (s.pdest != 0.0) s.pdlast = s.pdest;
                    converged = true;
                }
            }
            if (!converged) {
                s.m++;
                if (s.m == LS_MAXCOR || (s.m >= 2 && s.del > 2.0 * s.delp)) {
                    // :1871-1891
                    if (miter == 0 || s.jcur == 1) {
                        corfail = true;
                    } else {
                        s.ipup = miter;
                        s.m = 0;
                        s.rate = 0.0;
                        s.del = 0.0;
                        v.y_from_yh1();
                    }
                } else {
                    s.delp = s.del;
                }
            }
        }
        iterate = iterate && !converged && !corfail;
        LS_SYNC();
    }
    // ---- (5) the attempt is over: local error test, LSODA.cpp:1170-1177
    double dsm = 0.0, rhup = 0.0;
    bool accepted = false, select_order = false;
    if (converged) {
        s.jcur = 0;
        if (s.m == 0) dsm = ls_div(s.del, TESCO(s.nq, 2));
        if (s.m > 0) dsm = ls_div(v.norm_acor(), TESCO(s.nq, 2));
        accepted = dsm <= 1.0;
    }
    if (accepted) {
        // LSODA.cpp:1191-1277
        const int l = s.nq + 1;
        const int lmax = (s.meth == 1 ? LS_MXORDN : LS_MXORDS) + 1;
        s.kflag = 0;
        s.nst++;
        s.hu = s.h;
        s.nqu = s.nq;
        s.mused = s.meth;
        v.accept(l, &ELCO(s.nq, 0));  // yh[j] += el[j] * acor, j = 1..l
        s.flags |= LS_F_FINISH;
        // (the reference decrements on every step, LSODA.cpp:1203, and only asks for the sign; the
        // 16-bit field stops at -1 instead of wrapping after 32 k steps without a method switch)
        if (s.icount >= 0) s.icount--;
        bool switched = false;
        if (s.icount < 0) {
            ls_methodswitch(s, T, dsm, s.pnorm, pdh, s.rh);
            if (s.meth != s.mused) {
                s.rh = ls_max(s.rh, ls_hmin_ratio(s.h));
                s.flags |= LS_F_RESCALE | LS_F_RMAX10;
                switched = true;
            }
        }
        if (!switched) {
            s.ialth--;
            if (s.ialth == 0) {
                if (l != lmax) {
                    // savf = acor - yh[lmax], :1222-1228
                    const double dup = ls_div(v.norm_acor_minus_yh(lmax), TESCO(s.nq, 3));
                    const double exup = T->rcp[l + 1];
                    rhup = ls_rhfun(1.4, dup, exup, 0.0000014);
                }
                select_order = true;
            } else if (!(s.ialth > 1 || l == lmax)) {
                v.yh_from_acor(lmax);
            }
        }
    } else if (converged || corfail) {
        // the attempt failed: corfailure (:1904-1922) or error-test failure (:1286-1363);
        // both retract the prediction
        s.rmax = 2.0;
        s.tn = s.told;
        v.request_retract(s.nq);  // pascal<false>: :1290-1295, :1909-1912
        s.phase = LS_P_ATTEMPT;
        const bool h_is_zero = fabs(s.h) <= 0.0 * 1.00001;  // hmin = 0
        if (corfail) {
            // (the reference's mxncf limit never fires: it increments the pointer, :1906)
            if (h_is_zero) {
                s.kflag = -2;
                R.istate = -5;  // :976
                fin = true;
            } else {
                s.rh = 0.25;
                s.ipup = miter;
                s.rh = ls_max(s.rh, ls_hmin_ratio(s.h));
                s.flags |= LS_F_RESCALE;
            }
        } else {
            s.kflag--;
            if (h_is_zero || s.kflag == -10) {
                s.kflag = -1;
                R.istate = -4;  // :969
                fin = true;
            } else if (s.kflag > -3) {
                select_order = true;
            } else {
                // three or more failures: back to order 1 with a fresh derivative, :1333-1361
                s.rh = 0.1;
                s.rh = ls_max(ls_hmin_ratio(s.h), s.rh);
                s.h *= s.rh;
                if (V::kDeferred) {
                    s.flags |= LS_F_REFRESH;  // after the retraction requested above is served
                } else {
                    v.refresh_yh2(s.tn, s.h, pp);  // y = yh[1]; yh[2] = h f(tn, y)
                    s.nfe++;
                }
                s.ipup = miter;
                s.ialth = 5;
                if (s.nq != 1) {
                    s.nq = 1;
                    ls_resetcoeff(s, T);
                }
            }
        }
    }
    LS_SYNC();
    if (select_order) {
        const int orderflag = ls_orderswitch(s, v, T, rhup, dsm, pdh, s.rh);
        if (accepted) {
            if (orderflag != 0) {
                if (orderflag == 2) ls_resetcoeff(s, T);
                s.rh = ls_max(s.rh, ls_hmin_ratio(s.h));
                s.flags |= LS_F_RESCALE | LS_F_RMAX10;
            }
        } else {
            if (orderflag == 0) s.rh = ls_min(s.rh, 0.2);
            if (orderflag == 2) ls_resetcoeff(s, T);
            s.rh = ls_max(s.rh, ls_hmin_ratio(s.h));
            s.flags |= LS_F_RESCALE;
        }
    }
    LS_SYNC();
    return fin;
#undef LS_SYNC
}

// One pass for a backend that executes every operation on the spot.
template <class V>
LS_FN bool ls_advance(LsLane& s, LsRun& R, V& v, const B200LsodaTables* T, double* pp, int nt,
                      const double* teval, double* usol, const LsTol& tol,
                      unsigned long long mxstep_u, int exit_on_warning, unsigned lanes)
{
    const bool fin = ls_advance_a(s, R, v, T, nt, teval, usol, lanes);
    // (the lanes that go on: v.sync inside the second half must not wait for the finished ones)
    const unsigned lanes_b = v.vote(lanes, !fin);
    if (fin) return true;
    return ls_advance_b(s, R, v, T, pp, tol, mxstep_u, exit_on_warning, lanes_b);
}

#endif
