// LSODA vector backend: one warp owns one trajectory of 9 <= n <= 128 equations.
//
// The vector work of the reference's routines -- vmnorm / fnorm (LSODA.cpp:1705-1736),
// ewset (:1368-1390), the Pascal sweeps of stoda (:1141-1144, :1290-1295), scaleh's
// rescale (:1621-1627), prja (:1633-1696), dgefa / dgesl (:110-231), the corrector
// updates (:1792-1829), the Nordsieck update (:1197-1202) and intdy (:1435-1453) -- for
// the warp-per-trajectory kernel (lsoda_warp_kernel.cu).  Control flow: lsoda_core.cuh;
// all 32 lanes execute it redundantly on the same scalars, so there is no divergence.
//
// Storage (shared memory, per warp): yh[13][n], y, savf, acor, ewt, the iteration
// matrix stored transposed with a padded leading dimension (wmT[c*ld + r] = a[r][c],
// ld = n + 1: a lane per row is conflict free, a lane per column 2-way), one private
// copy of y per lane for the finite-difference Jacobian, the pivot vector.
//
// Parallel decomposition, chosen so that every floating-point result is the one the
// reference's sequential loops produce (bit-for-bit in strict mode):
//   * elementwise vector updates, Pascal sweeps, interpolation: lane i owns components
//     i, i + 32, ...;
//   * max-norms and the pivot search: per-lane partial results combined by shuffles
//     (max and first-arg-max are order independent);
//   * Jacobian (prja): the n perturbed RHS evaluations are independent -- 8 columns at
//     a time, one per lane, each lane calling the user's RHS on its private copy of y and
//     writing straight into the matrix column; fnorm's row sums run over j in order, a
//     lane per row;
//   * LU (dgefa): the row-k scaling and the rank-1 update touch independent elements:
//     a lane per row of the trailing block;
//   * forward substitution (dgesl): column sweep -- once b[c] is final every later row
//     adds a[k][c] * b[c] to its private running sum, which accumulates the terms in the
//     same order c = 0, 1, ... as the reference's dot product;
//   * back substitution: the reference's dot products run over ascending c with pivot
//     swaps in between, which fixes a sequential summation order: n^2/2 dependent adds.
//     Strict mode keeps it (bit parity); the fast mode (`fast`, the default of the
//     library) factors with full column interchanges -- the swaps are also applied to the
//     multipliers above the diagonal -- which turns the back substitution into a swap-free
//     column sweep (n dependent steps, lanes own rows) followed by one permutation, and
//     multiplies by precomputed reciprocals of the diagonal in the forward sweep.  Same
//     mathematics, different summation order: last-bit differences only;
//   * the RHS of a corrector iteration is one call of the user's scalar function: lane 0.
#ifndef B200ODE_LSODA_VEC_WARP_CUH
#define B200ODE_LSODA_VEC_WARP_CUH

#include "lsoda_core.cuh"

#include "b200ode_warp.cuh"

#define WV_NMAX 128
// Jacobian columns evaluated concurrently (each needs a private copy of y in shared
// memory): 8 keeps a trajectory at 47 KB for n = 64, i.e. 4 warps per SM instead of 3,
// and the n/8 rounds of RHS evaluations stay a small part of prja next to the LU
#define WV_JCOLS (WV_NL < B200ODE_LSODAW_JCOLS ? WV_NL : B200ODE_LSODAW_JCOLS)
#define WV_NPL (WV_NMAX / WV_NL)  // rows a lane can own
// (a lane owns at most WV_NPL = 4 components: unrolling these loops only grows the code, and the
// warp kernels are bound by instruction fetch -- 140 KB of SASS against a 32 KB instruction cache)
#if defined(__CUDA_ARCH__)
#define WV_FOR_OWN(i) _Pragma("unroll 1") for (int i = WV_LANE; i < n; i += WV_NL)
#else
#define WV_FOR_OWN(i) for (int i = WV_LANE; i < n; i += WV_NL)
#endif
#define WYH(j, i) yh[((j) - 1) * n + (i)]
#define WMT(r, c) wmT[(c) * ld + (r)]  // a[r][c]

#include "b200ode_launch.h"  // B200ODE_LSODAW_DOUBLES(n): shared memory of one trajectory

// vmnorm (LSODA.cpp:1705-1712): max_i |v_i| w_i; one out-of-line copy (it is used from a
// dozen places and the warp kernel is bound by instruction fetch)
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
static inline
#endif
double wv_wmax_norm(const double* v, const double* w, int n)
{
    double vm = 0.0;
    WV_FOR_OWN(i) vm = ls_max(vm, fabs(v[i]) * w[i]);
    return wv_max(vm);
}

// ---- tensor memory as row storage (sm_100a).  A warp owns the 32 TMEM lanes of its quadrant
// (warp index mod 4): lane l keeps rows l and l + 32 of the iteration matrix, row block q at
// columns 2 (64 q + c) .. + 1 (a double is two 32-bit columns).  tcgen05.ld / .st move the same
// column range of every lane at once: eight consecutive entries of each lane's row per
// instruction (.32x32b.x16), or one entry (.x2); the column address is a run-time value
// (tools/tmem_probe.cu checks all of this on the device).  Loads and stores are asynchronous:
// a load's wait is part of the same asm statement, so that no use of the registers can move above
// it; stores are followed by wv_tm_wait_st() before their columns are read again.
#define WV_TM_STRIDE 64  // doubles per row block in tensor memory (n <= 64)
#if defined(__CUDA_ARCH__)
// (the 32-bit halves are paired inside the asm statement: ptxas then loads straight into the
// register pairs of the doubles instead of moving sixteen words around)
#define WV_TM_REGS "t0,t1,t2,t3,t4,t5,t6,t7,t8,t9,t10,t11,t12,t13,t14,t15"
__device__ __forceinline__ void wv_tm_ld8(unsigned taddr, double (&v)[8])
{
    asm volatile("{\n .reg .b32 t<16>;\n"
                 "tcgen05.ld.sync.aligned.32x32b.x16.b32 {" WV_TM_REGS "}, [%8];\n"
                 "tcgen05.wait::ld.sync.aligned;\n"
                 "mov.b64 %0, {t0,t1};\n mov.b64 %1, {t2,t3};\n mov.b64 %2, {t4,t5};\n mov.b64 %3, {t6,t7};\n"
                 "mov.b64 %4, {t8,t9};\n mov.b64 %5, {t10,t11};\n mov.b64 %6, {t12,t13};\n mov.b64 %7, {t14,t15};\n}"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]), "=d"(v[4]), "=d"(v[5]), "=d"(v[6]), "=d"(v[7])
                 : "r"(taddr)
                 : "memory");
}
// two windows, one wait
__device__ __forceinline__ void wv_tm_ld8x2(unsigned ta, unsigned tb, double (&v)[8], double (&w)[8])
{
    asm volatile("{\n .reg .b32 t<16>;\n .reg .b32 s<16>;\n"
                 "tcgen05.ld.sync.aligned.32x32b.x16.b32 {" WV_TM_REGS "}, [%16];\n"
                 "tcgen05.ld.sync.aligned.32x32b.x16.b32 {s0,s1,s2,s3,s4,s5,s6,s7,s8,s9,s10,s11,s12,s13,s14,s15}, [%17];\n"
                 "tcgen05.wait::ld.sync.aligned;\n"
                 "mov.b64 %0, {t0,t1};\n mov.b64 %1, {t2,t3};\n mov.b64 %2, {t4,t5};\n mov.b64 %3, {t6,t7};\n"
                 "mov.b64 %4, {t8,t9};\n mov.b64 %5, {t10,t11};\n mov.b64 %6, {t12,t13};\n mov.b64 %7, {t14,t15};\n"
                 "mov.b64 %8, {s0,s1};\n mov.b64 %9, {s2,s3};\n mov.b64 %10, {s4,s5};\n mov.b64 %11, {s6,s7};\n"
                 "mov.b64 %12, {s8,s9};\n mov.b64 %13, {s10,s11};\n mov.b64 %14, {s12,s13};\n mov.b64 %15, {s14,s15};\n}"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]), "=d"(v[4]), "=d"(v[5]), "=d"(v[6]), "=d"(v[7]),
                   "=d"(w[0]), "=d"(w[1]), "=d"(w[2]), "=d"(w[3]), "=d"(w[4]), "=d"(w[5]), "=d"(w[6]), "=d"(w[7])
                 : "r"(ta), "r"(tb)
                 : "memory");
}
// (asynchronous: wv_tm_wait_st before the stored columns are loaded again)
__device__ __forceinline__ void wv_tm_st8(unsigned taddr, const double (&v)[8])
{
    asm volatile("{\n .reg .b32 t<16>;\n"
                 "mov.b64 {t0,t1}, %1;\n mov.b64 {t2,t3}, %2;\n mov.b64 {t4,t5}, %3;\n mov.b64 {t6,t7}, %4;\n"
                 "mov.b64 {t8,t9}, %5;\n mov.b64 {t10,t11}, %6;\n mov.b64 {t12,t13}, %7;\n mov.b64 {t14,t15}, %8;\n"
                 "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {" WV_TM_REGS "};\n}"
                 :
                 : "r"(taddr), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]), "d"(v[4]), "d"(v[5]), "d"(v[6]), "d"(v[7])
                 : "memory");
}
__device__ __forceinline__ void wv_tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ double wv_tm_ld1(unsigned taddr)
{
    double x;
    asm volatile("{\n .reg .b32 t<2>;\n"
                 "tcgen05.ld.sync.aligned.32x32b.x2.b32 {t0,t1}, [%1];\n"
                 "tcgen05.wait::ld.sync.aligned;\n"
                 "mov.b64 %0, {t0,t1};\n}"
                 : "=d"(x)
                 : "r"(taddr)
                 : "memory");
    return x;
}
__device__ __forceinline__ void wv_tm_st1(unsigned taddr, double x)
{
    asm volatile("{\n .reg .b32 t<2>;\n"
                 "mov.b64 {t0,t1}, %1;\n"
                 "tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {t0,t1};\n}"
                 :
                 : "r"(taddr), "d"(x)
                 : "memory");
}
#else
// (host builds instantiate the shared-memory backend only)
static inline void wv_tm_ld8(unsigned, double (&)[8]) {}
static inline void wv_tm_ld8x2(unsigned, unsigned, double (&)[8], double (&)[8]) {}
static inline void wv_tm_st8(unsigned, const double (&)[8]) {}
static inline void wv_tm_wait_st() {}
static inline double wv_tm_ld1(unsigned) { return 0.0; }
static inline void wv_tm_st1(unsigned, double) {}
#endif

// TM: the dense iteration matrix lives in tensor memory instead of shared memory (32 < n <= 64).
// At n = 64 the matrix is 32 KB of the 47 KB a trajectory needs, and shared memory holds four
// trajectories per SM -- one warp per scheduler, nothing to hide latency with.  Tensor memory is
// another 256 KB per SM that this kernel has no other use for: eight matrices, eight warps.
// BAND: 1 = the matrix is banded, 0 = dense, -1 = decided at run time by ml (host harness).  The
// three kernels each compile one of the two paths only: they are bound by instruction fetch.
template <bool TM, int BAND = -1>
struct LsWarpVecT {
    unsigned tm;  // TM: tensor-memory address of this warp's 32 lanes x 256 columns
    double* fcopy;  // TM: right-hand sides of the perturbed copies (the matrix columns' staging)
    LS_MFN unsigned tma(int q, int c) const { return tm + 2u * (unsigned)(q * WV_TM_STRIDE + c); }
    int n, ld;
    bool fast;  // parallel triangular solves (see above); false = the reference's summation order
    // Banded iteration matrix (jt = 5 with ml, mu, which the
    // reference's checker accepts, LSODA.cpp:369-392, and its prja / solsy do not implement):
    // ml >= 0 selects it.  The matrix then lives in LINPACK band storage, abd[(2 ml + mu + 1) x n]
    // column-major in the place of wmT, the Jacobian costs min(ml + mu + 1, n) right-hand-side
    // evaluations (columns j, j + mband, ... perturbed together) instead of n, and the LU and
    // the solves touch the band only: ODEPACK's DPRJA / DSOLSY miter = 5 with LINPACK's
    // dgbfa / dgbsl, element for element (oracle/lsoda_oracle.c restates the same routines).
    int ml, mu;
    double *yh, *y, *savf, *acor, *ewt, *rdiag, *wmT, *ucopy;
    int *ipvt, *perm;

    // `matrix`: storage of the iteration matrix outside `base` (the kernel's L2-resident
    // per-block scratch), or null to keep it with the vectors in shared memory
    LS_MFN void attach(double* base, int neq, bool fast_mode, double* matrix = nullptr,
                       int band_ml = -1, int band_mu = -1)
    {
        n = neq;
        ld = neq + 1;
        fast = fast_mode;
        ml = band_ml;
        mu = band_mu;
        yh = base;
        y = yh + 13 * n;
        savf = y + n;
        acor = savf + n;
        ewt = acor + n;
        rdiag = ewt + n;
        double* next = rdiag + n;
        if (TM) {
            wmT = nullptr;
            fcopy = next;
            next += B200ODE_LSODAW_JCOLS * ld;
        } else if (matrix) {
            wmT = matrix;
        } else {
            wmT = next;
            next += banded() ? B200ODE_LSODAW_BAND_MATRIX(n, ml, mu) : n * ld;
        }
        ucopy = next;
        ipvt = reinterpret_cast<int*>(ucopy + (banded() ? B200ODE_LSODAW_BAND_COPIES(n, ml, mu) * 2
                                                         : B200ODE_LSODAW_JCOLS) * ld);
        perm = ipvt + n;
    }
    LS_MFN bool banded() const { return BAND < 0 ? ml >= 0 : BAND == 1; }
    LS_MFN int jac_rhs_calls() const { return banded() && ml + mu + 1 < n ? ml + mu + 1 : n; }
    LS_MFN int size() const { return n; }
    LS_MFN void sync(unsigned) {}
    LS_MFN unsigned vote(unsigned lanes, bool) const { return lanes; }
    // every operation executes on the spot (lsoda_core.cuh: deferred backends)
    static const bool kDeferred = false;
    LS_MFN void request_jac(double, double, double) {}
    LS_MFN int take_jac(double&) { return 0; }
    LS_MFN void request_retract(int nq) { pascal<false>(nq); }
    LS_MFN void intdy_at(int nq, double tn, double h, double tout, int row, double* usol) const
    {
        intdy(nq, ls_div(tout - tn, h), usol + (size_t)row * n);
    }

    LS_MFN double norm_ptr(const double* v) const { return wv_wmax_norm(v, ewt, n); }
    LS_MFN double norm_yh(int j) const { return norm_ptr(yh + (j - 1) * n); }
    LS_MFN double norm_acor() const { return norm_ptr(acor); }
    LS_MFN double norm_acor_minus_yh(int j)
    {
        WV_FOR_OWN(i) savf[i] = acor[i] - WYH(j, i);
        return norm_ptr(savf);
    }

    LS_MFN bool set_ewt(const LsTol& tol)
    {
        bool bad = false;
        WV_FOR_OWN(i) {
            double w = tol.rt(i) * fabs(WYH(1, i)) + tol.at(i);
            if (w <= 0.0) bad = true;
            ewt[i] = ls_div(1.0, w);
        }
        bad = wv_any(bad);
        WV_SYNC();
        return bad;
    }

    template <bool FORWARD>
    LS_MFN void pascal(int nq)
    {
        WV_FOR_OWN(i) {
            for (int j = nq; j >= 1; j--) {
                double a = WYH(j, i);
                for (int i1 = j; i1 <= nq; i1++) {
                    const double b = WYH(i1 + 1, i);
                    WYH(i1, i) = FORWARD ? a + b : a - b;
                    a = b;
                }
            }
        }
        WV_SYNC();
    }

    LS_MFN void scale_yh(int l, double rh)
    {
        double r = 1.0;
        for (int j = 2; j <= l; j++) {
            r *= rh;
            WV_FOR_OWN(i) WYH(j, i) = WYH(j, i) * r;
        }
        WV_SYNC();
    }

    LS_MFN void y_from_yh1()
    {
        WV_FOR_OWN(i) y[i] = WYH(1, i);
        WV_SYNC();
    }
    LS_MFN void rhs_to(double t, double* out, double* pp)
    {
        wv_user_rhs_all(t, y, out, pp);  // (barriers inside; lane 0 alone without a lanes function)
    }
    LS_MFN void rhs_savf(double t, double* pp) { rhs_to(t, savf, pp); }
    LS_MFN void acor_zero()
    {
        WV_FOR_OWN(i) acor[i] = 0.0;
        WV_SYNC();
    }

    LS_MFN double functional(double h, double el1)
    {
        WV_FOR_OWN(i) {
            savf[i] = h * savf[i] - WYH(2, i);
            y[i] = savf[i] - acor[i];
        }
        const double del = norm_ptr(y);
        WV_FOR_OWN(i) {
            y[i] = WYH(1, i) + el1 * savf[i];
            acor[i] = savf[i];
        }
        WV_SYNC();
        return del;
    }

    LS_MFN double chord(double h, double el1)
    {
        WV_FOR_OWN(i) y[i] = h * savf[i] - (WYH(2, i) + acor[i]);
        WV_SYNC();
        if (banded()) dgbsl(y);
        else dgesl(y);
        const double del = norm_ptr(y);
        WV_FOR_OWN(i) {
            acor[i] += y[i];
            y[i] = WYH(1, i) + el1 * acor[i];
        }
        WV_SYNC();
        return del;
    }

    LS_MFN void accept(int l, const double* el)
    {
        for (int j = 1; j <= l; j++) {
            const double r = el[j];
            WV_FOR_OWN(i) WYH(j, i) = WYH(j, i) + r * acor[i];
        }
        WV_SYNC();
    }
    LS_MFN void yh_from_acor(int j)
    {
        WV_FOR_OWN(i) WYH(j, i) = acor[i];
        WV_SYNC();
    }
    LS_MFN void yh_from_acor_scaled(int j, double r)
    {
        WV_FOR_OWN(i) WYH(j, i) = acor[i] * r;
        WV_SYNC();
    }
    LS_MFN void refresh_yh2(double t, double h, double* pp)
    {
        y_from_yh1();
        rhs_savf(t, pp);
        WV_FOR_OWN(i) WYH(2, i) = h * savf[i];
        WV_SYNC();
    }

    LS_MFN void intdy(int nq, double sfrac, double* out) const
    {
        WV_FOR_OWN(i) {
            double v = 1.0 * WYH(nq + 1, i);
            for (int j = nq - 1; j >= 0; j--) v = 1.0 * WYH(j + 1, i) + sfrac * v;
            out[i] = v;
        }
    }
    LS_MFN void store_y(double* out) const { WV_FOR_OWN(i) out[i] = y[i]; }

    // ---- first call (ls_begin)
    LS_MFN void begin(const double* u0, double* usol)
    {
        WV_FOR_OWN(i) {
            y[i] = u0[i];
            usol[i] = y[i];
        }
        WV_SYNC();
    }
    LS_MFN void fill_row(double* out, const double* u0) const { WV_FOR_OWN(i) out[i] = u0[i]; }
    LS_MFN void rhs_yh2(double t0, double* pp)
    {
        rhs_to(t0, yh + n, pp);
        WV_FOR_OWN(i) WYH(1, i) = y[i];
        WV_SYNC();
    }
    LS_MFN bool tol_negative(const LsTol& tol) const
    {
        bool neg = false;
#pragma unroll 1
        WV_FOR_OWN(i) if (tol.rt(i) < 0.0 || tol.at(i) < 0.0) neg = true;
        return wv_any(neg);
    }
    LS_MFN double tol_rtol_max(const LsTol& tol) const
    {
        double t = tol.rt(0);
#pragma unroll 1
        WV_FOR_OWN(i) t = ls_max(t, tol.rt(i));
        return wv_max(t);
    }
    LS_MFN double tol_floor(const LsTol& tol, double t) const
    {
#pragma unroll 1
        WV_FOR_OWN(i) {
            const double ayi = fabs(y[i]);
            if (ayi != 0.0) t = ls_max(t, tol.at(i) / ayi);
        }
        return wv_max(t);
    }
    LS_MFN void scale_yh2(double h0)
    {
        WV_FOR_OWN(i) WYH(2, i) = WYH(2, i) * h0;
        WV_SYNC();
    }

    // ---- linear algebra
    // dgefa, LSODA.cpp:173-231.  Returns nonzero when a zero pivot was met.
    // The same factorisation with the matrix in tensor memory.  Row k is first copied to shared
    // memory by the lane that owns it (the pivot search and every row's update read it); then
    // each lane updates eight entries of its own rows per load / store.  Element for element the
    // arithmetic is the loop below's.  The pivots (the diagonal of the factors) are collected in
    // rdiag on the way: a lane's diagonal entries sit at different columns, which a tensor-memory
    // load cannot address at once.
    // eight entries of this lane's row r, columns right of the pivot column: the rank-one update
    // (rows below the pivot row), the scaled pivot row itself (row k), and row k + 1 filed for
    // the next step
    LS_MFN void tm_update8(double (&v)[8], const double* prow8, double* pnext8, double tr, bool below,
                           bool isk, bool isnext) const
    {
        double p[8];
#pragma unroll
        for (int e = 0; e < 8; e++) p[e] = prow8[e];
        if (below) {
#pragma unroll
            for (int e = 0; e < 8; e++) v[e] = tr * p[e] + v[e];
        } else if (isk) {
#pragma unroll
            for (int e = 0; e < 8; e++) v[e] = p[e];
        }
        if (isnext) {
#pragma unroll
            for (int e = 0; e < 8; e++) pnext8[e] = v[e];
        }
    }

    LS_MFN int dgefa_tm()
    {
        int info = 0;
        // row k in shared memory, double buffered: the lane that owns row k + 1 files it while it
        // updates it (ucopy is free between the Jacobian's evaluation and the next one)
        double* prow = ucopy;
        double* pnext = ucopy + WV_TM_STRIDE;
        bool have = false;
        const int nblk = (n + 31) >> 5;
        for (int k = 0; k < n - 1; k++) {
            const int qk = k >> 5, lk = k & 31, ck0 = k & ~7;
            if (!have) {
                for (int c0 = ck0; c0 < n; c0 += 8) {
                    double v[8];
                    wv_tm_ld8(tma(qk, c0), v);
                    if (WV_LANE == lk) {
#pragma unroll
                        for (int e = 0; e < 8; e++) prow[c0 + e] = v[e];
                    }
                }
                WV_SYNC();
            }
            have = false;
            unsigned long long best = 0;
            int bj = n;
            for (int c = k + WV_LANE; c < n; c += WV_NL) {
                const unsigned long long v = ls_trunc_u64(prow[c]);
                if (v > best) {
                    best = v;
                    bj = c;
                }
            }
            wv_argmax(best, bj);
            const int j = (best == 0) ? k : bj;
            const double akj = prow[j];
            WV_SYNC();
            if (WV_LANE == 0) {
                ipvt[k] = j;
                rdiag[k] = akj;
            }
            if (akj == 0.0) {
                info = k + 1;
                continue;
            }
            if (j != k && WV_LANE == 0) {
                prow[j] = prow[k];
                prow[k] = akj;
            }
            WV_SYNC();
            const double t = ls_div(-1.0, akj);
            for (int c = k + 1 + WV_LANE; c < n; c += WV_NL) prow[c] = t * prow[c];
            WV_SYNC();
            for (int q = 0; q < nblk; q++) {
                const bool has_rows = 32 * q + 31 >= k;  // rows k.. in this block
                if (!has_rows && !(fast && j != k)) continue;
                const int r = WV_LANE + 32 * q;
                const bool below = r > k && r < n, isk = r == k, isnext = r == k + 1;
                // the interchange of columns j and k: rows below the pivot row, and in fast mode the
                // multipliers of the rows above it; the entry left at column k is the multiplier
                const double tr = wv_tm_ld1(tma(q, j));
                if (j != k) {
                    const double ark = wv_tm_ld1(tma(q, k));
                    const bool swap = below || (fast && r < k);
                    wv_tm_st1(tma(q, k), swap ? tr : ark);
                    wv_tm_st1(tma(q, j), swap ? ark : tr);
                    wv_tm_wait_st();
                }
                if (!has_rows) continue;
                {
                    // the chunk that holds column k: entries up to it stay
                    double v[8], p[8];
                    wv_tm_ld8(tma(q, ck0), v);
#pragma unroll
                    for (int e = 0; e < 8; e++) p[e] = prow[ck0 + e];
                    if (below) {
#pragma unroll
                        for (int e = 0; e < 8; e++)
                            if (ck0 + e > k) v[e] = tr * p[e] + v[e];
                    } else if (isk) {
#pragma unroll
                        for (int e = 0; e < 8; e++)
                            if (ck0 + e >= k) v[e] = p[e];
                    }
                    if (isnext) {
#pragma unroll
                        for (int e = 0; e < 8; e++) pnext[ck0 + e] = v[e];
                    }
                    wv_tm_st8(tma(q, ck0), v);
                }
                int c0 = ck0 + 8;
                // ... then two chunks per round trip
                for (; c0 + 8 < n; c0 += 16) {
                    double v[8], w[8];
                    wv_tm_ld8x2(tma(q, c0), tma(q, c0 + 8), v, w);
                    tm_update8(v, prow + c0, pnext + c0, tr, below, isk, isnext);
                    tm_update8(w, prow + c0 + 8, pnext + c0 + 8, tr, below, isk, isnext);
                    wv_tm_st8(tma(q, c0), v);
                    wv_tm_st8(tma(q, c0 + 8), w);
                }
                if (c0 < n) {
                    double v[8];
                    wv_tm_ld8(tma(q, c0), v);
                    tm_update8(v, prow + c0, pnext + c0, tr, below, isk, isnext);
                    wv_tm_st8(tma(q, c0), v);
                }
            }
            wv_tm_wait_st();
            WV_SYNC();
            {
                double* sw = prow;
                prow = pnext;
                pnext = sw;
            }
            have = true;
        }
        // the last diagonal entry: its owner reads it, everybody learns it
        double ann = wv_tm_ld1(tma((n - 1) >> 5, n - 1));
        ann = wv_bcast(ann, (n - 1) & 31);
        if (WV_LANE == 0) {
            ipvt[n - 1] = n - 1;
            rdiag[n - 1] = ann;
        }
        if (ann == 0.0) info = n;
        WV_SYNC();
        if (fast) {
            WV_FOR_OWN(i) {
                rdiag[i] = ls_div(1.0, rdiag[i]);
                perm[i] = i;
            }
            WV_SYNC();
            if (WV_LANE == 0) {
                for (int k = n - 2; k >= 0; k--) {
                    const int j = ipvt[k];
                    if (j != k) {
                        const int pk = perm[k];
                        perm[k] = perm[j];
                        perm[j] = pk;
                    }
                }
            }
            WV_SYNC();
        }
        return info;
    }

    LS_MFN int dgefa()
    {
        if constexpr (TM) return dgefa_tm();
        int info = 0;
        for (int k = 0; k < n - 1; k++) {
            // idamax1 along row k (integer-truncated magnitudes, first maximum)
            unsigned long long best = 0;
            int bj = n;
            for (int c = k + WV_LANE; c < n; c += WV_NL) {
                const unsigned long long v = ls_trunc_u64(WMT(k, c));
                if (v > best) {
                    best = v;
                    bj = c;
                }
            }
            wv_argmax(best, bj);
            const int j = (best == 0) ? k : bj;
            const double akj = WMT(k, j);
            WV_SYNC();
            if (WV_LANE == 0) ipvt[k] = j;
            if (akj == 0.0) {
                info = k + 1;
                continue;
            }
            if (j != k && WV_LANE == 0) {
                WMT(k, j) = WMT(k, k);
                WMT(k, k) = akj;
            }
            if (fast && j != k) {
                // full column interchange: also the multipliers of the rows above
                for (int r = WV_LANE; r < k; r += WV_NL) {
                    const double tr = WMT(r, j);
                    WMT(r, j) = WMT(r, k);
                    WMT(r, k) = tr;
                }
            }
            WV_SYNC();
            const double t = ls_div(-1.0, akj);
            for (int c = k + 1 + WV_LANE; c < n; c += WV_NL) WMT(k, c) = t * WMT(k, c);
            WV_SYNC();
            for (int r = k + 1 + WV_LANE; r < n; r += WV_NL) {
                const double tr = WMT(r, j);
                if (j != k) {
                    WMT(r, j) = WMT(r, k);
                    WMT(r, k) = tr;
                }
                // (measured slower at n = 64: a hand-pipelined pointer-walking version of this
                // loop, -6 %; a lane per column with the pivot-row entry in a register, -7 %)
#pragma unroll 4
                for (int c = k + 1; c < n; c++) WMT(r, c) = tr * WMT(k, c) + WMT(r, c);
            }
            WV_SYNC();
        }
        if (WV_LANE == 0) ipvt[n - 1] = n - 1;
        if (WMT(n - 1, n - 1) == 0.0) info = n;
        WV_SYNC();
        if (fast) {
            // reciprocals of the diagonal, and the composition of the interchanges:
            // x[i] = w[perm[i]] after the swap-free back substitution
            WV_FOR_OWN(i) {
                rdiag[i] = ls_div(1.0, WMT(i, i));
                perm[i] = i;
            }
            WV_SYNC();
            if (WV_LANE == 0) {
                for (int k = n - 2; k >= 0; k--) {
                    const int j = ipvt[k];
                    if (j != k) {
                        const int pk = perm[k];
                        perm[k] = perm[j];
                        perm[j] = pk;
                    }
                }
            }
            WV_SYNC();
        }
        return info;
    }

    // dgesl with job = 0, LSODA.cpp:119-144
    // ... and the solve: the sweeps below with eight columns of each lane's rows per load.
    // rdiag holds the pivots (strict) or their reciprocals (fast).
    LS_MFN void dgesl_tm(double* b) const
    {
        const bool two = n > 32;
        // The right-hand side and the running sums stay in registers (lane l: rows l, l + 32): no
        // shared-memory round trip and no branch in a step of a sweep -- every lane forms the
        // candidate for its own row each step, the broadcast takes the lane whose row it is.
        const int i0 = WV_LANE, i1 = WV_LANE + 32;
        double b0 = b[i0 < n ? i0 : 0], b1 = b[i1 < n ? i1 : 0];
        const double d0 = rdiag[i0 < n ? i0 : 0], d1 = rdiag[i1 < n ? i1 : 0];
        const int n0 = n < 32 ? n : 32;
        if (fast) {
            // The sweeps as recurrences on the candidates themselves: forward s = (b - t) d is
            // kept as s -= (a[k][c] d) x_c, backward s = b + t as s += a[r][c] x_c -- one
            // multiply-add between two broadcasts instead of three operations (the scaled
            // entries are formed off the critical path).  Same mathematics as the sweeps of the
            // shared-memory kernel, different rounding: last-bit differences.
            double s0 = b0 * d0, s1 = b1 * d1;
            for (int c0 = 0; c0 < n0; c0 += 8) {
                double w0[8], w1[8];
                if (two) wv_tm_ld8x2(tma(0, c0), tma(1, c0), w0, w1);
                else wv_tm_ld8(tma(0, c0), w0);
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    w0[e] = -(w0[e] * d0);
                    if (two) w1[e] = -(w1[e] * d1);
                }
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const int c = c0 + e;
                    if (c >= n) break;
                    const double x = wv_bcast(s0, c);
                    if (WV_LANE == c) b0 = x;
                    s0 = fma(w0[e], x, s0);
                    if (two) s1 = fma(w1[e], x, s1);
                }
            }
            for (int c0 = 32; c0 < n; c0 += 8) {
                double w1[8];
                wv_tm_ld8(tma(1, c0), w1);
#pragma unroll
                for (int e = 0; e < 8; e++) w1[e] = -(w1[e] * d1);
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const int c = c0 + e;
                    if (c >= n) break;
                    const double x = wv_bcast(s1, c - 32);
                    if (WV_LANE == c - 32) b1 = x;
                    s1 = fma(w1[e], x, s1);
                }
            }
            s0 = b0;
            s1 = b1;
            for (int c0 = (n - 1) & ~7; c0 >= 32; c0 -= 8) {
                double w0[8], w1[8];
                wv_tm_ld8x2(tma(0, c0), tma(1, c0), w0, w1);
#pragma unroll
                for (int e = 7; e >= 0; e--) {
                    const int c = c0 + e;
                    if (c >= n) continue;
                    const double x = wv_bcast(s1, c - 32);
                    if (WV_LANE == c - 32) b1 = x;
                    s0 = fma(w0[e], x, s0);
                    s1 = fma(w1[e], x, s1);
                }
            }
            for (int c0 = (n0 - 1) & ~7; c0 >= 0; c0 -= 8) {
                double w0[8];
                wv_tm_ld8(tma(0, c0), w0);
#pragma unroll
                for (int e = 7; e >= 0; e--) {
                    const int c = c0 + e;
                    if (c >= n) continue;
                    const double x = wv_bcast(s0, c);
                    if (WV_LANE == c) b0 = x;
                    s0 = fma(w0[e], x, s0);
                }
            }
            // x[i] = w[perm[i]]
            if (i0 < n) b[i0] = b0;
            if (i1 < n) b[i1] = b1;
            WV_SYNC();
            double x[2];
#pragma unroll
            for (int q2 = 0; q2 < 2; q2++) {
                const int i = WV_LANE + 32 * q2;
                x[q2] = (i < n) ? b[perm[i]] : 0.0;
            }
            WV_SYNC();
#pragma unroll
            for (int q2 = 0; q2 < 2; q2++) {
                const int i = WV_LANE + 32 * q2;
                if (i < n) b[i] = x[q2];
            }
            WV_SYNC();
            return;
        }
        // strict: the reference's operations in its order.  t0 / t1: running sums of this lane's
        // rows; read when the row's column comes up, dead afterwards, so rows already solved may
        // go on accumulating.
        double t0 = 0.0, t1 = 0.0;
        for (int c0 = 0; c0 < n0; c0 += 8) {
            double w0[8], w1[8];
            if (two) wv_tm_ld8x2(tma(0, c0), tma(1, c0), w0, w1);
            else wv_tm_ld8(tma(0, c0), w0);
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int c = c0 + e;
                if (c >= n) break;
                const double cand = ls_div(b0 - t0, d0);
                const double bc = wv_bcast(cand, c);
                if (WV_LANE == c) b0 = cand;
                t0 += w0[e] * bc;
                if (two) t1 += w1[e] * bc;
            }
        }
        for (int c0 = 32; c0 < n; c0 += 8) {
            double w1[8];
            wv_tm_ld8(tma(1, c0), w1);
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int c = c0 + e;
                if (c >= n) break;
                const double cand = ls_div(b1 - t1, d1);
                const double bc = wv_bcast(cand, c - 32);
                if (WV_LANE == c - 32) b1 = cand;
                t1 += w1[e] * bc;
            }
        }
        if (i0 < n) b[i0] = b0;
        if (i1 < n) b[i1] = b1;
        WV_SYNC();
        // backward, the reference's order: the owner of row k forms the dot product over ascending
        // c (the other lanes run the same instructions on their own rows; their sums are unused)
        for (int k = n - 2; k >= 0; k--) {
            const int qk = k >> 5;
            double tk = 0.0;
            for (int c0 = (k + 1) & ~7; c0 < n; c0 += 8) {
                double v[8];
                wv_tm_ld8(tma(qk, c0), v);
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const int c = c0 + e;
                    if (c > k && c < n) tk += v[e] * b[c];
                }
            }
            tk = wv_bcast(tk, k & 31);
            WV_SYNC();
            if (WV_LANE == 0) {
                const double bk = b[k] + tk;
                const int j = ipvt[k];
                if (j != k) {
                    const double tt = b[j];
                    b[j] = bk;
                    b[k] = tt;
                } else {
                    b[k] = bk;
                }
            }
            WV_SYNC();
        }
    }

    LS_MFN void dgesl(double* b) const
    {
        if constexpr (TM) {
            dgesl_tm(b);
            return;
        }
        // forward: t_k = sum_{c<k} a[k][c] b[c] accumulated column by column
        double t[WV_NPL];
#pragma unroll
        for (int q = 0; q < WV_NPL; q++) t[q] = 0.0;
#pragma unroll
        for (int q = 0; q < WV_NPL; q++) {
            if (q * WV_NL < n) {
                for (int cc = 0; cc < WV_NL; cc++) {
                    const int c = q * WV_NL + cc;
                    if (c >= n) break;
                    double bc = 0.0;
                    if (WV_LANE == cc) {
                        bc = fast ? (b[c] - t[q]) * rdiag[c] : ls_div(b[c] - t[q], WMT(c, c));
                        b[c] = bc;
                    }
                    bc = wv_bcast(bc, cc);
#pragma unroll
                    for (int q2 = q; q2 < WV_NPL; q2++) {
                        const int k = WV_LANE + WV_NL * q2;
                        if (k > c && k < n) t[q2] += WMT(k, c) * bc;
                    }
                }
            }
        }
        WV_SYNC();
        if (fast) {
            // swap-free column sweep: once w[c] is final every row above adds its term
#pragma unroll
            for (int q = 0; q < WV_NPL; q++) t[q] = 0.0;
#pragma unroll
            for (int q = WV_NPL - 1; q >= 0; q--) {
                if (q * WV_NL < n) {
                    for (int cc = WV_NL - 1; cc >= 0; cc--) {
                        const int c = q * WV_NL + cc;
                        if (c >= n) continue;
                        double wc = 0.0;
                        if (WV_LANE == cc) {
                            wc = b[c] + t[q];
                            b[c] = wc;
                        }
                        wc = wv_bcast(wc, cc);
#pragma unroll
                        for (int q2 = 0; q2 <= q; q2++) {
                            const int r = WV_LANE + WV_NL * q2;
                            if (r < c) t[q2] += WMT(r, c) * wc;
                        }
                    }
                }
            }
            WV_SYNC();
            // x[i] = w[perm[i]]
            double x[WV_NPL];
#pragma unroll
            for (int q = 0; q < WV_NPL; q++) {
                const int i = WV_LANE + WV_NL * q;
                x[q] = (i < n) ? b[perm[i]] : 0.0;
            }
            WV_SYNC();
#pragma unroll
            for (int q = 0; q < WV_NPL; q++) {
                const int i = WV_LANE + WV_NL * q;
                if (i < n) b[i] = x[q];
            }
            WV_SYNC();
            return;
        }
        // backward: dot products over ascending c, pivot swaps in between
        for (int k = n - 2; k >= 0; k--) {
            double tk = 0.0;
            for (int c = k + 1; c < n; c++) tk += WMT(k, c) * b[c];
            WV_SYNC();
            if (WV_LANE == 0) {
                const double bk = b[k] + tk;
                const int j = ipvt[k];
                if (j != k) {
                    const double tt = b[j];
                    b[j] = bk;
                    b[k] = tt;
                } else {
                    b[k] = bk;
                }
            }
            WV_SYNC();
        }
    }

    // prja, LSODA.cpp:1633-1696, with fnorm (:1714-1736) and the LU factorisation.
    // Uses y, savf = f(tn, y), ewt.  Returns ierpj.
    // ---- banded iteration matrix: A(i,j) at abd[j * lda + (i - j + ml + mu)], lda = 2 ml + mu + 1
#define WABD(i, j) wmT[(j) * lda + ((i) - (j) + ml + mu)]
    // dgbfa: partial pivoting by row interchanges inside the band.  One elimination step per k;
    // the multipliers and the (<= ml x (ml + mu)) block they update are independent elements:
    // lanes own them, the arithmetic per element is LINPACK's.
    LS_MFN int dgbfa()
    {
        const int lda = 2 * ml + mu + 1, m = ml + mu;
        int info = 0, ju = 0;
        // zero the initial fill-in columns
        const int j1 = (n < m + 1 ? n : m + 1) - 1;
        for (int jz = mu + 1; jz < j1; jz++)
            for (int i = m - jz + WV_LANE; i < ml; i += WV_NL) wmT[jz * lda + i] = 0.0;
        int jz = j1 - 1;
        WV_SYNC();
        for (int k = 0; k < n - 1; k++) {
            jz++;
            if (jz < n)
                for (int i = WV_LANE; i < ml; i += WV_NL) wmT[jz * lda + i] = 0.0;
            const int lm = ml < n - 1 - k ? ml : n - 1 - k;
            double* ck = wmT + k * lda;
            // pivot: first largest magnitude among a(k..k+lm, k); every lane finds it
            int l = m;
            for (int i = 1; i <= lm; i++)
                if (fabs(ck[m + i]) > fabs(ck[l])) l = m + i;
            const double akl = ck[l];
            WV_SYNC();
            if (WV_LANE == 0) ipvt[k] = l + k - m;
            if (akl == 0.0) {
                info = k + 1;
                continue;
            }
            if (l != m && WV_LANE == 0) {
                ck[l] = ck[m];
                ck[m] = akl;
            }
            WV_SYNC();
            const double t = ls_div(-1.0, akl);
            for (int i = 1 + WV_LANE; i <= lm; i += WV_NL) ck[m + i] = ck[m + i] * t;
            const int jn = mu + (l + k - m);
            ju = ju > jn ? ju : jn;
            if (ju > n - 1) ju = n - 1;
            // the interchange in the columns to the right, one lane per column ...
            for (int j = k + 1 + WV_LANE; j <= ju; j += WV_NL) {
                const int lj = l - (j - k), mj = m - (j - k);
                if (lj != mj) {
                    double* cj = wmT + j * lda;
                    const double tj = cj[lj];
                    cj[lj] = cj[mj];
                    cj[mj] = tj;
                }
            }
            WV_SYNC();
            // ... and the rank-one update of the block below the pivot row, a lane per element
            const int nj = ju - k;
            for (int q = WV_LANE; q < nj * lm; q += WV_NL) {
                const int j = k + 1 + q / lm, i = 1 + q % lm;
                double* cj = wmT + j * lda;
                const int mj = m - (j - k);
                cj[mj + i] = cj[mj + i] + cj[mj] * ck[m + i];
            }
            WV_SYNC();
        }
        if (WV_LANE == 0) ipvt[n - 1] = n - 1;
        if (wmT[(n - 1) * lda + m] == 0.0) info = n;
        WV_SYNC();
        return info;
    }

    // dgbsl, job = 0.  Every step reads what it needs before anybody writes, so one barrier per
    // step suffices: forward, lane 0 stores the interchange while lanes 1..lm add the pivot row's
    // multiples (the lane whose row was interchanged takes the swapped-in value); backward, lane 0
    // stores the quotient while lanes 0..lm-1 subtract its multiples from the rows above.
    // (Measured and dropped: the same solve with b in registers and the pivot row's value
    // travelling by shuffle -- one shuffle + one multiply-add per step on the critical path instead
    // of a shared-memory round trip and a barrier -- ran the C4 banded workload at 128 k traj/s
    // against 199 k: with 11 warps per SM the barriers' latency is hidden and the ~30
    // select / shuffle instructions per step, executed by all 32 lanes, are not.  Nor does a
    // sliding window of b in registers with every lane running the same scalar recurrence -- no
    // barrier, no shuffle, bit-identical on the host harness: 259 k against 264 k, strict mode
    // 175 k against 209 k.  This solve is 40 % of the banded kernel's time
    // (profiles/r2_lsodaw_band_v1.txt) and bound by the instructions of its 2 n steps; multiplying
    // by stored reciprocals of the diagonal in the throughput mode changes nothing either
    // (280.2 k / 280.6 k against 280.6 k / 280.7 k, same box).)
    LS_MFN void dgbsl(double* b) const
    {
        const int lda = 2 * ml + mu + 1, m = ml + mu;
        if (ml != 0) {
            for (int k = 0; k < n - 1; k++) {
                const int lm = ml < n - 1 - k ? ml : n - 1 - k;
                const int l = ipvt[k];
                const double t = b[l], bk = b[k];
                const double* ck = wmT + k * lda;
                double upd = 0.0;
                const int i = 1 + WV_LANE;
                if (WV_NL > 1) {
                    if (i <= lm) upd = ((k + i == l) ? bk : b[k + i]) + t * ck[m + i];
                    WV_SYNC();
                    if (WV_LANE == 0) {
                        if (l != k && l > k + lm) b[l] = bk;  // (cannot happen: l <= k + lm)
                        b[k] = t;
                    }
                    if (i <= lm) b[k + i] = upd;
                    for (int i2 = i + WV_NL; i2 <= lm; i2 += WV_NL)  // ml > 32
                        b[k + i2] = ((k + i2 == l) ? bk : b[k + i2]) + t * ck[m + i2];
                } else {
                    if (l != k) {
                        b[l] = bk;
                        b[k] = t;
                    }
                    for (int i2 = 1; i2 <= lm; i2++) b[k + i2] = b[k + i2] + t * ck[m + i2];
                }
                WV_SYNC();
            }
        }
        for (int k = n - 1; k >= 0; k--) {
            const double* ck = wmT + k * lda;
            const double bk = ls_div(b[k], ck[m]);
            const int lm = k < m ? k : m;
            const int la = m - lm, lb = k - lm;
            const double t = -bk;
            if (WV_NL > 1) {
                double upd = 0.0;
                if (WV_LANE < lm) upd = b[lb + WV_LANE] + t * ck[la + WV_LANE];
                WV_SYNC();
                if (WV_LANE == 0) b[k] = bk;
                if (WV_LANE < lm) b[lb + WV_LANE] = upd;
                for (int i = WV_LANE + WV_NL; i < lm; i += WV_NL) b[lb + i] = b[lb + i] + t * ck[la + i];
            } else {
                b[k] = bk;
                for (int i = 0; i < lm; i++) b[lb + i] = b[lb + i] + t * ck[la + i];
            }
            WV_SYNC();
        }
    }

    // DPRJA, miter = 5, with DBNORM
    LS_MFN int prja_band(double tn, double h, double hl0, double* pp, double& pdnorm)
    {
        const int lda = 2 * ml + mu + 1, mband = ml + mu + 1;
        const int mba = mband < n ? mband : n;
        // column groups evaluated concurrently, one per lane (the host build has one lane)
        const int ncopy = B200ODE_LSODAW_BAND_COPIES(n, ml, mu) < WV_NL ? B200ODE_LSODAW_BAND_COPIES(n, ml, mu) : WV_NL;
        const double fac0 = norm_ptr(savf);
        double r0 = 1000.0 * fabs(h) * LS_ETA * ((double)n) * fac0;
        if (r0 == 0.0) r0 = 1.0;
        for (int j0 = 0; j0 < mba; j0 += ncopy) {
            const int j = j0 + WV_LANE;
            if (WV_LANE < ncopy && j < mba) {
                // this lane's group of columns j, j + mband, ...: one perturbed copy of y, one
                // right-hand-side evaluation
                double* u = ucopy + (2 * WV_LANE) * ld;
                double* fu = u + ld;
                for (int i = 0; i < n; i++) u[i] = y[i];
                for (int i = j; i < n; i += mband) {
                    const double r = ls_max(LS_SQRTETA * fabs(y[i]), ls_div(r0, ewt[i]));
                    u[i] = y[i] + r;
                }
                wv_user_rhs(tn, u, fu, pp);
                for (int jj = j; jj < n; jj += mband) {
                    const double r = ls_max(LS_SQRTETA * fabs(y[jj]), ls_div(r0, ewt[jj]));
                    const double fac = ls_div(-hl0, r);
                    const int i1 = jj - mu > 0 ? jj - mu : 0;
                    const int i2 = jj + ml < n - 1 ? jj + ml : n - 1;
                    for (int i = i1; i <= i2; i++) WABD(i, jj) = (fu[i] - savf[i]) * fac;
                }
            }
            WV_SYNC();
        }
        double an = 0.0;
        WV_FOR_OWN(i) {
            double sum = 0.0;
            const int i1 = i - ml > 0 ? i - ml : 0;
            const int i2 = i + mu < n - 1 ? i + mu : n - 1;
            for (int j = i1; j <= i2; j++) sum += ls_div(fabs(WABD(i, j)), ewt[j]);
            an = ls_max(an, sum * ewt[i]);
        }
        an = wv_max(an);
        pdnorm = ls_div(an, fabs(hl0));
        WV_SYNC();
        WV_FOR_OWN(i) WABD(i, i) = WABD(i, i) + 1.0;
        WV_SYNC();
        return dgbfa() != 0 ? 1 : 0;
    }
#undef WABD

    // prja with the matrix in tensor memory: the perturbed right-hand sides land in shared memory
    // (fcopy), and each lane then forms eight columns' entries of its own rows, adds them to its
    // rows' fnorm sums (ascending j, as the reference), adds the identity and stores them.
    LS_MFN int prja_tm(double tn, double h, double hl0, double* pp, double& pdnorm)
    {
        const double fac0 = norm_ptr(savf);
        double r0 = 1000.0 * fabs(h) * LS_ETA * ((double)n) * fac0;
        if (r0 == 0.0) r0 = 1.0;
        const int nblk = (n + 31) >> 5;
        double sum[2] = {0.0, 0.0};
        for (int j0 = 0; j0 < n; j0 += WV_JCOLS) {
            const int j = j0 + WV_LANE;
            double fac = 0.0;
            if (WV_LANE < WV_JCOLS && j < n) {
                double* u = ucopy + WV_LANE * ld;
                for (int i = 0; i < n; i++) u[i] = y[i];
                const double yj = y[j];
                const double r = ls_max(LS_SQRTETA * fabs(yj), ls_div(r0, ewt[j]));
                u[j] = yj + r;
                fac = ls_div(-hl0, r);
                wv_user_rhs(tn, u, fcopy + WV_LANE * ld, pp);
            }
            WV_SYNC();
            for (int g = 0; g < WV_JCOLS && j0 + g < n; g += 8) {
                double fe[8];
#pragma unroll
                for (int e = 0; e < 8; e++) fe[e] = wv_bcast(fac, (g + e) & 31);
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    if (q < nblk) {
                        const int i = WV_LANE + 32 * q;
                        const double si = (i < n) ? savf[i] : 0.0;
                        double v[8];
#pragma unroll
                        for (int e = 0; e < 8; e++) {
                            const int jj = j0 + g + e;
                            v[e] = 0.0;
                            if (i < n && jj < n) {
                                v[e] = (fcopy[(g + e) * ld + i] - si) * fe[e];
                                sum[q] += ls_div(fabs(v[e]), ewt[jj]);
                                if (i == jj) v[e] = v[e] + 1.0;
                            }
                        }
                        wv_tm_st8(tma(q, j0 + g), v);
                    }
                }
            }
            WV_SYNC();
        }
        wv_tm_wait_st();
        double an = 0.0;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int i = WV_LANE + 32 * q;
            if (i < n) an = ls_max(an, sum[q] * ewt[i]);
        }
        an = wv_max(an);
        pdnorm = ls_div(an, fabs(hl0));
        return dgefa() != 0 ? 1 : 0;
    }

    LS_MFN int prja(double tn, double h, double hl0, double* pp, double& pdnorm)
    {
        if constexpr (TM) return prja_tm(tn, h, hl0, pp, pdnorm);
        if (banded()) return prja_band(tn, h, hl0, pp, pdnorm);
        const double fac0 = norm_ptr(savf);
        double r0 = 1000.0 * fabs(h) * LS_ETA * ((double)n) * fac0;
        if (r0 == 0.0) r0 = 1.0;
        for (int j0 = 0; j0 < n; j0 += WV_JCOLS) {
            const int j = j0 + WV_LANE;
            if (WV_LANE < WV_JCOLS && j < n) {
                double* u = ucopy + WV_LANE * ld;
                for (int i = 0; i < n; i++) u[i] = y[i];
                const double yj = y[j];
                const double r = ls_max(LS_SQRTETA * fabs(yj), ls_div(r0, ewt[j]));
                u[j] = yj + r;
                const double fac = ls_div(-hl0, r);
                double* col = wmT + j * ld;
                wv_user_rhs(tn, u, col, pp);
                for (int i = 0; i < n; i++) col[i] = (col[i] - savf[i]) * fac;
            }
            WV_SYNC();
        }
        double an = 0.0;
        WV_FOR_OWN(i) {
            double sum = 0.0;
#pragma unroll 4
            for (int j = 0; j < n; j++) sum += ls_div(fabs(WMT(i, j)), ewt[j]);
            an = ls_max(an, sum * ewt[i]);
        }
        an = wv_max(an);
        pdnorm = ls_div(an, fabs(hl0));
        WV_FOR_OWN(i) WMT(i, i) = WMT(i, i) + 1.0;
        WV_SYNC();
        return dgefa() != 0 ? 1 : 0;
    }
};
typedef LsWarpVecT<false, -1> LsWarpVec;

#endif
