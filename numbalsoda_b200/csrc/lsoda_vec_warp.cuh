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
#define WV_FOR_OWN(i) for (int i = WV_LANE; i < n; i += WV_NL)
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

struct LsWarpVec {
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
        if (matrix) {
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
    LS_MFN bool banded() const { return ml >= 0; }
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
    LS_MFN int dgefa()
    {
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
    LS_MFN void dgesl(double* b) const
    {
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
    // select / shuffle instructions per step, executed by all 32 lanes, are not.)
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

    LS_MFN int prja(double tn, double h, double hl0, double* pp, double& pdnorm)
    {
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

#endif
