/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * CPU oracle for the LSODA half of the hot path: a plain-C restatement of what
 * the reference computes when Python calls numbalsoda.lsoda():
 *
 *   lsoda_wrapper          /root/reference/src/wrapper.cpp:10-57
 *   LSODA::lsoda_update    /root/reference/src/LSODA.cpp:2247-2277
 *   LSODA::lsoda           /root/reference/src/LSODA.cpp:287-993
 *   LSODA::stoda           /root/reference/src/LSODA.cpp:995-1366
 *   correction/corfailure  /root/reference/src/LSODA.cpp:1743-1922
 *   prja, dgefa, dgesl     /root/reference/src/LSODA.cpp:1633-1696, :173-231, :110-170
 *   methodswitch           /root/reference/src/LSODA.cpp:1946-2069
 *   orderswitch, scaleh    /root/reference/src/LSODA.cpp:2097-2209, :1594-1631
 *   cfode, resetcoeff      /root/reference/src/LSODA.cpp:1463-1592, :2211-2225
 *   intdy, ewset, norms    /root/reference/src/LSODA.cpp:1414-1461, :1368-1390, :1705-1736
 *
 * It is organised differently from the reference: one flat state record with
 * 0-based component storage, and a driver that steps freely and, after every
 * accepted step, interpolates every output time the step has passed (the
 * reference instead re-enters lsoda() once per output time; SURVEY.md section 7
 * step 1d shows the two are arithmetically identical, and
 * tests/test_oracle_lsoda.py checks this file bit-for-bit, including the step /
 * RHS / Jacobian / method-switch counters, against the reference's own C++
 * compiled by oracle/Makefile into oracle/_ref/).
 *
 * Build with -ffp-contract=off: the reference binaries are x86-64 SSE2 code
 * without fused multiply-adds.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

#define ETA 2.2204460492503131e-16 /* LSODA.cpp:33 */
#define MXORDN 12                  /* LSODA.cpp:38 */
#define MXORDS 5
#define MAXCOR 3 /* LSODA.cpp:564 */
#define MSBP 20  /* LSODA.cpp:565 */
#define MXHNIL 10 /* LSODA.cpp:294 */
#define CCMAX 0.3 /* LSODA.cpp:563 */

static const double SM1[13] = {0., 0.5, 0.575, 0.55, 0.45, 0.35, 0.25,
                               0.2, 0.15, 0.1, 0.075, 0.05, 0.025}; /* LSODA.cpp:39-43 */

typedef struct {
    int n;
    oracle_rhs_fn f;
    void *data;
    double rtol, atol;
    /* Nordsieck array yh[j][i], j = 1..13 used (order index 1-based as in the paper) */
    double *yh[MXORDN + 3];
    double *wm; /* n x n, row-major: wm[i*n+j] = d f_i / d y_j */
    double *ewt, *savf, *acor, *y;
    int *ipvt;
    double el[MXORDN + 2], elco[MXORDN + 1][MXORDN + 2], tesco[MXORDN + 1][4];
    double cm1[MXORDN + 1], cm2[MXORDS + 1];
    int nq, l, lmax, maxord, meth, mused, miter, ialth, ipup, nslp, icount, irflag;
    int kflag, jstart, ierpj, jcur;
    int nst, nfe, nje, nqu, nattempt;
    double h, hu, hold, tn, rc, el0, crate, rmax, pdest, pdlast, pdnorm, conit;
    double sqrteta;
    /* banded iteration matrix (jt = 5): ml / mu as the reference's checker takes them
     * (LSODA.cpp:369-392, iworks[0..1]); banded < 0 means the dense path (jt = 2) */
    int banded, ml, mu;
    double *abd; /* LINPACK band storage, (2 ml + mu + 1) x n, column-major */
} lsoda_ctx;

/* std::max / std::min semantics (they matter when an operand is NaN) */
static inline double dmax(double a, double b) { return (a < b) ? b : a; }
static inline double dmin(double a, double b) { return (b < a) ? b : a; }

/* vmnorm, LSODA.cpp:1705-1712 */
static double wmax_norm(int n, const double *v, const double *w)
{
    double vm = 0.;
    for (int i = 0; i < n; i++)
        vm = dmax(vm, fabs(v[i]) * w[i]);
    return vm;
}

/* cfode, LSODA.cpp:1463-1592 */
static void method_coefficients(lsoda_ctx *s, int meth)
{
    double pc[13], agamq, fnq, fnqm1, pint, ragq, rqfac, rq1fac, tsign, xpin;
    int i, nq, nqm1, nqp1;
    if (meth == 1) {
        s->elco[1][1] = 1.;
        s->elco[1][2] = 1.;
        s->tesco[1][1] = 0.;
        s->tesco[1][2] = 2.;
        s->tesco[2][1] = 1.;
        s->tesco[12][3] = 0.;
        pc[1] = 1.;
        rqfac = 1.;
        for (nq = 2; nq <= 12; nq++) {
            rq1fac = rqfac;
            rqfac = rqfac / (double)nq;
            nqm1 = nq - 1;
            fnqm1 = (double)nqm1;
            nqp1 = nq + 1;
            pc[nq] = 0.;
            for (i = nq; i >= 2; i--)
                pc[i] = pc[i - 1] + fnqm1 * pc[i];
            pc[1] = fnqm1 * pc[1];
            pint = pc[1];
            xpin = pc[1] / 2.;
            tsign = 1.;
            for (i = 2; i <= nq; i++) {
                tsign = -tsign;
                pint += tsign * pc[i] / (double)i;
                xpin += tsign * pc[i] / (double)(i + 1);
            }
            s->elco[nq][1] = pint * rq1fac;
            s->elco[nq][2] = 1.;
            for (i = 2; i <= nq; i++)
                s->elco[nq][i + 1] = rq1fac * pc[i] / (double)i;
            agamq = rqfac * xpin;
            ragq = 1. / agamq;
            s->tesco[nq][2] = ragq;
            if (nq < 12)
                s->tesco[nqp1][1] = ragq * rqfac / (double)nqp1;
            s->tesco[nqm1][3] = ragq;
        }
        return;
    }
    pc[1] = 1.;
    rq1fac = 1.;
    for (nq = 1; nq <= 5; nq++) {
        fnq = (double)nq;
        nqp1 = nq + 1;
        pc[nqp1] = 0.;
        for (i = nq + 1; i >= 2; i--)
            pc[i] = pc[i - 1] + fnq * pc[i];
        pc[1] *= fnq;
        for (i = 1; i <= nqp1; i++)
            s->elco[nq][i] = pc[i] / pc[2];
        s->elco[nq][2] = 1.;
        s->tesco[nq][1] = rq1fac;
        s->tesco[nq][2] = ((double)nqp1) / s->elco[nq][1];
        s->tesco[nq][3] = ((double)(nq + 2)) / s->elco[nq][1];
        rq1fac /= fnq;
    }
}

/* resetcoeff, LSODA.cpp:2211-2225 */
static void load_order_coefficients(lsoda_ctx *s)
{
    for (int i = 1; i <= s->l; i++)
        s->el[i] = s->elco[s->nq][i];
    s->rc = s->rc * s->el[1] / s->el0;
    s->el0 = s->el[1];
    s->conit = 0.5 / (double)(s->nq + 2);
}

/* scaleh, LSODA.cpp:1594-1631 (hmxi == 0, so the hmax clamp is a division by 1) */
static void rescale_step(lsoda_ctx *s, double *rh, double *pdh)
{
    double r;
    *rh = dmin(*rh, s->rmax);
    *rh = *rh / dmax(1., fabs(s->h) * 0. * *rh);
    if (s->meth == 1) {
        s->irflag = 0;
        *pdh = dmax(fabs(s->h) * s->pdlast, 0.000001);
        if ((*rh * *pdh * 1.00001) >= SM1[s->nq]) {
            *rh = SM1[s->nq] / *pdh;
            s->irflag = 1;
        }
    }
    r = 1.;
    for (int j = 2; j <= s->l; j++) {
        r *= *rh;
        for (int i = 0; i < s->n; i++)
            s->yh[j][i] *= r;
    }
    s->h *= *rh;
    s->rc *= *rh;
    s->ialth = s->l;
}

/* The Pascal-triangle sweeps: forward = prediction (LSODA.cpp:1141-1144),
 * backward = retraction (LSODA.cpp:1290-1295, :1909-1912). */
static void pascal_sweep(lsoda_ctx *s, int forward)
{
    for (int j = s->nq; j >= 1; j--)
        for (int i1 = j; i1 <= s->nq; i1++)
            for (int i = 0; i < s->n; i++) {
                if (forward)
                    s->yh[i1][i] += s->yh[i1 + 1][i];
                else
                    s->yh[i1][i] -= s->yh[i1 + 1][i];
            }
}

/* idamax1 with its integer truncation, LSODA.cpp:59-77: the "largest" entry is the
 * first one whose magnitude, converted to an unsigned integer, is largest. */
static int trunc_argmax(const double *row, int from, int to)
{
    unsigned long long v, vmax = 0;
    int idx = from;
    for (int i = from; i <= to; i++) {
        v = (unsigned long long)fabs(row[i]);
        if (v > vmax) {
            vmax = v;
            idx = i;
        }
    }
    return idx;
}

/* dgefa, LSODA.cpp:173-231: row-oriented elimination, pivot searched along row k,
 * columns swapped.  Returns nonzero when a zero pivot was met. */
static int lu_factor(lsoda_ctx *s)
{
    const int n = s->n;
    double *a = s->wm, t;
    int info = 0, i, j, k, c;
    for (k = 0; k < n - 1; k++) {
        j = trunc_argmax(a + (size_t)k * n, k, n - 1);
        s->ipvt[k] = j;
        if (a[k * n + j] == 0.) {
            info = k + 1;
            continue;
        }
        if (j != k) {
            t = a[k * n + j];
            a[k * n + j] = a[k * n + k];
            a[k * n + k] = t;
        }
        t = -1. / a[k * n + k];
        for (c = k + 1; c < n; c++)
            a[k * n + c] = t * a[k * n + c];
        for (i = k + 1; i < n; i++) {
            t = a[i * n + j];
            if (j != k) {
                a[i * n + j] = a[i * n + k];
                a[i * n + k] = t;
            }
            for (c = k + 1; c < n; c++)
                a[i * n + c] = t * a[k * n + c] + a[i * n + c];
        }
    }
    s->ipvt[n - 1] = n - 1;
    if (a[(n - 1) * n + (n - 1)] == 0.)
        info = n;
    return info;
}

/* dgesl with job = 0, LSODA.cpp:119-144 */
static void lu_solve(lsoda_ctx *s, double *b)
{
    const int n = s->n;
    const double *a = s->wm;
    double t;
    int j, k, c;
    for (k = 0; k < n; k++) {
        t = 0.0;
        for (c = 0; c < k; c++)
            t += a[k * n + c] * b[c];
        b[k] = (b[k] - t) / a[k * n + k];
    }
    for (k = n - 2; k >= 0; k--) {
        t = 0.0;
        for (c = k + 1; c < n; c++)
            t += a[k * n + c] * b[c];
        b[k] = b[k] + t;
        j = s->ipvt[k];
        if (j != k) {
            t = b[j];
            b[j] = b[k];
            b[k] = t;
        }
    }
}

/* ---- jt = 5: banded, internally generated Jacobian ---------------------------------
 * The reference accepts jt = 4 / 5 with ml, mu (LSODA.cpp:369-392) but its prja and solsy
 * implement miter = 2 only (:1656-1660, :1936-1942: "not implemented" and exit).  What follows
 * is therefore an EXTENSION, restating the routines the reference was translated from --
 * ODEPACK's DPRJA / DSOLSY for miter = 5 with DBNORM, and LINPACK's dgbfa / dgbsl -- in this
 * file's conventions (inverted ewt, 0-based).  Parity for it: the dense path on the same
 * problem within the tolerance (tests/test_banded.py), not bit identity with the reference. */
#define ABD(s, i, j) ((s)->abd[(size_t)(j) * (2 * (s)->ml + (s)->mu + 1) + ((i) - (j) + (s)->ml + (s)->mu)])

/* dgbfa: LU of the band matrix with partial pivoting (row interchanges inside the band) */
static int band_factor(lsoda_ctx *s)
{
    const int n = s->n, ml = s->ml, mu = s->mu;
    const int m = ml + mu; /* row of the diagonal in a column of abd */
    const int lda = 2 * ml + mu + 1;
    double *a = s->abd, t;
    int info = 0, i, j, k, l, lm, ju = 0, jz, j0, j1, mm;
    /* zero the initial fill-in columns */
    j0 = mu + 1;
    j1 = (n < m + 1 ? n : m + 1) - 1;
    for (jz = j0; jz < j1; jz++)
        for (i = m + 1 - (jz + 1); i < ml; i++)
            a[(size_t)jz * lda + i] = 0.;
    jz = j1 - 1;
    for (k = 0; k < n - 1; k++) {
        /* zero the next fill-in column */
        jz++;
        if (jz < n)
            for (i = 0; i < ml; i++)
                a[(size_t)jz * lda + i] = 0.;
        /* pivot: first largest |a(i,k)|, i = k .. k + lm */
        lm = ml < n - 1 - k ? ml : n - 1 - k;
        l = m;
        for (i = 1; i <= lm; i++)
            if (fabs(a[(size_t)k * lda + m + i]) > fabs(a[(size_t)k * lda + l]))
                l = m + i;
        s->ipvt[k] = l + k - m;
        if (a[(size_t)k * lda + l] == 0.) {
            info = k + 1;
            continue;
        }
        if (l != m) {
            t = a[(size_t)k * lda + l];
            a[(size_t)k * lda + l] = a[(size_t)k * lda + m];
            a[(size_t)k * lda + m] = t;
        }
        t = -1. / a[(size_t)k * lda + m];
        for (i = 1; i <= lm; i++)
            a[(size_t)k * lda + m + i] *= t;
        /* row elimination with column indexing */
        ju = ju > mu + s->ipvt[k] ? ju : mu + s->ipvt[k];
        if (ju > n - 1)
            ju = n - 1;
        mm = m;
        for (j = k + 1; j <= ju; j++) {
            l--;
            mm--;
            t = a[(size_t)j * lda + l];
            if (l != mm) {
                a[(size_t)j * lda + l] = a[(size_t)j * lda + mm];
                a[(size_t)j * lda + mm] = t;
            }
            for (i = 1; i <= lm; i++)
                a[(size_t)j * lda + mm + i] += t * a[(size_t)k * lda + m + i];
        }
    }
    s->ipvt[n - 1] = n - 1;
    if (a[(size_t)(n - 1) * lda + m] == 0.)
        info = n;
    return info;
}

/* dgbsl, job = 0 */
static void band_solve(lsoda_ctx *s, double *b)
{
    const int n = s->n, ml = s->ml, mu = s->mu, m = ml + mu, lda = 2 * ml + mu + 1;
    const double *a = s->abd;
    double t;
    int i, k, l, lm, la, lb;
    if (ml != 0)
        for (k = 0; k < n - 1; k++) {
            lm = ml < n - 1 - k ? ml : n - 1 - k;
            l = s->ipvt[k];
            t = b[l];
            if (l != k) {
                b[l] = b[k];
                b[k] = t;
            }
            for (i = 1; i <= lm; i++)
                b[k + i] += t * a[(size_t)k * lda + m + i];
        }
    for (k = n - 1; k >= 0; k--) {
        b[k] /= a[(size_t)k * lda + m];
        lm = (k < m ? k : m);
        la = m - lm;
        lb = k - lm;
        t = -b[k];
        for (i = 0; i < lm; i++)
            b[lb + i] += t * a[(size_t)k * lda + la + i];
    }
}

/* DPRJA, miter = 5: columns j, j + mband, j + 2 mband, ... are perturbed together, so the
 * Jacobian costs min(ml + mu + 1, n) right-hand-side evaluations instead of n */
static void build_band_iteration_matrix(lsoda_ctx *s)
{
    const int n = s->n, ml = s->ml, mu = s->mu, mband = ml + mu + 1;
    const int mba = mband < n ? mband : n;
    double fac, hl0, r, r0, an, sum;
    int i, j, jj, i1, i2;
    s->nje++;
    s->ierpj = 0;
    s->jcur = 1;
    hl0 = s->h * s->el0;
    fac = wmax_norm(n, s->savf, s->ewt);
    r0 = 1000. * fabs(s->h) * ETA * ((double)n) * fac;
    if (r0 == 0.)
        r0 = 1.;
    for (j = 0; j < mba; j++) {
        for (i = j; i < n; i += mband) {
            r = dmax(s->sqrteta * fabs(s->yh[1][i]), r0 / s->ewt[i]);
            s->y[i] += r;
        }
        s->f(s->tn, s->y, s->acor, s->data);
        for (jj = j; jj < n; jj += mband) {
            s->y[jj] = s->yh[1][jj];
            r = dmax(s->sqrteta * fabs(s->y[jj]), r0 / s->ewt[jj]);
            fac = -hl0 / r;
            i1 = jj - mu > 0 ? jj - mu : 0;
            i2 = jj + ml < n - 1 ? jj + ml : n - 1;
            for (i = i1; i <= i2; i++)
                ABD(s, i, jj) = (s->acor[i] - s->savf[i]) * fac;
        }
    }
    s->nfe += mba;
    /* DBNORM: the weighted max-norm of the band matrix, rows summed over the band */
    an = 0.;
    for (i = 0; i < n; i++) {
        sum = 0.;
        i1 = i - ml > 0 ? i - ml : 0;
        i2 = i + mu < n - 1 ? i + mu : n - 1;
        for (j = i1; j <= i2; j++)
            sum += fabs(ABD(s, i, j)) / s->ewt[j];
        an = dmax(an, sum * s->ewt[i]);
    }
    s->pdnorm = an / fabs(hl0);
    for (i = 0; i < n; i++)
        ABD(s, i, i) += 1.;
    if (band_factor(s) != 0)
        s->ierpj = 1;
}

/* prja, LSODA.cpp:1633-1696 (miter == 2 only: dense finite-difference Jacobian) */
static void build_iteration_matrix(lsoda_ctx *s)
{
    if (s->banded) {
        build_band_iteration_matrix(s);
        return;
    }
    const int n = s->n;
    double fac, hl0, r, r0, yj, an, sum;
    int i, j;
    s->nje++;
    s->ierpj = 0;
    s->jcur = 1;
    hl0 = s->h * s->el0;
    fac = wmax_norm(n, s->savf, s->ewt);
    r0 = 1000. * fabs(s->h) * ETA * ((double)n) * fac;
    if (r0 == 0.)
        r0 = 1.;
    for (j = 0; j < n; j++) {
        yj = s->y[j];
        r = dmax(s->sqrteta * fabs(yj), r0 / s->ewt[j]);
        s->y[j] += r;
        fac = -hl0 / r;
        s->f(s->tn, s->y, s->acor, s->data);
        for (i = 0; i < n; i++)
            s->wm[i * n + j] = (s->acor[i] - s->savf[i]) * fac;
        s->y[j] = yj;
    }
    s->nfe += n;
    /* fnorm, LSODA.cpp:1714-1736 */
    an = 0.;
    for (i = 0; i < n; i++) {
        sum = 0.;
        for (j = 0; j < n; j++)
            sum += fabs(s->wm[i * n + j]) / s->ewt[j];
        an = dmax(an, sum * s->ewt[i]);
    }
    s->pdnorm = an / fabs(hl0);
    for (i = 0; i < n; i++)
        s->wm[i * n + i] += 1.;
    if (lu_factor(s) != 0)
        s->ierpj = 1;
}

/* corfailure, LSODA.cpp:1904-1922.  The reference increments the *pointer* ncf
 * (:1906), so its mxncf limit reads past a local and never fires in practice;
 * with hmin == 0 the only way to corflag 2 is h == 0 exactly. */
static int corrector_failed(lsoda_ctx *s, double told, double *rh)
{
    s->rmax = 2.;
    s->tn = told;
    pascal_sweep(s, 0);
    if (fabs(s->h) <= 0. * 1.00001)
        return 2;
    *rh = 0.25;
    s->ipup = s->miter;
    return 1;
}

/* correction, LSODA.cpp:1743-1902.  Returns corflag. */
static int corrector(lsoda_ctx *s, double pnorm, double *del, double *delp, double told,
                     double *rh, int *m_out)
{
    const int n = s->n;
    double rm, rate = 0.0, dcon;
    int m = 0, i;
    *del = 0.;
    for (i = 0; i < n; i++)
        s->y[i] = s->yh[1][i];
    s->f(s->tn, s->y, s->savf, s->data);
    s->nfe++;
    for (;;) {
        if (m == 0) {
            if (s->ipup > 0) {
                build_iteration_matrix(s);
                s->ipup = 0;
                s->rc = 1.;
                s->nslp = s->nst;
                s->crate = 0.7;
                if (s->ierpj != 0) {
                    *m_out = m;
                    return corrector_failed(s, told, rh);
                }
            }
            for (i = 0; i < n; i++)
                s->acor[i] = 0.;
        }
        if (s->miter == 0) {
            for (i = 0; i < n; i++) {
                s->savf[i] = s->h * s->savf[i] - s->yh[2][i];
                s->y[i] = s->savf[i] - s->acor[i];
            }
            *del = wmax_norm(n, s->y, s->ewt);
            for (i = 0; i < n; i++) {
                s->y[i] = s->yh[1][i] + s->el[1] * s->savf[i];
                s->acor[i] = s->savf[i];
            }
        } else {
            for (i = 0; i < n; i++)
                s->y[i] = s->h * s->savf[i] - (s->yh[2][i] + s->acor[i]);
            if (s->banded)
                band_solve(s, s->y);
            else
                lu_solve(s, s->y);
            *del = wmax_norm(n, s->y, s->ewt);
            for (i = 0; i < n; i++) {
                s->acor[i] += s->y[i];
                s->y[i] = s->yh[1][i] + s->el[1] * s->acor[i];
            }
        }
        if (*del <= 100. * pnorm * ETA)
            break;
        if (m != 0 || s->meth != 1) {
            if (m != 0) {
                rm = 1024.0;
                if (*del <= (1024. * *delp))
                    rm = *del / *delp;
                rate = dmax(rate, rm);
                s->crate = dmax(0.2 * s->crate, rm);
            }
            dcon = *del * dmin(1., 1.5 * s->crate) / (s->tesco[s->nq][2] * s->conit);
            if (dcon <= 1.) {
                s->pdest = dmax(s->pdest, rate / fabs(s->h * s->el[1]));
                if (s->pdest != 0.)
                    s->pdlast = s->pdest;
                break;
            }
        }
        m++;
        if (m == MAXCOR || (m >= 2 && *del > 2. * *delp)) {
            if (s->miter == 0 || s->jcur == 1) {
                *m_out = m;
                return corrector_failed(s, told, rh);
            }
            s->ipup = s->miter;
            m = 0;
            rate = 0.;
            *del = 0.;
            for (i = 0; i < n; i++)
                s->y[i] = s->yh[1][i];
            s->f(s->tn, s->y, s->savf, s->data);
            s->nfe++;
        } else {
            *delp = *del;
            s->f(s->tn, s->y, s->savf, s->data);
            s->nfe++;
        }
    }
    *m_out = m;
    return 0;
}

/* methodswitch, LSODA.cpp:1946-2069 */
static void consider_method_switch(lsoda_ctx *s, double dsm, double pnorm, double *pdh,
                                   double *rh)
{
    double rh1, rh2, rh1it, exm2, dm2, exm1, dm1, alpha, exsm;
    int lm1, lm2, nqm1, nqm2;
    if (s->meth == 1) {
        if (s->nq > 5)
            return;
        if (dsm <= (100. * pnorm * ETA) || s->pdest == 0.) {
            if (s->irflag == 0)
                return;
            rh2 = 2.;
            nqm2 = s->nq < MXORDS ? s->nq : MXORDS;
        } else {
            exsm = 1. / (double)s->l;
            rh1 = 1. / (1.2 * pow(dsm, exsm) + 0.0000012);
            rh1it = 2. * rh1;
            *pdh = s->pdlast * fabs(s->h);
            if ((*pdh * rh1) > 0.00001)
                rh1it = SM1[s->nq] / *pdh;
            rh1 = dmin(rh1, rh1it);
            if (s->nq > MXORDS) {
                nqm2 = MXORDS;
                lm2 = MXORDS + 1;
                exm2 = 1. / (double)lm2;
                dm2 = wmax_norm(s->n, s->yh[lm2 + 1], s->ewt) / s->cm2[MXORDS];
                rh2 = 1. / (1.2 * pow(dm2, exm2) + 0.0000012);
            } else {
                dm2 = dsm * (s->cm1[s->nq] / s->cm2[s->nq]);
                rh2 = 1. / (1.2 * pow(dm2, exsm) + 0.0000012);
                nqm2 = s->nq;
            }
            if (rh2 < 5. * rh1)
                return;
        }
        *rh = rh2;
        s->icount = 20;
        s->meth = 2;
        s->miter = 2;
        s->pdlast = 0.;
        s->nq = nqm2;
        s->l = s->nq + 1;
        return;
    }
    exsm = 1. / (double)s->l;
    if (MXORDN < s->nq) {
        nqm1 = MXORDN;
        lm1 = MXORDN + 1;
        exm1 = 1. / (double)lm1;
        dm1 = wmax_norm(s->n, s->yh[lm1 + 1], s->ewt) / s->cm1[MXORDN];
        rh1 = 1. / (1.2 * pow(dm1, exm1) + 0.0000012);
    } else {
        dm1 = dsm * (s->cm2[s->nq] / s->cm1[s->nq]);
        rh1 = 1. / (1.2 * pow(dm1, exsm) + 0.0000012);
        nqm1 = s->nq;
        exm1 = exsm;
    }
    rh1it = 2. * rh1;
    *pdh = s->pdnorm * fabs(s->h);
    if ((*pdh * rh1) > 0.00001)
        rh1it = SM1[nqm1] / *pdh;
    rh1 = dmin(rh1, rh1it);
    rh2 = 1. / (1.2 * pow(dsm, exsm) + 0.0000012);
    if ((rh1 * 5.) < (5. * rh2))
        return;
    alpha = dmax(0.001, rh1);
    dm1 *= pow(alpha, exm1);
    if (dm1 <= 1000. * ETA * pnorm)
        return;
    *rh = rh1;
    s->icount = 20;
    s->meth = 1;
    s->miter = 0;
    s->pdlast = 0.;
    s->nq = nqm1;
    s->l = s->nq + 1;
}

/* orderswitch, LSODA.cpp:2097-2209.  Returns orderflag. */
static int choose_order(lsoda_ctx *s, double *rhup, double dsm, double *pdh, double *rh)
{
    int newq;
    double exsm, rhdn, rhsm, ddn, exdn, r;
    exsm = 1. / (double)s->l;
    rhsm = 1. / (1.2 * pow(dsm, exsm) + 0.0000012);
    rhdn = 0.;
    if (s->nq != 1) {
        ddn = wmax_norm(s->n, s->yh[s->l], s->ewt) / s->tesco[s->nq][1];
        exdn = 1. / (double)s->nq;
        rhdn = 1. / (1.3 * pow(ddn, exdn) + 0.0000013);
    }
    if (s->meth == 1) {
        *pdh = dmax(fabs(s->h) * s->pdlast, 0.000001);
        if (s->l < s->lmax)
            *rhup = dmin(*rhup, SM1[s->l] / *pdh);
        rhsm = dmin(rhsm, SM1[s->nq] / *pdh);
        if (s->nq > 1)
            rhdn = dmin(rhdn, SM1[s->nq - 1] / *pdh);
        s->pdest = 0.;
    }
    if (rhsm >= *rhup) {
        if (rhsm >= rhdn) {
            newq = s->nq;
            *rh = rhsm;
        } else {
            newq = s->nq - 1;
            *rh = rhdn;
            if (s->kflag < 0 && *rh > 1.)
                *rh = 1.;
        }
    } else {
        if (*rhup <= rhdn) {
            newq = s->nq - 1;
            *rh = rhdn;
            if (s->kflag < 0 && *rh > 1.)
                *rh = 1.;
        } else {
            *rh = *rhup;
            if (*rh >= 1.1) {
                r = s->el[s->l] / (double)s->l;
                s->nq = s->l;
                s->l = s->nq + 1;
                for (int i = 0; i < s->n; i++)
                    s->yh[s->l][i] = s->acor[i] * r;
                return 2;
            }
            s->ialth = 3;
            return 0;
        }
    }
    if (s->meth == 1) {
        if ((*rh * *pdh * 1.00001) < SM1[newq])
            if (s->kflag == 0 && *rh < 1.1) {
                s->ialth = 3;
                return 0;
            }
    } else {
        if (s->kflag == 0 && *rh < 1.1) {
            s->ialth = 3;
            return 0;
        }
    }
    if (s->kflag <= -2)
        *rh = dmin(*rh, 0.2);
    if (newq == s->nq)
        return 1;
    s->nq = newq;
    s->l = s->nq + 1;
    return 2;
}

/* endstoda, LSODA.cpp:2075-2082 */
static void finish_step(lsoda_ctx *s)
{
    double r = 1. / s->tesco[s->nqu][2];
    for (int i = 0; i < s->n; i++)
        s->acor[i] *= r;
    s->hold = s->h;
    s->jstart = 1;
}

/* stoda, LSODA.cpp:995-1366: one accepted step (or a failure flagged in kflag). */
static void one_step(lsoda_ctx *s)
{
    const int n = s->n;
    int corflag, orderflag, m = 0, i, j;
    double del = 0.0, delp = 0.0, dsm = 0.0, dup, exup, r, rh = 0.0, rhup, told;
    double pdh = 0.0, pnorm;

    s->kflag = 0;
    told = s->tn;
    s->ierpj = 0;
    s->jcur = 0;
    if (s->jstart == 0) {
        s->lmax = s->maxord + 1;
        s->nq = 1;
        s->l = 2;
        s->ialth = 2;
        s->rmax = 10000.;
        s->rc = 0.;
        s->el0 = 1.;
        s->crate = 0.7;
        s->hold = s->h;
        s->nslp = 0;
        s->ipup = s->miter;
        s->icount = 20;
        s->irflag = 0;
        s->pdest = 0.;
        s->pdlast = 0.;
        method_coefficients(s, 2);
        for (i = 1; i <= 5; i++)
            s->cm2[i] = s->tesco[i][2] * s->elco[i][i + 1];
        method_coefficients(s, 1);
        for (i = 1; i <= 12; i++)
            s->cm1[i] = s->tesco[i][2] * s->elco[i][i + 1];
        load_order_coefficients(s);
    }
    if (s->jstart == -1) {
        s->ipup = s->miter;
        s->lmax = s->maxord + 1;
        if (s->ialth == 1)
            s->ialth = 2;
        if (s->meth != s->mused) {
            method_coefficients(s, s->meth);
            s->ialth = s->l;
            load_order_coefficients(s);
        }
        if (s->h != s->hold) {
            rh = s->h / s->hold;
            s->h = s->hold;
            rescale_step(s, &rh, &pdh);
        }
    }
    /* jstart == -2 never occurs on this path (itask == 1, LSODA.cpp:2260). */

    for (;;) {
        for (;;) {
            if (fabs(s->rc - 1.) > CCMAX)
                s->ipup = s->miter;
            if (s->nst >= s->nslp + MSBP)
                s->ipup = s->miter;
            s->tn += s->h;
            pascal_sweep(s, 1);
            pnorm = wmax_norm(n, s->yh[1], s->ewt);
            s->nattempt++;
            corflag = corrector(s, pnorm, &del, &delp, told, &rh, &m);
            if (corflag == 0)
                break;
            if (corflag == 1) {
                rh = dmax(rh, 0. / fabs(s->h));
                rescale_step(s, &rh, &pdh);
                continue;
            }
            s->kflag = -2;
            s->hold = s->h;
            s->jstart = 1;
            return;
        }
        s->jcur = 0;
        if (m == 0)
            dsm = del / s->tesco[s->nq][2];
        if (m > 0)
            dsm = wmax_norm(n, s->acor, s->ewt) / s->tesco[s->nq][2];

        if (dsm <= 1.) {
            s->kflag = 0;
            s->nst++;
            s->hu = s->h;
            s->nqu = s->nq;
            s->mused = s->meth;
            for (j = 1; j <= s->l; j++) {
                r = s->el[j];
                for (i = 0; i < n; i++)
                    s->yh[j][i] += r * s->acor[i];
            }
            s->icount--;
            if (s->icount < 0) {
                consider_method_switch(s, dsm, pnorm, &pdh, &rh);
                if (s->meth != s->mused) {
                    rh = dmax(rh, 0. / fabs(s->h));
                    rescale_step(s, &rh, &pdh);
                    s->rmax = 10.;
                    finish_step(s);
                    return;
                }
            }
            s->ialth--;
            if (s->ialth == 0) {
                rhup = 0.;
                if (s->l != s->lmax) {
                    for (i = 0; i < n; i++)
                        s->savf[i] = s->acor[i] - s->yh[s->lmax][i];
                    dup = wmax_norm(n, s->savf, s->ewt) / s->tesco[s->nq][3];
                    exup = 1. / (double)(s->l + 1);
                    rhup = 1. / (1.4 * pow(dup, exup) + 0.0000014);
                }
                orderflag = choose_order(s, &rhup, dsm, &pdh, &rh);
                if (orderflag == 0) {
                    finish_step(s);
                    return;
                }
                if (orderflag == 2)
                    load_order_coefficients(s);
                rh = dmax(rh, 0. / fabs(s->h));
                rescale_step(s, &rh, &pdh);
                s->rmax = 10.;
                finish_step(s);
                return;
            }
            if (s->ialth > 1 || s->l == s->lmax) {
                finish_step(s);
                return;
            }
            for (i = 0; i < n; i++)
                s->yh[s->lmax][i] = s->acor[i];
            finish_step(s);
            return;
        }
        /* error test failed, LSODA.cpp:1286-1363 */
        s->kflag--;
        s->tn = told;
        pascal_sweep(s, 0);
        s->rmax = 2.;
        if (fabs(s->h) <= 0. * 1.00001) {
            s->kflag = -1;
            s->hold = s->h;
            s->jstart = 1;
            return;
        }
        if (s->kflag > -3) {
            rhup = 0.;
            orderflag = choose_order(s, &rhup, dsm, &pdh, &rh);
            if (orderflag == 1 || orderflag == 0) {
                if (orderflag == 0)
                    rh = dmin(rh, 0.2);
                rh = dmax(rh, 0. / fabs(s->h));
                rescale_step(s, &rh, &pdh);
            }
            if (orderflag == 2) {
                load_order_coefficients(s);
                rh = dmax(rh, 0. / fabs(s->h));
                rescale_step(s, &rh, &pdh);
            }
            continue;
        }
        if (s->kflag == -10) {
            s->kflag = -1;
            s->hold = s->h;
            s->jstart = 1;
            return;
        }
        rh = 0.1;
        rh = dmax(0. / fabs(s->h), rh);
        s->h *= rh;
        for (i = 0; i < n; i++)
            s->y[i] = s->yh[1][i];
        s->f(s->tn, s->y, s->savf, s->data);
        s->nfe++;
        for (i = 0; i < n; i++)
            s->yh[2][i] = s->h * s->savf[i];
        s->ipup = s->miter;
        s->ialth = 5;
        if (s->nq == 1)
            continue;
        s->nq = 1;
        s->l = 2;
        load_order_coefficients(s);
    }
}

/* intdy with k = 0, LSODA.cpp:1435-1453 */
static void interpolate(const lsoda_ctx *s, double t, double *out)
{
    const double sfrac = (t - s->tn) / s->h;
    int i, j;
    for (i = 0; i < s->n; i++)
        out[i] = 1.0 * s->yh[s->l][i];
    for (j = s->nq - 1; j >= 0; j--)
        for (i = 0; i < s->n; i++)
            out[i] = 1.0 * s->yh[j + 1][i] + sfrac * out[i];
}

/* rtol_v / atol_v: per-component tolerances (the reference's itol_ = 3 / 4 branches,
 * LSODA.cpp:500-520, :625-640, :1368-1390), or NULL: the scalar applies to every component
 * (what lsoda_update sets up, :2267-2270). */
#define RT(i) (rtol_v ? rtol_v[i] : rtol)
#define AT(i) (atol_v ? atol_v[i] : atol)
void oracle_lsoda_v(oracle_rhs_fn f, int neq, const double *u0, void *data, int nt,
                    const double *teval, double *usol, double rtol, double atol,
                    const double *rtol_v, const double *atol_v, int mxstep,
                    int *success, int exit_on_warning, oracle_lsoda_stats *st)
{
    oracle_lsoda_band(f, neq, u0, data, nt, teval, usol, rtol, atol, rtol_v, atol_v, -1, -1, mxstep,
                      success, exit_on_warning, st);
}

/* ml, mu >= 0: jt = 5, banded finite-difference Jacobian (see band_factor above); ml < 0: the
 * reference's jt = 2 */
void oracle_lsoda_band(oracle_rhs_fn f, int neq, const double *u0, void *data, int nt,
                       const double *teval, double *usol, double rtol, double atol,
                       const double *rtol_v, const double *atol_v, int ml, int mu, int mxstep,
                       int *success, int exit_on_warning, oracle_lsoda_stats *st)
{
    const int n = neq;
    /* wrapper.cpp:16 stores the int in a size_t member (LSODA.h:130) */
    const unsigned long long mxstep_u = (unsigned long long)(long long)mxstep;
    oracle_lsoda_stats local;
    lsoda_ctx S, *s = &S;
    double *w, t0, tout, tdist, w0, tol, sum, h0, tolsf, tp;
    int i, row, istate = 1, nslast = 0, nhnil = 0, nswitch = 0, first_in_step;

    if (!st)
        st = &local;
    memset(st, 0, sizeof(*st));
    memset(s, 0, sizeof(*s));
    *success = 1;
    for (i = 0; i < n; i++)
        usol[i] = u0[i]; /* wrapper.cpp:25-28 */
    st->rows = 1;
    st->istate = 2;
    if (nt < 2)
        return;
    if (teval[1] < teval[0]) { /* wrapper.cpp:33-36 */
        *success = 0;
        st->istate = 0;
        return;
    }

    w = (double *)calloc((size_t)n * (MXORDN + 2 + 4) + (size_t)n * n, sizeof(double));
    s->n = n;
    s->f = f;
    s->data = data;
    s->rtol = rtol;
    s->atol = atol;
    for (i = 0; i <= MXORDN + 1; i++)
        s->yh[i] = w + (size_t)i * n;
    s->ewt = w + (size_t)(MXORDN + 2) * n;
    s->savf = s->ewt + n;
    s->acor = s->savf + n;
    s->y = s->acor + n;
    s->wm = s->y + n;
    s->ipvt = (int *)calloc((size_t)n, sizeof(int));
    s->banded = (ml >= 0 && mu >= 0);
    if (s->banded) {
        /* the reference's checker, LSODA.cpp:376-391 */
        if (ml >= n || mu >= n) {
            istate = -3;
            goto done;
        }
        s->ml = ml;
        s->mu = mu;
        s->abd = (double *)calloc((size_t)n * (2 * ml + mu + 1), sizeof(double));
    }

    /* ---- first call, blocks b/c of LSODA::lsoda, LSODA.cpp:341-663 ---- */
    t0 = teval[0];
    tout = teval[1];
    s->sqrteta = sqrt(ETA);
    s->meth = 1;
    s->tn = t0;
    s->maxord = MXORDN;
    s->jstart = 0;
    s->miter = 0;
    for (i = 0; i < n; i++)
        if (RT(i) < 0. || AT(i) < 0.) { /* LSODA.cpp:503-520 */
            istate = -3;
            goto done;
        }
    for (i = 0; i < n; i++)
        s->y[i] = u0[i];
    f(t0, s->y, s->yh[2], data);
    s->nfe = 1;
    for (i = 0; i < n; i++)
        s->yh[1][i] = s->y[i];
    s->nq = 1;
    s->h = 1.;
    for (i = 0; i < n; i++) {
        s->ewt[i] = RT(i) * fabs(s->y[i]) + AT(i); /* ewset, LSODA.cpp:1368-1390 */
        if (s->ewt[i] <= 0.) {
            /* LSODA.cpp:585-590 returns through terminate2, which leaves istate at 1,
             * so wrapper.cpp:41 does not see a failure: every lsoda_update call
             * re-initialises, fails the same way and hands back u0.  Reproduced. */
            int r, c;
            for (r = 1; r < nt; r++) {
                if (teval[r] < teval[r - 1]) {
                    istate = 0;
                    goto done;
                }
                for (c = 0; c < n; c++)
                    usol[(size_t)r * n + c] = u0[c];
                st->rows = r + 1;
            }
            istate = 2;
            goto done;
        }
        s->ewt[i] = 1. / s->ewt[i];
    }
    tdist = fabs(tout - t0);
    w0 = dmax(fabs(t0), fabs(tout));
    if (tdist < 2. * ETA * w0) { /* LSODA.cpp:618-623 */
        istate = -3;
        goto done;
    }
    tol = RT(0); /* :624-628 */
    for (i = 1; i < n; i++)
        tol = dmax(tol, RT(i));
    if (tol <= 0.) {
        for (i = 0; i < n; i++) {
            double ayi = fabs(s->y[i]);
            if (ayi != 0.)
                tol = dmax(tol, AT(i) / ayi);
        }
    }
    tol = dmax(tol, 100. * ETA);
    tol = dmin(tol, 0.001);
    sum = wmax_norm(n, s->yh[2], s->ewt);
    sum = 1. / (tol * w0 * w0) + tol * sum * sum;
    h0 = 1. / sqrt(sum);
    h0 = dmin(h0, tdist);
    h0 = h0 * ((tout - t0 >= 0.) ? 1. : -1.);
    s->h = h0;
    for (i = 0; i < n; i++)
        s->yh[2][i] *= h0;

    /* ---- step loop, block e of LSODA::lsoda, LSODA.cpp:774-992 ---- */
    row = 1;
    for (;;) {
        if (s->nst != 0) {
            if ((unsigned long long)(s->nst - nslast) >= mxstep_u) { /* :778 */
                istate = -1;
                break;
            }
            for (i = 0; i < n; i++) {
                s->ewt[i] = RT(i) * fabs(s->yh[1][i]) + AT(i);
                if (s->ewt[i] <= 0.) {
                    istate = -6;
                    break;
                }
                s->ewt[i] = 1. / s->ewt[i];
            }
            if (istate == -6)
                break;
        }
        tolsf = ETA * wmax_norm(n, s->yh[1], s->ewt);
        if (tolsf > 0.01) { /* :801-818 */
            istate = (s->nst == 0) ? -3 : -2;
            break;
        }
        if ((s->tn + s->h) == s->tn) { /* :820-845 */
            nhnil++;
            if (nhnil <= MXHNIL && exit_on_warning) {
                istate = -2;
                break;
            }
        }
        one_step(s);
        if (s->kflag != 0) { /* :961-991 */
            istate = (s->kflag == -1) ? -4 : -5;
            break;
        }
        if (s->meth != s->mused) { /* :861-877 */
            s->maxord = (s->meth == 2) ? MXORDS : MXORDN;
            s->jstart = -1;
            nswitch++;
        }
        /* every output time this step has reached: the first one is what the
         * reference interpolates at :887, the others are its block-d returns
         * (:675-689) on the following lsoda_update calls, which also re-check the
         * order of t_eval (wrapper.cpp:33) and intdy's interval (:1426-1434). */
        first_in_step = 1;
        while (row < nt) {
            if (teval[row] < teval[row - 1]) {
                istate = 0;
                break;
            }
            tout = teval[row];
            if ((s->tn - tout) * s->h < 0.)
                break;
            /* intdy's interval test, LSODA.cpp:1426-1434.  After a step (:887) the
             * reference ignores iflag and returns the untouched y (the last corrector
             * iterate); on a later call (:677-684) it is an illegal-input failure. */
            tp = s->tn - s->hu - 100. * ETA * (s->tn + s->hu);
            if ((tout - tp) * (tout - s->tn) > 0.) {
                if (!first_in_step) {
                    istate = -3;
                    break;
                }
                for (i = 0; i < n; i++)
                    usol[(size_t)row * n + i] = s->y[i];
            } else {
                interpolate(s, tout, usol + (size_t)row * n);
            }
            first_in_step = 0;
            nslast = s->nst;
            row++;
        }
        st->rows = row;
        if (istate <= 0 || row >= nt)
            break;
    }
    if (istate == 1)
        istate = 2;
done:
    st->nst = s->nst;
    st->nfe = s->nfe;
    st->nje = s->nje;
    st->nswitch = nswitch;
    st->nqu = s->nqu;
    st->meth = s->meth;
    st->istate = istate;
    st->nattempt = s->nattempt;
    st->nhnil = nhnil;
    st->tn = s->tn;
    st->hu = s->hu;
    if (istate <= 0)
        *success = 0; /* wrapper.cpp:41-45 */
    free(w);
    free(s->ipvt);
    free(s->abd);
}
#undef RT
#undef AT

void oracle_lsoda(oracle_rhs_fn f, int neq, const double *u0, void *data, int nt,
                  const double *teval, double *usol, double rtol, double atol, int mxstep,
                  int *success, int exit_on_warning, oracle_lsoda_stats *st)
{
    oracle_lsoda_v(f, neq, u0, data, nt, teval, usol, rtol, atol, NULL, NULL, mxstep, success,
                   exit_on_warning, st);
}
