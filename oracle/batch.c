/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * Ensemble driver for the CPU oracle: one independent solver call per
 * trajectory, spread over host threads (pthreads, dynamic chunks of 8) -- the C
 * equivalent of the numba prange loop the reference's README (:60-67) recommends
 * for ensembles.
 */
#include <pthread.h>
#include <stdatomic.h>
#include <stddef.h>
#include <unistd.h>

#include "oracle.h"

typedef struct {
    int solver;
    oracle_rhs_fn f;
    long B;
    int neq;
    const double *u0;
    long u0_stride;
    const void *data;
    long data_stride;
    int nt;
    const double *teval;
    double *usol;
    double rtol, atol;
    int mxstep, exit_on_warning;
    int *success, *counters;
    /* per-component tolerance vectors of trajectory b: rtol_v + b * rtol_vs (NULL: the scalar) */
    const double *rtol_v, *atol_v;
    long rtol_vs, atol_vs;
    atomic_long next;
} job_t;

static int g_ml = -1, g_mu = -1; /* banded Jacobian of the LSODA calls (ml < 0: dense) */
void oracle_batch_set_band(int ml, int mu)
{
    g_ml = ml;
    g_mu = mu;
}

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    for (;;) {
        long b0 = atomic_fetch_add(&j->next, 8), b1 = b0 + 8;
        if (b0 >= j->B)
            break;
        if (b1 > j->B)
            b1 = j->B;
        for (long b = b0; b < b1; b++) {
            const double *u0b = j->u0 + b * j->u0_stride;
            void *db = (void *)((const char *)j->data + b * j->data_stride);
            double *out = j->usol + (size_t)b * j->nt * j->neq;
            int ok = 0;
            if (j->solver == 0) {
                oracle_dop853_stats st;
                oracle_dop853_v(j->f, j->neq, u0b, db, j->nt, j->teval, out, j->rtol, j->atol,
                                j->rtol_v ? j->rtol_v + b * j->rtol_vs : 0,
                                j->atol_v ? j->atol_v + b * j->atol_vs : 0, j->mxstep, &ok, &st);
                if (j->counters) {
                    int *c = j->counters + b * ORACLE_NCOUNTERS;
                    c[0] = st.nstep, c[1] = st.nfcn, c[2] = st.naccpt, c[3] = st.nrejct;
                    c[4] = st.nevent, c[5] = st.idid, c[6] = st.rows, c[7] = 0;
                }
            } else {
                oracle_lsoda_stats st;
                oracle_lsoda_band(j->f, j->neq, u0b, db, j->nt, j->teval, out, j->rtol, j->atol,
                                  j->rtol_v ? j->rtol_v + b * j->rtol_vs : 0,
                                  j->atol_v ? j->atol_v + b * j->atol_vs : 0, g_ml, g_mu, j->mxstep,
                                  &ok, j->exit_on_warning, &st);
                if (j->counters) {
                    int *c = j->counters + b * ORACLE_NCOUNTERS;
                    c[0] = st.nst, c[1] = st.nfe, c[2] = st.nje, c[3] = st.nswitch;
                    c[4] = st.nqu, c[5] = st.istate, c[6] = st.rows, c[7] = st.meth;
                }
            }
            j->success[b] = ok;
        }
    }
    return 0;
}

int oracle_solve_batch_v(int solver, oracle_rhs_fn f, long B, int neq, const double *u0,
                         long u0_stride, const void *data, long data_stride, int nt,
                         const double *teval, double *usol, double rtol, double atol,
                         const double *rtol_v, long rtol_vs, const double *atol_v, long atol_vs,
                         int mxstep, int exit_on_warning, int *success, int *counters,
                         int nthreads)
{
    job_t j = {solver, f,    B,    neq,  u0,     u0_stride,       data,    data_stride,
               nt,     teval, usol, rtol, atol,   mxstep, exit_on_warning, success, counters,
               rtol_v, atol_v, rtol_vs, atol_vs, 0};
    pthread_t th[256];
    if (nthreads <= 0)
        nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nthreads > 256)
        nthreads = 256;
    if (nthreads < 1)
        nthreads = 1;
    for (int i = 1; i < nthreads; i++)
        pthread_create(&th[i], 0, worker, &j);
    worker(&j);
    for (int i = 1; i < nthreads; i++)
        pthread_join(th[i], 0);
    return nthreads;
}

int oracle_solve_batch(int solver, oracle_rhs_fn f, long B, int neq, const double *u0,
                       long u0_stride, const void *data, long data_stride, int nt,
                       const double *teval, double *usol, double rtol, double atol,
                       int mxstep, int exit_on_warning, int *success, int *counters,
                       int nthreads)
{
    return oracle_solve_batch_v(solver, f, B, neq, u0, u0_stride, data, data_stride, nt, teval, usol,
                                rtol, atol, 0, 0, 0, 0, mxstep, exit_on_warning, success, counters,
                                nthreads);
}

/* ---- solve_ivp ensembles: t_eval given (y_out[B,nt,neq]) or `cap` rows per trajectory
 * (t_out[B,cap], y_out[B,cap,neq]); res[B], y_event[B,neq]. */
typedef struct {
    oracle_rhs_fn f, events;
    long B;
    int neq, nt, n_events, cap;
    const double *t_span;
    long t_span_stride;
    const double *y0;
    long y0_stride;
    const void *data;
    long data_stride;
    const double *teval;
    double first_step, max_step, rtol, atol;
    double *t_out, *y_out, *y_event;
    oracle_ivp_result *res;
    atomic_long next;
} ivp_job_t;

static void *ivp_worker(void *arg)
{
    ivp_job_t *j = (ivp_job_t *)arg;
    const long rows = j->nt > 0 ? j->nt : j->cap;
    for (;;) {
        long b0 = atomic_fetch_add(&j->next, 8), b1 = b0 + 8;
        if (b0 >= j->B)
            break;
        if (b1 > j->B)
            b1 = j->B;
        for (long b = b0; b < b1; b++)
            oracle_solve_ivp(j->f, j->t_span + b * j->t_span_stride, j->neq, j->y0 + b * j->y0_stride,
                             j->nt, j->teval, j->n_events, j->events,
                             (void *)((const char *)j->data + b * j->data_stride), j->first_step,
                             j->max_step, j->rtol, j->atol, j->cap,
                             j->t_out ? j->t_out + (size_t)b * rows : 0,
                             j->y_out + (size_t)b * rows * j->neq, j->y_event + (size_t)b * j->neq,
                             &j->res[b]);
    }
    return 0;
}

int oracle_solve_ivp_batch(oracle_rhs_fn f, oracle_rhs_fn events, long B, int neq, const double *t_span,
                           long t_span_stride, const double *y0, long y0_stride, const void *data,
                           long data_stride, int nt, const double *teval, int n_events,
                           double first_step, double max_step, double rtol, double atol, int cap,
                           double *t_out, double *y_out, double *y_event, oracle_ivp_result *res,
                           int nthreads)
{
    ivp_job_t j = {f, events, B, neq, nt, n_events, cap, t_span, t_span_stride, y0, y0_stride, data,
                   data_stride, teval, first_step, max_step, rtol, atol, t_out, y_out, y_event, res, 0};
    pthread_t th[256];
    if (nthreads <= 0)
        nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nthreads > 256)
        nthreads = 256;
    if (nthreads < 1)
        nthreads = 1;
    for (int i = 1; i < nthreads; i++)
        pthread_create(&th[i], 0, ivp_worker, &j);
    ivp_worker(&j);
    for (int i = 1; i < nthreads; i++)
        pthread_join(th[i], 0);
    return nthreads;
}
