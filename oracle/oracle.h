/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * Interface of the CPU oracle (plain C restatement of the reference's lsoda /
 * dop853 paths).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it.  The product (libb200ode.so)
 * never links or loads anything from oracle/.
 */
#ifndef ORACLE_H
#define ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* The reference's right-hand-side contract: numbalsoda/driver.py:8-11 (lsoda_sig),
 * src/LSODA.h:32. */
typedef void (*oracle_rhs_fn)(double t, double *u, double *du, void *data);

typedef struct {
    int nfcn;   /* RHS evaluations              (module :587,:629,:691) */
    int nstep;  /* attempted steps              (module :558) */
    int naccpt; /* accepted steps               (module :627) */
    int nrejct; /* rejected steps after the 1st accept (module :727) */
    int nevent; /* accepted steps that prepared dense output (module :657) */
    int idid;   /* 1 ok, -1 bad input, -2 nmax, -3 step too small */
    int rows;   /* usol rows written */
    int pad_;
    double hlast;
    double xlast;
} oracle_dop853_stats;

typedef struct {
    int nst;     /* accepted steps                    (LSODA.cpp:1193) */
    int nfe;     /* RHS evaluations                   (LSODA.cpp:573,:1350,:1678,:1766,:1890,:1899) */
    int nje;     /* Jacobian evaluations              (LSODA.cpp:1649) */
    int nswitch; /* Adams<->BDF switches              (LSODA.cpp:861) */
    int nqu;     /* order of the last accepted step   (LSODA.cpp:1195) */
    int meth;    /* method in use at exit: 1 Adams, 2 BDF */
    int istate;  /* 2 ok; <0 as LSODA.cpp:782,:793,:815,:833,:969,:976; -3 illegal input */
    int rows;    /* usol rows written */
    int nattempt; /* prediction + corrector passes (accepted or not) */
    int nhnil;   /* t+h==t warnings (LSODA.cpp:822) */
    double tn;
    double hu;
} oracle_lsoda_stats;

/* Same contract as the reference's dop853_wrapper (src/dop853_c_interface.f90:75-130)
 * plus an optional statistics block. */
void oracle_dop853(oracle_rhs_fn f, int neq, const double *u0, void *data, int nt,
                   const double *teval, double *usol, double rtol, double atol, int mxstep,
                   int *success, oracle_dop853_stats *stats);

/* ... with per-component tolerances rtol_v[neq] / atol_v[neq] (NULL: the scalar applies): the
 * reference's itol = 1 branches, dop853_module.f90:603-612, :779-803 */
void oracle_dop853_v(oracle_rhs_fn f, int neq, const double *u0, void *data, int nt,
                     const double *teval, double *usol, double rtol, double atol,
                     const double *rtol_v, const double *atol_v, int mxstep,
                     int *success, oracle_dop853_stats *stats);

/* test-only, see dop853_oracle.c */
void oracle_dop853_set_scipy_reject_rule(int on);
void oracle_dop853_set_fast_ctrl(int on);

/* solve_ivp (numbalsoda/driver_solve_ivp.py:86-218, src/dop853_solve_ivp.f90): see
 * solve_ivp_oracle.c.  message: which of the reference's strings. */
#define ORACLE_IVP_MSG_END 0      /* 'The solver successfully reached the end of the integration interval.' */
#define ORACLE_IVP_MSG_EVENT 1    /* 'A termination event occurred.' */
#define ORACLE_IVP_MSG_FAILED 2   /* 'Failed to complete integration.' */
#define ORACLE_IVP_MSG_TEVAL 3    /* '"t_eval" does not always increase or decrease' */
#define ORACLE_IVP_MSG_INIT 4     /* 'Integrator failed to initialize.' (unreachable from the driver) */
#define ORACLE_IVP_MSG_BRENT 5    /* the reference prints 'brent failed' and stops the process */
#define ORACLE_IVP_MSG_CAPACITY 6 /* not in the reference: more accepted steps than rows provided */
typedef struct {
    int success, message, nfev, nt_out, event_found, ind_event;
    int idid, nstep, naccpt, nrejct;
    int rows_needed; /* rows a complete result has (= accepted steps + 1 without t_eval) */
    int pad_;
    double t_event;
} oracle_ivp_result;

/* t_eval given (nt > 0): y_out[nt,neq], t_out unused.  Otherwise every accepted step is
 * saved: t_out[cap], y_out[cap,neq]. */
void oracle_solve_ivp(oracle_rhs_fn f, const double *t_span, int neq, const double *y0, int nt,
                      const double *teval, int n_events, oracle_rhs_fn events, void *data,
                      double first_step, double max_step, double rtol, double atol, int cap,
                      double *t_out, double *y_out, double *y_event, oracle_ivp_result *res);

void oracle_ivp_set_scipy_reject_rule(int on); /* test-only, see solve_ivp_oracle.c */

/* ensemble of independent oracle_solve_ivp calls over host threads (batch.c) */
int oracle_solve_ivp_batch(oracle_rhs_fn f, oracle_rhs_fn events, long B, int neq, const double *t_span,
                           long t_span_stride, const double *y0, long y0_stride, const void *data,
                           long data_stride, int nt, const double *teval, int n_events,
                           double first_step, double max_step, double rtol, double atol, int cap,
                           double *t_out, double *y_out, double *y_event, oracle_ivp_result *res,
                           int nthreads);

/* Same contract as the reference's lsoda_wrapper (src/wrapper.cpp:10-57) plus an
 * optional statistics block. */
void oracle_lsoda(oracle_rhs_fn f, int neq, const double *u0, void *data, int nt,
                  const double *teval, double *usol, double rtol, double atol, int mxstep,
                  int *success, int exit_on_warning, oracle_lsoda_stats *stats);

/* ... with per-component tolerances (NULL: the scalar applies): the reference's itol_ = 3 / 4
 * branches, LSODA.cpp:500-520, :624-640, :1368-1390 */
void oracle_lsoda_v(oracle_rhs_fn f, int neq, const double *u0, void *data, int nt,
                    const double *teval, double *usol, double rtol, double atol,
                    const double *rtol_v, const double *atol_v, int mxstep,
                    int *success, int exit_on_warning, oracle_lsoda_stats *stats);

/* ... and a banded finite-difference Jacobian (jt = 5 with ml, mu, which the reference's
 * checker accepts, LSODA.cpp:369-392, and its prja / solsy do not implement): an extension
 * restating ODEPACK's DPRJA / DSOLSY miter = 5 with LINPACK dgbfa / dgbsl.  ml < 0: jt = 2. */
void oracle_lsoda_band(oracle_rhs_fn f, int neq, const double *u0, void *data, int nt,
                       const double *teval, double *usol, double rtol, double atol,
                       const double *rtol_v, const double *atol_v, int ml, int mu, int mxstep,
                       int *success, int exit_on_warning, oracle_lsoda_stats *stats);
/* oracle_solve_batch_v uses these for its LSODA calls (test knob, set before the call) */
void oracle_batch_set_band(int ml, int mu);

/* Ensemble helpers: the way numbalsoda is used for ensembles today is a prange
 * loop of independent calls (README.md:60-67); these do the same with OpenMP.
 * u0 is [B,n] (u0_stride = n) or a single [n] vector (u0_stride = 0); data is
 * B records of data_stride bytes (0 = shared).  usol is [B,nt,n].
 * solver: 0 = dop853, 1 = lsoda.  Returns the number of threads used. */
int oracle_solve_batch(int solver, oracle_rhs_fn f, long B, int neq, const double *u0,
                       long u0_stride, const void *data, long data_stride, int nt,
                       const double *teval, double *usol, double rtol, double atol,
                       int mxstep, int exit_on_warning, int *success, int *counters,
                       int nthreads);
/* rtol_v / atol_v: trajectory b's per-component tolerances at rtol_v + b * rtol_vs (stride 0 =
 * one vector for all), NULL = the scalar */
int oracle_solve_batch_v(int solver, oracle_rhs_fn f, long B, int neq, const double *u0,
                         long u0_stride, const void *data, long data_stride, int nt,
                         const double *teval, double *usol, double rtol, double atol,
                         const double *rtol_v, long rtol_vs, const double *atol_v, long atol_vs,
                         int mxstep, int exit_on_warning, int *success, int *counters,
                         int nthreads);
#define ORACLE_NCOUNTERS 8 /* ints per trajectory written to `counters` (may be NULL) */

/* Built-in right-hand sides written with the operation order numba produces for
 * the reference's benchmark scripts (benchmark/benchmark_lorenz.py:9-16,
 * benchmark/benchmark_robber.py:8-14, README.md:26-31). */
void oracle_rhs_lotka_volterra(double t, double *u, double *du, void *p);
void oracle_rhs_lorenz(double t, double *u, double *du, void *p);
void oracle_rhs_robertson(double t, double *u, double *du, void *p);
void oracle_rhs_brusselator1d(double t, double *u, double *du, void *p);
void oracle_rhs_brusselator1d_il(double t, double *u, double *du, void *p);
oracle_rhs_fn oracle_builtin_rhs(const char *name);

#ifdef __cplusplus
}
#endif
#endif
