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

/* test-only, see dop853_oracle.c */
void oracle_dop853_set_scipy_reject_rule(int on);
void oracle_dop853_set_fast_ctrl(int on);

/* Same contract as the reference's lsoda_wrapper (src/wrapper.cpp:10-57) plus an
 * optional statistics block. */
void oracle_lsoda(oracle_rhs_fn f, int neq, const double *u0, void *data, int nt,
                  const double *teval, double *usol, double rtol, double atol, int mxstep,
                  int *success, int exit_on_warning, oracle_lsoda_stats *stats);

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
#define ORACLE_NCOUNTERS 8 /* ints per trajectory written to `counters` (may be NULL) */

/* Built-in right-hand sides written with the operation order numba produces for
 * the reference's benchmark scripts (benchmark/benchmark_lorenz.py:9-16,
 * benchmark/benchmark_robber.py:8-14, README.md:26-31). */
void oracle_rhs_lotka_volterra(double t, double *u, double *du, void *p);
void oracle_rhs_lorenz(double t, double *u, double *du, void *p);
void oracle_rhs_robertson(double t, double *u, double *du, void *p);
void oracle_rhs_brusselator1d(double t, double *u, double *du, void *p);
oracle_rhs_fn oracle_builtin_rhs(const char *name);

#ifdef __cplusplus
}
#endif
#endif
