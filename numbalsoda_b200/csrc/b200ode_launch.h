/* Kernel argument block shared by the host library (b200ode.cpp) and the device
 * code (b200ode_device.cuh).  Plain C, fixed layout. */
#ifndef B200ODE_LAUNCH_H
#define B200ODE_LAUNCH_H

#define B200ODE_NPMAX 16        /* parameters staged into registers per trajectory */
#define B200ODE_NCOUNTERS_DEV 8 /* ints per trajectory in the counters array */

/* DOP853 dense-output queue (dop853_kernel.cu): per warp, QCAP records of QFIELDS
 * doubles in dynamic shared memory.  QCAP is a build-time choice of each kernel image
 * (the host asks the image: b200ode_dop853_config). */
#define B200ODE_DOP853_QFIELDS(n) (4 + 11 * (n))
#define B200ODE_THREADS 128 /* threads per block of every solver kernel */

/* Systems of up to B200ODE_THREAD_NMAX equations run one trajectory per thread
 * (dop853_kernel.cu, lsoda_kernel.cu: one image per size), larger ones -- up to
 * B200ODE_WARP_NMAX -- one trajectory per warp (dop853_warp_kernel.cu, lsoda_warp_kernel.cu:
 * a single image each, size at run time) with this many doubles of dynamic shared memory
 * per warp. */
#define B200ODE_DOP853W_DOUBLES(n) (22 * (n)) /* dop853_warp.cuh: shared memory of one trajectory */
#define B200ODE_THREAD_NMAX 8
#define B200ODE_WARP_NMAX 128
#ifndef B200ODE_LSODAW_JCOLS
#define B200ODE_LSODAW_JCOLS 16 /* Jacobian columns evaluated concurrently (<= 32): measured at
                                   n = 64 -- 8: 77.0 k traj/s, 16: 80.1 k (still 4 warps per SM),
                                   32: 61.5 k (3 warps per SM); profiles/r1_lsodaw_jcols_sweep.log */
#endif
#define B200ODE_LSODAW_MATRIX(n) ((n) * ((n) + 1)) /* the transposed, padded iteration matrix */
#define B200ODE_LSODAW_DOUBLES(n) \
    (18 * (n) + B200ODE_LSODAW_MATRIX(n) + B200ODE_LSODAW_JCOLS * ((n) + 1) + (n) + 1)

/* ... with the dense matrix in tensor memory (32 < n <= 64; lsoda_vec_warp.cuh): blocks of eight
 * independent warps, each with 32 lanes x 256 columns of the block's 512-column allocation; shared
 * memory keeps the vectors, the perturbed copies of y and their right-hand sides */
#define B200ODE_LSODAW_TMEM_NMAX 64
#define B200ODE_LSODAW_TMEM_WARPS 8
#define B200ODE_LSODAW_TMEM_DOUBLES(n) (18 * (n) + 2 * B200ODE_LSODAW_JCOLS * ((n) + 1) + (n) + 1)

/* ... and with a banded iteration matrix (ml, mu >= 0): band storage (2 ml + mu + 1) x n, and two
 * vectors (perturbed state, its right-hand side) for each column group evaluated concurrently */
#define B200ODE_LSODAW_BAND_MATRIX(n, ml, mu) ((2 * (ml) + (mu) + 1) * (n))
#define B200ODE_LSODAW_BAND_COPIES(n, ml, mu)                                             \
    (((ml) + (mu) + 1 < (n) ? (ml) + (mu) + 1 : (n)) < B200ODE_LSODAW_JCOLS              \
         ? ((ml) + (mu) + 1 < (n) ? (ml) + (mu) + 1 : (n))                                \
         : B200ODE_LSODAW_JCOLS)
#define B200ODE_LSODAW_BAND_DOUBLES(n, ml, mu)                         \
    (18 * (n) + B200ODE_LSODAW_BAND_MATRIX(n, ml, mu) +                \
     2 * B200ODE_LSODAW_BAND_COPIES(n, ml, mu) * ((n) + 1) + (n) + 1)

typedef struct B200OdeLaunch {
    long long B;              /* trajectories in this launch */
    const double* u0;         /* [B,n] (u0_stride = n) or [n] shared (u0_stride = 0) */
    long long u0_stride;      /* in doubles */
    const char* data;         /* B records of data_stride bytes, or one shared record */
    long long data_stride;    /* in bytes; 0 = shared */
    int np;                   /* doubles per record staged in registers; < 0: pass the pointer */
    int nt;
    const double* t_eval;     /* [nt] */
    double* usol;             /* [B,nt,n], the reference's row layout per trajectory */
    double rtol, atol;
    const double* rtol_b;     /* tolerance arrays, or null: the scalars apply.  Trajectory b's */
    const double* atol_b;     /* value is rtol_b[b * rtol_sb] -- or, with rtol_vec, its value for */
    long long rtol_sb;        /* component i is rtol_b[b * rtol_sb + i]: [B] = (1, 0), [neq] shared */
    long long atol_sb;        /* by all = (0, 1), [B,neq] = (neq, 1) as (sb, vec) */
    int rtol_vec, atol_vec;
    int mxstep;
    int exit_on_warning;      /* lsoda only */
    int neq;                  /* system size (run-time for the warp-per-trajectory kernel) */
    int strict;               /* handle linked with B200ODE_STRICT_FP: keep the reference's
                                 summation order everywhere (warp kernel's triangular solves) */
    int* success;             /* [B] 1 / 0, as the reference's *success */
    int* counters;            /* [B,8] or null */
    unsigned long long* next; /* work queue head (zeroed before the launch) */
    double* scratch;          /* warp kernel: per-block iteration matrices in global memory
                                 (B200ODE_LSODAW_MATRIX(neq) doubles each), or null = shared */
    int ml, mu;               /* lsoda: banded finite-difference Jacobian (jt = 5) when ml >= 0 -- used
                                 by the warp kernel; the thread kernel (n <= 8) keeps the dense one */
    int te_shared;            /* dop853 thread kernel: the launch carries nt more doubles of dynamic
                                 shared memory and every block keeps its own copy of t_eval there */
} B200OdeLaunch;

/* ---- solve_ivp (solve_ivp_core.cuh, solve_ivp_kernel.cu) */
#define B200ODE_IVP_NEVMAX 8   /* event functions per problem (values kept in registers) */
#define B200ODE_IVP_NRESULT 12 /* ints per trajectory in the result array (include/b200ode.h) */
/* per thread: 8 n dense-output coefficients + the previous event values */
#define B200ODE_IVP_DOUBLES(n) (8 * (n) + B200ODE_IVP_NEVMAX)
/* warp per trajectory (solve_ivp_warp.cuh): 12 stage vectors, 8 n coefficients, 2 n norm terms,
 * one interpolated state, three event vectors */
#define B200ODE_IVPW_DOUBLES(n) (23 * (n) + 3 * B200ODE_IVP_NEVMAX)
/* message codes = which of the reference's strings (dop853_solve_ivp.f90:273,:300,:351,:364,:390) */
#define B200ODE_IVP_MSG_END 0
#define B200ODE_IVP_MSG_EVENT 1
#define B200ODE_IVP_MSG_FAILED 2
#define B200ODE_IVP_MSG_TEVAL 3
#define B200ODE_IVP_MSG_INIT 4
#define B200ODE_IVP_MSG_BRENT 5
#define B200ODE_IVP_MSG_CAPACITY 6

typedef struct B200IvpLaunch {
    long long B;
    const double* t_span;    /* [B,2] (t_span_stride = 2) or [2] shared (0) */
    long long t_span_stride; /* in doubles */
    const double* y0;        /* [B,n] (y0_stride = n) or [n] shared (0) */
    long long y0_stride;
    const char* data;        /* as B200OdeLaunch */
    long long data_stride;
    int np;
    int nt;                  /* 0: no t_eval, every accepted step is a row */
    const double* t_eval;    /* [nt] */
    int n_events;
    long long cap;           /* without t_eval: rows per trajectory of t_out / y_out ... */
    const long long* row_offset; /* ... or, if not null, [B+1] ragged row offsets */
    double first_step, max_step, rtol, atol;
    const double* rtol_b;    /* [B] or null */
    const double* atol_b;
    double* t_out;           /* [B,cap] or [row_offset[B]], written without t_eval only */
    double* y_out;           /* [B,nt,n], [B,cap,n] or [row_offset[B],n] */
    int* result;             /* [B,12] */
    double* t_event;         /* [B] */
    double* y_event;         /* [B,n] */
    unsigned long long* next;
} B200IvpLaunch;

#endif
