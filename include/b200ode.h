/* libb200ode -- C ABI of the B200-native batched ODE integrator.
 *
 * This is the drop-in boundary for the reference's two native entry points
 *
 *   void lsoda_wrapper (rhs, neq, u0, data, nt, teval, usol, rtol, atol, mxstep,
 *                       success, exit_on_warning)   src/wrapper.cpp:10-13
 *   void dop853_wrapper(rhs, neq, u0, data, nt, teval, usol, rtol, atol, mxstep,
 *                       success)                    src/dop853_c_interface.f90:75-87
 *
 * (ctypes prototypes: numbalsoda/driver.py:22-26, numbalsoda/driver_dop853.py:17-21).
 * The argument meaning, the [nt,neq] row layout of usol, the 1/0 success flag and
 * the failure semantics of those two functions are kept; two things change because
 * the solver now runs on the GPU:
 *
 *   - `rhs` cannot be an x86 function pointer.  The right-hand side is handed over
 *     once, as device code (LTO-IR / PTX produced by numba.cuda or nvcc, or the name
 *     of a built-in problem) and linked into the solver kernels: b200ode_link_rhs()
 *     returns a handle that takes the place of the function pointer;
 *   - every array gains a leading ensemble dimension B: u0[B,neq] (or one shared
 *     u0[neq]), data[B] records (or one shared record), usol[B,nt,neq], success[B].
 *
 * Plain pointers and sizes only; no C++ or torch types.  All functions return 0 on
 * success and a negative B200ODE_E* code otherwise; b200ode_last_error() gives the
 * message of the last failure on the calling thread.
 */
#ifndef B200ODE_H
#define B200ODE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200ODE_VERSION 101

/* solver */
#define B200ODE_DOP853 0 /* replaces dop853_wrapper, src/dop853_c_interface.f90:75 */
#define B200ODE_LSODA 1  /* replaces lsoda_wrapper,  src/wrapper.cpp:10 */
#define B200ODE_SOLVE_IVP 2 /* replaces dop853_solve_ivp_wrapper + dop853_solve_ivp_result,
                               src/dop853_solve_ivp.f90:211-434 (b200ode_link_ivp) */

/* kind of right-hand-side image passed to b200ode_link_rhs() */
#define B200ODE_IMG_LTOIR 0   /* NVVM LTO-IR, e.g. numba.cuda.compile(..., output='ltoir') */
#define B200ODE_IMG_PTX 1     /* PTX text (not inlined: slower, state spills to local memory) */
#define B200ODE_IMG_FATBIN 2  /* nvcc -gencode arch=compute_100a,code=lto_100a -fatbin */
#define B200ODE_IMG_BUILTIN 3 /* image = name of a built-in problem: "lorenz", "robertson",
                                 "lotka_volterra", "brusselator1d" */

/* flags of b200ode_link_rhs() */
#define B200ODE_STRICT_FP 1u /* no FMA contraction anywhere: arithmetic identical to the
                                reference's x86-64 build (parity mode).  Default (0) lets
                                the compiler fuse multiply-adds (throughput mode). */
#define B200ODE_EXACT_CTRL 2u /* throughput mode only: keep the reference's step-size
                                controller arithmetic (IEEE divisions, sqrt, libm-exact pow).
                                Default (0): DOP853's error norm and controller
                                (dop853_module.f90:603-623, :725) use reciprocals and a
                                single-precision err**(1/8) -- the accept/reject test stays in
                                double precision, the new step size differs from the
                                reference's by ~1e-7 relative.  Implied by STRICT_FP. */

/* error codes */
#define B200ODE_EINVAL -1  /* bad argument (also: more than 2^31 - 1 trajectories in one device-resident
                              launch -- the host-buffer entry points chunk the ensemble themselves) */
#define B200ODE_ELINK -2   /* nvJitLink failed (message has the linker log) */
#define B200ODE_ECUDA -3   /* CUDA driver error */
#define B200ODE_ENODEV -4  /* no CUDA driver / device: there is NO CPU fallback */
#define B200ODE_ENOSYS -5  /* unsupported combination (e.g. neq out of range) */

#define B200ODE_NCOUNTERS 8
/* counters[b*8 + k], optional per-trajectory statistics:
 *   dop853: nstep, nfcn, naccpt, nrejct, nevent, idid, rows written, 0
 *           (dop853_module.f90:141-164 `info`; idid as :340-347)
 *   lsoda : nst, nfe, nje, method switches, nqu, istate, rows written, meth
 *           (LSODA.h:128; istate as LSODA.cpp:782-976)                         */

typedef struct b200ode_rhs_s* b200ode_rhs_t;

/* The right-hand side  void f(double t, double* u, double* du, void* data)
 * (numbalsoda/driver.py:8-11, src/LSODA.h:32) must be a device function named
 * `b200ode_user_rhs` with C linkage taking (double, double*, double*, double*);
 * a 64-bit dummy return value (what numba's C ABI emits) is allowed.
 * Links it with the solver kernels for systems of `neq` equations; needs no GPU. */
int b200ode_link_rhs(int solver, int neq, const void* image, size_t size, int kind,
                     unsigned flags, b200ode_rhs_t* out);
/* The same with an optional second device function for systems of more than 8 equations, which
 * run one trajectory per WARP:
 * `b200ode_user_rhs_lanes`, a device function with C linkage taking
 * (double t, double* u, double* du, double* data, int lane, int nlanes) and returning long long,
 * is called by all `nlanes` (32) lanes of the warp at once and must, between them, write every
 * du[i] exactly as the scalar function does -- e.g. a method-of-lines loop `for i in
 * range(lane, m, nlanes)`.  The scalar function is still needed (the finite-difference Jacobian
 * evaluates one column per lane with it).  lanes_image = NULL: lane 0 calls the scalar function. */
int b200ode_link_rhs_lanes(int solver, int neq, const void* image, size_t size, int kind,
                           const void* lanes_image, size_t lanes_size, int lanes_kind,
                           unsigned flags, b200ode_rhs_t* out);
void b200ode_rhs_free(b200ode_rhs_t rhs);
/* The linked sm_100a cubin (for inspection: cuobjdump -sass / -res-usage). */
int b200ode_rhs_cubin(b200ode_rhs_t rhs, const void** cubin, size_t* size);

/* Batched dop853_wrapper.  Host pointers; copies inputs to `device`, integrates all
 * B trajectories there and copies usol / success (/ counters) back.
 *   u0        [B,neq] with u0_stride = neq, or one shared [neq] with u0_stride = 0
 *   data      data_stride > 0: B records of data_stride bytes, one per trajectory;
 *             data_stride < 0: one record of -data_stride bytes shared by all;
 *             data_stride = 0: one shared record of np doubles.
 *             np = number of leading doubles of a record the RHS may read through
 *             its `p` argument from registers (<= 16), or -1 to pass the record's
 *             device address unchanged (any size / layout, e.g. a C struct)
 *   usol      [B,nt,neq]; rows a failed trajectory did not reach are left untouched,
 *             as in the reference (np.empty memory there)
 *   success   [B], 1 or 0 exactly like the reference's *success
 *   counters  [B,8] or NULL                                                     */
int b200ode_dop853_batched(b200ode_rhs_t rhs, long long B, int neq, const double* u0,
                           long long u0_stride, const void* data, long long data_stride,
                           int np, int nt, const double* teval, double* usol, double rtol,
                           double atol, int mxstep, int* success, int* counters, int device);

/* Batched lsoda_wrapper; same conventions, plus exit_on_warning (wrapper.cpp:13). */
int b200ode_lsoda_batched(b200ode_rhs_t rhs, long long B, int neq, const double* u0,
                          long long u0_stride, const void* data, long long data_stride,
                          int np, int nt, const double* teval, double* usol, double rtol,
                          double atol, int mxstep, int* success, int* counters,
                          int exit_on_warning, int device);

/* Device-resident variants: every pointer is a device pointer on the current
 * device of the calling thread; `stream` is a CUstream / cudaStream_t (NULL = default
 * stream).  Asynchronous: returns after enqueueing.  Nothing is copied. */
int b200ode_dop853_batched_device(b200ode_rhs_t rhs, long long B, int neq, const double* u0,
                                  long long u0_stride, const void* data,
                                  long long data_stride, int np, int nt, const double* teval,
                                  double* usol, double rtol, double atol, int mxstep,
                                  int* success, int* counters, void* stream);
int b200ode_lsoda_batched_device(b200ode_rhs_t rhs, long long B, int neq, const double* u0,
                                 long long u0_stride, const void* data,
                                 long long data_stride, int np, int nt, const double* teval,
                                 double* usol, double rtol, double atol, int mxstep,
                                 int* success, int* counters, int exit_on_warning,
                                 void* stream);

/* The same two calls with every argument in one structure, plus what the positional
 * entry points cannot express: per-trajectory and per-component tolerances (SURVEY.md
 * section 8f, rank 1 -- the reference's drivers take one scalar rtol / atol,
 * numbalsoda/driver.py:29; its solvers support one value per component: LSODA's itol_ = 3 / 4,
 * LSODA.cpp:500-520, :624-640, :1368-1390, and dop853's itol = 1, dop853_module.f90:603-612,
 * :779-803).  The solver is the one `rhs` was linked for.  Zero-initialise the structure and
 * set `size = sizeof(b200ode_problem)`. */
#define B200ODE_TOL_PER_TRAJECTORY 0 /* rtol_b is [B] */
#define B200ODE_TOL_PER_COMPONENT 1  /* rtol_b is [neq], shared by every trajectory */
#define B200ODE_TOL_FULL 2           /* rtol_b is [B,neq] */
typedef struct b200ode_problem {
    size_t size;
    long long B;
    int neq;
    const double* u0;       /* [B,neq] (u0_stride = neq) or [neq] shared (u0_stride = 0) */
    long long u0_stride;
    const void* data;       /* as b200ode_dop853_batched */
    long long data_stride;
    int np;
    int nt;
    const double* teval;    /* [nt] */
    double* usol;           /* [B,nt,neq] */
    double rtol, atol;      /* used where the arrays below are NULL */
    const double* rtol_b;   /* [B], [neq] or [B,neq] as rtol_dims says, or NULL */
    const double* atol_b;   /* likewise with atol_dims */
    int mxstep;
    int exit_on_warning;    /* lsoda only */
    int* success;           /* [B] */
    int* counters;          /* [B,8] or NULL */
    int rtol_dims;          /* B200ODE_TOL_*: shape of rtol_b (0 = [B]) */
    int atol_dims;
    /* lsoda: how the iteration matrix is formed.  The reference
     * fixes jt = 2 (LSODA.cpp:2263: dense, by finite differences: neq right-hand-side calls
     * and a dense LU per Jacobian); its checker also accepts jt = 4 / 5 with ml, mu
     * (:369-392) but prja / solsy implement miter = 2 only (:1656-1660, :1936-1942).
     *   jt = 0 or 2        the reference's dense finite-difference Jacobian
     *   jt = 5, ml, mu     banded finite-difference Jacobian (ODEPACK's DPRJA / DSOLSY miter = 5
     *                      with LINPACK dgbfa / dgbsl): ml + mu + 1 right-hand-side calls and a
     *                      band LU.  Requires 0 <= ml, mu < neq.  Systems of up to 8 equations
     *                      (thread-per-trajectory kernels) keep the dense matrix. */
    int jt, ml, mu;
} b200ode_problem;
/* host pointers; copies to / from `device` */
int b200ode_solve(b200ode_rhs_t rhs, const b200ode_problem* problem, int device);
/* device pointers on the current device; asynchronous on `stream` */
int b200ode_solve_device(b200ode_rhs_t rhs, const b200ode_problem* problem, void* stream);

/* ---- solve_ivp: the reference's third entry point,
 *
 *   void dop853_solve_ivp_wrapper(rhs, t_span, neq, y0, nt, t_eval, n_events, events_fcn,
 *                                 data, first_step, max_step, rtol, atol, len_message,
 *                                 nt_out, n_events_found, d_p)   src/dop853_solve_ivp.f90:211-215
 *   void dop853_solve_ivp_result (d_p, neq, len_message, nt_out, message, success, nfev,
 *                                 t, y, event_found, ind_event, t_event, y_event)   :398-401
 *
 * (ctypes prototypes: numbalsoda/driver_solve_ivp.py:19-56), for ensembles: DOP853 over
 * t_span (forward or backward) with output at t_eval or at every accepted step, first_step /
 * max_step, and terminal events located by Brent's method on the dense interpolant.  The
 * reference's two-phase call (solve, then fetch the variable-length result through an opaque
 * pointer) becomes one call with caller-owned arrays.
 *
 * The events function  void events(double t, double* y, double* out, void* data)
 * (src/dop853_solve_ivp.f90:9-17) is handed over like the right-hand side: a device function
 * named `b200ode_user_events` with the same signature as `b200ode_user_rhs`, writing
 * n_events <= 8 values.  events_image = NULL with n_events = 0 links a handle without events. */
#define B200ODE_IVP_NEVENTS_MAX 8
int b200ode_link_ivp(int neq, const void* rhs_image, size_t rhs_size, int rhs_kind,
                     int n_events, const void* events_image, size_t events_size,
                     int events_kind, unsigned flags, b200ode_rhs_t* out);

/* result[b*12 + k]: what the reference's ODEResult holds (driver_solve_ivp.py:59-69) and
 * the solver's private counters */
#define B200ODE_IVP_NRESULT 12
#define B200ODE_IVP_SUCCESS 0     /* 1 / 0 */
#define B200ODE_IVP_MESSAGE 1     /* b200ode_ivp_message() gives the reference's string */
#define B200ODE_IVP_NFEV 2        /* right-hand-side evaluations (0 on failure, as the reference) */
#define B200ODE_IVP_NT_OUT 3      /* rows of t / y that are valid */
#define B200ODE_IVP_ROWS_NEEDED 4 /* rows a complete result has (no t_eval: accepted steps + 1) */
#define B200ODE_IVP_EVENT_FOUND 5
#define B200ODE_IVP_IND_EVENT 6   /* 0-based index of the event function that fired */
#define B200ODE_IVP_IDID 7        /* dop853_module.f90:340-347; 2 = stopped by an event */
#define B200ODE_IVP_NSTEP 8
#define B200ODE_IVP_NACCPT 9
#define B200ODE_IVP_NREJCT 10
const char* b200ode_ivp_message(int code);

/* Zero-initialise and set size = sizeof(b200ode_ivp_problem). */
typedef struct b200ode_ivp_problem {
    size_t size;
    long long B;
    int neq;
    const double* t_span;     /* [B,2] (t_span_stride = 2) or one shared [2] (0) */
    long long t_span_stride;
    const double* y0;         /* [B,neq] (y0_stride = neq) or one shared [neq] (0) */
    long long y0_stride;
    const void* data;         /* as b200ode_dop853_batched */
    long long data_stride;
    int np;
    int nt;                   /* 0: no t_eval -- every accepted step is an output row */
    const double* teval;      /* [nt], shared by the ensemble; must not decrease (increase) for a
                                 trajectory that integrates forward (backward) */
    int n_events;             /* as given to b200ode_link_ivp */
    double first_step;        /* 0: the solver chooses (hinit) */
    double max_step;          /* 0: no limit */
    double rtol, atol;
    const double* rtol_b;     /* [B] or NULL */
    const double* atol_b;
    /* Output rows.  With t_eval: y_out[B,nt,neq], t_out unused.  Without: trajectory b owns
     * `cap` rows, t_out[B,cap], y_out[B,cap,neq] -- or, with row_offset[B+1], the rows
     * [row_offset[b], row_offset[b+1]) of t_out[row_offset[B]], y_out[row_offset[B],neq]
     * (row_offset[0] need not be 0: a shard passes a slice of the offsets and the same arrays).
     * A trajectory with more accepted steps than rows ends with success = 0, message
     * "output capacity exceeded", its rows filled and ROWS_NEEDED = what it takes: a first
     * pass with cap = 0 (t_out = y_out = NULL) sizes an exact second one. */
    long long cap;
    const long long* row_offset;
    double* t_out;
    double* y_out;
    int* result;              /* [B,12] */
    double* t_event;          /* [B]; 0 without an event */
    double* y_event;          /* [B,neq]; 0 without an event */
} b200ode_ivp_problem;
/* host pointers; copies to / from `device` */
int b200ode_solve_ivp(b200ode_rhs_t ivp, const b200ode_ivp_problem* problem, int device);
/* device pointers on the current device; asynchronous on `stream` */
int b200ode_solve_ivp_device(b200ode_rhs_t ivp, const b200ode_ivp_problem* problem, void* stream);

/* b200ode_solve_ivp with positional arguments, for callers that cannot build the structure
 * (numba nopython code calls C through ctypes function objects: numbalsoda_b200/njit_api.py). */
int b200ode_solve_ivp_args(b200ode_rhs_t ivp, long long B, int neq, const double* t_span,
                           long long t_span_stride, const double* y0, long long y0_stride,
                           const void* data, long long data_stride, int np, int nt,
                           const double* teval, int n_events, double first_step, double max_step,
                           double rtol, double atol, long long cap, const long long* row_offset,
                           double* t_out, double* y_out, int* result, double* t_event,
                           double* y_event, int device);

/* Page-locked host memory for usol / success / counters: with it the device-to-host
 * copies of b200ode_*_batched run at full PCIe speed and overlap the kernels (with pageable
 * memory the driver stages them).  The reference's driver allocates usol with np.empty
 * (numbalsoda/driver.py:35); numbalsoda_b200 allocates it here.  Needs a CUDA driver. */
int b200ode_host_alloc(size_t bytes, void** ptr);
int b200ode_host_free(void* ptr);

/* Device copies of the arrays a `data` record points to.  The
 * reference lets the right-hand side reach further arrays through addresses stored in its data
 * structure (address_as_void_pointer, numbalsoda/driver.py:45-53;
 * passing_data_to_rhs_function.ipynb); on the GPU those addresses must be device addresses:
 * allocate a buffer on `device`, copy the array into it and store *dptr in the record that is
 * then passed by address (np = -1).  Synchronous. */
int b200ode_device_alloc(int device, size_t bytes, void** dptr);
int b200ode_device_upload(int device, void* dptr, const void* host, size_t bytes);
int b200ode_device_free(int device, void* dptr);

/* Launch geometry and resource usage of the kernel behind `rhs` on the current
 * device: info = {registers/thread, local bytes/thread, static smem bytes,
 * threads/block, blocks/SM, SM count, grid size of the last launch, kernel launches
 * so far}. */
int b200ode_kernel_info(b200ode_rhs_t rhs, int info[8]);

/* Self test: out[i] = device pow(x[i], y) for device arrays of n doubles.  The solvers'
 * step-size controllers must reproduce the host libm pow bit-for-bit (the reference
 * calls it at dop853_module.f90:618,:817 and LSODA.cpp:1229,:1983-2113). */
int b200ode_selftest_pow(const double* x_dev, double y, double* out_dev, long long n,
                         void* stream);

int b200ode_device_count(void);
const char* b200ode_last_error(void);
int b200ode_version(void);

#ifdef __cplusplus
}
#endif
#endif
