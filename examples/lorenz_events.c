/* libb200ode's solve_ivp from plain C: what the reference's dop853_solve_ivp_wrapper +
 * dop853_solve_ivp_result (src/dop853_solve_ivp.f90:211-434) do for one trajectory, for an
 * ensemble on the GPU, with terminal events.  No Python, no torch.
 *
 *   nvcc ... -DLORENZ_RHS -o lorenz_rhs.fatbin lorenz_events.cu     (see lorenz_events.cu)
 *   nvcc ... -DLORENZ_EVENTS -o lorenz_events.fatbin lorenz_events.cu
 *   gcc -O2 -I../include -o lorenz_events lorenz_events.c -L../numbalsoda_b200 -lb200ode -lm
 *   LD_LIBRARY_PATH=../numbalsoda_b200 ./lorenz_events lorenz_rhs.fatbin lorenz_events.fatbin 1000
 *
 * Every accepted step is an output row (no t_eval): a first pass with cap = 0 counts the rows
 * of each trajectory, a second one writes them into exactly sized ragged storage.
 * Prints, per trajectory: success, message code, nfev, rows, event_found, ind_event, t_event,
 * y_event and the last row.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200ode.h"

static void* slurp(const char* path, size_t* size)
{
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    fseek(f, 0, SEEK_END);
    *size = (size_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    void* p = malloc(*size);
    if (fread(p, 1, *size, f) != *size) { perror("read"); exit(2); }
    fclose(f);
    return p;
}

int main(int argc, char** argv)
{
    if (argc < 4) {
        fprintf(stderr, "usage: %s rhs.fatbin events.fatbin B [strict]\n", argv[0]);
        return 2;
    }
    const long long B = atoll(argv[3]);
    const unsigned flags = (argc > 4 && !strcmp(argv[4], "strict")) ? B200ODE_STRICT_FP : 0;
    size_t rhs_size, ev_size;
    void* rhs_img = slurp(argv[1], &rhs_size);
    void* ev_img = slurp(argv[2], &ev_size);
    enum { N = 3, NEV = 2 };

    b200ode_rhs_t ivp = NULL;
    if (b200ode_link_ivp(N, rhs_img, rhs_size, B200ODE_IMG_FATBIN, NEV, ev_img, ev_size,
                         B200ODE_IMG_FATBIN, flags, &ivp)) {
        fprintf(stderr, "link: %s\n", b200ode_last_error());
        return 1;
    }
    double t_span[2] = {0.0, 5.0}, y0[N] = {1.0, 0.0, 0.0};
    double* data = malloc(sizeof(double) * 4 * (size_t)B);
    for (long long b = 0; b < B; b++) { /* a deterministic sweep */
        const double s = (double)b / (double)(B > 1 ? B - 1 : 1);
        data[4 * b + 0] = 9.0 + 2.0 * s;
        data[4 * b + 1] = 20.0 + 16.0 * s;
        data[4 * b + 2] = 2.0 + 1.33 * s;
        data[4 * b + 3] = 25.0 + 30.0 * (1.0 - s); /* the z level of event 0 */
    }
    int* result = malloc(sizeof(int) * B200ODE_IVP_NRESULT * (size_t)B);
    double* t_event = malloc(sizeof(double) * (size_t)B);
    double* y_event = malloc(sizeof(double) * N * (size_t)B);
    long long* row_offset = malloc(sizeof(long long) * (size_t)(B + 1));

    b200ode_ivp_problem p;
    memset(&p, 0, sizeof p);
    p.size = sizeof p;
    p.B = B; p.neq = N;
    p.t_span = t_span; p.t_span_stride = 0;   /* one span and one initial state for all */
    p.y0 = y0; p.y0_stride = 0;
    p.data = data; p.data_stride = 4 * sizeof(double); p.np = 4;
    p.nt = 0; p.teval = NULL;                 /* rows = the solver's own steps */
    p.n_events = NEV;
    p.rtol = 1.0e-8; p.atol = 1.0e-8;
    p.result = result; p.t_event = t_event; p.y_event = y_event;
    p.cap = 0;                                /* pass 1: count */
    if (b200ode_solve_ivp(ivp, &p, 0)) {
        fprintf(stderr, "solve_ivp: %s\n", b200ode_last_error());
        return 1;
    }
    row_offset[0] = 0;
    for (long long b = 0; b < B; b++)
        row_offset[b + 1] = row_offset[b] + result[B200ODE_IVP_NRESULT * b + B200ODE_IVP_ROWS_NEEDED];
    const long long R = row_offset[B];
    double* t = malloc(sizeof(double) * (size_t)(R ? R : 1));
    double* y = malloc(sizeof(double) * N * (size_t)(R ? R : 1));
    p.row_offset = row_offset; p.t_out = t; p.y_out = y; /* pass 2: write */
    if (b200ode_solve_ivp(ivp, &p, 0)) {
        fprintf(stderr, "solve_ivp: %s\n", b200ode_last_error());
        return 1;
    }
    for (long long b = 0; b < B; b++) {
        const int* r = result + B200ODE_IVP_NRESULT * b;
        const long long last = row_offset[b] + r[B200ODE_IVP_NT_OUT] - 1;
        printf("%lld %d %d %d %d %d %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", b,
               r[B200ODE_IVP_SUCCESS], r[B200ODE_IVP_MESSAGE], r[B200ODE_IVP_NFEV],
               r[B200ODE_IVP_NT_OUT], r[B200ODE_IVP_EVENT_FOUND], r[B200ODE_IVP_IND_EVENT],
               t_event[b], y_event[N * b], y_event[N * b + 1], y_event[N * b + 2], t[last],
               y[N * last], y[N * last + 1], y[N * last + 2]);
    }
    if (B > 0) fprintf(stderr, "first message: %s\n", b200ode_ivp_message(result[B200ODE_IVP_MESSAGE]));
    b200ode_rhs_free(ivp);
    free(rhs_img); free(ev_img); free(data); free(result); free(t_event); free(y_event);
    free(row_offset); free(t); free(y);
    return 0;
}
