/* libb200ode from plain C: what src/wrapper.cpp's lsoda_wrapper does for one trajectory,
 * for an ensemble on the GPU.  No Python, no torch.
 *
 *   nvcc -gencode arch=compute_100a,code=lto_100a -rdc=true -fatbin -o robertson_rhs.fatbin robertson_rhs.cu
 *   gcc -O2 -I../include -o robertson_ensemble robertson_ensemble.c -L../numbalsoda_b200 -lb200ode -lm
 *   LD_LIBRARY_PATH=../numbalsoda_b200 ./robertson_ensemble robertson_rhs.fatbin 1000
 *
 * Prints, per trajectory, success, the counters and the final state.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200ode.h"

int main(int argc, char** argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s rhs.fatbin B [strict]\n", argv[0]);
        return 2;
    }
    const long long B = atoll(argv[2]);
    const unsigned flags = (argc > 3 && !strcmp(argv[3], "strict")) ? B200ODE_STRICT_FP : 0;
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    fseek(f, 0, SEEK_END);
    const long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    void* image = malloc((size_t)size);
    if (fread(image, 1, (size_t)size, f) != (size_t)size) { perror("read"); return 2; }
    fclose(f);

    b200ode_rhs_t rhs = NULL;
    if (b200ode_link_rhs(B200ODE_LSODA, 3, image, (size_t)size, B200ODE_IMG_FATBIN, flags, &rhs)) {
        fprintf(stderr, "link: %s\n", b200ode_last_error());
        return 1;
    }
    enum { NT = 100, N = 3 };
    double teval[NT], u0[N] = {1.0, 0.0, 0.0};
    for (int i = 0; i < NT; i++) teval[i] = 1.0e5 * i / (NT - 1);
    double* data = malloc(sizeof(double) * 3 * (size_t)B);
    for (long long b = 0; b < B; b++) { /* a deterministic sweep of the rate constants */
        const double s = (double)b / (double)(B > 1 ? B - 1 : 1);
        data[3 * b + 0] = 0.04 * (0.5 + s);
        data[3 * b + 1] = 3.0e7 * (1.5 - s);
        data[3 * b + 2] = 1.0e4 * (0.75 + 0.5 * s);
    }
    double* usol = NULL; /* page-locked: the copies back run at PCIe speed */
    if (b200ode_host_alloc(sizeof(double) * NT * N * (size_t)B, (void**)&usol)) {
        fprintf(stderr, "host_alloc: %s\n", b200ode_last_error());
        return 1;
    }
    int* success = malloc(sizeof(int) * (size_t)B);
    int* counters = malloc(sizeof(int) * B200ODE_NCOUNTERS * (size_t)B);

    b200ode_problem p;
    memset(&p, 0, sizeof p);
    p.size = sizeof p;
    p.B = B; p.neq = N;
    p.u0 = u0; p.u0_stride = 0;                 /* one initial state shared by the ensemble */
    p.data = data; p.data_stride = 3 * sizeof(double); p.np = 3;
    p.nt = NT; p.teval = teval; p.usol = usol;
    p.rtol = 1.0e-8; p.atol = 1.0e-8;
    p.mxstep = 10000; p.exit_on_warning = 0;
    p.success = success; p.counters = counters;
    if (b200ode_solve(rhs, &p, 0)) {
        fprintf(stderr, "solve: %s\n", b200ode_last_error());
        return 1;
    }
    for (long long b = 0; b < B; b++) {
        const int* c = counters + B200ODE_NCOUNTERS * b;
        const double* y = usol + ((size_t)b * NT + (NT - 1)) * N;
        printf("%lld %d %d %d %d %d %.17g %.17g %.17g\n", b, success[b], c[0], c[1], c[2], c[3],
               y[0], y[1], y[2]);
    }
    b200ode_host_free(usol);
    b200ode_rhs_free(rhs);
    free(image); free(data); free(success); free(counters);
    return 0;
}
