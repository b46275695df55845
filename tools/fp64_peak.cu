// FP64 pipe peak on the device: independent DFMA chains per thread, swept over
// ILP and occupancy.  Prints one JSON line; bench.py's roofline uses the best rate
// (profiles/fp64_peak_*.json).  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int MODE>
__global__ void chains(double* out, int iters, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                if (MODE == 0) x[i] = fma(x[i], a, b);
                else if (MODE == 1) x[i] = x[i] + b;
                else x[i] = x[i] * a;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    if (s == 123.456) out[0] = s;
}

template <int ILP, int MODE>
double run(int blocks_per_sm, int threads, int sms, double* d)
{
    const int iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    chains<ILP, MODE><<<sms * blocks_per_sm, threads>>>(d, 16, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        chains<ILP, MODE><<<sms * blocks_per_sm, threads>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double ops = (double)sms * blocks_per_sm * threads * (double)iters * 16.0 * ILP;
    return ops / (best * 1e-3);  // instructions (lane-ops) per second
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* d;
    cudaMalloc(&d, 8);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double best_fma = 0;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"sweep\": [", p.name, sms, clk);
    bool first = true;
    for (int bps : {1, 2, 4, 8}) {
        double r1 = run<1, 0>(bps, 256, sms, d), r2 = run<2, 0>(bps, 256, sms, d),
               r4 = run<4, 0>(bps, 256, sms, d), r8 = run<8, 0>(bps, 256, sms, d);
        for (double r : {r1, r2, r4, r8}) if (r > best_fma) best_fma = r;
        printf("%s{\"warps_per_sm\": %d, \"dfma_per_s_ilp1\": %.4g, \"ilp2\": %.4g, \"ilp4\": %.4g, \"ilp8\": %.4g}",
               first ? "" : ", ", bps * 8, r1, r2, r4, r8);
        first = false;
    }
    double add = run<8, 1>(4, 256, sms, d), mul = run<8, 2>(4, 256, sms, d);
    printf("], \"dfma_per_s\": %.5g, \"dadd_per_s\": %.5g, \"dmul_per_s\": %.5g, "
           "\"fp64_tflops\": %.4f, \"dfma_per_clk_per_sm_at_boost\": %.2f}\n",
           best_fma, add, mul, 2.0 * best_fma / 1e12, best_fma / sms / (clk * 1e3));
    return 0;
}
