// Tensor memory as per-lane row storage: does a CTA of four independent warps get 32 lanes x 256
// columns each, addressable with run-time column offsets, and what do a load and a store cost?
// (The dense warp-per-trajectory LSODA kernel keeps its 64 x 64 iteration matrix there:
// numbalsoda_b200/csrc/lsoda_vec_warp.cuh.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_probe tools/tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define COLS 256

__device__ __forceinline__ void tm_ld16(uint32_t taddr, double (&v)[8])
{
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}

__device__ __forceinline__ void tm_st16(uint32_t taddr, const double (&v)[8])
{
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r[2 * i] = (uint32_t)__double2loint(v[i]);
        r[2 * i + 1] = (uint32_t)__double2hiint(v[i]);
    }
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
        "tcgen05.wait::st.sync.aligned;\n"
        :
        : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

__device__ __forceinline__ double tm_ld1(uint32_t taddr)
{
    uint32_t lo, hi;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(lo), "=r"(hi)
        : "r"(taddr)
        : "memory");
    return __hiloint2double((int)hi, (int)lo);
}

__device__ __forceinline__ void tm_st1(uint32_t taddr, double x)
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};\n"
        "tcgen05.wait::st.sync.aligned;\n"
        :
        : "r"(taddr), "r"((uint32_t)__double2loint(x)), "r"((uint32_t)__double2hiint(x))
        : "memory");
}

__device__ double pattern(int cta, int warp, int lane, int col) { return cta * 1.0e6 + warp * 1.0e4 + lane * 1.0e2 + col + 0.25; }

__global__ void __launch_bounds__(128, 2) probe(int* errors, long long* cycles, int rounds)
{
    __shared__ uint32_t tm_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(&tm_base)),
                     "n"(COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t base = tm_base + ((uint32_t)(warp * 32) << 16);
    int bad = 0;
    // 128 doubles per lane: columns 0..255
    for (int c0 = 0; c0 < 128; c0 += 8) {
        double v[8];
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = pattern(blockIdx.x, warp, lane, c0 + e);
        tm_st16(base + 2 * c0, v);
    }
    // the other warps' stores must not have touched this warp's lanes
    __syncthreads();
    for (int c0 = 0; c0 < 128; c0 += 8) {
        double v[8];
        tm_ld16(base + 2 * c0, v);
#pragma unroll
        for (int e = 0; e < 8; e++) bad += v[e] != pattern(blockIdx.x, warp, lane, c0 + e);
    }
    // single elements at run-time columns, read-modify-write by one lane (everybody stores back what it holds)
    for (int k = 0; k < 128; k += 5) {
        double x = tm_ld1(base + 2 * k);
        bad += x != pattern(blockIdx.x, warp, lane, k);
        if (lane == (k & 31)) x = -x;
        tm_st1(base + 2 * k, x);
    }
    for (int k = 0; k < 128; k++) {
        const double x = tm_ld1(base + 2 * k);
        const bool flipped = (k % 5 == 0) && lane == (k & 31);
        bad += x != (flipped ? -1.0 : 1.0) * pattern(blockIdx.x, warp, lane, k);
    }
    // unaligned 8-double window (column offset not a multiple of 16)
    {
        double v[8];
        tm_ld16(base + 2 * 3, v);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int k = 3 + e;
            const bool flipped = (k % 5 == 0) && lane == (k & 31);
            bad += v[e] != (flipped ? -1.0 : 1.0) * pattern(blockIdx.x, warp, lane, k);
        }
    }
    if (bad) atomicAdd(errors, bad);
    // cost of dependent load -> multiply-add -> store round trips on one window, and of loads alone
    __syncwarp();
    long long t0 = clock64();
    double acc[8];
    tm_ld16(base, acc);
    for (int r = 0; r < rounds; r++) {
        double v[8];
        tm_ld16(base + 16 * (r & 7), v);
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = fma(v[e], 1.0000001, 1.0e-9);
        tm_st16(base + 16 * (r & 7), v);
    }
    long long t1 = clock64();
    for (int r = 0; r < rounds; r++) {
        double v[8];
        tm_ld16(base + 16 * (r & 7), v);
#pragma unroll
        for (int e = 0; e < 8; e++) acc[e] += v[e];
    }
    long long t2 = clock64();
    if (acc[0] == 123.0) errors[1] = 1;
    if (blockIdx.x == 0 && lane == 0) {
        cycles[2 * warp] = (t1 - t0) / rounds;
        cycles[2 * warp + 1] = (t2 - t1) / rounds;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm_base), "n"(COLS));
}

int main()
{
    int* errors;
    long long* cycles;
    cudaMallocManaged(&errors, 2 * sizeof(int));
    cudaMallocManaged(&cycles, 8 * sizeof(long long));
    errors[0] = errors[1] = 0;
    int sms = 0, occ = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe, 128, 0);
    probe<<<2 * sms, 128>>>(errors, cycles, 1000);
    cudaError_t e = cudaDeviceSynchronize();
    printf("{\"tmem_probe\": \"%s\", \"blocks_per_sm_by_registers\": %d, \"grid\": %d, \"mismatches\": %d, "
           "\"cycles_ld16_fma_st16\": [%lld, %lld, %lld, %lld], \"cycles_ld16\": [%lld, %lld, %lld, %lld]}\n",
           cudaGetErrorString(e), occ, 2 * sms, errors[0], cycles[0], cycles[2], cycles[4], cycles[6], cycles[1],
           cycles[3], cycles[5], cycles[7]);
    return e != cudaSuccess || errors[0] != 0;
}
