// Variants of the IMAD.WIDE ceiling measurement (development aid for bbp_int_peak).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
template <int V, int ILP>
__global__ void __launch_bounds__(1024, 2) k(uint32_t *out, uint32_t seed) {
    uint64_t acc[ILP];
    uint32_t m[ILP];
    uint32_t x = seed + threadIdx.x, y = seed * 7 + blockIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; i++) { acc[i] = ((uint64_t)(y + 7 * i) << 32) | (x + i); m[i] = x * (i + 3) + y; }
#pragma unroll 1
    for (int it = 0; it < ITERS; it += 4) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                if (V == 0) acc[i] += (uint64_t)(uint32_t)acc[i] * x;                       // own low word
                if (V == 1) acc[i] += (uint64_t)(uint32_t)acc[(i + 1) % ILP] * x;           // neighbour's low word
                if (V == 2) acc[i] += (uint64_t)m[i] * (uint32_t)(acc[(i + 1) % ILP] >> 32);  // distinct multiplicand, neighbour's high word
                if (V == 3) { uint32_t lo = (uint32_t)acc[i], hi = (uint32_t)(acc[i] >> 32);
                    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(m[i]), "r"(x));
                    acc[i] = ((uint64_t)hi << 32) | lo; }
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) r ^= (uint32_t)acc[i] ^ (uint32_t)(acc[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int V, int ILP>
void run(const char *name, int nsm) {
    int blocks = nsm * 2, threads = 1024;
    uint32_t *out; cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0); k<V, ILP><<<blocks, threads>>>(out, rep + 1); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    double total = (double)blocks * threads * (double)ITERS * ILP;
    printf("{\"variant\": \"%s\", \"ilp\": %d, \"T_wide_per_s\": %.3f, \"ms\": %.4f}\n", name, ILP, total / (best * 1e-3) / 1e12, best);
    cudaFree(out);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int nsm = p.multiProcessorCount;
    run<0, 8>("own-low-word", nsm); run<1, 8>("neighbour-low-word", nsm); run<2, 8>("distinct-multiplicand", nsm); run<3, 8>("asm lo.cc/hi pair", nsm);
    run<0, 4>("own-low-word", nsm); run<2, 4>("distinct-multiplicand", nsm); run<2, 12>("distinct-multiplicand", nsm); run<3, 12>("asm lo.cc/hi pair", nsm);
    return 0;
}
