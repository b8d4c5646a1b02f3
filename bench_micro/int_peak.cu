// Integer-pipe microbenchmark for the roofline denominator (MEASURED_PEAKS.json has no integer peak).
// Each variant runs ILP independent dependency chains per thread so that latency is hidden at full occupancy;
// the rate is reported as thread-instructions per clock per SM (from clock64 inside the kernel) and as T instr/s
// (from CUDA events). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_peak int_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ILP 8
#define ITERS 4096

template <int OP>
__global__ void __launch_bounds__(1024, 2) k(uint32_t *out, unsigned long long *cyc, uint32_t seed) {
    uint32_t a[ILP], b[ILP];
    uint64_t w[ILP];
    uint32_t x = seed + threadIdx.x, y = seed * 3 + blockIdx.x + 1;
#pragma unroll
    for (int i = 0; i < ILP; i++) { a[i] = x + i; b[i] = y + 7 * i; w[i] = ((uint64_t)a[i] << 32) | b[i]; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(x), "r"(y));
            if (OP == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(x), "r"(y));
            if (OP == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x), "r"(a[i]));
            if (OP == 3) asm volatile("{ .reg .u32 lo, hi;\n\tmov.b64 {lo,hi}, %0;\n\tmad.lo.cc.u32 lo, %1, %2, lo;\n\tmadc.hi.u32 hi, %1, %2, hi;\n\tmov.b64 %0, {lo,hi}; }" : "+l"(w[i]) : "r"(x), "r"(a[i]));
            if (OP == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(x));
            if (OP == 5) asm volatile("mad.wide.u32 %0, %2, %3, %0;\n\tadd.u32 %1, %1, %2;" : "+l"(w[i]), "+r"(b[i]) : "r"(x), "r"(a[i]));
            if (OP == 6) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.cc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(x), "r"(y));
            if (OP == 7) asm volatile("mad.wide.u32 %0, %2, %3, %0;\n\tadd.u32 %1, %1, %2;\n\txor.b32 %1, %1, %3;" : "+l"(w[i]), "+r"(b[i]) : "r"(x), "r"(a[i]));
        }
    }
    long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) r ^= a[i] ^ b[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int OP>
void run(const char *name, int instr_per_slot, int nsm) {
    int blocks = nsm * 2, threads = 1024;
    uint32_t *out;
    unsigned long long *cyc;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaMalloc(&cyc, blocks * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, cyc, 1);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k<OP><<<blocks, threads>>>(out, cyc, rep + 2);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    unsigned long long h[4096];
    cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; i++) avg += (double)h[i];
    avg /= blocks;
    double slots_per_sm = 2.0 * threads * (double)ITERS * ILP;   // 2 resident CTAs per SM
    double total = (double)blocks * threads * (double)ITERS * ILP * instr_per_slot;
    printf("{\"op\": \"%s\", \"instr_per_slot\": %d, \"thread_instr_per_clk_per_sm\": %.2f, \"T_instr_per_s\": %.3f, \"ms\": %.4f, \"eff_clock_mhz\": %.0f}\n",
           name, instr_per_slot, slots_per_sm * instr_per_slot / avg, total / (best * 1e-3) / 1e12, best, avg / (best * 1e-3) / 1e6);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, nsm, p.clockRate);
    run<0>("mad.lo.u32 (IMAD)", 1, nsm);
    run<1>("mad.hi.u32 (IMAD.HI)", 1, nsm);
    run<2>("mad.wide.u32 (IMAD.WIDE.U32)", 1, nsm);
    run<3>("mad.lo.cc+madc.hi.cc pair", 2, nsm);
    run<4>("add.u32 (IADD3)", 1, nsm);
    run<5>("mad.wide + add (1:1 mix)", 2, nsm);
    run<6>("add.cc+addc.cc pair", 2, nsm);
    run<7>("mad.wide + add + xor (1:2 mix)", 3, nsm);
    return 0;
}
