// microbench.cu — integer-pipe issue rates on the box (NVIDIA publishes no INT32 figure for B200).
// Measures, per SM per clock, the throughput of the instruction mixes the NTT butterfly is made of.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64; typedef unsigned int u32;

template <int OP> __global__ void k(u64 *out, u64 seed, int iters) {
    u64 a[8]; u32 b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed * (threadIdx.x + 1 + i * 977); b[i] = (u32)(seed >> i) + threadIdx.x; }
    const u64 w = seed | 1, wq = ~seed, p = (seed >> 4) | 1;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) b[i] = b[i] * (u32)w + (u32)p;                        // IMAD (32-bit lo)
            if (OP == 1) a[i] = (u64)(u32)a[i] * (u32)w + a[i];                 // IMAD.WIDE.U32
            if (OP == 2) b[i] = __umulhi(b[i], (u32)w) + (u32)p;               // IMAD.HI.U32
            if (OP == 3) a[i] = __umul64hi(a[i], wq) + w;                      // 64-bit mulhi
            if (OP == 4) a[i] = a[i] * w + p;                                  // 64-bit mullo (mad.lo.u64)
            if (OP == 5) { u64 q = __umul64hi(a[i], wq); a[i] = a[i] * w - q * p; }  // Shoup lazy modmul
            if (OP == 6) a[i] = a[i] + w + (a[i] >> 63);                       // 64-bit adds (ALU pipe)
            if (OP == 7) { u64 x = a[i], y = a[(i + 4) & 7]; u64 q = __umul64hi(y, wq); u64 t = y * w - q * p; u64 x2 = x >= 2 * p ? x - 2 * p : x; a[i] = x2 + t; a[(i + 4) & 7] = x2 - t + 2 * p; } // full butterfly (strict)
            if (OP == 8) { u64 x = a[i], y = a[(i + 4) & 7]; u64 q = __umul64hi(y, wq); u64 t = y * w - q * p; a[i] = x + t; a[(i + 4) & 7] = x - t + 2 * p; } // lazy butterfly
        }
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP> void run(const char *name, double ops_per_iter_thread) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int threads = 1024, blocks = sms * 2, iters = 4096;
    u64 *out; cudaMalloc(&out, (size_t)threads * blocks * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, 0x9e3779b97f4a7c15ull, 64);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); k<OP><<<blocks, threads>>>(out, 0x9e3779b97f4a7c15ull + r, iters); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double total = (double)threads * blocks * iters * 8.0 * ops_per_iter_thread;
    printf("%-28s %8.3f ms  %9.2f Gop/s  (%6.2f op/clk/SM at max clock %d MHz)\n", name, best, total / best * 1e-6,
           total / (best * 1e-3) / sms / (clk * 1e3), clk / 1000);
    cudaFree(out);
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    printf("device: %s, %d SMs\n", pr.name, pr.multiProcessorCount);
    run<0>("IMAD (32-bit)", 1); run<1>("IMAD.WIDE.U32", 1); run<2>("IMAD.HI.U32", 1);
    run<3>("mulhi64", 1); run<4>("mullo64+add", 1); run<5>("shoup_lazy modmul", 1);
    run<6>("add64 x2", 1); run<7>("strict butterfly (pair)", 0.5 * 2 / 2.0 * 1); run<8>("lazy butterfly (pair)", 0.5);
    return 0;
}
