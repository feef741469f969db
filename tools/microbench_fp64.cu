// microbench_fp64.cu — can the FP64 pipe carry the modular multiply for moduli below 2^50?
// (1) DFMA/DADD/DMUL issue rates, (2) throughput + exactness of a double-double modmul butterfly,
// (3) IMAD-path and FP64-path butterflies running side by side on the same SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64; typedef unsigned int u32;

__device__ __forceinline__ double modmul_dd(double x, double w, double p, double pinv) {
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: (b + M) - M == rint(b) for |b| < 2^51
    double h = x * w;
    double l = fma(x, w, -h);
    double c = (h * pinv + MAGIC) - MAGIC;
    double d = fma(-c, p, h);
    return d + l;                       // == x*w - c*p exactly, |.| <= ~p/2 + eps
}

template <int OP> __global__ void k(double *out, double seed, int iters, u64 pi, int mixmask) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed * (threadIdx.x + 1 + i * 977);
    const double w = 977.0 + seed, p = (double)pi, pinv = 1.0 / p;
    u64 ia[8];
#pragma unroll
    for (int i = 0; i < 8; i++) ia[i] = (u64)(threadIdx.x + 1 + i * 977) * 0x9e3779b97f4a7c15ull;
    const u64 iw = 0x9e3779b97f4a7c15ull | 1, iwq = ~iw, ip = pi;
    const bool int_warp = ((threadIdx.x >> 5) & mixmask) != 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) a[i] = fma(a[i], w, p);
            if (OP == 1) a[i] = a[i] + w;
            if (OP == 2) a[i] = a[i] * w;
            if (OP == 3) a[i] = modmul_dd(a[i], w, p, pinv);
            if (OP == 4 && i < 4) { double t = modmul_dd(a[i + 4], w, p, pinv); double x = a[i]; a[i] = x + t; a[i + 4] = x - t; }
            if (OP == 5 && i < 4) {   // mixed: some warps integer, some FP64
                if (int_warp) { u64 x = ia[i], y = ia[i + 4]; u64 q = __umul64hi(y, iwq); u64 t = y * iw - q * ip; ia[i] = x + t; ia[i + 4] = x - t + 2 * ip; }
                else { double t = modmul_dd(a[i + 4], w, p, pinv); double x = a[i]; a[i] = x + t; a[i + 4] = x - t; }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i] + (double)ia[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP> void run(const char *name, double ops, int mixmask = 0) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int threads = 1024, blocks = sms * 2, iters = 2048;
    double *out; cudaMalloc(&out, (size_t)threads * blocks * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, 1.0000001, 64, 1099511480321ull, mixmask); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); k<OP><<<blocks, threads>>>(out, 1.0000001 + r, iters, 1099511480321ull, mixmask); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double total = (double)threads * blocks * iters * ops;
    printf("%-36s %8.3f ms  %9.2f Gop/s  (%6.2f op/clk/SM at %d MHz)\n", name, best, total / best * 1e-6, total / (best * 1e-3) / sms / (clk * 1e3), clk / 1000);
    cudaFree(out);
}

// exactness: compare modmul_dd against 128-bit integer arithmetic on random signed inputs
__global__ void check(u64 p, u64 *bad, u64 seed) {
    u64 s = seed + blockIdx.x * blockDim.x + threadIdx.x;
    const double pd = (double)p, pinv = 1.0 / pd;
    for (int it = 0; it < 4096; it++) {
        s = s * 6364136223846793005ull + 1442695040888963407ull; u64 xr = (s >> 12) % (40 * p);      // |x| < 40 p
        s = s * 6364136223846793005ull + 1442695040888963407ull; u64 w = (s >> 12) % p;
        bool neg = s & 1;
        double x = neg ? -(double)xr : (double)xr;
        double g = modmul_dd(x, (double)w, pd, pinv);
        // exact: x*w mod p
        unsigned __int128 prod = (unsigned __int128)xr * w; u64 r = (u64)(prod % p);
        long long gi = (long long)g; long long expect = neg ? (r ? (long long)(p - r) : 0) : (long long)r;
        long long diff = gi - expect;               // must be a multiple of p with |g| <= p
        bool ok = (g == (double)gi) && (diff == 0 || diff == -(long long)p || diff == (long long)p) && (fabs(g) <= 0.5 * pd + 2.0);
        if (!ok) atomicAdd(bad, 1ull);
    }
}

int main() {
    run<0>("DFMA", 8); run<1>("DADD", 8); run<2>("DMUL", 8);
    run<3>("modmul_dd (double-double)", 8); run<4>("fp64 butterfly", 4);
    run<5>("mixed: all warps fp64", 4, 0); run<5>("mixed: 1/2 warps int, 1/2 fp64", 4, 1); run<5>("mixed: 1/4 int (mask 3 -> 3/4 int)", 4, 3);
    u64 *bad; cudaMalloc(&bad, 8);
    u64 primes[4] = {1099511480321ull /*40-bit*/, 1125899906826241ull /*50-bit*/, 36028797017456641ull /*55-bit?*/, 1032193ull};
    for (int i = 0; i < 4; i++) {
        cudaMemset(bad, 0, 8); check<<<296, 256>>>(primes[i], bad, 12345 + i); u64 h; cudaMemcpy(&h, bad, 8, cudaMemcpyDeviceToHost);
        printf("exactness p=%llu (%d bits): %llu bad of %d\n", primes[i], 64 - __builtin_clzll(primes[i]), h, 296 * 256 * 4096);
    }
    return 0;
}
