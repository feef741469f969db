// b2_ntt_launch.cuh — host-side launcher template for the two-pass NTT with fused load / store functors.
// Included by every translation unit that launches a transform (b2_ntt.cu, b2_keyswitch.cu, b2_rescale.cu).
#pragma once
#include <algorithm>
#include <vector>
#include "b2_ctx.h"
#include "b2_ntt.cuh"

namespace b2 {

// ---- tiny transforms (N <= 2048): one CTA per limb, everything in shared memory -----------------
template <bool FWD, class Load, class Store>
__global__ void ntt_small_kernel(const __grid_constant__ NttLaunch P, Load ld, Store st) {
    extern __shared__ u64 sm[];
    const LimbGroup &grp = P.groups[blockIdx.y];
    const int c = P.c_begin + blockIdx.x;
    if (c >= (int)grp.count) return;
    const ModInfo M = P.mods[grp.table];
    const u64 off = grp.base + (u64)c * grp.stride;
    const int logn = P.logn, n = 1 << logn, half = n >> 1;
    const u64 p = M.p, p2 = p << 1;
    const ArithInt<false> ar(M);
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = ld(P.g_base + (int)blockIdx.y, c, off, (u32)i);
    __syncthreads();
    for (int layer = 0; layer < logn; layer++) {
        for (int i = threadIdx.x; i < half; i += blockDim.x) {
            if (FWD) {
                const int gp = logn - layer - 1, gap = 1 << gp;
                const int xi = ((i >> gp) << (gp + 1)) + (i & (gap - 1));
                Tw w = M.tw[(1 << layer) + (i >> gp)];
                u64 X = sm[xi], Y = sm[xi + gap];
                ar.bf_fwd(X, Y, w);
                sm[xi] = X; sm[xi + gap] = Y;
            } else {
                const int gp = layer, gap = 1 << gp;
                const int xi = ((i >> gp) << (gp + 1)) + (i & (gap - 1));
                u64 U = sm[xi], V = sm[xi + gap];
                if (layer == logn - 1) {
                    ar.bf_inv_last(U, V, layer);
                    sm[xi] = U; sm[xi + gap] = V;
                } else {
                    Tw w = M.itw[n - (n >> layer) + 1 + (i >> gp)];
                    ar.bf_inv(U, V, w, layer);
                    sm[xi] = U; sm[xi + gap] = V;
                }
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) st(P.g_base + (int)blockIdx.y, c, off, (u32)i, FWD ? csub(csub(sm[i], p2), p) : sm[i]);
}

// ---- pass dispatch ---------------------------------------------------------------------------------
template <int A, class Load, class Store>
static void launch_fwd_col(const NttLaunch &P, dim3 grid, cudaStream_t s, Load ld, Store st) {
    using Pl = FwdColPlan<A>;
    ntt_pass_kernel<A, true, true, true, false, Pl::LO0, Pl::NL0, Pl::LO1, Pl::NL1, Pl::LO2, Pl::NL2><<<grid, NTT_THREADS, 0, s>>>(P, ld, st);
}
template <int A, class Load, class Store>
static void launch_inv_col(const NttLaunch &P, dim3 grid, cudaStream_t s, Load ld, Store st) {
    using Pl = InvColPlan<A>;
    ntt_pass_kernel<A, true, false, false, true, Pl::LO0, Pl::NL0, Pl::LO1, Pl::NL1, Pl::LO2, Pl::NL2><<<grid, NTT_THREADS, 0, s>>>(P, ld, st);
}
template <class Load, class Store>
static void launch_fwd_row(const NttLaunch &P, dim3 grid, cudaStream_t s, Load ld, Store st) {
    ntt_pass_kernel<8, false, true, false, true, 4, 4, 0, 4, 0, 0><<<grid, NTT_THREADS, 0, s>>>(P, ld, st);
}
template <class Load, class Store>
static void launch_inv_row(const NttLaunch &P, dim3 grid, cudaStream_t s, Load ld, Store st) {
    ntt_pass_kernel<8, false, false, true, false, 0, 4, 4, 4, 0, 0><<<grid, NTT_THREADS, 0, s>>>(P, ld, st);
}

#define B2_DISPATCH_A(A_, FN, ...)                                     \
    switch (A_) {                                                      \
    case 4: FN<4>(__VA_ARGS__); break;                                 \
    case 5: FN<5>(__VA_ARGS__); break;                                 \
    case 6: FN<6>(__VA_ARGS__); break;                                 \
    case 7: FN<7>(__VA_ARGS__); break;                                 \
    case 8: FN<8>(__VA_ARGS__); break;                                 \
    default: FN<9>(__VA_ARGS__); break;                                \
    }

// L2 budget for the intermediate between the two passes: limbs are processed in waves small
// enough that pass 2 finds pass 1's output still resident in the 126 MB L2.
static constexpr size_t L2_WAVE_BYTES = 40u << 20;

// groups' base/stride address the intermediate buffer `mid` (raw values between the passes); the
// functors map (group, limb, x) to wherever the external operands live.
template <class Load, class Store>
static b2_status run_ntt(b2_ctx *ctx, bool forward, const std::vector<LimbGroup> &groups, u64 *mid, Load ld, Store st,
                         cudaStream_t stream) {
    const unsigned logn = ctx->logn;
    const size_t n = ctx->n;
    NttLaunch P;
    P.mods = ctx->d_mods;
    P.logn = (int)logn;
    for (size_t g0 = 0; g0 < groups.size(); g0 += NTT_MAX_GROUPS) {
        const size_t ng = std::min(groups.size() - g0, (size_t)NTT_MAX_GROUPS);
        u32 maxc = 0;
        P.g_base = (int)g0;
        for (size_t g = 0; g < ng; g++) maxc = std::max(maxc, groups[g0 + g].count);
        if (logn <= 11) {
            P.n_groups = (int)ng; P.chunk = 1; P.c_begin = 0; P.c_end = (int)maxc;
            for (size_t g = 0; g < ng; g++) P.groups[g] = groups[g0 + g];
            const int threads = (int)std::max<size_t>(32, std::min<size_t>(n / 2, 1024));
            dim3 grid(maxc, (unsigned)ng);
            if (forward) ntt_small_kernel<true><<<grid, threads, n * 8, stream>>>(P, ld, st);
            else ntt_small_kernel<false><<<grid, threads, n * 8, stream>>>(P, ld, st);
            count_launches(1);
            B2_CUDA(cudaGetLastError());
            continue;
        }
        const int A = (int)logn - 8;
        const unsigned tiles = 1u << (A - 4);
        // wave over the limb index so that ng * wave * N * 8 bytes stay L2-resident between passes
        size_t wave = std::max<size_t>(1, L2_WAVE_BYTES / (ng * n * 8));
        for (u32 cs = 0; cs < maxc; cs += (u32)wave) {
            const u32 ce = (u32)std::min<size_t>(maxc, (size_t)cs + wave);
            const u32 wmax = ce - cs;
            for (size_t g = 0; g < ng; g++) P.groups[g] = groups[g0 + g];
            P.c_begin = (int)cs; P.c_end = (int)ce;
            P.n_groups = (int)ng;
            // enough CTAs to fill the machine a few times over, but keep limbs of a modulus together
            const size_t want_ctas = 148 * 8;
            size_t chunk = std::max<size_t>(1, ((size_t)tiles * ng * wmax) / want_ctas);
            chunk = std::min<size_t>(chunk, 8);
            P.chunk = (int)chunk;
            dim3 grid(tiles, (unsigned)ng, (unsigned)((wmax + chunk - 1) / chunk));
            if (forward) {
                B2_DISPATCH_A(A, launch_fwd_col, P, grid, stream, ld, PlainStore{mid});
                launch_fwd_row(P, grid, stream, PlainLoad{mid}, st);
            } else {
                launch_inv_row(P, grid, stream, ld, PlainStore{mid});
                B2_DISPATCH_A(A, launch_inv_col, P, grid, stream, PlainLoad{mid}, st);
            }
            count_launches(2);
            B2_CUDA(cudaGetLastError());
        }
    }
    return B2_OK;
}


}  // namespace b2
