// b2_ntt.cu — launchers for the two-pass NTT (b2_ntt.cuh) and the C-ABI entry points
// b2_ntt_forward / b2_ntt_inverse (+ _ptrs, _host).  Reference seam: utils::ntt_transfer_to_rev /
// ntt_transfer_from_rev and their _batched forms (src/utils/ntt.h:132-144, src/fgk/ntt_grouped.cu:258-344,
// 597-685).
#include <algorithm>
#include <vector>
#include "b2_ctx.h"
#include "b2_ntt.cuh"

namespace b2 {

// ---- tiny transforms (N <= 2048): one CTA per limb, everything in shared memory -----------------
template <bool FWD>
__global__ void ntt_small_kernel(const __grid_constant__ NttLaunch P, const u64 *in, u64 *out) {
    extern __shared__ u64 sm[];
    const LimbGroup &grp = P.groups[blockIdx.y];
    const int c = blockIdx.x;
    if (c >= (int)grp.count) return;
    const ModInfo M = P.mods[grp.table];
    const u64 off = grp.base + (u64)c * grp.stride;
    const int logn = P.logn, n = 1 << logn, half = n >> 1;
    const u64 p = M.p, p2 = p << 1;
    const ArithInt<false> ar(M);
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = in[off + i];
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
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[off + i] = FWD ? csub(csub(sm[i], p2), p) : sm[i];
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

static b2_status run_groups(b2_ctx *ctx, bool forward, const u64 *in, u64 *out, std::vector<LimbGroup> &groups,
                            cudaStream_t stream) {
    const unsigned logn = ctx->logn;
    const size_t n = ctx->n;
    NttLaunch P;
    P.mods = ctx->d_mods;
    P.logn = (int)logn;
    for (size_t g0 = 0; g0 < groups.size(); g0 += NTT_MAX_GROUPS) {
        const size_t ng = std::min(groups.size() - g0, (size_t)NTT_MAX_GROUPS);
        u32 maxc = 0;
        for (size_t g = 0; g < ng; g++) maxc = std::max(maxc, groups[g0 + g].count);
        if (logn <= 11) {
            P.n_groups = (int)ng; P.chunk = 1;
            for (size_t g = 0; g < ng; g++) P.groups[g] = groups[g0 + g];
            const int threads = (int)std::max<size_t>(32, std::min<size_t>(n / 2, 1024));
            dim3 grid(maxc, (unsigned)ng);
            if (forward) ntt_small_kernel<true><<<grid, threads, n * 8, stream>>>(P, in, out);
            else ntt_small_kernel<false><<<grid, threads, n * 8, stream>>>(P, in, out);
            B2_CUDA(cudaGetLastError());
            continue;
        }
        const int A = (int)logn - 8;
        const unsigned tiles = 1u << (A - 4);
        // wave over the limb index so that ng * wave * N * 8 bytes stay L2-resident between passes
        size_t wave = std::max<size_t>(1, L2_WAVE_BYTES / (ng * n * 8));
        for (u32 cs = 0; cs < maxc; cs += (u32)wave) {
            u32 wmax = 0;
            for (size_t g = 0; g < ng; g++) {
                LimbGroup lg = groups[g0 + g];
                u32 cnt = lg.count > cs ? std::min<u32>(lg.count - cs, (u32)wave) : 0;
                lg.base += (u64)cs * lg.stride; lg.count = cnt;
                P.groups[g] = lg; wmax = std::max(wmax, cnt);
            }
            P.n_groups = (int)ng;
            // enough CTAs to fill the machine a few times over, but keep limbs of a modulus together
            const size_t want_ctas = 148 * 8;
            size_t chunk = std::max<size_t>(1, ((size_t)tiles * ng * wmax) / want_ctas);
            chunk = std::min<size_t>(chunk, 8);
            P.chunk = (int)chunk;
            dim3 grid(tiles, (unsigned)ng, (unsigned)((wmax + chunk - 1) / chunk));
            if (forward) {
                B2_DISPATCH_A(A, launch_fwd_col, P, grid, stream, PlainLoad{in}, PlainStore{out});
                launch_fwd_row(P, grid, stream, PlainLoad{out}, PlainStore{out});
            } else {
                launch_inv_row(P, grid, stream, PlainLoad{in}, PlainStore{out});
                B2_DISPATCH_A(A, launch_inv_col, P, grid, stream, PlainLoad{out}, PlainStore{out});
            }
            B2_CUDA(cudaGetLastError());
        }
    }
    return B2_OK;
}

static b2_status check_ntt_args(b2_ctx *ctx, size_t batch, size_t pcount, size_t ncomp, int mode, size_t table_base, size_t decomp) {
    B2_REQUIRE(ctx, "null ctx");
    B2_REQUIRE(mode >= 0 && mode <= 2, "unknown table mode");
    B2_REQUIRE(batch * pcount * ncomp > 0, "empty operand");   // callers filter empties; reference would loop zero times
    if (mode == B2_TABLES_COMPONENTWISE) B2_REQUIRE(table_base + ncomp <= ctx->K, "table_base + ncomp exceeds modulus count");
    if (mode == B2_TABLES_KS_SET_PRODUCTS) B2_REQUIRE(decomp < ctx->K && pcount <= decomp + 1, "bad decomp for KS_SET_PRODUCTS");
    if (mode == B2_TABLES_KS_SKIP_FINALS) B2_REQUIRE(decomp < ctx->K && ncomp <= decomp + 1, "bad decomp for KS_SKIP_FINALS");
    return B2_OK;
}

b2_status ntt_launch(b2_ctx *ctx, bool forward, const uint64_t *in, uint64_t *out, size_t batch, size_t pcount,
                     size_t ncomp, int mode, size_t table_base, size_t decomp, cudaStream_t stream) {
    if (batch * pcount * ncomp == 0) return B2_OK;
    b2_status st = check_ntt_args(ctx, batch, pcount, ncomp, mode, table_base, decomp);
    if (st != B2_OK) return st;
    B2_REQUIRE(in && out, "null data pointer");
    B2_CUDA(cudaSetDevice(ctx->device));
    const size_t n = ctx->n, K = ctx->K;
    std::vector<LimbGroup> groups;
    if (mode == B2_TABLES_KS_SET_PRODUCTS) {
        for (size_t b = 0; b < batch; b++)
            for (size_t k = 0; k < pcount; k++)
                groups.push_back(LimbGroup{(u32)(k == decomp ? K - 1 : k), (u32)ncomp, (u64)((b * pcount + k) * ncomp * n), (u64)n});
    } else {
        for (size_t j = 0; j < ncomp; j++) {
            u32 table = mode == B2_TABLES_COMPONENTWISE ? (u32)(table_base + j) : (u32)(j == decomp ? K - 1 : j);
            groups.push_back(LimbGroup{table, (u32)(batch * pcount), (u64)(j * n), (u64)(ncomp * n)});
        }
    }
    return run_groups(ctx, forward, (const u64 *)in, (u64 *)out, groups, stream);
}

}  // namespace b2

using namespace b2;

extern "C" {

b2_status b2_ntt_forward(b2_ctx *ctx, const uint64_t *in, uint64_t *out, size_t batch, size_t pcount, size_t ncomp,
                         int mode, size_t table_base, size_t decomp, void *stream) {
    return ntt_launch(ctx, true, in, out, batch, pcount, ncomp, mode, table_base, decomp, (cudaStream_t)stream);
}
b2_status b2_ntt_inverse(b2_ctx *ctx, const uint64_t *in, uint64_t *out, size_t batch, size_t pcount, size_t ncomp,
                         int mode, size_t table_base, size_t decomp, void *stream) {
    return ntt_launch(ctx, false, in, out, batch, pcount, ncomp, mode, table_base, decomp, (cudaStream_t)stream);
}

static b2_status ntt_ptrs(b2_ctx *ctx, bool fwd, const uint64_t *const *in, uint64_t *const *out, size_t batch, size_t pcount,
                          size_t ncomp, int mode, size_t table_base, size_t decomp, void *stream) {
    B2_REQUIRE(ctx && (batch == 0 || (in && out)), "null argument");
    for (size_t b = 0; b < batch; b++) {
        b2_status st = ntt_launch(ctx, fwd, in[b], out[b], 1, pcount, ncomp, mode, table_base, decomp, (cudaStream_t)stream);
        if (st != B2_OK) return st;
    }
    return B2_OK;
}
b2_status b2_ntt_forward_ptrs(b2_ctx *ctx, const uint64_t *const *in, uint64_t *const *out, size_t batch, size_t pcount,
                              size_t ncomp, int mode, size_t table_base, size_t decomp, void *stream) {
    return ntt_ptrs(ctx, true, in, out, batch, pcount, ncomp, mode, table_base, decomp, stream);
}
b2_status b2_ntt_inverse_ptrs(b2_ctx *ctx, const uint64_t *const *in, uint64_t *const *out, size_t batch, size_t pcount,
                              size_t ncomp, int mode, size_t table_base, size_t decomp, void *stream) {
    return ntt_ptrs(ctx, false, in, out, batch, pcount, ncomp, mode, table_base, decomp, stream);
}

static b2_status ntt_host(b2_ctx *ctx, bool fwd, const uint64_t *in, uint64_t *out, size_t batch, size_t pcount, size_t ncomp,
                          int mode, size_t table_base, size_t decomp, void *stream) {
    B2_REQUIRE(ctx && in && out, "null argument");
    const size_t bytes = batch * pcount * ncomp * ctx->n * 8;
    if (bytes == 0) return B2_OK;
    B2_CUDA(cudaSetDevice(ctx->device));
    b2_status st = ensure_stage(ctx, bytes);
    if (st != B2_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    B2_CUDA(cudaMemcpyAsync(ctx->stage, in, bytes, cudaMemcpyHostToDevice, s));
    st = ntt_launch(ctx, fwd, (const uint64_t *)ctx->stage, (uint64_t *)ctx->stage, batch, pcount, ncomp, mode, table_base, decomp, s);
    if (st != B2_OK) return st;
    B2_CUDA(cudaMemcpyAsync(out, ctx->stage, bytes, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    return B2_OK;
}
b2_status b2_ntt_forward_host(b2_ctx *ctx, const uint64_t *in, uint64_t *out, size_t batch, size_t pcount, size_t ncomp,
                              int mode, size_t table_base, size_t decomp, void *stream) {
    return ntt_host(ctx, true, in, out, batch, pcount, ncomp, mode, table_base, decomp, stream);
}
b2_status b2_ntt_inverse_host(b2_ctx *ctx, const uint64_t *in, uint64_t *out, size_t batch, size_t pcount, size_t ncomp,
                              int mode, size_t table_base, size_t decomp, void *stream) {
    return ntt_host(ctx, false, in, out, batch, pcount, ncomp, mode, table_base, decomp, stream);
}

}  // extern "C"
