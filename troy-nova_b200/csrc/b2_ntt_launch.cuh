// b2_ntt_launch.cuh — host-side launcher template for the persistent NTT with fused load / store functors.
// Included by every translation unit that launches a transform (b2_ntt.cu, b2_keyswitch.cu, b2_rescale.cu).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "b2_ctx.h"
#include "b2_ntt.cuh"
#include "b2_ntt_cluster.cuh"

namespace b2 {

// ---- tiny transforms (N <= 2048): one CTA per limb, everything in shared memory -----------------
template <bool FWD, class Load, class Store>
__global__ void ntt_small_kernel(const __grid_constant__ NttJob P, Load ld, Store st) {
    extern __shared__ u64 sm[];
    const LimbGroup &grp = P.groups[blockIdx.y];
    const int c = blockIdx.x;
    if (c >= (int)grp.count) return;
    const ModInfo M = P.mods[group_table(grp.table)];
    const u64 off = limb_offset(P, grp, (u32)c);
    const int logn = P.logn, n = 1 << logn, half = n >> 1;
    const int g = P.g_base + (int)blockIdx.y;
    const u64 p = M.p, p2 = p << 1;
    const ArithInt<false> ar(M);
    const u64 *src = ld.src(g, c, off);
    const auto lp = ld.prep(g, c);
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = ld.xf(lp, (u32)i, src[i]);
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
                    Tw w = M.itw[n - (n >> layer) + (i >> gp)];
                    ar.bf_inv(U, V, w, layer);
                    sm[xi] = U; sm[xi + gap] = V;
                }
            }
        }
        __syncthreads();
    }
    const auto sp = st.prep(g, c, off);
    for (int i = threadIdx.x; i < n; i += blockDim.x) st(sp, (u32)i, FWD ? csub(csub(sm[i], p2), p) : sm[i]);
}

// second-pass items trail first-pass items by this many limbs: at N = 2^16 the in-flight intermediate is
// NTT_LAG x 512 KiB = 24 MiB, far inside the 126 MB L2; 888 items in flight (one being computed + one prefetched per
// resident CTA) are ~28 limbs, so a lag of 32 left 29 % of the second-pass items waiting on their dependency (ncu, r01c).
static constexpr u32 NTT_LAG = 48;

#ifndef B2_NTT_OCC
#define B2_NTT_OCC 3
#endif
constexpr int NTT_OCC = B2_NTT_OCC;     // resident CTAs per SM (register allocation is held to 65536 / (256 * NTT_OCC))

constexpr int B2_MAX_DEVICES = 64;      // function attributes are per device: one flag per (instantiation, device)
template <int A, bool FWD, class Load, class Store>
static cudaError_t launch_persistent(const NttJob &J, int device, unsigned grid, cudaStream_t s, Load ld, Store st) {
    static bool configured[B2_MAX_DEVICES] = {};
    if (!configured[device & (B2_MAX_DEVICES - 1)]) {
        cudaError_t e = cudaFuncSetAttribute(ntt_persistent_kernel<A, FWD, NTT_OCC, Load, Store>, cudaFuncAttributeMaxDynamicSharedMemorySize, NTT_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured[device & (B2_MAX_DEVICES - 1)] = true;
    }
    ntt_persistent_kernel<A, FWD, NTT_OCC, Load, Store><<<grid, NTT_THREADS, NTT_SMEM_BYTES, s>>>(J, ld, st);
    return cudaGetLastError();
}

// N = 2^16, few limbs: one 16-CTA cluster per limb, intermediate in distributed shared memory (b2_ntt_cluster.cuh)
template <bool FWD, class Load, class Store>
static cudaError_t launch_cluster(const NttJob &J, int device, cudaStream_t s, Load ld, Store st) {
    static int resident[B2_MAX_DEVICES] = {};       // co-resident clusters of this instantiation on the device (0: not configured yet)
    int &res = resident[device & (B2_MAX_DEVICES - 1)];
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(J.n_limbs * CL_SIZE, 1, 1);
    cfg.blockDim = dim3(NTT_THREADS, 1, 1);
    cfg.dynamicSmemBytes = CL_SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL_SIZE; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (!res) {
        cudaError_t e = cudaFuncSetAttribute(ntt_cluster_kernel<FWD, Load, Store>, cudaFuncAttributeMaxDynamicSharedMemorySize, CL_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ntt_cluster_kernel<FWD, Load, Store>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
        int n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, ntt_cluster_kernel<FWD, Load, Store>, &cfg);
        if (e != cudaSuccess) return e;
        if (n < 1) return cudaErrorInvalidConfiguration;
        if (const char *cap = getenv("B2_NTT_CLUSTERS")) n = std::max(1, std::min(n, atoi(cap)));
        res = n;
    }
    cfg.gridDim.x = std::min<unsigned>(J.n_limbs, (unsigned)res) * CL_SIZE;      // persistent above that: cluster k takes limbs k, k + res, ...
    return cudaLaunchKernelEx(&cfg, ntt_cluster_kernel<FWD, Load, Store>, J, ld, st);
}
// launches of at most this many limbs take the cluster kernel (latency-bound regime); B2_NTT_CLUSTER_MAX overrides (0 = never)
static inline unsigned cluster_max_limbs() {
    static const unsigned v = [] { const char *e = getenv("B2_NTT_CLUSTER_MAX"); return e ? (unsigned)atoi(e) : 32u; }();
    return v;
}

// groups' base/stride address the intermediate buffer `mid` (raw values between the passes); the
// functors map (group, limb) to wherever the external operands live.  All groups must have the same count.
// The direction is a template parameter so that a call site instantiates only the kernels it launches.
struct NttScatter {
    const u64 *const *out_tab = nullptr; u32 pc = 0; u64 pstride = 0;   // see NttJob::out_tab
    unsigned max_ctas = 0;      // 0: fill the machine; else cap the persistent grid (a launch that shares the SMs with another kernel)
    unsigned *sync = nullptr;   // work queue / counters to use instead of the scratch record's (concurrent launches need their own)
};
template <bool FWD, class Load, class Store>
static b2_status run_ntt_dir(b2_ctx *ctx, b2_scratch *sc, const std::vector<LimbGroup> &groups, u64 *mid, Load ld, Store st, cudaStream_t stream,
                             const NttScatter &scatter = NttScatter()) {
    const unsigned logn = ctx->logn;
    const size_t n = ctx->n;
    if (groups.empty() || groups[0].count == 0) return B2_OK;
    NttJob P;
    P.mods = ctx->d_mods;
    P.mid = mid;
    P.out_tab = scatter.out_tab; P.tab_pc = scatter.pc; P.tab_pstride = scatter.pstride; P.dbg = 0;
    P.logn = (int)logn;
    P.count = groups[0].count;
    for (const LimbGroup &g : groups) B2_REQUIRE(g.count == P.count, "internal: limb groups of one transform must have equal counts");
    for (size_t g0 = 0; g0 < groups.size(); g0 += NTT_MAX_GROUPS) {
        const size_t ng = std::min(groups.size() - g0, (size_t)NTT_MAX_GROUPS);
        P.g_base = (int)g0;
        P.n_groups = (int)ng;
        for (size_t g = 0; g < ng; g++) {
            P.groups[g] = groups[g0 + g];
            const ModInfo &M = ctx->h_mods[groups[g0 + g].table];
            P.groups[g].table = tag_table(groups[g0 + g].table, M.use_f64 ? PATH_F64 : (M.fwd_lazy ? PATH_INT_LAZY : PATH_INT),
                                          M.use_f64 ? PATH_F64 : (M.inv_lazy ? PATH_INT_LAZY : PATH_INT));
        }
        if (logn <= 11) {
            const int threads = (int)std::max<size_t>(32, std::min<size_t>(n / 2, 1024));
            dim3 grid(P.count, (unsigned)ng);
            ntt_small_kernel<FWD><<<grid, threads, n * 8, stream>>>(P, ld, st);
            count_launches(1);
            B2_CUDA(cudaGetLastError());
            continue;
        }
        const int A = (int)logn - 8;
        P.tile_shift = logn - NTT_TB;
        P.n_limbs = (u32)(ng * P.count);
        // 16-CTA clusters are a non-portable size: if this device / partition cannot co-schedule one (the launch is refused
        // synchronously), the persistent kernel below serves small launches too
        static bool cluster_refused[B2_MAX_DEVICES] = {};
        if (logn == 16 && P.n_limbs <= cluster_max_limbs() && !cluster_refused[ctx->device & (B2_MAX_DEVICES - 1)] && !scatter.max_ctas) {
            P.sync = nullptr; P.lag = 0; P.total_items = 0;
            cudaError_t e = launch_cluster<FWD>(P, ctx->device, stream, ld, st);
            if (e == cudaSuccess) { count_launches(1); continue; }
            cudaGetLastError();
            cluster_refused[ctx->device & (B2_MAX_DEVICES - 1)] = true;
        }
        P.lag = std::min<u32>(NTT_LAG, P.n_limbs);
        P.total_items = (P.n_limbs << P.tile_shift) * 2;
        if (scatter.sync) P.sync = scatter.sync;
        else {
            b2_status s = ensure_sync(sc, P.n_limbs, stream);
            if (s != B2_OK) return s;
            P.sync = sc->d_sync;
        }
        unsigned grid = std::min<unsigned>(P.total_items, (unsigned)NTT_OCC * (unsigned)ctx->sm_count);
        if (scatter.max_ctas) grid = std::min(grid, scatter.max_ctas);
        cudaError_t e;
        switch (A) {
        case 4: e = launch_persistent<4, FWD>(P, ctx->device, grid, stream, ld, st); break;
        case 5: e = launch_persistent<5, FWD>(P, ctx->device, grid, stream, ld, st); break;
        case 6: e = launch_persistent<6, FWD>(P, ctx->device, grid, stream, ld, st); break;
        case 7: e = launch_persistent<7, FWD>(P, ctx->device, grid, stream, ld, st); break;
        case 8: e = launch_persistent<8, FWD>(P, ctx->device, grid, stream, ld, st); break;
        default: e = launch_persistent<9, FWD>(P, ctx->device, grid, stream, ld, st); break;
        }
        count_launches(1);
        B2_CUDA(e);
    }
    return B2_OK;
}
// run-time direction (the plain transforms of b2_ntt.cu)
template <class Load, class Store>
static b2_status run_ntt(b2_ctx *ctx, b2_scratch *sc, bool forward, const std::vector<LimbGroup> &groups, u64 *mid, Load ld, Store st, cudaStream_t stream,
                         const NttScatter &scatter = NttScatter()) {
    return forward ? run_ntt_dir<true>(ctx, sc, groups, mid, ld, st, stream, scatter) : run_ntt_dir<false>(ctx, sc, groups, mid, ld, st, stream, scatter);
}


}  // namespace b2
