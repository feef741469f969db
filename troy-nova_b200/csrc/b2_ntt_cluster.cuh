// b2_ntt_cluster.cuh — low-latency N = 2^16 transform for SMALL launches: one thread-block cluster per limb, the
// intermediate between the two passes lives in distributed shared memory (DSMEM) instead of L2.
//
// The persistent kernel (b2_ntt.cuh) is built for throughput: a launch over a few limbs (keyswitch: INTT of the target,
// INTT of the special limb, the mod-down transform; every stage of the limb-sharded path) is bound by its serial chain
// queue claim -> tile from HBM -> first pass -> completion counter -> second-pass tile from L2 -> store, ~19 us even
// for one limb.  Here a limb is owned by a cluster of 16 CTAs (non-portable cluster size, 3 CTAs per SM):
//   CTA r stages column tile r (32 KiB, cp.async), runs the first pass on it (same pass_body as the persistent kernel)
//   and stores every intermediate word straight into the receive buffer of the CTA that owns it in the second pass
//   (st.shared::cluster, already in the stage layout the second pass reads); one cluster barrier; CTA r runs the
//   second pass on row tile r out of its receive buffer.  No work queue, no global counters, no L2 round trip.
// Same load / store functor interface, same arithmetic back ends, bit-identical results.
// The receive buffer IS the stage buffer (64 KiB per CTA -> 3 CTAs per SM, so the 20-limb transforms of a C3 keyswitch are
// one wave of 320 CTAs instead of two): a CTA arrives on a split-phase cluster barrier as soon as all its threads have pulled
// the staged tile into registers, and every CTA waits on that barrier before its first remote store.
#pragma once
#include "b2_ntt.cuh"

namespace b2 {

constexpr int CL_SIZE = 16;                               // CTAs per cluster = tiles per limb at N = 2^16
constexpr int CL_SMEM_BYTES = 2 * NTT_TILE * 8;           // stage (= receive) | exchange = 64 KiB (3 CTAs per SM)

__device__ __forceinline__ u32 cluster_ctarank() { u32 r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// first-pass store: word at limb position x -> receive buffer of the CTA whose second-pass tile contains x
//   forward (second pass = row tiles of 16 rows):    tile = x >> 12,          local index = x & 4095
//   inverse (second pass = column tiles of 16 cols): tile = (x & 255) >> 4,  local index = (row << 4) | (col & 15)
template <bool FWD>
struct DsmemStore {
    u32 recv;      // shared-window address of this CTA's receive buffer (identical offset in every CTA of the cluster)
    struct Prep {};
    // called once per thread, by every thread of the CTA, right before its stores: all CTAs of the cluster have consumed their stage
    __device__ __forceinline__ Prep prep(int, int, u64) const { asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory"); return Prep{}; }
    __device__ __forceinline__ void operator()(const Prep &, u32 x, u64 v) const {
        const u32 tile = FWD ? x >> 12 : (x & 255u) >> 4;
        const u32 idx = FWD ? x & 4095u : ((x >> 8) << 4) | (x & 15u);
        u32 remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(remote) : "r"(recv + 8u * (u32)swz16((int)idx)), "r"(tile));
        asm volatile("st.shared::cluster.u64 [%0], %1;\n" ::"r"(remote), "l"(v) : "memory");
    }
    __device__ __forceinline__ void st2(const Prep &pc, u32 x, u64 v0, u64 v1) const { (*this)(pc, x, v0); (*this)(pc, x + 1, v1); }
};

template <class Next>
struct ClusterHooks {
    NoHook mid;
    bool first;        // first pass: its stage buffer doubles as the receive buffer of the second pass
    Next next;         // second pass: stages the first-pass tile of the cluster's next limb
    __device__ __forceinline__ void sync_stage() const { cp_async_wait_all(); __syncthreads(); }
    __device__ __forceinline__ void pre_exchange() const {}
    // every thread of the CTA has read the staged tile.  First pass: the buffer may now receive the other CTAs' intermediates.
    // Second pass: the buffer is free for good (every remote store landed before the cluster barrier that precedes the pass).
    __device__ __forceinline__ void stage_free() const {
        if (first) asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        else next();
    }
};

// The grid is min(limbs, co-resident clusters) clusters: cluster k transforms limbs k, k + clusters, ... and stages the next limb's
// tile under the second pass of the current one.
template <bool FWD, class Load, class Store>
__global__ void __launch_bounds__(NTT_THREADS, 3) ntt_cluster_kernel(const __grid_constant__ NttJob J, Load ld, Store st) {
    extern __shared__ __align__(128) u64 smem[];
    u64 *const stage = smem, *const xch = smem + NTT_TILE, *const recv = stage;
    const u32 r = cluster_ctarank();
    const u32 n_cl = gridDim.x / CL_SIZE;
    u32 limb = blockIdx.x / CL_SIZE;
    // first-pass tile from the operand: forward = column tile r, inverse = row tile r
    const auto stage_limb = [&](u32 lb) {
        const u32 c = lb / (u32)J.n_groups, g = lb - c * (u32)J.n_groups;
        stage_tile<8, FWD>(stage, ld.src(J.g_base + (int)g, (int)c, limb_offset(J, J.groups[g], c)), (int)r);
        cp_async_commit();
    };
    stage_limb(limb);
    cluster_sync();            // every CTA of the cluster is resident before anyone writes into its shared memory
    const DsmemStore<FWD> xst{(u32)__cvta_generic_to_shared(recv)};
    const PlainLoad none{nullptr};
    for (; limb < J.n_limbs; limb += n_cl) {
        const u32 c = limb / (u32)J.n_groups, g = limb - c * (u32)J.n_groups;
        const LimbGroup &grp = J.groups[g];
        const u64 off = limb_offset(J, grp, c);
        const ModInfo &M = J.mods[group_table(grp.table)];
        const u32 path = group_path(grp.table, FWD);
        const int ga = J.g_base + (int)g;
        const u32 nxt = limb + n_cl;
        const auto next = [&]() { if (nxt < J.n_limbs) stage_limb(nxt); };
        ClusterHooks<decltype(next)> hk{NoHook(), true, next};
        if (FWD) {
            using Pl = FwdColPlan<8>;
            pass_dispatch<8, true, true, true, false, Pl::LO0, Pl::NL0, Pl::LO1, Pl::NL1, Pl::LO2, Pl::NL2>(path, stage, xch, M, (int)r, ga, (int)c, off, 16, ld, xst, hk);
        } else {
            pass_dispatch<8, false, false, true, false, 0, 4, 4, 4, 0, 0>(path, stage, xch, M, (int)r, ga, (int)c, off, 16, ld, xst, hk);
        }
        cluster_sync();        // all intermediates have landed in the receive buffers
        hk.first = false;
        if (FWD) {
            pass_dispatch<8, false, true, false, true, 4, 4, 0, 4, 0, 0>(path, recv, xch, M, (int)r, ga, (int)c, off, 16, none, st, hk);
        } else {
            using Pl = InvColPlan<8>;
            pass_dispatch<8, true, false, false, true, Pl::LO0, Pl::NL0, Pl::LO1, Pl::NL1, Pl::LO2, Pl::NL2>(path, recv, xch, M, (int)r, ga, (int)c, off, 16, none, st, hk);
        }
    }
}

}  // namespace b2
