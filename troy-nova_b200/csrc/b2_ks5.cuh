// b2_ks5.cuh — keyswitch steps 2 + 3 as ONE persistent launch at N = 2^16 (K2 + K3 of SURVEY.md §8a; reference:
// fgk::switch_key::set_accumulate + the (L+1) x L forward transforms + fgk::switch_key::accumulate_products,
// src/evaluator_keyswitching_core.cu:827-919, src/fgk/switch_key.cu:6-132).
//
// Round 1 ran them as two launches — a 420-limb transform writing temp[(L+1)][L][N] (210 MiB) and an inner-product kernel
// reading it back with the 440 MB key — and the ncu launch list showed the price: 1002 MB of DRAM traffic per keyswitch
// against 493 MB compulsory (profiles/r01n_ks_launches.csv).  Here the transforms and the inner product share one work
// queue on the warp-specialised TMA structure of b2_ntt5.cuh, ordered by OUTPUT MODULUS g:
//     step s:   first passes of group s  |  second passes of group s-2  |  inner product of group s-4
// and temp is a RING of KS5_RING groups (6 x 10 MiB at C3) that lives in L2: a group's transformed digits are consumed by
// its inner product two steps after they were written and the slot is overwritten two steps later, so the lines are
// re-dirtied in L2 instead of being written back.  HBM sees the coefficient-form target once, the key once and prod.
//   P1(g, j, tile)  column pass of digit j re-reduced mod m_g            needs  ipdone[g - RING] == IPI   (slot free)
//   P2(g, j, tile)  row pass -> canonical temp[slot][j]                  needs  done1[g][j] == 16
//   IP(g, q)        prod[k][g][512 q ..] = sum_j temp[slot][j] o key_j[k][g]   needs  gdone[g] == 16 C
// One dependency per item, every dependency points backwards in the queue, the producer publishes everything it holds before
// it spins: no cycles.  Results are the same canonical residues as the two-launch path (bit-exact, tests/test_keyswitch_gpu.py).
#pragma once
#include "b2_ntt5.cuh"

namespace b2 {

// The claimed-but-unfinished window of the machine is ~1350 items (444 CTAs x 2 stages + 1 claim ahead) = 1.8 steps of
// 2 x 320 + 128 items, so every dependency is placed at least two steps behind its consumer: second passes trail first
// passes by KS5_LAG1 steps, inner products trail second passes by KS5_LAG2, and a ring slot is reused KS5_RING groups later
// (6 x 10 MiB at C3, inside the 126 MB L2; the key is streamed with evict-first loads).
constexpr int KS5_LAG1 = 2, KS5_LAG2 = 2;
constexpr int KS5_RING = 6;
constexpr int KS5_IPI = 128;              // inner-product items per output modulus: 512 coefficients each, 2 per thread
constexpr int KS5_MAX_GROUPS = 40;

struct Ks5Job {
    const ModInfo *mods;
    u32 *sync;                 // [0] queue head, [1] finished CTAs, then done1[G*C], gdone[G], ipdone[G]
    const u64 *const *keys;    // device array of C digit pointers, each [2][key_rows][N]
    u64 *ring;                 // temp ring: [KS5_RING][C][N]
    u64 *prod;                 // [2][G][N]
    u32 G, C;                  // output moduli (decomp + 1), digits (decomp)
    u32 K, key_rows;           // key-level modulus count; rows per key polynomial
    u32 total_items;
    u32 f64_lazy;              // FP64 moduli take lazy mod-up inputs (b2_keyswitch.cu: ModUpLoad)
    u32 table[KS5_MAX_GROUPS];     // tagged table index of output modulus g (tag_table)
};

struct Ks5Item { u32 phase, g, j, tile; };      // phase 2: tile = inner-product item q
// queue order: step s = 0 .. G-1+LAG1+LAG2 :  P1(s) [C*16 items, digit-major]  P2(s-LAG1) [C*16]  IP(s-LAG1-LAG2) [IPI]
__device__ __forceinline__ Ks5Item ks5_decode(const Ks5Job &J, u32 id) {
    const u32 a = J.C * 16;
    Ks5Item r = {0, 0, 0, 0};
    for (u32 s = 0; s < J.G + KS5_LAG1 + KS5_LAG2; s++) {
        const u32 n1 = s < J.G ? a : 0, n2 = (s >= (u32)KS5_LAG1 && s < J.G + KS5_LAG1) ? a : 0, n3 = s >= (u32)(KS5_LAG1 + KS5_LAG2) ? (u32)KS5_IPI : 0;
        if (id < n1) { r.phase = 0; r.g = s; r.j = id >> 4; r.tile = id & 15; return r; }
        id -= n1;
        if (id < n2) { r.phase = 1; r.g = s - KS5_LAG1; r.j = id >> 4; r.tile = id & 15; return r; }
        id -= n2;
        if (id < n3) { r.phase = 2; r.g = s - KS5_LAG1 - KS5_LAG2; r.tile = id; return r; }
        id -= n3;
    }
    return r;
}

// mod-up re-reduction of a coefficient of digit j modulo the output modulus (src/fgk/switch_key.cu:22-26)
struct Ks5Reduce {
    u64 m, ratio; bool reduce;
    __device__ __forceinline__ u64 operator()(u64 v) const { return reduce ? barrett64(v, m, ratio) : v; }
};

// inner product of one output modulus over 512 coefficients: thread = one 16-byte chunk (2 coefficients), all digits
__device__ __forceinline__ void ks5_inner_product(const Ks5Job &J, const Ks5Item &it, const ModInfo &M, u32 slot, u32 ki) {
    const u32 n = 1u << N5_LOGN;
    const u32 x = it.tile * 512 + 2 * threadIdx.x;
    const u64 *tp = J.ring + ((u64)slot * J.C) * n + x;
    const u32 krow = J.key_rows == J.K ? ki : it.g;
    const u64 koff = (u64)krow * n + x, kpoly = (u64)J.key_rows * n;
    ulonglong2 o0, o1;
    if (M.use_f64) {
        // p < 2^41: exact double-double products reduced to (-p/2, p/2); twenty of them stay far below 2^52
        const double pd = M.pd, pinv = M.pinv;
        double a0 = 0, a1 = 0, b0 = 0, b1 = 0;
#pragma unroll 4
        for (u32 j = 0; j < J.C; j++) {
            const ulonglong2 d = __ldcg(reinterpret_cast<const ulonglong2 *>(tp + (u64)j * n));          // L2: written moments ago
            const u64 *kp = J.keys[j] + koff;
            const ulonglong2 k0 = __ldcs(reinterpret_cast<const ulonglong2 *>(kp)), k1 = __ldcs(reinterpret_cast<const ulonglong2 *>(kp + kpoly));   // streamed once
            const double dx = u64_to_dd(d.x), dy = u64_to_dd(d.y);
            a0 += mulmod_dd(dx, u64_to_dd(k0.x), pd, pinv); a1 += mulmod_dd(dy, u64_to_dd(k0.y), pd, pinv);
            b0 += mulmod_dd(dx, u64_to_dd(k1.x), pd, pinv); b1 += mulmod_dd(dy, u64_to_dd(k1.y), pd, pinv);
        }
        o0.x = dd_to_canonical(reduce_dd(a0, pd, pinv), pd); o0.y = dd_to_canonical(reduce_dd(a1, pd, pinv), pd);
        o1.x = dd_to_canonical(reduce_dd(b0, pd, pinv), pd); o1.y = dd_to_canonical(reduce_dd(b1, pd, pinv), pd);
    } else {
        const u64 p = M.p, r0 = M.r0, r1 = M.r1;
        u64 a0l = 0, a0h = 0, a1l = 0, a1h = 0, b0l = 0, b0h = 0, b1l = 0, b1h = 0;
        for (u32 j = 0; j < J.C; j++) {
            const ulonglong2 d = __ldcg(reinterpret_cast<const ulonglong2 *>(tp + (u64)j * n));
            const u64 *kp = J.keys[j] + koff;
            const ulonglong2 k0 = __ldcs(reinterpret_cast<const ulonglong2 *>(kp)), k1 = __ldcs(reinterpret_cast<const ulonglong2 *>(kp + kpoly));
            mac128(a0l, a0h, d.x, k0.x); mac128(a1l, a1h, d.y, k0.y);
            mac128(b0l, b0h, d.x, k1.x); mac128(b1l, b1h, d.y, k1.y);
            if ((j & 7) == 7) {     // keep the lazy 128-bit sums below 2^125 (8 products of 61-bit residues)
                a0l = barrett128(a0l, a0h, p, r0, r1); a0h = 0; a1l = barrett128(a1l, a1h, p, r0, r1); a1h = 0;
                b0l = barrett128(b0l, b0h, p, r0, r1); b0h = 0; b1l = barrett128(b1l, b1h, p, r0, r1); b1h = 0;
            }
        }
        o0.x = barrett128(a0l, a0h, p, r0, r1); o0.y = barrett128(a1l, a1h, p, r0, r1);
        o1.x = barrett128(b0l, b0h, p, r0, r1); o1.y = barrett128(b1l, b1h, p, r0, r1);
    }
    *reinterpret_cast<ulonglong2 *>(J.prod + ((u64)0 * J.G + it.g) * n + x) = o0;
    *reinterpret_cast<ulonglong2 *>(J.prod + ((u64)1 * J.G + it.g) * n + x) = o1;
}

// tensor maps: tc = NATURAL view of the coefficient-form target ([C] limbs); the ring in its three views
struct Ks5Maps { CUtensorMap tc, ring_perm, ring_rows; };

__global__ void __launch_bounds__(N5_THREADS, 3) ks5_kernel(const __grid_constant__ Ks5Job J, const __grid_constant__ Ks5Maps MP) {
    extern __shared__ char smem_raw[];
    __shared__ __align__(8) u64 s_full[2], s_empty[2];
    __shared__ u32 s_id[2];
    const u32 stage0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const u32 tw0 = stage0 + 2 * N5_STAGE_BYTES;
    const int t = threadIdx.x;
    const u32 total = J.total_items;
    u32 *const queue = J.sync, *const finished = J.sync + 1, *const done1 = J.sync + 2, *const gdone = done1 + J.G * J.C, *const ipdone = gdone + J.G;

    if (t == 0) {
        mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1); mbar_init(&s_empty[0], 8); mbar_init(&s_empty[1], 8);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        fence_proxy_async();
    }
    __syncthreads();

    if (t < NTT_THREADS) {
        // ================================ consumers ================================
        const N5Lane L = n5_lane();
        u32 parity = 0, s = 0;
        while (true) {
            mbar_wait(&s_full[s], (parity >> s) & 1u, 11);
            parity ^= 1u << s;
            const u32 cur_id = s_id[s];
            if (cur_id >= total) break;
            const Ks5Item it = ks5_decode(J, cur_id);
            const u32 tagged = J.table[it.g];
            const u32 ki = group_table(tagged);
            const ModInfo &M = J.mods[ki];
            const u32 stage = stage0 + s * N5_STAGE_BYTES, tw = tw0 + s * N5_TW_BYTES;
            if (it.phase == 2) {
                ks5_inner_product(J, it, M, it.g % KS5_RING, ki);
            } else {
                const u32 path = group_path(tagged, true);
                Ks5Reduce xf;
                {
                    const ModInfo &Mj = J.mods[it.j];
                    xf.m = M.p; xf.ratio = M.ratio;
                    xf.reduce = !(Mj.p <= M.p || (J.f64_lazy && M.use_f64 && Mj.bits <= M.bits + 4));
                }
                auto run = [&](const auto &ar) {
                    typedef typename std::remove_reference<decltype(ar)>::type AR;
                    if (it.phase == 0) n5_fwd_col(stage, tw, ar, L, xf, cur_id);
                    else n5_fwd_row(stage, tw, reinterpret_cast<const char *>(M.n5_frow) + (size_t)it.tile * 4096 * sizeof(typename AR::TwT), ar, L, cur_id);
                };
                if (path == PATH_F64) run(ArithF64(M));
                else if (path == PATH_INT_LAZY) run(ArithInt<true>(M));
                else run(ArithInt<false>(M));
                fence_proxy_async_smem();
            }
            __syncwarp();
            if ((t & 31) == 0) mbar_arrive(&s_empty[s]);
            s ^= 1;
        }
    } else {
      if (t == NTT_THREADS) {
        // ================================ producer ================================
        auto dep_counter = [&](const Ks5Item &it, u32 &need) -> const u32 * {
            if (it.phase == 0) { need = (u32)KS5_IPI; return it.g >= (u32)KS5_RING ? ipdone + (it.g - KS5_RING) : nullptr; }
            if (it.phase == 1) { need = 16; return done1 + it.g * J.C + it.j; }
            need = J.C * 16; return gdone + it.g;
        };
        auto pub_counter = [&](const Ks5Item &it) -> u32 * {
            return it.phase == 0 ? done1 + it.g * J.C + it.j : it.phase == 1 ? gdone + it.g : ipdone + it.g;
        };
        auto issue = [&](const Ks5Item &it, u32 s) {
            u64 *bar = &s_full[s];
            if (it.phase == 2) { mbar_arrive(bar); return; }          // nothing to stage: the inner product reads global memory itself
            const u32 tagged = J.table[it.g];
            const ModInfo &M = J.mods[group_table(tagged)];
            const u32 stage = stage0 + s * N5_STAGE_BYTES, tw = tw0 + s * N5_TW_BYTES;
            const u32 twb = group_path(tagged, true) == PATH_F64 ? 2048u : 4096u;
            const u32 limb = (it.g % KS5_RING) * J.C + it.j;
            mbar_expect_tx(bar, (u32)N5_STAGE_BYTES + twb);
            if (it.phase == 0) {
                tma_load_3d(stage, &MP.tc, bar, (int)(it.tile * 16), 0, (int)it.j);
                bulk_load(tw, M.n5_fcol, twb, bar);
            } else {
                tma_load_2d(stage, &MP.ring_rows, bar, 0, (int)(limb * 4096 + it.tile * 256));
                bulk_load(tw, reinterpret_cast<const char *>(M.n5_frow_u) + (size_t)it.tile * twb, twb, bar);
            }
        };
        auto store = [&](const Ks5Item &it, u32 s) {
            if (it.phase == 2) return;
            const u32 stage = stage0 + s * N5_STAGE_BYTES;
            const u32 limb = (it.g % KS5_RING) * J.C + it.j;
            if (it.phase == 0) tma_store_4d(&MP.ring_perm, stage, (int)(it.tile * 16), 0, 0, (int)limb);
            else tma_store_2d(&MP.ring_rows, stage, 0, (int)(limb * 4096 + it.tile * 256));
        };
        bool have[2] = {false, false};
        Ks5Item cur[2] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
        u32 eparity = 0;
        u32 *pub_new = nullptr, *pub_old = nullptr;       // counters of tiles whose store is in flight and not yet published
        u32 id = claim(queue);
        Ks5Item it = {0, 0, 0, 0};
        const u32 *depc = nullptr; u32 need = 0, dv = 0;
        if (id < total) { it = ks5_decode(J, id); depc = dep_counter(it, need); dv = depc ? ld_relaxed(depc) : need; }
        for (u32 k = 0;; k++) {
            const u32 s = k & 1;
            bool stored = false;
            if (have[s]) {
                mbar_wait(&s_empty[s], (eparity >> s) & 1u, 12);
                eparity ^= 1u << s;
                store(cur[s], s);
                bulk_commit();
                pub_old = pub_new;
                pub_new = pub_counter(cur[s]);
                have[s] = false; stored = true;
            }
            if (id >= total) {
                s_id[s] = total;
                mbar_arrive(&s_full[s]);
                bulk_wait0();
                if (pub_old) red_release(pub_old);
                if (pub_new) red_release(pub_new);
                pub_old = pub_new = nullptr;
                if (!have[s ^ 1]) break;
                continue;
            }
            if (stored) bulk_wait_read0();
            if (dv < need && ld_relaxed(depc) < need) {
                // dependency not met: nothing of ours may stay unpublished while we wait
                u32 *pub_x = nullptr;
                const u32 o = s ^ 1;
                if (have[o]) {
                    mbar_wait(&s_empty[o], (eparity >> o) & 1u, 13);
                    eparity ^= 1u << o;
                    store(cur[o], o);
                    bulk_commit();
                    pub_x = pub_counter(cur[o]);
                    have[o] = false;
                }
                bulk_wait0();
                if (pub_old) red_release(pub_old);
                if (pub_new) red_release(pub_new);
                if (pub_x) red_release(pub_x);
                pub_old = pub_new = nullptr;
                for (u32 spins = 0; ld_relaxed(depc) < need; spins++) {
                    __nanosleep(64);
                    if (spins == (1u << 24)) { printf("b2 ks5: dependency of item %u (phase %u group %u digit %u) stuck at %u of %u (block %d)\n", id, it.phase, it.g, it.j, ld_relaxed(depc), need, (int)blockIdx.x); __trap(); }
                }
            }
            if (depc) { if (it.phase == 2) __threadfence(); else fence_proxy_async_global(); }     // order the consumers' reads behind the observed counter
            s_id[s] = id;
            issue(it, s);
            have[s] = true; cur[s] = it;
            // (no claim-ahead here: the claimed-but-unfinished window must stay well inside the two-step lags, see KS5_LAG1)
            const u32 next_id = claim(queue);
            asm volatile("cp.async.bulk.wait_group 1;\n" ::: "memory");
            if (pub_old) { red_release(pub_old); pub_old = nullptr; }
            id = next_id;
            if (id < total) { it = ks5_decode(J, id); depc = dep_counter(it, need); dv = depc ? ld_relaxed(depc) : need; }
        }
        bulk_wait0();
        if (pub_old) red_release(pub_old);
        if (pub_new) red_release(pub_new);
      }
      __syncwarp();
    }
    __shared__ u32 s_last;
    __syncthreads();
    if (t == 0) {
        __threadfence();
        s_last = (atomicAdd(finished, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {
        const u32 cnt = J.G * J.C + 2 * J.G;
        for (u32 i = t; i < cnt; i += N5_THREADS) done1[i] = 0;
        if (t == 0) { *queue = 0; *finished = 0; }
    }
}

}  // namespace b2
