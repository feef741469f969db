// b2_ntt_warp.cuh — warp-autonomous persistent NTT for N = 2^16 (the BASELINE headline size).
//
// Same mathematical definition and the same load / store functor interface as ntt_persistent_kernel (b2_ntt.cuh;
// reference seam src/utils/ntt.h:132-144 -> src/fgk/ntt_grouped.cu:159-344,489-685), different execution model:
//
//   * The 2^16 coefficients of a limb are a 256 x 256 matrix.  A pass (8 butterfly layers) is cut into 128 SLABS of 512
//     coefficients — two columns (column pass) or two rows (row pass) — and a slab is transformed by ONE WARP: 16
//     coefficients per lane, two radix-16 rounds, one exchange through 4 KiB of the warp's own shared memory.  There is
//     no CTA barrier anywhere on the data path (ncu on the CTA-tile kernel: 21 % of all stall samples were barrier
//     waits, r01c); warps only meet in bar.warp.sync.
//   * Every warp runs its own software pipeline: the next slab's 4 KiB (and, for the row pass, the 4 KiB of twiddles of
//     its two rows, from a row-major copy of the table) is in flight through 16-byte cp.async while the current slab is
//     computed; the only synchronous global loads left are the 255 column-pass twiddles, which are L1-resident.
//   * One persistent CTA of 16 warps per SM (12 KiB of shared memory per warp, <= 128 registers).  Warps pull items
//     (pass, limb, pair of slabs) from the same in-order queue with per-limb completion counters as the CTA-tile kernel,
//     so the second pass of a limb still finds its intermediate in L2.
#pragma once
#include "b2_ntt.cuh"

namespace b2 {

constexpr int W_WARPS = 16;
constexpr int W_THREADS = W_WARPS * 32;
constexpr int W_SLAB = 512;                          // coefficients per slab
constexpr int W_ITEM_SLABS = 4;                      // slabs per queue item
constexpr int W_ITEM_SHIFT = 5;                      // 32 items per limb and pass
constexpr int W_WARP_WORDS = 3 * W_SLAB;             // two data buffers + one twiddle buffer
constexpr int W_SMEM_BYTES = W_WARPS * W_WARP_WORDS * 8;   // 192 KiB

// shared-memory layouts of a slab (8-byte word index -> word index), all GF(2)-linear:
//   column slab, local index = row * 2 + col: stage and exchange share one 16-byte-granular layout
__host__ __device__ constexpr int wl_col(int i) { return i ^ (((i >> 5) & 15) << 1); }
//   row slab, local index = position in the two rows: 16-byte-granular stage layout, 8-byte-granular exchange layout
__host__ __device__ constexpr int wl_rs(int i) { return i ^ (((i >> 4) & 7) << 1); }
__host__ __device__ constexpr int wl_rx(int i) { return i ^ ((i >> 4) & 15); }

template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ---- one radix-16 round (four butterfly layers), the ONLY hot code of the kernel -------------------------------------
// The round is generic: which layers it runs and where their twiddles live are run-time parameters, so the column
// and the row pass, their first and second round, all execute the same ~450 instructions (the 16 warps of an SM run
// unsynchronised; ncu on a fully specialised version showed 29 % of the stall samples in instruction fetch — the
// unrolled bodies did not fit the 32 KB L1.5 instruction cache, r01e).
//   p0..p3: twiddles of the four layers in execution order, GENERIC pointers (shared-memory staged copy or global table)
//   l0:     index of the first layer (layer s is l0 + s);  last: inverse only — layer 3 is the final layer (N^-1 folded)
template <class TwT>
__device__ __forceinline__ void w_ld_tw(const TwT *p, int cnt, TwT (&w)[8]);
template <>
__device__ __forceinline__ void w_ld_tw<double>(const double *p, int cnt, double (&w)[8]) {
    if (cnt == 1) { w[0] = p[0]; return; }
#pragma unroll
    for (int a = 0; a < 4; a++) if (2 * a < cnt) {
        const double2 d = reinterpret_cast<const double2 *>(p)[a];
        w[2 * a] = d.x; w[2 * a + 1] = d.y;
    }
}
template <>
__device__ __forceinline__ void w_ld_tw<Tw>(const Tw *p, int cnt, Tw (&w)[8]) {
#pragma unroll
    for (int a = 0; a < 8; a++) if (a < cnt) w[a] = p[a];
}

template <bool FWD, class AR, class Hook>
__device__ __forceinline__ void w_round(const AR &ar, typename AR::T (&v)[16], const typename AR::TwT *p0, const typename AR::TwT *p1,
                                        const typename AR::TwT *p2, const typename AR::TwT *p3, int l0, bool last, const Hook &after_tw) {
    typedef typename AR::TwT TwT;
    const TwT *pp[4] = {p0, p1, p2, p3};
    TwT w[4][8];
    if (AR::PREFETCH) {
#pragma unroll
        for (int s = 0; s < 4; s++) w_ld_tw<TwT>(pp[s], 16 >> ((FWD ? 3 - s : s) + 1), w[s]);
        after_tw();
    }
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int j = FWD ? 3 - s : s;
        const int l = l0 + s;
        if (!FWD && s == 3 && last) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (k & (1 << j)) continue;
                ar.bf_inv_last(v[k], v[k | (1 << j)], l);
            }
            continue;
        }
        if (!AR::PREFETCH) w_ld_tw<TwT>(pp[s], 16 >> (j + 1), w[s]);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (k & (1 << j)) continue;
            if (FWD) ar.bf_fwd(v[k], v[k | (1 << j)], w[s][k >> (j + 1)]);
            else ar.bf_inv(v[k], v[k | (1 << j)], w[s][k >> (j + 1)], l);
        }
    }
    if (!AR::PREFETCH) after_tw();
}

// ---- staging ----------------------------------------------------------------------------------------------------------
// column slab s of a limb: 256 rows x 16 bytes
__device__ __forceinline__ void w_stage_col(u64 *buf, const u64 *limb, int s, int lane) {
#pragma unroll
    for (int m = 0; m < 8; m++) {
        const int j = lane + 32 * m;                              // row
        cp_async16(buf + wl_col(2 * j), limb + (j << 8) + 2 * s);
    }
}
// row slab s of a limb: 4 KiB contiguous
__device__ __forceinline__ void w_stage_row(u64 *buf, const u64 *limb, int s, int lane) {
#pragma unroll
    for (int m = 0; m < 8; m++) {
        const int j = lane + 32 * m;                              // 16-byte chunk
        cp_async16(buf + wl_rs(2 * j), limb + s * W_SLAB + 2 * j);
    }
}
__device__ __forceinline__ void w_stage_tw(double *twb, const double *rowtab, int s, int lane) {
#pragma unroll
    for (int m = 0; m < 8; m++) {
        const int j = lane + 32 * m;
        cp_async16(twb + 2 * j, rowtab + s * W_SLAB + 2 * j);
    }
}

// ---- a slab ---------------------------------------------------------------------------------------------------------------
// forward = column pass then row pass; inverse = row pass then column pass.  Lane geometry of the four round types:
//   column slab:  A: window = row bits 4..7, lane = local bits 0..4        i = lane + 32 k        x = xA + (k << 12)
//                 B: window = row bits 0..3, lane = (col, row bits 4..7)   i = col + 2 k + 32 hi  x = xB + (k << 8)
//   row slab:     C: window = col bits 4..7, lane = (col bits 0..3, row)   i = lo + 16 k + 256 rl x = xC + (k << 4)
//                 D: window = col bits 0..3, lane = (col bits 4..7, row)   i = k + 16 lo + 256 rl x = xD + k
// forward: A, B | C, D ; inverse: D, C | B, A.  The first round reads the staged slab (pass-specific code), then the
// generic round runs twice around one exchange through the same buffer, then a pass-specific store.
//   mid_hook: runs after the exchange (the other data buffer is free: prefetch the next slab)
//   tw_hook:  runs once every twiddle of the slab's second round has been fetched (the twiddle buffer is free)
template <bool FWD, class AR, class Load, class Store, class MidHook, class TwHook>
__device__ __forceinline__ void w_slab(u64 *buf, const double *twb, const ModInfo &M, bool p2, int s, int g, int c, u64 off, u64 *mid,
                                       const Load &ld, const Store &st, const MidHook &mid_hook, const TwHook &tw_hook) {
    typedef typename AR::T T;
    typedef typename AR::TwT TwT;
    constexpr bool STAGED = AR::PREFETCH;             // FP64 path: row-pass twiddles come from the staged row-major copy
    const bool col = FWD ? !p2 : p2;
    const AR ar(M);
    const int t = threadIdx.x & 31;
    T v[16];
    // run-time description of the two rounds and of the exchange between them
    const TwT *q0, *q1, *q2, *q3, *r0, *r1, *r2, *r3;    // twiddles: first round, second round
    int la, lb;                                       // first layer of each round
    int wbase, rbase;                                 // exchange: word index of register 0 when writing / reading
    int wsh, rsh;                                     // exchange: register k sits at base ^ (k << sh) ^ (k << sh2)
    int wsh2, rsh2;                                   //   (sh2 < 0: no second term)
    const TwT *tab = FWD ? ar.ftab : ar.itab;
    const int cl = t & 1, hi = t >> 1, lo = t & 15, rl = t >> 4;
    const u32 xA = ((u32)hi << 8) + 2 * s + cl, xB = ((u32)hi << 12) + 2 * s + cl;
    const int iC = lo + (rl << 8), iD = (lo << 4) + (rl << 8);
    const u32 xC = (u32)s * W_SLAB + iC, xD = (u32)s * W_SLAB + iD;
    if (col) {
        const int pA = t, pB = wl_col(cl + (hi << 5));
        if (FWD) {          // rounds A (layers 0..3), B (layers 4..7)
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = ar.from_ext(ld.xf(g, c, xA + ((u32)k << 12), buf[pA ^ wl_col(k << 5)]));
            q0 = tab + 1; q1 = tab + 2; q2 = tab + 4; q3 = tab + 8;
            r0 = tab + 16 + hi; r1 = tab + 32 + 2 * hi; r2 = tab + 64 + 4 * hi; r3 = tab + 128 + 8 * hi;
            la = 0; lb = 4;
            wbase = pA; wsh = 5; wsh2 = 1; rbase = pB; rsh = 1; rsh2 = -1;
        } else {            // rounds B (layers 8..11), A (layers 12..15, the last)
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = ar.from_raw(buf[pB ^ wl_col(k << 1)]);
            // natural inverse table: layer l entry N - (N >> l) + (x >> (l + 1))
            q0 = tab + (65536 - 256) + 8 * hi; q1 = tab + (65536 - 128) + 4 * hi; q2 = tab + (65536 - 64) + 2 * hi; q3 = tab + (65536 - 32) + hi;
            r0 = tab + (65536 - 16); r1 = tab + (65536 - 8); r2 = tab + (65536 - 4); r3 = tab + (65536 - 2);
            la = 8; lb = 12;
            wbase = pB; wsh = 1; wsh2 = -1; rbase = pA; rsh = 5; rsh2 = 1;
        }
    } else {
        const TwT *twr = reinterpret_cast<const TwT *>(twb) + (rl << 8);     // only meaningful when STAGED
        if (FWD) {          // rounds C (layers 8..11), D (layers 12..15)
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = ar.from_raw(buf[iC ^ wl_rs(k << 4)]);
            if (STAGED) {
                q0 = twr + 1; q1 = twr + 2; q2 = twr + 4; q3 = twr + 8;
                r0 = twr + 16 + lo; r1 = twr + 32 + 2 * lo; r2 = twr + 64 + 4 * lo; r3 = twr + 128 + 8 * lo;
            } else {
                q0 = tab + 256 + (xC >> 8); q1 = tab + 512 + (xC >> 7); q2 = tab + 1024 + (xC >> 6); q3 = tab + 2048 + (xC >> 5);
                r0 = tab + 4096 + (xD >> 4); r1 = tab + 8192 + (xD >> 3); r2 = tab + 16384 + (xD >> 2); r3 = tab + 32768 + (xD >> 1);
            }
            la = 8; lb = 12;
            wbase = iC; wsh = 4; wsh2 = 0; rbase = iD ^ lo; rsh = 0; rsh2 = -1;
        } else {            // rounds D (layers 0..3), C (layers 4..7)
            const int sw = (lo & 7) << 1;                   // wl_rs(iD + k) = iD + (k ^ sw): 16-byte pairs stay adjacent
            u64 raw[16];
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                const ulonglong2 d = *reinterpret_cast<const ulonglong2 *>(buf + iD + (k ^ sw));
                raw[k] = d.x; raw[k + 1] = d.y;
            }
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = ar.from_ext(ld.xf(g, c, xD + (u32)k, raw[k]));
            if (STAGED) {
                q0 = twr + 128 + 8 * lo; q1 = twr + 64 + 4 * lo; q2 = twr + 32 + 2 * lo; q3 = twr + 16 + lo;
                r0 = twr + 8; r1 = twr + 4; r2 = twr + 2; r3 = twr + 1;
            } else {
                q0 = tab + (xD >> 1); q1 = tab + 32768 + (xD >> 2); q2 = tab + 49152 + (xD >> 3); q3 = tab + 57344 + (xD >> 4);
                r0 = tab + 61440 + (xC >> 5); r1 = tab + 63488 + (xC >> 6); r2 = tab + 64512 + (xC >> 7); r3 = tab + 65024 + (xC >> 8);
            }
            la = 0; lb = 4;
            wbase = iD ^ lo; wsh = 0; wsh2 = -1; rbase = iC; rsh = 4; rsh2 = 0;
        }
    }
    const NoHook nohook;
#pragma unroll 1
    for (int r = 0; r < 2; r++) {
        if (r == 0) w_round<FWD>(ar, v, q0, q1, q2, q3, la, false, nohook);
        else w_round<FWD>(ar, v, q0, q1, q2, q3, la, !FWD && col, tw_hook);
        if (r == 0) {
            if (!col) __syncwarp();     // row slabs: stage and exchange layouts differ — every lane must have read its staged words
            {
                const int m2 = wsh2 < 0 ? 0 : ~0, s2 = wsh2 < 0 ? 0 : wsh2;
#pragma unroll
                for (int k = 0; k < 16; k++) buf[wbase ^ (k << wsh) ^ ((k << s2) & m2)] = ar.to_raw(v[k]);
            }
            __syncwarp();
            {
                const int m2 = rsh2 < 0 ? 0 : ~0, s2 = rsh2 < 0 ? 0 : rsh2;
#pragma unroll
                for (int k = 0; k < 16; k++) v[k] = ar.from_raw(buf[rbase ^ (k << rsh) ^ ((k << s2) & m2)]);
            }
            mid_hook();
            q0 = r0; q1 = r1; q2 = r2; q3 = r3; la = lb;
        }
    }
    if (col) {
        if (FWD) {
#pragma unroll
            for (int k = 0; k < 16; k++) mid[off + xB + ((u32)k << 8)] = ar.to_raw(v[k]);
        } else {
#pragma unroll
            for (int k = 0; k < 16; k++) st(g, c, off, xA + ((u32)k << 12), ar.final_inv(v[k]));
        }
    } else {
        if (FWD) {
            u64 o[16];
#pragma unroll
            for (int k = 0; k < 16; k++) o[k] = ar.final_fwd(v[k]);
#pragma unroll
            for (int k = 0; k < 16; k += 2) st.st2(g, c, off, xD + (u32)k, o[k], o[k + 1]);
        } else {
#pragma unroll
            for (int k = 0; k < 16; k++) { ar.pass_boundary_inv(v[k]); mid[off + xC + ((u32)k << 4)] = ar.to_raw(v[k]); }
        }
    }
}

template <bool FWD, class Load, class Store, class MidHook, class TwHook>
__device__ __forceinline__ void w_slab_dispatch(u32 path, u64 *buf, const double *twb, const ModInfo &M, bool p2, int s, int g, int c, u64 off,
                                                u64 *mid, const Load &ld, const Store &st, const MidHook &mh, const TwHook &th) {
    if (path == PATH_F64) w_slab<FWD, ArithF64>(buf, twb, M, p2, s, g, c, off, mid, ld, st, mh, th);
    else if (path == PATH_INT_LAZY) w_slab<FWD, ArithInt<true>>(buf, twb, M, p2, s, g, c, off, mid, ld, st, mh, th);
    else w_slab<FWD, ArithInt<false>>(buf, twb, M, p2, s, g, c, off, mid, ld, st, mh, th);
}

// The kernel.  J.tile_shift == W_ITEM_SHIFT: an item is W_ITEM_SLABS consecutive slabs of one (pass, limb); done[l] counts
// the finished first-pass items of limb l.  Per warp: the next item is claimed while the first slab of the current one is
// transformed; its dependency counter is requested at the start of the last slab and examined half a slab later, when
// the prefetch of the next slab is issued (all slabs: after the exchange of the current one); the completion signal of
// a first-pass item (MEMBAR + RED) is deferred to the middle of the following slab, when its stores have long drained.
template <bool FWD, class Load, class Store>
__global__ void __launch_bounds__(W_THREADS, 1) ntt_warp_kernel(const __grid_constant__ NttJob J, Load ld, Store st) {
    extern __shared__ __align__(128) u64 smem[];
    __shared__ u32 s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 *const wbuf = smem + warp * W_WARP_WORDS;
    double *const twb = reinterpret_cast<double *>(wbuf + 2 * W_SLAB);
    const u32 total = J.total_items, per_limb = 1u << W_ITEM_SHIFT;
    u32 *const queue = J.sync, *const finished = J.sync + 1, *const done = J.sync + 2;

    struct Slab { u32 phase, limb, s; };
    struct LimbRef { u32 g, c, table, path; u64 off; const double *rowtab; };
    auto limb_ref = [&](u32 limb, u32 phase) {
        LimbRef r;
        r.c = limb / (u32)J.n_groups; r.g = limb - r.c * (u32)J.n_groups;
        const LimbGroup &grp = J.groups[r.g];
        r.off = grp.base + (u64)r.c * grp.stride;
        r.table = group_table(grp.table); r.path = group_path(grp.table, FWD);
        const bool row = FWD ? phase == 1 : phase == 0;
        r.rowtab = (row && r.path == PATH_F64) ? (FWD ? J.mods[r.table].twrow : J.mods[r.table].itwrow) : nullptr;
        return r;
    };
    auto stage_data = [&](const Slab &sl, const LimbRef &r, u64 *buf) {
        if (sl.phase == 0) {
            const u64 *src = ld.src(J.g_base + (int)r.g, (int)r.c, r.off);
            if (FWD) w_stage_col(buf, src, (int)sl.s, lane); else w_stage_row(buf, src, (int)sl.s, lane);
        } else {
            if (FWD) w_stage_row(buf, J.mid + r.off, (int)sl.s, lane); else w_stage_col(buf, J.mid + r.off, (int)sl.s, lane);
        }
    };
    auto wait_dep = [&](u32 limb) {
        if (lane == 0) while (ld_relaxed(done + limb) < per_limb) __nanosleep(64);
        __syncwarp();
    };

    u32 id = 0;
    if (lane == 0) id = claim(queue);
    id = __shfl_sync(0xffffffffu, id, 0);
    if (id < total) {
        const NttItem it = ntt_decode(J, id);
        Slab cur{it.phase, it.limb, it.tile * W_ITEM_SLABS};
        if (cur.phase) wait_dep(cur.limb);
        LimbRef cr = limb_ref(cur.limb, cur.phase);
        int b = 0;
        stage_data(cur, cr, wbuf);
        if (cr.rowtab) w_stage_tw(twb, cr.rowtab, (int)cur.s, lane);
        cp_async_commit();
        u32 claimed = 0, depv = 0;  // lane 0: id of the item after the current one; its dependency counter
        bool pub = false; u32 pub_limb = 0;
        while (true) {
            const u32 sub = cur.s & (W_ITEM_SLABS - 1);
            const bool last_sub = sub == W_ITEM_SLABS - 1;
            if (sub == 0 && lane == 0) claimed = claim(queue);
            Slab nxt = cur; LimbRef nr = cr; bool have_next = true, need_dep = false, staged = false;
            if (!last_sub) nxt.s = cur.s + 1;
            else {
                const u32 nid = __shfl_sync(0xffffffffu, claimed, 0);
                have_next = nid < total;
                if (have_next) {
                    const NttItem ni = ntt_decode(J, nid);
                    nxt = Slab{ni.phase, ni.limb, ni.tile * W_ITEM_SLABS};
                    nr = limb_ref(nxt.limb, nxt.phase);
                    need_dep = nxt.phase != 0;
                    if (need_dep && lane == 0) depv = ld_relaxed(done + nxt.limb);
                }
            }
            cp_async_wait<0>();      // this slab's data and twiddles (issued half a slab ago or earlier) have landed
            __syncwarp();
            {
                const ModInfo &M = J.mods[cr.table];
                const int ga = J.g_base + (int)cr.g;
                auto mid_hook = [&]() {
                    if (pub) { if (lane == 0) red_release(done + pub_limb); pub = false; }
                    if (have_next) {
                        u32 ok = 1;
                        if (need_dep) { ok = depv >= per_limb ? 1u : 0u; ok = __shfl_sync(0xffffffffu, ok, 0); }
                        if (ok) { stage_data(nxt, nr, wbuf + (b ^ 1) * W_SLAB); staged = true; }
                    }
                    cp_async_commit();
                };
                auto tw_hook = [&]() {
                    __syncwarp();
                    if (staged && nr.rowtab) w_stage_tw(twb, nr.rowtab, (int)nxt.s, lane);
                    cp_async_commit();
                };
                w_slab_dispatch<FWD>(cr.path, wbuf + b * W_SLAB, twb, M, cur.phase != 0, (int)cur.s, ga, (int)cr.c, cr.off, J.mid, ld, st, mid_hook, tw_hook);
            }
            if (last_sub) {
                if (cur.phase == 0) { __syncwarp(); pub = true; pub_limb = cur.limb; }
                if (!have_next) break;
                if (!staged) {       // dependency was not met at prefetch time: signal our own completion, wait, stage
                    if (pub) { if (lane == 0) red_release(done + pub_limb); pub = false; }
                    wait_dep(nxt.limb);
                    stage_data(nxt, nr, wbuf + (b ^ 1) * W_SLAB);
                    if (nr.rowtab) w_stage_tw(twb, nr.rowtab, (int)nxt.s, lane);
                    cp_async_commit();
                }
            }
            cur = nxt; cr = nr; b ^= 1;
        }
        if (pub && lane == 0) red_release(done + pub_limb);
    }
    cp_async_wait<0>();
    // ---- teardown: the last CTA to finish resets the queue and the counters for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(finished, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {
        for (u32 i = threadIdx.x; i < J.n_limbs; i += W_THREADS) done[i] = 0;
        if (threadIdx.x == 0) { *queue = 0; *finished = 0; }
    }
}

}  // namespace b2
