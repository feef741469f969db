// b2_ntt.cuh — negacyclic NTT / inverse NTT over RNS limbs for sm_100a.
//
// Replaces the reference's utils::ntt_transfer_to_rev / ntt_transfer_from_rev
// (src/utils/ntt.h:132-144 -> src/fgk/ntt_grouped.cu:159-344,489-685).  Same mathematical
// definition (psi = minimal primitive 2N-th root, natural-order input, bit-reversed output,
// canonical output), different algorithm:
//
//   * N = 2^logN is viewed as a (2^A rows) x (256 columns) matrix, A = logN - 8.  The forward
//     transform is a COLUMN pass (layers 0..A-1, strides >= 256) followed by a ROW pass (layers
//     A..logN-1 inside contiguous 2 KiB rows); the inverse runs the two passes in the opposite
//     order.  Each pass works on 4096-coefficient tiles (32 KiB): 2^(12-A) columns x 2^A rows
//     (128-byte row segments when N = 2^16) or 16 whole rows.
//   * ONE persistent launch per transform call (ntt_persistent_kernel): 3 CTAs per SM pull (pass,
//     limb, tile) items from an in-order work queue.  The second-pass items of limb l are queued
//     `lag` limbs behind its first-pass items and wait on a per-limb completion counter, so the
//     intermediate of a limb (512 KiB at N = 2^16) is still in L2 when the second pass reads it:
//     HBM sees one read and one write per coefficient, there are no per-wave launches and no tails.
//   * Tiles are staged into shared memory with 16-byte cp.async (LDGSTS) copies: as soon as every thread has
//     pulled item i out of the stage buffer (the pass's first barrier) the copy of item i+1 is issued and is in
//     flight while item i is computed; 3 CTAs per SM (64 KiB shared memory, <= 80 registers) cover the rest.
//     The stage uses a 16-byte-granular XOR swizzle, the exchange buffer an 8-byte-granular one; every
//     access pattern of every round is bank-conflict free.
//   * Inside a pass every thread keeps 16 coefficients in registers and performs a radix-16
//     (four-layer) butterfly network on them; an 8-layer pass touches the exchange buffer once.
//   * Two arithmetic back ends (b2_modarith.cuh): ArithF64 — exact double-double modular multiply on
//     the FP64 pipe for p < 2^41, signed lazy values, 8 DP instructions per butterfly; ArithInt —
//     Harvey/Shoup lazy butterflies on the integer pipe for larger p, with no per-layer correction
//     at all when the modulus leaves enough headroom (ModInfo::fwd_lazy / inv_lazy).
//   * N^-1 is folded into the last inverse layer's twiddles (no extra pass).  Twiddles of one
//     butterfly layer are contiguous per thread and fetched with 16-byte loads.
//   * Loads and stores go through functors so callers can fuse a prologue / epilogue into the
//     transform (keyswitch mod-up, rescale, mod-down: b2_keyswitch.cu, b2_rescale.cu).
#pragma once
#include "b2_modarith.cuh"

namespace b2 {

constexpr int NTT_TB = 12;                  // tile = 2^12 coefficients = 32 KiB
constexpr int NTT_TILE = 1 << NTT_TB;
constexpr int NTT_THREADS = NTT_TILE / 16;  // 256
constexpr int NTT_MAX_GROUPS = 72;

// A set of limbs that share one modulus: limb c (0 <= c < count) starts at element offset
// base + c * stride of the operand.
struct LimbGroup {
    u32 table;   // index into ModInfo[]
    u32 count;
    u64 base;
    u64 stride;
};

// One transform call = one persistent launch over n_groups * count limbs.
struct NttJob {
    const ModInfo *mods;
    u64 *mid;           // raw intermediate between the two passes; groups[].base/stride address it
    u32 *sync;          // [0] work-queue head, [1] finished CTAs, [2 + l] first-pass tiles of limb l completed
    int logn;
    u32 count;          // limbs per group (equal for all groups of a call)
    u32 n_limbs;        // n_groups * count
    u32 lag;            // second-pass items trail first-pass items by this many limbs (<= n_limbs)
    u32 tile_shift;     // log2(tiles per limb) = logN - 12
    u32 total_items;    // 2 * n_limbs * tiles
    int g_base;         // index of groups[0] in the caller's group list (functors see absolute indices)
    int n_groups;
    // scattered batches (b2_ntt_*_ptrs, the reference's SliceVec operands): limb c of a group belongs to batch item
    // c / tab_pc; its offset is that item's distance from `mid` (= out_tab[0]) plus the in-item offset.  Null: contiguous.
    const u64 *const *out_tab;
    u32 tab_pc;         // limbs of one group per batch item (pcount)
    u64 tab_pstride;    // elements between consecutive polynomials of an item (ncomp * N)
    u32 dbg;            // developer switches of the TMA kernel (B2_NTT5_DBG): 1 skip the arithmetic, 2 skip the stores
    LimbGroup groups[NTT_MAX_GROUPS];
};
// element offset (relative to J.mid) of limb c of a group
__device__ __forceinline__ u64 limb_offset(const NttJob &J, const LimbGroup &grp, u32 c) {
    if (J.out_tab == nullptr) return grp.base + (u64)c * grp.stride;
    const u32 b = c / J.tab_pc, k = c - b * J.tab_pc;
    return (u64)(J.out_tab[b] - J.mid) + grp.base + (u64)k * J.tab_pstride;
}

// XOR swizzle of a tile-local coefficient index (8-byte units).  It is GF(2)-linear, so
// swz(a | b) == swz(a) ^ swz(b) for disjoint a, b: per-register offsets are compile-time constants.
// For every round's access pattern a half-warp lands on 16 distinct 8-byte slots.
__host__ __device__ constexpr int swz(int idx) { return idx ^ ((idx >> 4) & 15); }
// Stage-buffer swizzle: 16-byte granular (cp.async writes whole 16-byte chunks): chunk index within a
// 128-byte row XOR the low three bits of the row index.  Also GF(2)-linear.
__host__ __device__ constexpr int swz16(int idx) { return idx ^ (((idx >> 4) & 7) << 1); }
constexpr int NTT_SMEM_BYTES = 2 * NTT_TILE * 8;   // one stage buffer + one exchange buffer = 64 KiB (3 CTAs per SM)

// tile-local index of register 0 of thread t in a round whose 4-bit window starts at bit LO
template <int LO>
__device__ __forceinline__ int round_base(int t) { return ((t >> LO) << (LO + 4)) | (t & ((1 << LO) - 1)); }

// ---- arithmetic back ends ----------------------------------------------------------------------
template <bool LAZY>
struct ArithInt {
    typedef u64 T;
    typedef Tw TwT;
    static constexpr bool PREFETCH = false;     // 16-byte twiddles: prefetching a layer ahead of the barrier costs too many registers
    u64 p, p2, ratio;
    Tw inv_n, inv_n_w;
    const Tw *ftab, *itab;
    __device__ __forceinline__ explicit ArithInt(const ModInfo &M)
        : p(M.p), p2(M.p << 1), ratio(M.ratio), inv_n(M.inv_n), inv_n_w(M.inv_n_w), ftab(M.tw), itab(M.itw) {}
    __device__ __forceinline__ T from_ext(u64 x) const { return x; }
    __device__ __forceinline__ T from_raw(u64 x) const { return x; }
    __device__ __forceinline__ u64 to_raw(T v) const { return v; }
    // cnt is a compile-time constant after unrolling at every call site
    __device__ __forceinline__ void ldtw_n(const TwT *t, u32 i, int cnt, TwT (&w)[8]) const {
#pragma unroll
        for (int a = 0; a < 8; a++) if (a < cnt) w[a] = __ldg(t + i + a);
    }
    __device__ __forceinline__ void bf_fwd(T &X, T &Y, TwT w) const {
        u64 x = LAZY ? X : csub(X, p2);
        u64 t = shoup_lazy(Y, w, p);
        X = x + t;
        Y = x - t + p2;
    }
    // l = layer index: in LAZY mode operands are below p * 2^(l+1)
    __device__ __forceinline__ void bf_inv(T &U, T &V, TwT w, int l) const {
        u64 s = U + V;
        u64 d = U - V + (LAZY ? (p << (l + 1)) : p2);
        U = LAZY ? s : csub(s, p2);
        V = shoup_lazy(d, w, p);
    }
    __device__ __forceinline__ void bf_inv_last(T &U, T &V, int l) const {   // N^-1 folded, canonical out
        u64 s = U + V, d = U - V + (LAZY ? (p << (l + 1)) : p2);
        U = csub(shoup_lazy(s, inv_n, p), p);
        V = csub(shoup_lazy(d, inv_n_w, p), p);
    }
    __device__ __forceinline__ void pass_boundary_inv(T &) const {}
    __device__ __forceinline__ u64 final_fwd(T v) const { return LAZY ? barrett64(v, p, ratio) : csub(csub(v, p2), p); }
    __device__ __forceinline__ u64 final_inv(T v) const { return v; }
};

struct ArithF64 {
    typedef double T;
    typedef double TwT;
    static constexpr bool PREFETCH = true;
    double p, pinv, inv_n, inv_n_w;
    const double *ftab, *itab;
    __device__ __forceinline__ explicit ArithF64(const ModInfo &M)
        : p(M.pd), pinv(M.pinv), inv_n(M.inv_n_d), inv_n_w(M.inv_n_w_d), ftab(M.twd), itab(M.itwd) {}
    __device__ __forceinline__ T from_ext(u64 x) const { return u64_to_dd(x); }
    __device__ __forceinline__ T from_raw(u64 x) const { return __longlong_as_double((long long)x); }
    __device__ __forceinline__ u64 to_raw(T v) const { return (u64)__double_as_longlong(v); }
    // CNT consecutive table entries starting at an index that is a multiple of CNT: 16-byte loads
    // (cnt is a compile-time constant after unrolling at every call site)
    __device__ __forceinline__ void ldtw_n(const TwT *t, u32 i, int cnt, TwT (&w)[8]) const {
        if (cnt == 1) { w[0] = __ldg(t + i); return; }
#pragma unroll
        for (int a = 0; a < 4; a++) if (2 * a < cnt) {
            const double2 d = __ldg(reinterpret_cast<const double2 *>(t + i) + a);
            w[2 * a] = d.x; w[2 * a + 1] = d.y;
        }
    }
    __device__ __forceinline__ void bf_fwd(T &X, T &Y, TwT w) const {
        double t = mulmod_dd(Y, w, p, pinv);
        Y = X - t;
        X = X + t;
    }
    __device__ __forceinline__ void bf_inv(T &U, T &V, TwT w, int) const {
        double s = U + V, d = U - V;
        U = s;
        V = mulmod_dd(d, w, p, pinv);
    }
    __device__ __forceinline__ void bf_inv_last(T &U, T &V, int) const {
        double s = U + V, d = U - V;
        U = mulmod_dd(s, inv_n, p, pinv);
        V = mulmod_dd(d, inv_n_w, p, pinv);
    }
    // sums double per inverse layer; one reduction between the two passes keeps |v| < 2^10 p
    __device__ __forceinline__ void pass_boundary_inv(T &v) const { v = reduce_dd(v, p, pinv); }
    __device__ __forceinline__ u64 final_fwd(T v) const { return dd_to_canonical(reduce_dd(v, p, pinv), p); }
    __device__ __forceinline__ u64 final_inv(T v) const { return dd_to_canonical(v, p); }
};

// ---- radix-16 rounds ------------------------------------------------------------------------------
struct NoHook { __device__ __forceinline__ void operator()() const {} };

// Twiddles of the first layers of a round, fetched BEFORE the barrier that precedes the round so that their L2 latency
// overlaps the barrier wait and the shared-memory reads (forward: layers 0 and 1 = 1 + 2 entries; inverse: layer 0 = 8).
template <int NL, int XLO, bool FWD, bool LAST, class AR>
struct RoundTw {
    static constexpr int PF = !AR::PREFETCH ? 0 : (FWD ? (NL < 2 ? NL : 2) : ((LAST && NL == 1) ? 0 : 1));
    typename AR::TwT w[PF > 0 ? PF : 1][8];
    // L1 prefetch hints for the layers load() does not fetch (register budget): their per-thread runs of 2..8 entries are
    // distinct per thread in the row pass and would otherwise be long-scoreboard waits on L2 inside the round
    __device__ __forceinline__ void hint_rest(const AR &ar, u32 xt, int logn) const {
        if (!AR::PREFETCH) return;
#pragma unroll
        for (int s = PF; s < NL; s++) {
            if (!FWD && LAST && s == NL - 1) continue;
            const typename AR::TwT *p;
            if (FWD) {
                const int j = NL - 1 - s, l = logn - 1 - (XLO + j);
                p = ar.ftab + (1u << l) + (xt >> (XLO + j + 1));
            } else {
                const int j = (4 - NL) + s, l = XLO + j;
                const u32 n = 1u << logn;
                p = ar.itab + n - (n >> l) + (xt >> (l + 1));
            }
            asm volatile("prefetch.global.L1 [%0];\n" ::"l"(p));
        }
    }
    __device__ __forceinline__ void load(const AR &ar, u32 xt, int logn) {
#pragma unroll
        for (int s = 0; s < PF; s++) {
            if (FWD) {
                const int j = NL - 1 - s;
                const int l = logn - 1 - (XLO + j);
                ar.ldtw_n(ar.ftab, (1u << l) + (xt >> (XLO + j + 1)), 16 >> (j + 1), w[s]);
            } else {
                const int j = (4 - NL) + s;
                const int l = XLO + j;
                const u32 n = 1u << logn;
                ar.ldtw_n(ar.itab, n - (n >> l) + (xt >> (l + 1)), 16 >> (j + 1), w[s]);
            }
        }
    }
};

// Forward (Cooley-Tukey) network on the low NL window bits (free window bits above them).
// xt: limb-local position of register 0 (window bits clear); XLO: limb-local position of window bit 0.
// `mid` runs once, about half way through the round (thread 0 uses it to advance the work-queue bookkeeping).
template <int NL, int XLO, class AR, class RT, class Mid>
__device__ __forceinline__ void fwd_round(const AR &ar, typename AR::T (&v)[16], u32 xt, int logn, const RT &rt, const Mid &mid) {
#pragma unroll
    for (int s = 0; s < NL; s++) {
        const int j = NL - 1 - s;
        const int l = logn - 1 - (XLO + j);
        const u32 tb = (1u << l) + (xt >> (XLO + j + 1));
        typename AR::TwT w[8];
        if (s < RT::PF) {
#pragma unroll
            for (int a = 0; a < 8; a++) if (a < (16 >> (j + 1))) w[a] = rt.w[s < RT::PF ? s : 0][a];
        } else ar.ldtw_n(ar.ftab, tb, 16 >> (j + 1), w);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (k & (1 << j)) continue;
            ar.bf_fwd(v[k], v[k | (1 << j)], w[k >> (j + 1)]);
        }
        if (s == (NL >= 2 ? 1 : NL - 1)) mid();
    }
}

// Inverse (Gentleman-Sande) network on the high NL window bits (free window bits below them).
template <int NL, int XLO, bool LAST, class AR, class RT, class Mid>
__device__ __forceinline__ void inv_round(const AR &ar, typename AR::T (&v)[16], u32 xt, int logn, const RT &rt, const Mid &mid) {
    constexpr int OFF = 4 - NL;
    const u32 n = 1u << logn;
#pragma unroll
    for (int s = 0; s < NL; s++) {
        const int j = OFF + s;
        const int l = XLO + j;
        if (LAST && s == NL - 1) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (k & (1 << j)) continue;
                ar.bf_inv_last(v[k], v[k | (1 << j)], l);
            }
        } else {
            const u32 tb = n - (n >> l) + (xt >> (l + 1));     // device table = inv_root_powers shifted down by one
            typename AR::TwT w[8];
            if (s < RT::PF) {
#pragma unroll
                for (int a = 0; a < 8; a++) if (a < (16 >> (j + 1))) w[a] = rt.w[s < RT::PF ? s : 0][a];
            } else ar.ldtw_n(ar.itab, tb, 16 >> (j + 1), w);
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (k & (1 << j)) continue;
                ar.bf_inv(v[k], v[k | (1 << j)], w[k >> (j + 1)], l);
            }
        }
        if (s == (NL >= 2 ? 1 : NL - 1)) mid();
    }
}

// ---- load / store functors -------------------------------------------------------------------------
// A LOAD functor names, for limb c of group g, where the limb's N source words live (src) and an
// element-wise transform applied when the staged word is read into registers: prep(g, c) is evaluated ONCE per
// work item and returns whatever per-limb constants the transform needs, xf(prep, x, raw) is evaluated per word
// (the per-limb part used to be re-evaluated per word — ptxas does not hoist it out of the unrolled loop — which made
// the fused first passes 70 % - 400 % longer in instructions, r01h SASS).  `off` is the limb's element offset in the
// intermediate buffer (LimbGroup base + c * stride); plain transforms use it for the operand as well.  Sources must be
// 16-byte aligned (cp.async).
#define B2_IDENTITY_XF                                                                   \
    struct Prep {};                                                                      \
    __device__ __forceinline__ Prep prep(int, int) const { return Prep{}; }              \
    __device__ __forceinline__ u64 xf(const Prep &, u32, u64 raw) const { return raw; }
struct PlainLoad {
    const u64 *base;
    __device__ __forceinline__ const u64 *src(int, int, u64 off) const { return base + off; }
    B2_IDENTITY_XF
};
// scattered batch: limb c of group g lives in batch item c / pc at [c % pc][g0 + g][N] of that item
struct PtrTabLoad {
    const u64 *const *tab; u32 pc, n; u64 pstride;
    __device__ __forceinline__ const u64 *src(int g, int c, u64) const {
        const u32 b = (u32)c / pc, k = (u32)c - b * pc;
        return tab[b] + (u64)k * pstride + (u64)g * n;
    }
    B2_IDENTITY_XF
};
// A STORE functor receives canonical results.  prep(g, c, off) is evaluated once per work item (row pointers, per-limb
// constants), st(prep, x, v) stores one word at limb position x, st2(prep, x, v0, v1) two adjacent words (x even).
struct PlainStore {
    u64 *base;
    struct Prep { u64 *p; };
    __device__ __forceinline__ Prep prep(int, int, u64 off) const { return Prep{base + off}; }
    __device__ __forceinline__ void operator()(const Prep &pc, u32 x, u64 v) const { pc.p[x] = v; }
    __device__ __forceinline__ void st2(const Prep &pc, u32 x, u64 v0, u64 v1) const {
        *reinterpret_cast<ulonglong2 *>(pc.p + x) = make_ulonglong2(v0, v1);
    }
};

// Tile geometry.  COL: tile = 2^A rows x 2^(12-A) columns of the 2^A x 256 matrix, local index =
// row * C + col.  ROW: tile = 16 consecutive rows, local index = position inside the tile.
template <int A, bool COL>
struct Geo {
    static constexpr int CB = COL ? (NTT_TB - A) : 0;
    __device__ static __forceinline__ u32 pos(int tile, int idx) {
        if (COL) return (u32)(((idx >> CB) << 8) + (tile << CB) + (idx & ((1 << CB) - 1)));
        return (u32)((tile << NTT_TB) | idx);
    }
    static constexpr int xbit(int b) { return COL ? (b >= CB ? b + 8 - CB : b) : b; }
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Stage one tile of a limb (N contiguous words at `limb`) into `stage`: 2048 16-byte chunks, 8 per thread.
template <int A, bool COL>
__device__ __forceinline__ void stage_tile(u64 *stage, const u64 *limb, int tile) {
    using G = Geo<A, COL>;
#pragma unroll
    for (int m = 0; m < 8; m++) {
        const int idx = 2 * ((int)threadIdx.x + NTT_THREADS * m);
        cp_async16(stage + swz16(idx), limb + G::pos(tile, idx));
    }
}

// One pass = one to three radix-16 rounds over a staged tile.  (LOr, NLr) per round; NLr == 0: round absent.
//   FIRST: the pass reads external data (xf -> from_ext); otherwise it reads the raw intermediate
//   FINAL: the pass writes canonical results; otherwise it writes the raw intermediate
// Hooks (the persistent loop's bookkeeping, woven into the pass so that none of it sits between two barriers):
//   hk.sync_stage()    after the modulus constants and the first twiddles have been requested: waits for the staged tile
//                      (cp.async group + CTA barrier); thread 0 then signals the previous item and claims the next one
//   hk.mid()           half way through round 0: thread 0 decodes the claimed item and requests its dependency counter
//   hk.pre_exchange()  before the barrier that ends round 0: thread 0 posts (next item, ready) to shared memory
//   hk.stage_free()    after that barrier: the stage buffer is free -> the next item's tile is staged
template <int A, bool COL, bool FWD, bool FIRST, bool FINAL, class AR, int LO0, int NL0, int LO1, int NL1, int LO2, int NL2,
          class Load, class Store, class Hooks>
__device__ __forceinline__ void pass_body(const u64 *stage, u64 *sm, const ModInfo &M, int tile, int g, int c, u64 off, int logn,
                                          const Load &ld, const Store &st, Hooks &hk) {
    using G = Geo<A, COL>;
    typedef typename AR::T T;
    const AR ar(M);
    const int t = threadIdx.x;
    T v[16];
    constexpr int NR = NL2 > 0 ? 3 : (NL1 > 0 ? 2 : 1);
    constexpr int LOL = NR == 3 ? LO2 : (NR == 2 ? LO1 : LO0);
    constexpr bool LASTL = FINAL && !FWD;    // the pass contains the last inverse layer
    constexpr int N1 = NL1 > 0 ? NL1 : 1, N2 = NL2 > 0 ? NL2 : 1;
    const NoHook nohook;

    RoundTw<N1, G::xbit(LO1), FWD, LASTL && NR == 2, AR> tw1;
    {   // ---- round 0: from the stage buffer
        const int it = round_base<LO0>(t);
        const u32 xt = G::pos(tile, it);
        const int sb = swz16(it);
        RoundTw<NL0, G::xbit(LO0), FWD, LASTL && NR == 1, AR> tw0;
        tw0.load(ar, xt, logn);
        if (!COL) {      // row pass: twiddles are distinct per thread — hint the ones that are loaded inside the rounds
            tw0.hint_rest(ar, xt, logn);
            if (NR > 1) tw1.hint_rest(ar, G::pos(tile, round_base<LO1>(t)), logn);
        }
        hk.sync_stage();
        u64 raw[16];
        if (LO0 == 0) {
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                const ulonglong2 d = *reinterpret_cast<const ulonglong2 *>(stage + (sb ^ k));
                raw[k] = d.x; raw[k + 1] = d.y;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 16; k++) raw[k] = stage[sb ^ swz16(k << LO0)];
        }
        if (FIRST) {
            const auto lp = ld.prep(g, c);
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = ar.from_ext(ld.xf(lp, xt + ((u32)k << G::xbit(LO0)), raw[k]));
        } else {
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = ar.from_raw(raw[k]);
        }
        if (FWD) fwd_round<NL0, G::xbit(LO0)>(ar, v, xt, logn, tw0, hk.mid);
        else inv_round<NL0, G::xbit(LO0), LASTL && NR == 1>(ar, v, xt, logn, tw0, hk.mid);
        if (NR > 1) {
            const int xb = swz(it);
#pragma unroll
            for (int k = 0; k < 16; k++) sm[xb ^ swz(k << LO0)] = ar.to_raw(v[k]);
            tw1.load(ar, G::pos(tile, round_base<LO1>(t)), logn);
        }
    }
    hk.pre_exchange();
    __syncthreads();   // every thread has consumed the stage buffer (and, for NR > 1, written the exchange buffer)
    hk.stage_free();   // -> the next item's tile is staged while this one is computed
    RoundTw<N2, G::xbit(LO2), FWD, LASTL && NR == 3, AR> tw2;
    if (NR > 1) {   // ---- round 1
        const int it = round_base<LO1>(t);
        const u32 xt = G::pos(tile, it);
        const int sb = swz(it);
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = ar.from_raw(sm[sb ^ swz(k << LO1)]);
        if (FWD) fwd_round<N1, G::xbit(LO1)>(ar, v, xt, logn, tw1, nohook);
        else inv_round<N1, G::xbit(LO1), LASTL && NR == 2>(ar, v, xt, logn, tw1, nohook);
        if (NR > 2) {     // each thread overwrites exactly the slots it has just read
#pragma unroll
            for (int k = 0; k < 16; k++) sm[sb ^ swz(k << LO1)] = ar.to_raw(v[k]);
            tw2.load(ar, G::pos(tile, round_base<LO2>(t)), logn);
        }
    }
    if (NR > 2) {   // ---- round 2
        __syncthreads();
        const int it = round_base<LO2>(t);
        const u32 xt = G::pos(tile, it);
        const int sb = swz(it);
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = ar.from_raw(sm[sb ^ swz(k << LO2)]);
        if (FWD) fwd_round<N2, G::xbit(LO2)>(ar, v, xt, logn, tw2, nohook);
        else inv_round<N2, G::xbit(LO2), LASTL>(ar, v, xt, logn, tw2, nohook);
    }
    {   // ---- to global memory
        const int it = round_base<LOL>(t);
        const u32 xt = G::pos(tile, it);
        u64 o[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (FINAL) o[k] = FWD ? ar.final_fwd(v[k]) : ar.final_inv(v[k]);
            else {
                if (!FWD) ar.pass_boundary_inv(v[k]);
                o[k] = ar.to_raw(v[k]);
            }
        }
        const auto sp = st.prep(g, c, off);
        if (G::xbit(LOL) == 0 && FINAL && !COL && NR > 1 && LOL == 0) {
            // Final row pass: a thread holds 16 adjacent words, so a direct store instruction would touch 32 different 128-byte
            // lines with 16 bytes each (and a fused store functor would read its operands the same way).  The warp's 512 words
            // are contiguous, and the warp's part of the exchange buffer is its own from here on (its last reads came from
            // exactly these 512 slots): transpose through it so that consecutive lanes store consecutive 16-byte chunks.
            const int l = t & 31;
            u64 *wr = sm + 512 * (t >> 5);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; j++)
                *reinterpret_cast<ulonglong2 *>(wr + 2 * (8 * l + (j ^ (l & 7)))) = make_ulonglong2(o[2 * j], o[2 * j + 1]);
            __syncwarp();
            const u32 xw = G::pos(tile, 512 * (t >> 5));
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const int ch = l + 32 * m, tq = ch >> 3;
                const ulonglong2 d = *reinterpret_cast<const ulonglong2 *>(wr + 2 * (8 * tq + ((ch & 7) ^ (tq & 7))));
                st.st2(sp, xw + 2u * (u32)ch, d.x, d.y);
            }
        } else if (G::xbit(LOL) == 0) {     // 16 adjacent words per thread: 16-byte stores
#pragma unroll
            for (int k = 0; k < 16; k += 2) st.st2(sp, xt + (u32)k, o[k], o[k + 1]);
        } else {
#pragma unroll
            for (int k = 0; k < 16; k++) st(sp, xt + ((u32)k << G::xbit(LOL)), o[k]);
        }
    }
}

// Round plans (window low bit, layers).  A = logN - 8 in [4, 9].
//   forward column: tile bits 11 .. CB descending;  forward row: bits 7 .. 0 descending
//   inverse row:    bits 0 .. 7 ascending;          inverse column: bits CB .. 11 ascending
template <int A> struct FwdColPlan {
    static constexpr int CB = NTT_TB - A;
    static constexpr int LO0 = 8, NL0 = 4;
    static constexpr int LO1 = (A >= 8) ? 4 : CB, NL1 = (A >= 8) ? 4 : A - 4;
    static constexpr int LO2 = CB, NL2 = (A > 8) ? A - 8 : 0;
};
template <int A> struct InvColPlan {
    static constexpr int CB = NTT_TB - A;
    static constexpr int LO0 = CB, NL0 = 4;
    static constexpr int LO1 = (A >= 8) ? CB + 4 : 8, NL1 = (A >= 8) ? 4 : A - 4;
    static constexpr int LO2 = 8, NL2 = (A > 8) ? A - 8 : 0;
};

// ---- work items ----------------------------------------------------------------------------------------
// Item order (TP = tiles per limb, D = lag, NL = limbs):  P1(0..D-1) | [P1(D+b), P2(b)] for b < NL-D | P2(NL-D..NL-1)
// where Px(l) stands for the TP tiles of pass x of limb l.  Every P2(l) item comes after all P1(l) items.
struct NttItem { u32 phase, limb, tile; };
__device__ __forceinline__ NttItem ntt_decode(const NttJob &J, u32 item) {
    const u32 ts = J.tile_shift, tm = (1u << ts) - 1, D = J.lag;
    NttItem r;
    if (item < (D << ts)) { r.phase = 0; r.limb = item >> ts; r.tile = item & tm; return r; }
    u32 rem = item - (D << ts);
    const u32 mid_items = (J.n_limbs - D) << (ts + 1);
    if (rem < mid_items) {
        const u32 b = rem >> (ts + 1), w = rem & ((2u << ts) - 1);
        r.phase = w >> ts; r.tile = w & tm; r.limb = r.phase ? b : D + b;
        return r;
    }
    rem -= mid_items;
    r.phase = 1; r.limb = J.n_limbs - D + (rem >> ts); r.tile = rem & tm;
    return r;
}

// arithmetic path of a limb group, decided on the host (run_ntt) and carried in the upper half of LimbGroup::table so that
// the kernel does not wait on a ModInfo load before it can branch
constexpr u32 PATH_INT = 0, PATH_INT_LAZY = 1, PATH_F64 = 2;
__host__ __device__ __forceinline__ u32 group_table(u32 tagged) { return tagged & 0xffffu; }
__host__ __device__ __forceinline__ u32 group_path(u32 tagged, bool fwd) { return (tagged >> (fwd ? 16 : 20)) & 0xfu; }
__host__ __device__ __forceinline__ u32 tag_table(u32 table, u32 fwd_path, u32 inv_path) { return table | (fwd_path << 16) | (inv_path << 20); }

template <int A, bool COL, bool FWD, bool FIRST, bool FINAL, int LO0, int NL0, int LO1, int NL1, int LO2, int NL2, class Load, class Store, class Hooks>
__device__ __forceinline__ void pass_dispatch(u32 path, const u64 *stage, u64 *xch, const ModInfo &M, int tile, int g, int c, u64 off, int logn,
                                              const Load &ld, const Store &st, Hooks &hk) {
    if (path == PATH_F64)
        pass_body<A, COL, FWD, FIRST, FINAL, ArithF64, LO0, NL0, LO1, NL1, LO2, NL2>(stage, xch, M, tile, g, c, off, logn, ld, st, hk);
    else if (path == PATH_INT_LAZY)
        pass_body<A, COL, FWD, FIRST, FINAL, ArithInt<true>, LO0, NL0, LO1, NL1, LO2, NL2>(stage, xch, M, tile, g, c, off, logn, ld, st, hk);
    else
        pass_body<A, COL, FWD, FIRST, FINAL, ArithInt<false>, LO0, NL0, LO1, NL1, LO2, NL2>(stage, xch, M, tile, g, c, off, logn, ld, st, hk);
}

__device__ __forceinline__ u32 ld_relaxed(const u32 *p) {      // L2 load, no L1 invalidation (cf. ld.acquire -> CCTL.IVALL)
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release(u32 *p) {          // MEMBAR.ALL.GPU + REDG: orders the CTA's stores (seen through the barrier) before the signal
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(p) : "memory");
}
__device__ __forceinline__ u32 claim(u32 *queue) {
    u32 v;
    asm volatile("atom.relaxed.gpu.global.add.u32 %0, [%1], 1;\n" : "=r"(v) : "l"(queue) : "memory");
    return v;
}

template <class F0, class F1, class F2, class F3>
struct PassHooks { F0 sync_stage; F1 mid; F2 pre_exchange; F3 stage_free; };
template <class F0, class F1, class F2, class F3>
__device__ __forceinline__ PassHooks<F0, F1, F2, F3> make_hooks(F0 a, F1 b, F2 c, F3 d) { return PassHooks<F0, F1, F2, F3>{a, b, c, d}; }

// The persistent transform kernel.  FWD: first pass = column pass (reads the operand through `ld`), second
// pass = row pass (writes through `st`); inverse: first pass = row pass, second pass = column pass.
// OCC = resident CTAs per SM the register allocation is held to.
//
// Synchronisation between CTAs: done[l] counts the finished first-pass tiles of limb l.  A producer signals with
// red.release after a CTA barrier; a consumer polls with a relaxed L2 load, passes a CTA barrier and then reads the
// intermediate with cp.async.cg (L2, never L1), so no L1 invalidation is needed (ld.acquire would flush the SM's L1 — and
// with it the twiddles — once per item).
// Per item there are two CTA barriers (tile staged / exchange) and the queue bookkeeping of thread 0 is spread over the
// item: claim after the first barrier, dependency load in the middle of round 0, hand-over at the second barrier.
template <int A, bool FWD, int OCC, class Load, class Store>
__global__ void __launch_bounds__(NTT_THREADS, OCC) ntt_persistent_kernel(const __grid_constant__ NttJob J, Load ld, Store st) {
    extern __shared__ __align__(128) u64 smem[];
    __shared__ u32 s_item, s_ready;
    u64 *const stage = smem, *const xch = smem + NTT_TILE;
    const int t = threadIdx.x;
    const u32 total = J.total_items, tiles = 1u << J.tile_shift;
    u32 *const queue = J.sync, *const finished = J.sync + 1, *const done = J.sync + 2;
    const PlainLoad mid_ld{J.mid};
    const PlainStore mid_st{J.mid};

    // limb l -> (group, limb in group): consecutive limbs cycle through the groups, so CTAs working on FP64-pipe
    // moduli and on integer-pipe moduli are resident on an SM at the same time and the two pipes overlap
    auto stage_item = [&](const NttItem &it) {
        const u32 c = it.limb / (u32)J.n_groups, g = it.limb - c * (u32)J.n_groups;
        const LimbGroup &grp = J.groups[g];
        const u64 off = limb_offset(J, grp, c);
        if (it.phase == 0) {
            const u64 *src = ld.src(J.g_base + (int)g, (int)c, off);
            stage_tile<A, FWD>(stage, src, (int)it.tile);          // forward: column tile; inverse: row tile
        } else {
            stage_tile<A, !FWD>(stage, J.mid + off, (int)it.tile);
        }
        cp_async_commit();
    };

    // ---- prologue: claim the first item and stage it (blocking on its dependency)
    if (t == 0) {
        const u32 i0 = claim(queue);
        s_item = i0;
        if (i0 < total) {
            const NttItem it = ntt_decode(J, i0);
            if (it.phase) while (ld_relaxed(done + it.limb) < tiles) __nanosleep(100);
        }
    }
    __syncthreads();
    const u32 first_id = s_item;
    if (first_id < total) {
        NttItem cur = ntt_decode(J, first_id), nxt = cur;
        stage_item(cur);
        __syncthreads();              // s_item has been read by everyone before thread 0 posts the next one
        // thread-0 bookkeeping registers
        bool publish = false; u32 publish_limb = 0;       // a finished first-pass item whose completion is not yet signalled
        u32 nx = 0, dep = 0; bool has_dep = false;
        u32 nxt_id = 0; bool pre = false;
        auto hooks = make_hooks(
            [&]() {   // sync_stage
                cp_async_wait_all();
                __syncthreads();      // tile staged and visible; the previous item's stores and exchange reads are complete CTA-wide
                if (t == 0) {
                    if (publish) red_release(done + publish_limb);
                    nx = claim(queue);
                }
            },
            [&]() {   // mid
                if (t == 0) {
                    has_dep = false;
                    if (nx < total) {
                        const NttItem ni = ntt_decode(J, nx);
                        if (ni.phase) { has_dep = true; dep = ld_relaxed(done + ni.limb); }
                    }
                }
            },
            [&]() {   // pre_exchange
                if (t == 0) { s_item = nx; s_ready = (nx < total && (!has_dep || dep >= tiles)) ? 1u : 0u; }
            },
            [&]() {   // stage_free
                nxt_id = s_item; pre = s_ready != 0;
                if (nxt_id < total) nxt = ntt_decode(J, nxt_id);
                if (pre) stage_item(nxt);
            });
        while (true) {
            {
                const u32 c = cur.limb / (u32)J.n_groups, g = cur.limb - c * (u32)J.n_groups;
                const LimbGroup &grp = J.groups[g];
                const u64 off = limb_offset(J, grp, c);
                const ModInfo &M = J.mods[group_table(grp.table)];
                const u32 path = group_path(grp.table, FWD);
                const int ga = J.g_base + (int)g;
                if (FWD) {
                    using Pl = FwdColPlan<A>;
                    if (cur.phase == 0)
                        pass_dispatch<A, true, true, true, false, Pl::LO0, Pl::NL0, Pl::LO1, Pl::NL1, Pl::LO2, Pl::NL2>(path, stage, xch, M, (int)cur.tile, ga, (int)c, off, J.logn, ld, mid_st, hooks);
                    else
                        pass_dispatch<8, false, true, false, true, 4, 4, 0, 4, 0, 0>(path, stage, xch, M, (int)cur.tile, ga, (int)c, off, J.logn, mid_ld, st, hooks);
                } else {
                    using Pl = InvColPlan<A>;
                    if (cur.phase == 0)
                        pass_dispatch<8, false, false, true, false, 0, 4, 4, 4, 0, 0>(path, stage, xch, M, (int)cur.tile, ga, (int)c, off, J.logn, ld, mid_st, hooks);
                    else
                        pass_dispatch<A, true, false, false, true, Pl::LO0, Pl::NL0, Pl::LO1, Pl::NL1, Pl::LO2, Pl::NL2>(path, stage, xch, M, (int)cur.tile, ga, (int)c, off, J.logn, mid_ld, st, hooks);
                }
            }
            publish = cur.phase == 0; publish_limb = cur.limb;
            if (nxt_id >= total) break;
            if (!pre) {               // dependency not met at prefetch time: signal our own completion, then wait
                __syncthreads();
                if (t == 0) {
                    if (publish) red_release(done + publish_limb);
                    while (ld_relaxed(done + nxt.limb) < tiles) __nanosleep(100);
                }
                publish = false;
                __syncthreads();
                stage_item(nxt);
            }
            cur = nxt;
        }
        __syncthreads();
        if (t == 0 && publish) red_release(done + publish_limb);
    }
    // ---- teardown: the last CTA to finish resets the queue and the counters for the next launch
    if (t == 0) {
        __threadfence();
        s_ready = (atomicAdd(finished, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_ready) {
        for (u32 i = t; i < J.n_limbs; i += NTT_THREADS) done[i] = 0;
        if (t == 0) { *queue = 0; *finished = 0; }
    }
}

}  // namespace b2
