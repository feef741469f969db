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
//     order.  Each pass moves a 4096-coefficient tile (32 KiB) per CTA: 16 columns x 2^A rows
//     (128-byte row segments) or 16 whole rows.
//   * Inside a pass every thread keeps 16 coefficients in registers and performs a radix-16
//     (four-layer) butterfly network on them; the tile is re-distributed between radix-16 rounds
//     through XOR-swizzled shared memory (conflict-free for every round's access pattern).  An
//     8-layer pass therefore touches shared memory once and global memory once in / once out.
//   * Two arithmetic back ends (b2_modarith.cuh): ArithF64 — exact double-double modular multiply on
//     the FP64 pipe for p < 2^41, signed lazy values, 8 DP instructions per butterfly; ArithInt —
//     Harvey/Shoup lazy butterflies on the integer pipe for larger p, with no per-layer correction
//     at all when the modulus leaves enough headroom (ModInfo::fwd_lazy / inv_lazy).
//   * N^-1 is folded into the last inverse layer's twiddles (no extra pass).
//   * Limbs that share a modulus are processed by the same CTA back to back ("limb groups"), so
//     the twiddles of a tile are fetched once per group chunk and then hit in L1.
//   * Loads and stores go through functors so callers can fuse a prologue / epilogue into the
//     transform (keyswitch mod-up, key inner product, rescale, mod-down: b2_keyswitch.cu).
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

struct NttLaunch {
    const ModInfo *mods;
    int logn;
    int chunk;        // limbs of one group handled by one CTA
    int c_begin, c_end;   // absolute limb-index range of this launch (L2 wave)
    int g_base;           // index of groups[0] in the caller's group list (functors see absolute indices)
    int n_groups;
    LimbGroup groups[NTT_MAX_GROUPS];
};

// XOR swizzle of a tile-local coefficient index (8-byte units).  It is GF(2)-linear, so
// swz(a | b) == swz(a) ^ swz(b) for disjoint a, b: per-register offsets are compile-time constants.
// For every round's access pattern a half-warp lands on 16 distinct 8-byte slots.
__host__ __device__ constexpr int swz(int idx) { return idx ^ ((idx >> 4) & 15); }

// tile-local index of register 0 of thread t in a round whose 4-bit window starts at bit LO
template <int LO>
__device__ __forceinline__ int round_base(int t) { return ((t >> LO) << (LO + 4)) | (t & ((1 << LO) - 1)); }

// ---- arithmetic back ends ----------------------------------------------------------------------
template <bool LAZY>
struct ArithInt {
    typedef u64 T;
    typedef Tw TwT;
    u64 p, p2, ratio;
    Tw inv_n, inv_n_w;
    const Tw *ftab, *itab;
    __device__ __forceinline__ explicit ArithInt(const ModInfo &M)
        : p(M.p), p2(M.p << 1), ratio(M.ratio), inv_n(M.inv_n), inv_n_w(M.inv_n_w), ftab(M.tw), itab(M.itw) {}
    __device__ __forceinline__ T from_ext(u64 x) const { return x; }
    __device__ __forceinline__ T from_raw(u64 x) const { return x; }
    __device__ __forceinline__ u64 to_raw(T v) const { return v; }
    __device__ __forceinline__ TwT ldtw(const TwT *t, u32 i) const { return __ldg(t + i); }
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
    double p, pinv, inv_n, inv_n_w;
    const double *ftab, *itab;
    __device__ __forceinline__ explicit ArithF64(const ModInfo &M)
        : p(M.pd), pinv(M.pinv), inv_n(M.inv_n_d), inv_n_w(M.inv_n_w_d), ftab(M.twd), itab(M.itwd) {}
    __device__ __forceinline__ T from_ext(u64 x) const { return u64_to_dd(x); }
    __device__ __forceinline__ T from_raw(u64 x) const { return __longlong_as_double((long long)x); }
    __device__ __forceinline__ u64 to_raw(T v) const { return (u64)__double_as_longlong(v); }
    __device__ __forceinline__ TwT ldtw(const TwT *t, u32 i) const { return __ldg(t + i); }
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
// Forward (Cooley-Tukey) network on the low NL window bits (free window bits above them).
// xt: limb-local position of register 0 (window bits clear); XLO: limb-local position of window bit 0.
template <int NL, int XLO, class AR>
__device__ __forceinline__ void fwd_round(const AR &ar, typename AR::T (&v)[16], u32 xt, int logn) {
#pragma unroll
    for (int s = 0; s < NL; s++) {
        const int j = NL - 1 - s;
        const int l = logn - 1 - (XLO + j);
        const u32 tb = (1u << l) + (xt >> (XLO + j + 1));
        typename AR::TwT w[8];
#pragma unroll
        for (int i = 0; i < (16 >> (j + 1)); i++) w[i] = ar.ldtw(ar.ftab, tb + i);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (k & (1 << j)) continue;
            ar.bf_fwd(v[k], v[k | (1 << j)], w[k >> (j + 1)]);
        }
    }
}

// Inverse (Gentleman-Sande) network on the high NL window bits (free window bits below them).
template <int NL, int XLO, bool LAST, class AR>
__device__ __forceinline__ void inv_round(const AR &ar, typename AR::T (&v)[16], u32 xt, int logn) {
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
            const u32 tb = n - (n >> l) + 1 + (xt >> (l + 1));
            typename AR::TwT w[8];
#pragma unroll
            for (int i = 0; i < (16 >> (j + 1)); i++) w[i] = ar.ldtw(ar.itab, tb + i);
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (k & (1 << j)) continue;
                ar.bf_inv(v[k], v[k | (1 << j)], w[k >> (j + 1)], l);
            }
        }
    }
}

// ---- default load / store functors ------------------------------------------------------------
// ld(group, limb_in_group, element offset of the limb, x) -> external (integer) coefficient at position x
struct PlainLoad {
    const u64 *base;
    __device__ __forceinline__ u64 operator()(int, int, u64 off, u32 x) const { return base[off + x]; }
};
struct PlainStore {
    u64 *base;
    __device__ __forceinline__ void operator()(int, int, u64 off, u32 x, u64 v) const { base[off + x] = v; }
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

// One pass = one to three radix-16 rounds over a tile.  (LOr, NLr) per round; NLr == 0: round absent.
//   FIRST: the pass reads external data (ld -> from_ext); otherwise it reads the raw intermediate
//   FINAL: the pass writes canonical results; otherwise it writes the raw intermediate
template <int A, bool COL, bool FWD, bool FIRST, bool FINAL, class AR, int LO0, int NL0, int LO1, int NL1, int LO2, int NL2,
          class Load, class Store>
__device__ __forceinline__ void pass_body(u64 *sm, const ModInfo &M, int tile, int g, int c, u64 off, int logn,
                                          const Load &ld, const Store &st) {
    using G = Geo<A, COL>;
    typedef typename AR::T T;
    const AR ar(M);
    const int t = threadIdx.x;
    T v[16];
    constexpr int NR = NL2 > 0 ? 3 : (NL1 > 0 ? 2 : 1);
    constexpr int LOL = NR == 3 ? LO2 : (NR == 2 ? LO1 : LO0);
    constexpr bool LASTL = FINAL && !FWD;    // the pass contains the last inverse layer

    {   // ---- round 0: from global memory
        const int it = round_base<LO0>(t);
        const u32 xt = G::pos(tile, it);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            u64 raw = ld(g, c, off, xt + ((u32)k << G::xbit(LO0)));
            v[k] = FIRST ? ar.from_ext(raw) : ar.from_raw(raw);
        }
        if (FWD) fwd_round<NL0, G::xbit(LO0)>(ar, v, xt, logn);
        else inv_round<NL0, G::xbit(LO0), LASTL && NR == 1>(ar, v, xt, logn);
        if (NR > 1) {
            const int sb = swz(it);
#pragma unroll
            for (int k = 0; k < 16; k++) sm[sb ^ swz(k << LO0)] = ar.to_raw(v[k]);
        }
    }
    if (NR > 1) {   // ---- round 1
        __syncthreads();
        const int it = round_base<LO1>(t);
        const u32 xt = G::pos(tile, it);
        const int sb = swz(it);
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = ar.from_raw(sm[sb ^ swz(k << LO1)]);
        if (FWD) fwd_round<(NL1 > 0 ? NL1 : 1), G::xbit(LO1)>(ar, v, xt, logn);
        else inv_round<(NL1 > 0 ? NL1 : 1), G::xbit(LO1), LASTL && NR == 2>(ar, v, xt, logn);
        if (NR > 2) {
#pragma unroll
            for (int k = 0; k < 16; k++) sm[sb ^ swz(k << LO1)] = ar.to_raw(v[k]);
        }
    }
    if (NR > 2) {   // ---- round 2
        __syncthreads();
        const int it = round_base<LO2>(t);
        const u32 xt = G::pos(tile, it);
        const int sb = swz(it);
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = ar.from_raw(sm[sb ^ swz(k << LO2)]);
        if (FWD) fwd_round<(NL2 > 0 ? NL2 : 1), G::xbit(LO2)>(ar, v, xt, logn);
        else inv_round<(NL2 > 0 ? NL2 : 1), G::xbit(LO2), LASTL>(ar, v, xt, logn);
    }
    {   // ---- to global memory
        const int it = round_base<LOL>(t);
        const u32 xt = G::pos(tile, it);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            u64 o;
            if (FINAL) o = FWD ? ar.final_fwd(v[k]) : ar.final_inv(v[k]);
            else {
                if (!FWD) ar.pass_boundary_inv(v[k]);
                o = ar.to_raw(v[k]);
            }
            st(g, c, off, xt + ((u32)k << G::xbit(LOL)), o);
        }
    }
}

template <int A, bool COL, bool FWD, bool FIRST, bool FINAL, int LO0, int NL0, int LO1, int NL1, int LO2, int NL2, class Load,
          class Store>
__global__ void __launch_bounds__(NTT_THREADS) ntt_pass_kernel(const __grid_constant__ NttLaunch P, Load ld, Store st) {
    __shared__ __align__(16) u64 sm[NTT_TILE];
    const LimbGroup &grp = P.groups[blockIdx.y];
    const int c0 = P.c_begin + blockIdx.z * P.chunk;
    const int cend = min((int)grp.count, P.c_end);
    if (c0 >= cend) return;
    const int c1 = min(cend, c0 + P.chunk);
    const ModInfo &M = P.mods[grp.table];
    const int path = M.use_f64 ? 2 : ((FWD ? M.fwd_lazy : M.inv_lazy) ? 1 : 0);
    constexpr bool MULTI = NL1 > 0;
    for (int c = c0; c < c1; c++) {
        const u64 off = grp.base + (u64)c * grp.stride;
        if (path == 2)
            pass_body<A, COL, FWD, FIRST, FINAL, ArithF64, LO0, NL0, LO1, NL1, LO2, NL2>(sm, M, blockIdx.x, P.g_base + blockIdx.y, c, off, P.logn, ld, st);
        else if (path == 1)
            pass_body<A, COL, FWD, FIRST, FINAL, ArithInt<true>, LO0, NL0, LO1, NL1, LO2, NL2>(sm, M, blockIdx.x, P.g_base + blockIdx.y, c, off, P.logn, ld, st);
        else
            pass_body<A, COL, FWD, FIRST, FINAL, ArithInt<false>, LO0, NL0, LO1, NL1, LO2, NL2>(sm, M, blockIdx.x, P.g_base + blockIdx.y, c, off, P.logn, ld, st);
        if (MULTI && c + 1 < c1) __syncthreads();
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

}  // namespace b2
