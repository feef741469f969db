// b2_fused.cuh — load / store functors shared by the fused transforms of keyswitch mod-down and
// rescale / modswitch (rounding division by the last modulus).
#pragma once
#include "b2_ntt.cuh"
#include "../../include/b2rns.h"

namespace b2 {

// ---- step 5 functors ---------------------------------------------------------------------------------
// BFV/CKKS: t = ((u + P/2) mod P) mod q_j + q_j - (P/2 mod q_j)      (ski_util6, :570-598)
// BGV:      t = ((-u P^-1 mod t) mod q_j) * P + u  mod q_j           (ski_util5 step 1, :436-474)
// BGV is a template parameter (and the per-modulus constants are computed once per work item, Pre) so that a kernel
// carries only the formula it uses: with run-time flags every instantiation contained both, per word, and the mod-down
// transform's code grew to 42 k SASS instructions — past the instruction caches (r01h).
template <bool BGV>
struct ModDownTerm {
    const u64 *u;         // [count][N]: coefficient form of the limb being divided out
    const ModInfo *mods;
    u32 last, n;     // last: index of the modulus being divided out (P for keyswitch, q_{nl-1} for rescale)
    u64 t, t_ratio, t_r0, t_r1, pinv_t;
    struct Pre { u64 q, ratio, r0, r1, P, half, c; bool big; };      // c: CKKS/BFV q - (P/2 mod q) ; BGV P mod q
    __device__ __forceinline__ Pre pre(int j) const {
        const ModInfo &Q = mods[j];
        Pre r;
        r.q = Q.p; r.ratio = Q.ratio; r.r0 = Q.r0; r.r1 = Q.r1; r.P = mods[last].p; r.half = r.P >> 1; r.big = r.P > r.q;
        r.c = BGV ? barrett64(r.P, r.q, r.ratio) : r.q - barrett64(r.half, r.q, r.ratio);
        return r;
    }
    __device__ __forceinline__ u64 apply(const Pre &c, u64 uv) const {
        if (!BGV) {
            u64 r = uv + c.half; r = r >= c.P ? r - c.P : r;
            const u64 rq = c.big ? barrett64(r, c.q, c.ratio) : r;
            return rq + c.c;
        }
        u64 kt = barrett64(uv, t, t_ratio);
        kt = kt ? t - kt : 0;
        if (pinv_t != 1) kt = mulmod(kt, pinv_t, t, t_r0, t_r1);
        const u64 a = mulmod(barrett64(kt, c.q, c.ratio), c.c, c.q, c.r0, c.r1);
        return addmod(a, barrett64(uv, c.q, c.ratio), c.q);
    }
    __device__ __forceinline__ u64 term_of(int j, u64 uv) const { return apply(pre(j), uv); }
};
template <bool BGV>
struct ModDownLoad {      // NTT-form path: the transform's input IS the term
    ModDownTerm<BGV> T;
    u32 jbase = 0;        // modulus index of group 0 (limb-sharded calls transform a sub-range of the moduli)
    typedef typename ModDownTerm<BGV>::Pre Prep;
    __device__ __forceinline__ const u64 *src(int, int c, u64) const { return T.u + (u64)c * T.n; }
    __device__ __forceinline__ Prep prep(int g, int) const { return T.pre(g + (int)jbase); }
    __device__ __forceinline__ u64 xf(const Prep &pc, u32, u64 raw) const { return T.apply(pc, raw); }
};
// batches are passed to the kernels BY VALUE as small pointer tables (no per-call device upload of pointer arrays,
// cf. the reference's construct_batch, src/utils/box_batch.h:222-247); larger batches are processed in chunks
constexpr int KS_MAXB = 128;
struct PtrTab { const u64 *p[KS_MAXB]; };
struct PtrTabW { u64 *p[KS_MAXB]; };

// final store: out = (s - t) * P^-1 mod q_j ; then Overwrite / AddInplace / OverwriteExceptFirst.
// limb index c = 2 b + k (ciphertext b of the chunk, polynomial k)
template <bool BGV, bool TFROM>      // TFROM: the transformed value is t (NTT-form ciphertext); else it is s and t is computed here
struct ModDownStore {
    ModDownTerm<BGV> T;
    const u64 *prod;      // [B][2][R][N]
    const Tw *pinv;       // [K] : P^-1 mod q_j  (row K-1 of the inverse table)
    PtrTabW dest;         // per ciphertext: [2][decomp][N]
    u32 decomp, assign;
    // geometry (defaults = the whole modulus range on one GPU): group g is modulus jbase + g, row g of the prod_rows rows
    // of each polynomial of `prod` and row g of the dest_rows rows of each polynomial of the destination
    u32 jbase, prod_rows, dest_rows;
    // per work item: the limb's modulus, scaling constant and row pointers (computed once, not per word)
    struct Prep {
        u64 q; Tw pinv; bool add;
        u64 *d; const u64 *s, *uu;                     // rows of dest, prod, u for this (modulus, ciphertext, polynomial)
        typename ModDownTerm<BGV>::Pre tp;              // only used when the term is computed here (!TFROM)
    };
    __device__ __forceinline__ Prep prep(int g, int c, u64) const {
        const int j = g + (int)jbase, k = c & 1;
        const u64 n = T.n;
        Prep r;
        r.q = T.mods[j].p; r.pinv = pinv[j];
        r.add = assign == B2_ASSIGN_ADD_INPLACE || (k == 0 && assign == B2_ASSIGN_OVERWRITE_EXCEPT_FIRST);
        r.d = dest.p[c >> 1] + ((u64)k * dest_rows + g) * n;
        r.s = prod + ((u64)c * prod_rows + g) * n;
        r.uu = T.u + (u64)c * n;
        if (!TFROM) r.tp = T.pre(j);
        return r;
    }
    __device__ __forceinline__ u64 one(const Prep &pc, u64 s, u64 v, u64 uv, u64 d) const {
        u64 t;
        if (TFROM) t = v;
        else { s = v; t = csub(T.apply(pc.tp, uv), pc.q); }           // term < 2q -> canonical
        const u64 r = shoup(submod(s, t, pc.q), pc.pinv, pc.q);
        return pc.add ? addmod(d, r, pc.q) : r;
    }
    // read-only-path loads (ld.global.nc): every word is read at most once, before this thread's own store to it, and by no
    // other thread
    __device__ __forceinline__ void operator()(const Prep &pc, u32 x, u64 v) const {
        const u64 s = TFROM ? __ldg(pc.s + x) : 0;
        const u64 uv = TFROM ? 0 : __ldg(pc.uu + x);
        pc.d[x] = one(pc, s, v, uv, pc.add ? __ldg(pc.d + x) : 0);
    }
    // two adjacent coefficients: 16-byte accesses to prod / u / dest
    __device__ __forceinline__ void st2(const Prep &pc, u32 x, u64 v0, u64 v1) const {
        ulonglong2 *d = reinterpret_cast<ulonglong2 *>(pc.d + x);
        ulonglong2 s = make_ulonglong2(0, 0), uv = make_ulonglong2(0, 0), dv = make_ulonglong2(0, 0);
        if (TFROM) s = __ldg(reinterpret_cast<const ulonglong2 *>(pc.s + x));
        else uv = __ldg(reinterpret_cast<const ulonglong2 *>(pc.uu + x));
        if (pc.add) dv = __ldg(d);
        *d = make_ulonglong2(one(pc, s.x, v0, uv.x, dv.x), one(pc, s.y, v1, uv.y, dv.y));
    }
};

// rescale / modswitch final store: out[p][i] = (in[p][i] - t) * q_last^-1 mod q_i
// (src/utils/rns_tool.cu:629-644 CKKS step 2, :1680-1711 BGV step 2)
struct RescaleStore {
    const u64 *in;        // [count][nl][N]
    u64 *out;             // [count][nl-1][N]
    const ModInfo *mods;
    const Tw *inv;        // row `last` of the inverse table: q_last^-1 mod q_i
    u32 nl, n;
    struct Prep { u64 q; Tw inv; const u64 *s; u64 *o; };
    __device__ __forceinline__ Prep prep(int i, int p, u64) const {
        return Prep{mods[i].p, inv[i], in + ((u64)p * nl + i) * n, out + ((u64)p * (nl - 1) + i) * n};
    }
    __device__ __forceinline__ void operator()(const Prep &pc, u32 x, u64 t) const {
        pc.o[x] = shoup(submod(__ldg(pc.s + x), t, pc.q), pc.inv, pc.q);
    }
    __device__ __forceinline__ void st2(const Prep &pc, u32 x, u64 t0, u64 t1) const {
        const ulonglong2 s = __ldg(reinterpret_cast<const ulonglong2 *>(pc.s + x));
        *reinterpret_cast<ulonglong2 *>(pc.o + x) = make_ulonglong2(shoup(submod(s.x, t0, pc.q), pc.inv, pc.q), shoup(submod(s.y, t1, pc.q), pc.inv, pc.q));
    }
};
struct LastLimbLoad {     // the limb being divided out, for its inverse transform
    const u64 *in; u32 nl, n;
    __device__ __forceinline__ const u64 *src(int, int p, u64) const { return in + ((u64)p * nl + nl - 1) * n; }
    B2_IDENTITY_XF
};

}  // namespace b2
