// b2_fused.cuh — load / store functors shared by the fused transforms of keyswitch mod-down and
// rescale / modswitch (rounding division by the last modulus).
#pragma once
#include "b2_ntt.cuh"
#include "../../include/b2rns.h"

namespace b2 {

// ---- step 5 functors ---------------------------------------------------------------------------------
// BFV/CKKS: t = ((u + P/2) mod P) mod q_j + q_j - (P/2 mod q_j)      (ski_util6, :570-598)
// BGV:      t = ((-u P^-1 mod t) mod q_j) * P + u  mod q_j           (ski_util5 step 1, :436-474)
struct ModDownTerm {
    const u64 *u;         // [count][N]: coefficient form of the limb being divided out
    const ModInfo *mods;
    u32 last, n;     // last: index of the modulus being divided out (P for keyswitch, q_{nl-1} for rescale)
    int bgv;
    u64 t, t_ratio, t_r0, t_r1, pinv_t;
    __device__ __forceinline__ u64 term(int j, int k, u32 x) const { return term_of(j, u[(u64)k * n + x]); }
    __device__ __forceinline__ u64 term_of(int j, u64 uv) const {
        const ModInfo &Q = mods[j];
        const u64 q = Q.p, P = mods[last].p;
        if (!bgv) {
            const u64 half = P >> 1;
            u64 r = uv + half; r = r >= P ? r - P : r;
            const u64 rq = P > q ? barrett64(r, q, Q.ratio) : r;
            return rq + q - barrett64(half, q, Q.ratio);
        }
        u64 kt = barrett64(uv, t, t_ratio);
        kt = kt ? t - kt : 0;
        if (pinv_t != 1) kt = mulmod(kt, pinv_t, t, t_r0, t_r1);
        const u64 a = mulmod(barrett64(kt, q, Q.ratio), barrett64(P, q, Q.ratio), q, Q.r0, Q.r1);
        return addmod(a, barrett64(uv, q, Q.ratio), q);
    }
};
struct ModDownLoad {      // NTT-form path: the transform's input IS the term
    ModDownTerm T;
    u32 jbase = 0;        // modulus index of group 0 (limb-sharded calls transform a sub-range of the moduli)
    __device__ __forceinline__ const u64 *src(int, int c, u64) const { return T.u + (u64)c * T.n; }
    __device__ __forceinline__ u64 xf(int g, int, u32, u64 raw) const { return T.term_of(g + (int)jbase, raw); }
};
// batches are passed to the kernels BY VALUE as small pointer tables (no per-call device upload of pointer arrays,
// cf. the reference's construct_batch, src/utils/box_batch.h:222-247); larger batches are processed in chunks
constexpr int KS_MAXB = 16;
struct PtrTab { const u64 *p[KS_MAXB]; };
struct PtrTabW { u64 *p[KS_MAXB]; };

// final store: out = (s - t) * P^-1 mod q_j ; then Overwrite / AddInplace / OverwriteExceptFirst.
// limb index c = 2 b + k (ciphertext b of the chunk, polynomial k)
struct ModDownStore {
    ModDownTerm T;
    const u64 *prod;      // [B][2][R][N]
    const Tw *pinv;       // [K] : P^-1 mod q_j  (row K-1 of the inverse table)
    PtrTabW dest;         // per ciphertext: [2][decomp][N]
    u32 decomp, assign;
    int t_from_transform; // 1: transformed value is t (NTT-form ciphertext); 0: transformed value is s, t computed here
    // geometry (defaults = the whole modulus range on one GPU): group g is modulus jbase + g, row g of the prod_rows rows
    // of each polynomial of `prod` and row g of the dest_rows rows of each polynomial of the destination
    u32 jbase, prod_rows, dest_rows;
    __device__ __forceinline__ u64 one(int j, u64 q, u64 s, u64 v, u64 uv, u64 d, bool add) const {
        u64 t;
        if (t_from_transform) t = v;
        else { s = v; t = csub(T.term_of(j, uv), q); }           // term < 2q -> canonical
        const u64 r = shoup(submod(s, t, q), pinv[j], q);
        return add ? addmod(d, r, q) : r;
    }
    __device__ __forceinline__ void operator()(int g, int c, u64, u32 x, u64 v) const {
        const int j = g + (int)jbase;
        const u64 q = T.mods[j].p, n = T.n;
        const int k = c & 1;
        const bool add = assign == B2_ASSIGN_ADD_INPLACE || (k == 0 && assign == B2_ASSIGN_OVERWRITE_EXCEPT_FIRST);
        u64 *d = dest.p[c >> 1] + ((u64)k * dest_rows + g) * n + x;
        // read-only-path loads (ld.global.nc): every word is read at most once, before this thread's own store to it, and by no
        // other thread — and the compiler may hoist the loads of the following registers above this register's store
        const u64 s = t_from_transform ? __ldg(prod + ((u64)c * prod_rows + g) * n + x) : 0;
        const u64 uv = t_from_transform ? 0 : __ldg(T.u + (u64)c * n + x);
        *d = one(j, q, s, v, uv, add ? __ldg(d) : 0, add);
    }
    // two adjacent coefficients: 16-byte accesses to prod / u / dest
    __device__ __forceinline__ void st2(int g, int c, u64, u32 x, u64 v0, u64 v1) const {
        const int j = g + (int)jbase;
        const u64 q = T.mods[j].p, n = T.n;
        const int k = c & 1;
        const bool add = assign == B2_ASSIGN_ADD_INPLACE || (k == 0 && assign == B2_ASSIGN_OVERWRITE_EXCEPT_FIRST);
        ulonglong2 *d = reinterpret_cast<ulonglong2 *>(dest.p[c >> 1] + ((u64)k * dest_rows + g) * n + x);
        ulonglong2 s = make_ulonglong2(0, 0), uv = make_ulonglong2(0, 0), dv = make_ulonglong2(0, 0);
        if (t_from_transform) s = __ldg(reinterpret_cast<const ulonglong2 *>(prod + ((u64)c * prod_rows + g) * n + x));
        else uv = __ldg(reinterpret_cast<const ulonglong2 *>(T.u + (u64)c * n + x));
        if (add) dv = __ldg(d);
        *d = make_ulonglong2(one(j, q, s.x, v0, uv.x, dv.x, add), one(j, q, s.y, v1, uv.y, dv.y, add));
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
    __device__ __forceinline__ void operator()(int i, int p, u64, u32 x, u64 t) const {
        const u64 q = mods[i].p;
        const u64 s = __ldg(in + ((u64)p * nl + i) * n + x);
        out[((u64)p * (nl - 1) + i) * n + x] = shoup(submod(s, t, q), inv[i], q);
    }
    __device__ __forceinline__ void st2(int i, int p, u64, u32 x, u64 t0, u64 t1) const {
        const u64 q = mods[i].p;
        const ulonglong2 s = __ldg(reinterpret_cast<const ulonglong2 *>(in + ((u64)p * nl + i) * n + x));
        *reinterpret_cast<ulonglong2 *>(out + ((u64)p * (nl - 1) + i) * n + x) =
            make_ulonglong2(shoup(submod(s.x, t0, q), inv[i], q), shoup(submod(s.y, t1, q), inv[i], q));
    }
};
struct LastLimbLoad {     // the limb being divided out, for its inverse transform
    const u64 *in; u32 nl, n;
    __device__ __forceinline__ const u64 *src(int, int p, u64) const { return in + ((u64)p * nl + nl - 1) * n; }
    __device__ __forceinline__ u64 xf(int, int, u32, u64 raw) const { return raw; }
};

}  // namespace b2
