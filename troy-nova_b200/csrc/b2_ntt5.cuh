// b2_ntt5.cuh — TMA-fed N = 2^16 transform for FP64-pipe moduli (p < 2^41): the headline kernel.
//
// Same mathematical definition as b2_ntt.cuh (reference: src/fgk/ntt_grouped.cu:159-344,489-685) and the same persistent
// work queue (one launch per call, second-pass items trail first-pass items so that the intermediate stays in L2), but the
// data movement is rebuilt around the sm_100 copy engine, because the ncu profile of the cp.async kernel (profiles/r01n_*)
// showed 58 % of its issue slots going to addressing / staging / twiddle loads rather than to FP64 butterflies:
//
//   * tiles travel by TMA: one cp.async.bulk.tensor load and one store per 32 KiB tile (SWIZZLE_128B tensor maps over
//     the operand, completion on an mbarrier) — no per-thread LDGSTS / STG, no 64-bit address arithmetic;
//   * twiddles travel with the tile: N = 256 x 256, every pass is sixteen (per warp: two) independent 256-point
//     transforms whose 255 twiddles are one 2 KiB table — per LIMB for the column pass, per ROW for the row pass
//     (row-major table copies built at context creation) — fetched by one cp.async.bulk per item into shared memory, in
//     a layout permuted so that every per-thread read is a conflict-free LDS.128.  No LDG, no long-scoreboard stalls;
//   * each warp owns two columns (column pass) or two rows (row pass) of the tile END TO END: the radix-16 exchange
//     between the two rounds happens in the warp's own 4 KiB of the stage buffer behind __syncwarp — no exchange buffer
//     and no CTA barrier inside a pass;
//   * shared-memory addresses are (per-thread base) + immediate, or one XOR with an immediate: the tensor maps of the
//     column pass are declared with permuted dimensions (col, row_hi, row_lo) where that makes the pattern additive.
//
// Shared memory per CTA: two stages of (32 KiB tile + 2 KiB twiddles) = 68 KiB -> 3 CTAs per SM, every load a whole item ahead.
#pragma once
#include <cstdio>
#include <type_traits>
#include <cuda.h>
#include "b2_ntt.cuh"

namespace b2 {

constexpr int N5_LOGN = 16;
constexpr int N5_STAGE_BYTES = 32768, N5_TW_BYTES = 4096;      // per stage: tile + 256 staged twiddles (8-byte or 16-byte entries)
constexpr int N5_SMEM_BYTES = 2 * (N5_STAGE_BYTES + N5_TW_BYTES) + 1024;      // two stages; + slack to align them to 1024 B (SWIZZLE_128B)

// first pass loads `first` (the operand), stores `mid_st`; second pass loads `mid_ld`, stores `last` (same buffer as mid)
struct Ntt5Maps { CUtensorMap first, mid_st, mid_ld, last; };

// ---- PTX wrappers ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(u64 *bar, u32 parity) {
    u32 ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol error must surface as a trapped kernel with a message, never as a hung GPU (try_wait itself blocks for
// a hardware-defined interval per call, so 2^24 failed attempts are seconds, far beyond any legitimate wait).
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity, int tag = 0) {
    for (u32 spins = 0; !mbar_try_wait(bar, parity); spins++)
        if (spins == (1u << 24)) { printf("b2 ntt5: wait %d stuck (block %d thread %d parity %u)\n", tag, (int)blockIdx.x, (int)threadIdx.x, parity); __trap(); }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(u32 dst, const CUtensorMap *m, u64 *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
                 ::"r"(dst), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(u32 dst, const CUtensorMap *m, u64 *bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                 ::"r"(dst), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(u32 dst, const CUtensorMap *m, u64 *bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n"
                 ::"r"(dst), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *m, u32 src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(m), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *m, u32 src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];\n" ::"l"(m), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *m, u32 src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n"
                 ::"l"(m), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_load(u32 dst, const void *src, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }

// explicit shared-window accesses (the generic-address forms would cost an address-space check per access)
__device__ __forceinline__ u64 lds64(u32 a) { u64 v; asm volatile("ld.shared.u64 %0, [%1];\n" : "=l"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts64(u32 a, u64 v) { asm volatile("st.shared.u64 [%0], %1;\n" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ ulonglong2 lds128(u32 a) { ulonglong2 v; asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];\n" : "=l"(v.x), "=l"(v.y) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts128(u32 a, u64 x, u64 y) { asm volatile("st.shared.v2.u64 [%0], {%1, %2};\n" ::"r"(a), "l"(x), "l"(y) : "memory"); }
// twiddle reads: the table is written only by the copy engine, between items — not volatile, so that the compiler may schedule
// them freely inside an item; `tok` (the item id) keeps it from reusing or hoisting them across items
__device__ __forceinline__ double ldsd(u32 a, u32 tok) { double v; asm("ld.shared.f64 %0, [%1]; // %2\n" : "=d"(v) : "r"(a), "r"(tok)); return v; }
__device__ __forceinline__ double2 ldsd2(u32 a, u32 tok) { double2 v; asm("ld.shared.v2.f64 {%0, %1}, [%2]; // %3\n" : "=d"(v.x), "=d"(v.y) : "r"(a), "r"(tok)); return v; }

// ---- 256-point transforms on 16 registers ---------------------------------------------------------------------------
// Twiddle table of one 256-point transform (256 entries; built by b2_ctx.cu: n5_table_slot).  Entries [1, 16): the 15 twiddles
// of the four layers whose twiddle depends on the REGISTER index only (heap order).  Then the four layers with 1 / 2 / 4 / 8
// twiddles per thread y (0..15), laid out so that the sixteen lanes of a row / column group read consecutive 16-byte chunks:
//   FP64 path (8-byte entries)                           integer path (16-byte Shoup pairs)
//   [16 + y]                                             [16 + y]
//   [32 + 2y + i]                                        [32 + 16 i + y]
//   [64 + 32 q + 2y + i]    entry 2q + i of thread y     [64 + 16 i + y]
//   [128 + 32 q + 2y + i]   entry 2q + i                 [128 + 16 i + y]
// An accessor hides where the table lives (shared memory for the column passes, global memory for the per-thread part of
// the row passes: 30 KiB per tile that stay out of shared memory; coalesced read-only loads) and the entry type.
__device__ __forceinline__ ulonglong2 lds128_tok(u32 a, u32 tok) { ulonglong2 v; asm("ld.shared.v2.u64 {%0, %1}, [%2]; // %3\n" : "=l"(v.x), "=l"(v.y) : "r"(a), "r"(tok)); return v; }

template <class TwT> struct N5Shared;
template <> struct N5Shared<double> {
    u32 T, Ty, tok;
    __device__ __forceinline__ N5Shared(u32 T_, u32 y, u32 tok_) : T(T_), Ty(T_ + 16 * y), tok(tok_) {}
    __device__ __forceinline__ double uni(int idx) const { return ldsd(T + 8 * idx, tok); }
    __device__ __forceinline__ double one() const { return ldsd(T + 8 * 16 + (Ty - T) / 2, tok); }
    template <int M, int I0, int CNT> __device__ __forceinline__ void layer(double (&w)[CNT]) const {      // entries I0 .. I0+CNT-1 of the layer with 2^M twiddles
#pragma unroll
        for (int i = 0; i < CNT; i += 2) { const double2 d = ldsd2(Ty + 16 * ((1 << (M - 1)) + 16 * ((I0 + i) >> 1)), tok); w[i] = d.x; w[i + 1] = d.y; }
    }
};
template <> struct N5Shared<Tw> {
    u32 T, Ty, tok;
    __device__ __forceinline__ N5Shared(u32 T_, u32 y, u32 tok_) : T(T_), Ty(T_ + 16 * y), tok(tok_) {}
    __device__ __forceinline__ Tw uni(int idx) const { const ulonglong2 d = lds128_tok(T + 16 * idx, tok); return d; }
    __device__ __forceinline__ Tw one() const { return lds128_tok(Ty + 16 * 16, tok); }
    template <int M, int I0, int CNT> __device__ __forceinline__ void layer(Tw (&w)[CNT]) const {
#pragma unroll
        for (int i = 0; i < CNT; i++) w[i] = lds128_tok(Ty + 16 * ((1 << M) + 16 * (I0 + i)), tok);
    }
};
template <class TwT> struct N5Global;
template <> struct N5Global<double> {
    const double *T; const double2 *T2;
    __device__ __forceinline__ N5Global(const void *T_, u32 y) : T(reinterpret_cast<const double *>(T_) + y), T2(reinterpret_cast<const double2 *>(T_) + y) {}
    __device__ __forceinline__ double one() const { return __ldg(T + 16); }
    template <int M, int I0, int CNT> __device__ __forceinline__ void layer(double (&w)[CNT]) const {
#pragma unroll
        for (int i = 0; i < CNT; i += 2) { const double2 d = __ldg(T2 + (1 << (M - 1)) + 16 * ((I0 + i) >> 1)); w[i] = d.x; w[i + 1] = d.y; }
    }
};
template <> struct N5Global<Tw> {
    const Tw *T;
    __device__ __forceinline__ N5Global(const void *T_, u32 y) : T(reinterpret_cast<const Tw *>(T_) + y) {}
    __device__ __forceinline__ Tw one() const { return __ldg(T + 16); }
    template <int M, int I0, int CNT> __device__ __forceinline__ void layer(Tw (&w)[CNT]) const {
#pragma unroll
        for (int i = 0; i < CNT; i++) w[i] = __ldg(T + (1 << M) + 16 * (I0 + i));
    }
};

// layers whose twiddle depends on the register index only (shared-memory table)
template <class AR>
__device__ __forceinline__ void n5_fwd_uniform(const AR &ar, typename AR::T (&v)[16], const N5Shared<typename AR::TwT> &tw) {
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int j = 3 - s;
#pragma unroll
        for (int a = 0; a < 8; a++) if (a < (1 << s)) {
            const typename AR::TwT w = tw.uni((1 << s) + a);
#pragma unroll
            for (int k = 0; k < 16; k++) if (!(k & (1 << j)) && (k >> (j + 1)) == a) ar.bf_fwd(v[k], v[k | (1 << j)], w);
        }
    }
}
// LB: index of the first of the four layers inside the whole 16-layer inverse transform (the lazy integer path needs it)
template <bool LAST, int LB, class AR>
__device__ __forceinline__ void n5_inv_uniform(const AR &ar, typename AR::T (&v)[16], const N5Shared<typename AR::TwT> &tw) {
#pragma unroll
    for (int s = 0; s < 4; s++) {         // pairs (k, k | 1 << s); 8 >> s twiddles at heap index (8 >> s) + (k >> (s + 1))
        if (LAST && s == 3) {
#pragma unroll
            for (int k = 0; k < 8; k++) ar.bf_inv_last(v[k], v[k | 8], LB + 3);
        } else {
#pragma unroll
            for (int a = 0; a < 8; a++) if (a < (8 >> s)) {
                const typename AR::TwT w = tw.uni((8 >> s) + a);
#pragma unroll
                for (int k = 0; k < 16; k++) if (!(k & (1 << s)) && (k >> (s + 1)) == a) ar.bf_inv(v[k], v[k | (1 << s)], w, LB + s);
            }
        }
    }
}
// layers with per-thread twiddles.  Loads are staged (each layer's twiddles are requested while the previous layer computes,
// the eight-twiddle layer in two halves) so that the kernel fits 72 registers on both arithmetic paths.
template <class AR, class TW>
__device__ __forceinline__ void n5_fwd_thread(const AR &ar, typename AR::T (&v)[16], const TW &tw) {
    typedef typename AR::TwT W;
    const W w1 = tw.one();
    W w2[2]; tw.template layer<5, 0, 2>(w2);
#pragma unroll
    for (int k = 0; k < 8; k++) ar.bf_fwd(v[k], v[k | 8], w1);
    W w4[4]; tw.template layer<6, 0, 4>(w4);
#pragma unroll
    for (int k = 0; k < 16; k++) if (!(k & 4)) ar.bf_fwd(v[k], v[k | 4], w2[k >> 3]);
    W wa[4]; tw.template layer<7, 0, 4>(wa);
#pragma unroll
    for (int k = 0; k < 16; k++) if (!(k & 2)) ar.bf_fwd(v[k], v[k | 2], w4[k >> 2]);
    W wb[4]; tw.template layer<7, 4, 4>(wb);
#pragma unroll
    for (int k = 0; k < 8; k++) if (!(k & 1)) ar.bf_fwd(v[k], v[k | 1], wa[k >> 1]);
#pragma unroll
    for (int k = 8; k < 16; k++) if (!(k & 1)) ar.bf_fwd(v[k], v[k | 1], wb[(k >> 1) - 4]);
}
// wa: entries 0..3 of the first layer's eight twiddles, requested by the caller before it reads the data
template <int LB, class AR, class TW>
__device__ __forceinline__ void n5_inv_thread(const AR &ar, typename AR::T (&v)[16], const TW &tw, const typename AR::TwT (&wa)[4]) {
    typedef typename AR::TwT W;
    W wb[4]; tw.template layer<7, 4, 4>(wb);
#pragma unroll
    for (int k = 0; k < 8; k++) if (!(k & 1)) ar.bf_inv(v[k], v[k | 1], wa[k >> 1], LB);
    W w4[4]; tw.template layer<6, 0, 4>(w4);
#pragma unroll
    for (int k = 8; k < 16; k++) if (!(k & 1)) ar.bf_inv(v[k], v[k | 1], wb[(k >> 1) - 4], LB);
    W w2[2]; tw.template layer<5, 0, 2>(w2);
    const W w1 = tw.one();
#pragma unroll
    for (int k = 0; k < 16; k++) if (!(k & 2)) ar.bf_inv(v[k], v[k | 2], w4[k >> 2], LB + 1);
#pragma unroll
    for (int k = 0; k < 16; k++) if (!(k & 4)) ar.bf_inv(v[k], v[k | 4], w2[k >> 3], LB + 2);
#pragma unroll
    for (int k = 0; k < 8; k++) ar.bf_inv(v[k], v[k | 8], w1, LB + 3);
}

// ---- the four pass bodies (byte offsets into the stage buffer; see the header comment and DESIGN.md §4.1) ------------------
// Lane -> (hb, x).  Shared memory serves 8-byte accesses half-warp by half-warp, so the sixteen lanes of a half-warp must
// fall on sixteen different 8-byte bank pairs: in a COLUMN tile the bank pair is (chunk = w ^ (x & 7), half = hb), so a
// half-warp holds x = 0..7 with both hb; in a ROW tile it is (chunk = f(x >> 1), half = x & 1), so a half-warp is one row.
struct N5Lane {
    u32 base0;     // column tile, TMA layout: x*128 + ((w ^ (x&7)) << 4) + hb*8
    u32 pbw, pbr;  // exchange layout of the column passes: (P | Bw), (P | Br)
    u32 q0, q1;    // row tile: q0 one-word pattern (thread = low column bits), q1 16-byte pattern (thread = high column bits)
    u32 xc, xr, row;   // xc: x of the column passes; xr: x of the row passes; row = 2 * warp + hb of the row passes
};
__device__ __forceinline__ N5Lane n5_lane() {
    const u32 t = threadIdx.x, w = t >> 5, lane = t & 31;
    N5Lane L;
    {   // column passes: hb = lane & 1, x = lane >> 1
        const u32 hb = lane & 1, x = lane >> 1;
        L.xc = x;
        L.base0 = x * 128 + ((w ^ (x & 7)) << 4) + hb * 8;
        const u32 P = ((x & 7) * 144) ^ (w << 4);
        L.pbw = P | ((x >> 3) * 1024 + hb * 8);
        L.pbr = P | (x * 2048 + hb * 8);
    }
    {   // row passes: hb = lane >> 4, x = lane & 15
        const u32 hb = lane >> 4, x = lane & 15;
        L.xr = x; L.row = 2 * w + hb;
        L.q0 = (2 * w + hb) * 2048 + (x & 1) * 8 + ((x >> 1) << 4);
        L.q1 = (2 * w + hb) * 2048 + x * 128 + ((x & 7) << 4);
    }
    return L;
}
#define N5_LD64(off) lds64(stage + (off))
#define N5_ST64(off, val) sts64(stage + (off), (val))

// forward, first pass: column tile (natural layout in, (col, row_hi, row_lo) layout out), external words in
template <class AR, class XF>
__device__ __forceinline__ void n5_fwd_col(u32 stage, u32 T, const AR &ar, const N5Lane &L, const XF &xf, u32 tok) {
    typename AR::T v[16];
    const N5Shared<typename AR::TwT> tw(T, L.xc, tok);
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = ar.from_ext(xf(N5_LD64(L.base0 + k * 2048)));
    n5_fwd_uniform(ar, v, tw);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; k++) N5_ST64((L.pbw ^ ((k & 7) * 144)) + k * 2048, ar.to_raw(v[k]));
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = ar.from_raw(N5_LD64((L.pbr ^ ((k & 7) * 144)) + (k >> 3) * 1024));
    n5_fwd_thread(ar, v, tw);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; k++) N5_ST64(L.base0 + k * 2048, ar.to_raw(v[k]));
}
// forward, second pass: row tile, canonical words out.  Tu: the rows' 16 register-indexed twiddles (shared memory, 16 entries
// per row); Tg: the rows' full 256-entry tables in global memory
template <class AR>
__device__ __forceinline__ void n5_fwd_row(u32 stage, u32 Tu, const void *Tg, const AR &ar, const N5Lane &L, u32 tok) {
    typedef typename AR::TwT W;
    const N5Shared<W> twu(Tu + L.row * 16 * (u32)sizeof(W), 0, tok);
    typename AR::T v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = ar.from_raw(N5_LD64((L.q0 ^ ((k & 7) << 4)) + k * 128));
    n5_fwd_uniform(ar, v, twu);
#pragma unroll
    for (int k = 0; k < 16; k++) N5_ST64((L.q0 ^ ((k & 7) << 4)) + k * 128, ar.to_raw(v[k]));
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const ulonglong2 d = lds128(stage + (L.q1 ^ (j << 4)));
        v[2 * j] = ar.from_raw(d.x); v[2 * j + 1] = ar.from_raw(d.y);
    }
    n5_fwd_thread(ar, v, N5Global<W>(reinterpret_cast<const W *>(Tg) + L.row * 256, L.xr));
#pragma unroll
    for (int j = 0; j < 8; j++) sts128(stage + (L.q1 ^ (j << 4)), ar.final_fwd(v[2 * j]), ar.final_fwd(v[2 * j + 1]));
}
// inverse, first pass: row tile, external words in (layers 0..7 of the inverse transform)
template <class AR, class XF>
__device__ __forceinline__ void n5_inv_row(u32 stage, u32 Tu, const void *Tg, const AR &ar, const N5Lane &L, const XF &xf, u32 tok) {
    typedef typename AR::TwT W;
    const N5Shared<W> twu(Tu + L.row * 16 * (u32)sizeof(W), 0, tok);
    typename AR::T v[16];
    {
        const N5Global<W> tw(reinterpret_cast<const W *>(Tg) + L.row * 256, L.xr);
        W wa[4]; tw.template layer<7, 0, 4>(wa);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const ulonglong2 d = lds128(stage + (L.q1 ^ (j << 4)));
            v[2 * j] = ar.from_ext(xf(d.x)); v[2 * j + 1] = ar.from_ext(xf(d.y));
        }
        n5_inv_thread<0>(ar, v, tw, wa);
    }
#pragma unroll
    for (int j = 0; j < 8; j++) sts128(stage + (L.q1 ^ (j << 4)), ar.to_raw(v[2 * j]), ar.to_raw(v[2 * j + 1]));
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = ar.from_raw(N5_LD64((L.q0 ^ ((k & 7) << 4)) + k * 128));
    n5_inv_uniform<false, 4>(ar, v, twu);
#pragma unroll
    for (int k = 0; k < 16; k++) { ar.pass_boundary_inv(v[k]); N5_ST64((L.q0 ^ ((k & 7) << 4)) + k * 128, ar.to_raw(v[k])); }
}
// inverse, second pass: column tile ((col, row_hi, row_lo) layout in, natural layout out), canonical words out (layers 8..15)
template <class AR>
__device__ __forceinline__ void n5_inv_col(u32 stage, u32 T, const AR &ar, const N5Lane &L, u32 tok) {
    typedef typename AR::TwT W;
    typename AR::T v[16];
    const N5Shared<W> tw(T, L.xc, tok);
    {
        W wa[4]; tw.template layer<7, 0, 4>(wa);
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = ar.from_raw(N5_LD64(L.base0 + k * 2048));
        n5_inv_thread<8>(ar, v, tw, wa);
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; k++) N5_ST64((L.pbr ^ ((k & 7) * 144)) + (k >> 3) * 1024, ar.to_raw(v[k]));
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = ar.from_raw(N5_LD64((L.pbw ^ ((k & 7) * 144)) + k * 2048));
    n5_inv_uniform<true, 12>(ar, v, tw);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; k++) N5_ST64(L.base0 + k * 2048, ar.final_inv(v[k]));
}

// element-wise transform applied to the operand's words as they enter the first pass (identity for plain transforms;
// the keyswitch mod-up re-reduction otherwise)
struct N5Identity {
    struct Prep {};
    __device__ __forceinline__ Prep prep(int, int) const { return Prep{}; }
    __device__ __forceinline__ u32 src_limb(int, int, u32 limb) const { return limb; }
    struct Fn { __device__ __forceinline__ u64 operator()(u64 x) const { return x; } };
    __device__ __forceinline__ Fn fn(const Prep &) const { return Fn{}; }
};

// ---- the persistent kernel ------------------------------------------------------------------------------------------------
// Work items, queue, lag and per-limb completion counters exactly as in ntt_persistent_kernel (b2_ntt.cuh); limb offsets
// (LimbGroup base / stride, in elements) must be multiples of N: they are turned into limb coordinates of the tensor maps.
//
// Warp-specialised: warps 0..7 transform tiles, warp 8 (one lane) is the PRODUCER — it claims items, waits for their
// dependency, starts the TMA loads, issues the TMA stores and publishes completion.  Two stage buffers per CTA:
//   full[s]   (1 arrival + transaction bytes)  producer -> consumers: the tile (and its shared-memory twiddles) landed in stage s
//   empty[s]  (8 arrivals, one per warp)       consumers -> producer: every warp has finished its part of stage s
// There is NO CTA-wide barrier in the steady state: a warp that finishes early goes straight on to the tile in the other
// stage (which was requested a whole item earlier), so skew between warps is absorbed instead of being waited out at a
// barrier (the ncu profiles of the barrier-per-item variants spent 36-39 % of all warp cycles there: profiles/r02f, r02h).
//
// Cross-CTA hand-over of the intermediate: the producer's TMA store is complete (cp.async.bulk.wait_group) before its
// red.release on the limb's counter; the consumer side observes the counter, then fence.proxy.async orders its TMA load
// behind that observation.  The producer never spins on a counter while it holds an unpublished first-pass tile.
constexpr int N5_THREADS = NTT_THREADS + 32;
__device__ __forceinline__ void mbar_arrive(u64 *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory"); }

template <bool FWD, class Pro>
__global__ void __launch_bounds__(N5_THREADS, 3) ntt5_kernel(const __grid_constant__ NttJob J, const __grid_constant__ Ntt5Maps MP, Pro pro) {
    extern __shared__ char smem_raw[];
    __shared__ __align__(8) u64 s_full[2], s_empty[2];
    __shared__ u32 s_id[2];
    const u32 stage0 = (smem_u32(smem_raw) + 1023u) & ~1023u;       // shared-window addresses; SWIZZLE_128B wants 1024-byte alignment
    const u32 tw0 = stage0 + 2 * N5_STAGE_BYTES;
    const int t = threadIdx.x;
    const u32 total = J.total_items, tiles = 1u << J.tile_shift;
    u32 *const queue = J.sync, *const finished = J.sync + 1, *const done = J.sync + 2;

    auto limb_of = [&](const NttItem &it, u32 &g, u32 &c) -> u32 {
        c = it.limb / (u32)J.n_groups; g = it.limb - c * (u32)J.n_groups;
        const LimbGroup &grp = J.groups[g];
        return (u32)((grp.base + (u64)c * grp.stride) >> N5_LOGN);
    };
    if (t == 0) {
        mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1); mbar_init(&s_empty[0], 8); mbar_init(&s_empty[1], 8);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        fence_proxy_async();
    }
    __syncthreads();

    if (t < NTT_THREADS) {
        // ================================ consumers ================================
        const N5Lane L = n5_lane();
        u32 parity = 0, s = 0;               // bit s of parity: phase of full[s]
        while (true) {
            mbar_wait(&s_full[s], (parity >> s) & 1u, 1);
            parity ^= 1u << s;
            const u32 cur_id = s_id[s];
            if (cur_id >= total) break;
            const NttItem it = ntt_decode(J, cur_id);
            u32 g, c;
            limb_of(it, g, c);
            const ModInfo &M = J.mods[group_table(J.groups[g].table)];
            const u32 path = group_path(J.groups[g].table, FWD);
            const u32 stage = stage0 + s * N5_STAGE_BYTES, tw = tw0 + s * N5_TW_BYTES;
            const void *rowtab = FWD ? M.n5_frow : M.n5_irow;
            const bool first = it.phase == 0;
            auto run = [&](const auto &ar) {
                typedef typename std::remove_reference<decltype(ar)>::type AR;
                const size_t rt = (size_t)it.tile * 4096 * sizeof(typename AR::TwT);
                if (FWD) {
                    if (first) { const auto pp = pro.prep(J.g_base + (int)g, (int)c); n5_fwd_col(stage, tw, ar, L, pro.fn(pp), cur_id); }
                    else n5_fwd_row(stage, tw, reinterpret_cast<const char *>(rowtab) + rt, ar, L, cur_id);
                } else {
                    if (first) { const auto pp = pro.prep(J.g_base + (int)g, (int)c); n5_inv_row(stage, tw, reinterpret_cast<const char *>(rowtab) + rt, ar, L, pro.fn(pp), cur_id); }
                    else n5_inv_col(stage, tw, ar, L, cur_id);
                }
            };
            if (J.dbg & 1) {}
            else if (path == PATH_F64) run(ArithF64(M));
            else if (path == PATH_INT_LAZY) run(ArithInt<true>(M));
            else run(ArithInt<false>(M));
            fence_proxy_async_smem();     // this thread's shared-memory writes -> visible to the copy engine
            __syncwarp();
            if ((t & 31) == 0) mbar_arrive(&s_empty[s]);
            s ^= 1;
        }
    } else {
      // ================================ producer warp: one lane works, the warp stays convergent for the final barrier ===========
      if (t == NTT_THREADS) {
        // start the copies of one item (tile + its shared-memory twiddles) into stage s
        auto issue = [&](const NttItem &it, u32 s) {
            u32 g, c;
            const u32 limb = limb_of(it, g, c);
            const ModInfo &M = J.mods[group_table(J.groups[g].table)];
            const u32 stage = stage0 + s * N5_STAGE_BYTES, tw = tw0 + s * N5_TW_BYTES;
            u64 *bar = &s_full[s];
            const u32 twb = group_path(J.groups[g].table, FWD) == PATH_F64 ? 2048u : 4096u;      // 256 staged twiddles of 8 or 16 bytes
            mbar_expect_tx(bar, (u32)N5_STAGE_BYTES + twb);
            const char *colt = reinterpret_cast<const char *>(FWD ? M.n5_fcol : M.n5_icol);
            const char *rowu = reinterpret_cast<const char *>(FWD ? M.n5_frow_u : M.n5_irow_u) + (size_t)it.tile * twb;
            if (FWD) {
                if (it.phase == 0) {
                    tma_load_3d(stage, &MP.first, bar, (int)(it.tile * 16), 0, (int)pro.src_limb(J.g_base + (int)g, (int)c, limb));
                    bulk_load(tw, colt, twb, bar);
                } else {
                    tma_load_2d(stage, &MP.mid_ld, bar, 0, (int)(limb * 4096 + it.tile * 256));
                    bulk_load(tw, rowu, twb, bar);
                }
            } else {
                if (it.phase == 0) {
                    tma_load_2d(stage, &MP.first, bar, 0, (int)(pro.src_limb(J.g_base + (int)g, (int)c, limb) * 4096 + it.tile * 256));
                    bulk_load(tw, rowu, twb, bar);
                } else {
                    tma_load_4d(stage, &MP.mid_ld, bar, (int)(it.tile * 16), 0, 0, (int)limb);
                    bulk_load(tw, colt, twb, bar);
                }
            }
        };
        auto store = [&](const NttItem &it, u32 s) {
            if (J.dbg & 2) return;
            u32 g, c;
            const u32 limb = limb_of(it, g, c);
            const u32 stage = stage0 + s * N5_STAGE_BYTES;
            if (FWD) {
                if (it.phase == 0) tma_store_4d(&MP.mid_st, stage, (int)(it.tile * 16), 0, 0, (int)limb);
                else tma_store_2d(&MP.last, stage, 0, (int)(limb * 4096 + it.tile * 256));
            } else {
                if (it.phase == 0) tma_store_2d(&MP.mid_st, stage, 0, (int)(limb * 4096 + it.tile * 256));
                else tma_store_3d(&MP.last, stage, (int)(it.tile * 16), 0, (int)limb);
            }
        };
        bool have[2] = {false, false};
        NttItem cur[2] = {{0, 0, 0}, {0, 0, 0}};
        u32 eparity = 0;                     // bit s: phase of empty[s]
        // first-pass tiles whose store is in flight and whose completion is not yet published: the newest and the one before
        bool pub_new = false, pub_old = false; u32 limb_new = 0, limb_old = 0;
        // Every global round trip of this loop is started one step before its result is needed: the queue claim for item k + 1
        // is issued at the top of iteration k and read at its end; the dependency counter of item k + 1 is requested at the end
        // of iteration k and read in iteration k + 1 (after the wait for the warps).  One iteration then costs the read-out of
        // the stored tile plus the publish fence, ~1.5 us, against ~5 us of arithmetic per tile and CTA.
        u32 id = claim(queue);
        NttItem it = {0, 0, 0};
        u32 dv = tiles;
        if (id < total) { it = ntt_decode(J, id); if (it.phase) dv = ld_relaxed(done + it.limb); }
        for (u32 k = 0;; k++) {
            const u32 s = k & 1;
            bool stored = false;
            if (have[s]) {                   // the tile that occupied this stage two items ago: wait for the warps, write it out
                mbar_wait(&s_empty[s], (eparity >> s) & 1u, 2);
                eparity ^= 1u << s;
                store(cur[s], s);
                bulk_commit();
                pub_old = pub_new; limb_old = limb_new;        // (the older one was published in the previous iteration)
                pub_new = cur[s].phase == 0; limb_new = cur[s].limb;
                have[s] = false; stored = true;
            }
            if (id >= total) {               // queue exhausted: wake the consumers on this stage so that they leave, drain the other stage
                s_id[s] = total;
                mbar_arrive(&s_full[s]);
                bulk_wait0();                // nothing left to overlap with: publish what has been stored before the flags shift again
                if (pub_old) red_release(done + limb_old);
                if (pub_new) red_release(done + limb_new);
                pub_old = pub_new = false;
                if (!have[s ^ 1]) break;
                continue;
            }
            const u32 next_id = claim(queue);                  // in flight until the end of this iteration
            if (stored) bulk_wait_read0();                     // the stage has been read out
            if (it.phase && dv < tiles && ld_relaxed(done + it.limb) < tiles) {
                // Dependency not met.  Nothing of OURS may stay unpublished while we wait (the missing first-pass tile could be
                // the one in our other stage): drain that stage too, publish every first-pass tile we have stored, then spin.
                bool pub_x = false; u32 limb_x = 0;
                const u32 o = s ^ 1;
                if (have[o]) {
                    mbar_wait(&s_empty[o], (eparity >> o) & 1u, 3);
                    eparity ^= 1u << o;
                    store(cur[o], o);
                    bulk_commit();
                    pub_x = cur[o].phase == 0; limb_x = cur[o].limb;
                    have[o] = false;
                }
                bulk_wait0();
                if (pub_old) red_release(done + limb_old);
                if (pub_new) red_release(done + limb_new);
                if (pub_x) red_release(done + limb_x);
                pub_old = pub_new = false;
                for (u32 spins = 0; ld_relaxed(done + it.limb) < tiles; spins++) {
                    __nanosleep(64);
                    if (spins == (1u << 24)) { printf("b2 ntt5: dependency of item %u (limb %u) stuck at %u (block %d)\n", id, it.limb, ld_relaxed(done + it.limb), (int)blockIdx.x); __trap(); }
                }
            }
            if (it.phase) fence_proxy_async_global();
            s_id[s] = id;
            issue(it, s);
            have[s] = true; cur[s] = it;
            // every store except the one issued in this iteration has landed by now (it was issued a whole item ago)
            asm volatile("cp.async.bulk.wait_group 1;\n" ::: "memory");
            if (pub_old) { red_release(done + limb_old); pub_old = false; }
            id = next_id;
            dv = tiles;
            if (id < total) { it = ntt_decode(J, id); if (it.phase) dv = ld_relaxed(done + it.limb); }
        }
        bulk_wait0();
        if (pub_old) red_release(done + limb_old);
        if (pub_new) red_release(done + limb_new);
      }
      __syncwarp();
    }
    // ---- teardown: the last CTA to finish resets the queue and the counters for the next launch
    __shared__ u32 s_last;
    __syncthreads();
    if (t == 0) {
        __threadfence();
        s_last = (atomicAdd(finished, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {
        for (u32 i = t; i < J.n_limbs; i += N5_THREADS) done[i] = 0;
        if (t == 0) { *queue = 0; *finished = 0; }
    }
}

}  // namespace b2
