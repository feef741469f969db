// b2_elementwise.cu — element-wise RNS kernels (M1, M3, E1, G1 in SURVEY.md §8a): ct x ct dyadic
// convolution / square (src/fgk/dyadic_convolute.cu:6-148), ct x pt broadcast product with optional
// accumulate (:151-292), add / sub / negate / negacyclic shift (src/utils/poly_small_mod.cu:263-949) and
// the Galois automorphisms (src/utils/galois.cu:141-300).  All HBM-bound: 16-byte vector accesses, one
// pass over the data, the modulus record read once per thread.  Layout [count][ncomp][N]; the modulus
// of limb j is table_base + j.
#include <algorithm>
#include <map>
#include <vector>
#include "b2_ctx.h"

using namespace b2;

namespace b2 {

struct EwShape { u32 ncomp, n2, table_base; u64 total2; };   // n2 = N/2 (vector units), total2 = count*ncomp*n2

__device__ __forceinline__ const ModInfo &ew_mod(const ModInfo *mods, const EwShape &S, u64 v) {
    return mods[S.table_base + (u32)((v / S.n2) % S.ncomp)];
}

// canonical a * b mod p on the FP64 pipe for canonical a, b < p < 2^41 (6 DP instructions + conversions instead of the
// ~40 integer-multiplier slots of a 128-bit Barrett reduction: the element-wise kernels stay HBM-bound).  Operands must be
// canonical, as every stored output of the hot path is.
__device__ __forceinline__ u64 mulmod_f64(u64 a, u64 b, double p, double pinv) {
    return dd_to_canonical(mulmod_dd(u64_to_dd(a), u64_to_dd(b), p, pinv), p);
}

template <int OP>   // 0 add, 1 sub, 2 negate
__global__ void __launch_bounds__(256) ew_addsub_kernel(const ulonglong2 *__restrict__ a, const ulonglong2 *__restrict__ b, ulonglong2 *__restrict__ o,
                                                          const ModInfo *__restrict__ mods, EwShape S) {
    for (u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < S.total2; v += (u64)gridDim.x * blockDim.x) {
        const u64 p = ew_mod(mods, S, v).p;
        ulonglong2 x = a[v], y, r;
        if (OP != 2) y = b[v];
        if (OP == 0) { r.x = addmod(x.x, y.x, p); r.y = addmod(x.y, y.y, p); }
        if (OP == 1) { r.x = submod(x.x, y.x, p); r.y = submod(x.y, y.y, p); }
        if (OP == 2) { r.x = negmod(x.x, p); r.y = negmod(x.y, p); }
        o[v] = r;
    }
}

// out[b][pidx] = sum_{x+y=pidx} a[b][x] o b[b][y]     (kernel_dyadic_convolute, dyadic_convolute.cu:6-42)
// NA, NB > 0: operand sizes known at compile time (operands stay in registers); 0: taken from na_, nb_
template <int NA, int NB>
__global__ void __launch_bounds__(256) ew_convolute_kernel(const ulonglong2 *__restrict__ a, u32 na_, const ulonglong2 *__restrict__ b, u32 nb_,
                                                             ulonglong2 *__restrict__ o, const ModInfo *__restrict__ mods, u32 ncomp, u32 n2, u32 batch) {
    const u32 na = NA ? (u32)NA : na_, nb = NB ? (u32)NB : nb_;
    const u64 per = (u64)ncomp * n2;                      // vectors per polynomial
    const u64 total = (u64)batch * per;
    for (u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (u64)gridDim.x * blockDim.x) {
        const u64 bi = v / per, r = v % per;
        const ModInfo &M = mods[(u32)(r / n2)];
        const u64 p = M.p, r0 = M.r0, r1 = M.r1;
        ulonglong2 av[3], bv[3];
#pragma unroll
        for (u32 x = 0; x < 3; x++) if (x < na) av[x] = a[(bi * na + x) * per + r];
#pragma unroll
        for (u32 y = 0; y < 3; y++) if (y < nb) bv[y] = b[(bi * nb + y) * per + r];
        const u32 no = na + nb - 1;
        if (M.use_f64) {        // p < 2^41: exact double-double products on the FP64 pipe, one reduction per output
            const double pd = M.pd, pinv = M.pinv;
            double ax[3], ay[3], bx[3], by[3];
#pragma unroll
            for (u32 x = 0; x < 3; x++) if (x < na) { ax[x] = u64_to_dd(av[x].x); ay[x] = u64_to_dd(av[x].y); }
#pragma unroll
            for (u32 y = 0; y < 3; y++) if (y < nb) { bx[y] = u64_to_dd(bv[y].x); by[y] = u64_to_dd(bv[y].y); }
#pragma unroll
            for (u32 pidx = 0; pidx < 5; pidx++) {
                if (pidx >= no) continue;
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (u32 x = 0; x < 3; x++) {
                    if (x >= na || pidx < x || pidx - x >= nb) continue;
                    s0 += mulmod_dd(ax[x], bx[pidx - x], pd, pinv);
                    s1 += mulmod_dd(ay[x], by[pidx - x], pd, pinv);
                }
                ulonglong2 res;
                res.x = dd_to_canonical(reduce_dd(s0, pd, pinv), pd); res.y = dd_to_canonical(reduce_dd(s1, pd, pinv), pd);
                o[(bi * no + pidx) * per + r] = res;
            }
            continue;
        }
#pragma unroll
        for (u32 pidx = 0; pidx < 5; pidx++) {
            if (pidx >= no) continue;
            u64 l0 = 0, h0 = 0, l1 = 0, h1 = 0;
#pragma unroll
            for (u32 x = 0; x < 3; x++) {
                if (x >= na || pidx < x || pidx - x >= nb) continue;
                mac128(l0, h0, av[x].x, bv[pidx - x].x);
                mac128(l1, h1, av[x].y, bv[pidx - x].y);
            }
            ulonglong2 res;
            res.x = barrett128(l0, h0, p, r0, r1); res.y = barrett128(l1, h1, p, r0, r1);
            o[(bi * no + pidx) * per + r] = res;
        }
    }
}

// (c0^2, 2 c0 c1, c1^2)     (kernel_dyadic_square, dyadic_convolute.cu:92-112)
__global__ void __launch_bounds__(256) ew_square_kernel(const ulonglong2 *__restrict__ a, ulonglong2 *__restrict__ o, const ModInfo *__restrict__ mods,
                                                          u32 ncomp, u32 n2, u32 batch) {
    const u64 per = (u64)ncomp * n2, total = (u64)batch * per;
    for (u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (u64)gridDim.x * blockDim.x) {
        const u64 bi = v / per, r = v % per;
        const ModInfo &M = mods[(u32)(r / n2)];
        const u64 p = M.p, r0 = M.r0, r1 = M.r1;
        const ulonglong2 c0 = a[(bi * 2) * per + r], c1 = a[(bi * 2 + 1) * per + r];
        ulonglong2 o0, o1, o2;
        if (M.use_f64) {
            const double pd = M.pd, pinv = M.pinv;
            const double a0 = u64_to_dd(c0.x), a1 = u64_to_dd(c0.y), b0 = u64_to_dd(c1.x), b1 = u64_to_dd(c1.y);
            o0.x = dd_to_canonical(mulmod_dd(a0, a0, pd, pinv), pd); o0.y = dd_to_canonical(mulmod_dd(a1, a1, pd, pinv), pd);
            const double t0 = mulmod_dd(a0, b0, pd, pinv), t1 = mulmod_dd(a1, b1, pd, pinv);
            o1.x = dd_to_canonical(reduce_dd(t0 + t0, pd, pinv), pd); o1.y = dd_to_canonical(reduce_dd(t1 + t1, pd, pinv), pd);
            o2.x = dd_to_canonical(mulmod_dd(b0, b0, pd, pinv), pd); o2.y = dd_to_canonical(mulmod_dd(b1, b1, pd, pinv), pd);
            o[(bi * 3) * per + r] = o0; o[(bi * 3 + 1) * per + r] = o1; o[(bi * 3 + 2) * per + r] = o2;
            continue;
        }
        o0.x = mulmod(c0.x, c0.x, p, r0, r1); o0.y = mulmod(c0.y, c0.y, p, r0, r1);
        u64 cx = mulmod(c0.x, c1.x, p, r0, r1), cy = mulmod(c0.y, c1.y, p, r0, r1);
        o1.x = addmod(cx, cx, p); o1.y = addmod(cy, cy, p);
        o2.x = mulmod(c1.x, c1.x, p, r0, r1); o2.y = mulmod(c1.y, c1.y, p, r0, r1);
        o[(bi * 3) * per + r] = o0; o[(bi * 3 + 1) * per + r] = o1; o[(bi * 3 + 2) * per + r] = o2;
    }
}

// out[k] (+)= a[k] o plain , k < pcount          (kernel_dyadic_broadcast_product_ps, dyadic_convolute.cu:151-170)
template <bool ACC>
__global__ void __launch_bounds__(256) ew_broadcast_kernel(const ulonglong2 *__restrict__ a, const ulonglong2 *__restrict__ pl, ulonglong2 *__restrict__ o,
                                                             const ModInfo *__restrict__ mods, EwShape S, u64 pstride) {
    // pstride = vector units of one ciphertext (pcount * ncomp * n2): ciphertext v / pstride of a batch uses plaintext v / pstride
    const u64 per = (u64)S.ncomp * S.n2;
    for (u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < S.total2; v += (u64)gridDim.x * blockDim.x) {
        const ModInfo &M = ew_mod(mods, S, v);
        const u64 p = M.p, r0 = M.r0, r1 = M.r1;
        const ulonglong2 x = a[v], y = pl[(v / pstride) * per + v % per];
        ulonglong2 r;
        if (M.use_f64) { r.x = mulmod_f64(x.x, y.x, M.pd, M.pinv); r.y = mulmod_f64(x.y, y.y, M.pd, M.pinv); }
        else { r.x = mulmod(x.x, y.x, p, r0, r1); r.y = mulmod(x.y, y.y, p, r0, r1); }
        if (ACC) { const ulonglong2 d = o[v]; r.x = addmod(d.x, r.x, p); r.y = addmod(d.y, r.y, p); }
        o[v] = r;
    }
}


// ---- multiply_plain over scattered pointers, ONE launch (M3 / B1) -----------------------------------------------------
// Reference: Evaluator::multiply_plain_ntt_batched / multiply_plain_ntt_accumulate (src/evaluator_multiply_plain.cu:216-307)
// -> kernel_dyadic_broadcast_product_bps (src/fgk/dyadic_convolute.cu:167-171), whose in-thread loop over the batch makes
// REPEATED destinations well defined: dest[i] (+)= ct[i] o pt[i] for i = 0, 1, ... in order.  Here the host groups the terms
// by destination (CSR: terms row_start[d] .. row_start[d+1] all write dest d) and a thread accumulates all terms of one
// destination word in registers: every ciphertext / plaintext word is read once, every destination word is read at most
// once and written once, however many terms it has.  The sums are the same canonical residues (modular addition is
// associative), so the result is bit-identical to the reference's term-by-term order.
//   ct[i] [pcount][ncomp][N], pt[i] [ncomp][N], dest[d] [pcount][ncomp][N]; limb j uses modulus table_base + j
struct TermTables { const u64 *const *ct; const u64 *const *pt; u64 *const *dest; const u32 *row_start; };
template <int PC>
__global__ void __launch_bounds__(256) ew_terms_kernel(TermTables T, const ModInfo *__restrict__ mods, u32 ncomp, u32 n2, u32 table_base, u32 pcount_,
                                                         int add_to_dest) {
    const u32 pcount = PC ? (u32)PC : pcount_;
    const u32 d = blockIdx.y;
    const u32 t0 = T.row_start[d], t1 = T.row_start[d + 1];
    const u64 per = (u64)ncomp * n2;
    ulonglong2 *__restrict__ dp = reinterpret_cast<ulonglong2 *>(T.dest[d]);
    for (u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < per; v += (u64)gridDim.x * blockDim.x) {
        const ModInfo &M = mods[table_base + (u32)(v / n2)];
        const u64 p = M.p;
        ulonglong2 res[3];
        if (M.use_f64) {          // p < 2^41: exact double-double products; partial sums stay below 2^10 p between reductions
            const double pd = M.pd, pinv = M.pinv;
            double s[3][2] = {{0, 0}, {0, 0}, {0, 0}};
            for (u32 t = t0; t < t1; t++) {
                const ulonglong2 y = __ldg(reinterpret_cast<const ulonglong2 *>(T.pt[t]) + v);
                const double y0 = u64_to_dd(y.x), y1 = u64_to_dd(y.y);
                const ulonglong2 *cp = reinterpret_cast<const ulonglong2 *>(T.ct[t]) + v;
#pragma unroll
                for (u32 k = 0; k < 3; k++) if (k < pcount) {
                    const ulonglong2 x = __ldcs(cp + k * per);
                    s[k][0] += mulmod_dd(u64_to_dd(x.x), y0, pd, pinv);
                    s[k][1] += mulmod_dd(u64_to_dd(x.y), y1, pd, pinv);
                }
                if (((t - t0) & 511) == 511) {
#pragma unroll
                    for (u32 k = 0; k < 3; k++) { s[k][0] = reduce_dd(s[k][0], pd, pinv); s[k][1] = reduce_dd(s[k][1], pd, pinv); }
                }
            }
#pragma unroll
            for (u32 k = 0; k < 3; k++) { res[k].x = dd_to_canonical(reduce_dd(s[k][0], pd, pinv), pd); res[k].y = dd_to_canonical(reduce_dd(s[k][1], pd, pinv), pd); }
        } else {                  // 128-bit lazy sums, one Barrett reduction per 64 terms (64 products of 61-bit residues < 2^128)
            const u64 r0 = M.r0, r1 = M.r1;
            u64 lo[3][2] = {{0, 0}, {0, 0}, {0, 0}}, hi[3][2] = {{0, 0}, {0, 0}, {0, 0}};
            for (u32 t = t0; t < t1; t++) {
                const ulonglong2 y = __ldg(reinterpret_cast<const ulonglong2 *>(T.pt[t]) + v);
                const ulonglong2 *cp = reinterpret_cast<const ulonglong2 *>(T.ct[t]) + v;
#pragma unroll
                for (u32 k = 0; k < 3; k++) if (k < pcount) {
                    const ulonglong2 x = __ldcs(cp + k * per);
                    mac128(lo[k][0], hi[k][0], x.x, y.x);
                    mac128(lo[k][1], hi[k][1], x.y, y.y);
                }
                if (((t - t0) & 63) == 63) {
#pragma unroll
                    for (u32 k = 0; k < 3; k++) {
                        lo[k][0] = barrett128(lo[k][0], hi[k][0], p, r0, r1); hi[k][0] = 0;
                        lo[k][1] = barrett128(lo[k][1], hi[k][1], p, r0, r1); hi[k][1] = 0;
                    }
                }
            }
#pragma unroll
            for (u32 k = 0; k < 3; k++) { res[k].x = barrett128(lo[k][0], hi[k][0], p, r0, r1); res[k].y = barrett128(lo[k][1], hi[k][1], p, r0, r1); }
        }
#pragma unroll
        for (u32 k = 0; k < 3; k++) if (k < pcount) {
            ulonglong2 r = res[k];
            if (add_to_dest) { const ulonglong2 o = dp[k * per + v]; r.x = addmod(o.x, r.x, p); r.y = addmod(o.y, r.y, p); }
            dp[k * per + v] = r;
        }
    }
}

// multiply by x^shift in Z[x]/(x^N+1)             (host_negacyclic_shift_ps, poly_small_mod.cu:902-925)
__global__ void __launch_bounds__(256) ew_negacyclic_shift_kernel(const u64 *__restrict__ a, u64 *__restrict__ o, const ModInfo *__restrict__ mods,
                                                                    u32 ncomp, u32 n, u32 table_base, u64 total, u32 shift) {
    for (u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (u64)gridDim.x * blockDim.x) {
        const u32 x = (u32)(v % n);                      // destination index
        const u64 limb = v / n;
        const u64 p = mods[table_base + (u32)(limb % ncomp)].p;
        const u32 src = (x - shift) & (n - 1);           // source index: src + shift = x (mod N)
        const bool neg = ((src + shift) & n) != 0;
        const u64 val = a[limb * n + src];
        o[v] = (neg && val) ? p - val : val;
    }
}

// coefficient form: coefficient i moves to i*g mod N with sign (-1)^floor(i*g/N)   (galois.cu:141-156)
// written as a GATHER (source index = dst * g^-1 mod 2N) so that stores are coalesced.
__global__ void __launch_bounds__(256) ew_galois_coeff_kernel(const u64 *__restrict__ a, u64 *__restrict__ o, const ModInfo *__restrict__ mods,
                                                                u32 ncomp, u32 logn, u32 table_base, u64 total, u32 ginv) {
    const u32 n = 1u << logn;
    for (u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (u64)gridDim.x * blockDim.x) {
        const u32 dst = (u32)(v & (n - 1));
        const u64 limb = v >> logn;
        const u64 p = mods[table_base + (u32)(limb % ncomp)].p;
        const u32 raw = (dst * ginv) & (2 * n - 1);      // i or i + N
        const u32 src = raw & (n - 1);
        // i*g mod 2N = dst (+N if sign): with i = src: sign bit = bit logN of src*g = bit logN of (raw*g - (raw>=n? n*g:0))...
        // equivalently x^(src*g) = +-x^dst ; raw >= n  <=>  src*g mod 2N = dst + N  <=> negative
        const u64 val = a[(limb << logn) + src];
        const bool neg = raw >= n;
        o[v] = (neg && val) ? p - val : val;
    }
}

// NTT form: out[i] = in[pi_g(i)], pi_g(i) = bitrev(((g * bitrev_{logN+1}(i + N)) >> 1) & (N-1))   (galois.cu:25-42)
// The permutation is computed with BREV instead of being read from an 8-byte-per-index table.
__global__ void __launch_bounds__(256) ew_galois_ntt_kernel(const u64 *__restrict__ a, u64 *__restrict__ o, u32 logn, u64 total, u32 g) {
    const u32 n = 1u << logn;
    for (u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (u64)gridDim.x * blockDim.x) {
        const u32 i = (u32)(v & (n - 1));
        const u32 rev = __brev(i + n) >> (31 - logn);
        const u32 raw = ((g * rev) >> 1) & (n - 1);
        const u32 src = __brev(raw) >> (32 - logn);
        o[v] = a[(v >> logn << logn) + src];
    }
}

// x <- x * (N^-1 * scalar) mod p      (kernel_ntt_multiply_inv_degree, src/utils/ntt.cu:109-117)
__global__ void __launch_bounds__(256) ew_scale_inv_degree_kernel(ulonglong2 *__restrict__ d, const ModInfo *__restrict__ mods, EwShape S, u64 scalar) {
    for (u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < S.total2; v += (u64)gridDim.x * blockDim.x) {
        const ModInfo &M = ew_mod(mods, S, v);
        const u64 p = M.p, f = mulmod(M.inv_n.x, barrett64(scalar, p, M.ratio), p, M.r0, M.r1);
        ulonglong2 x = d[v];
        x.x = mulmod(x.x, f, p, M.r0, M.r1); x.y = mulmod(x.y, f, p, M.r0, M.r1);
        d[v] = x;
    }
}

static inline unsigned ew_grid(u64 total) { return (unsigned)std::min<u64>((total + 255) / 256, 148u * 32u); }

}  // namespace b2

#define EW_PROLOG(cond, msg)                                  \
    B2_REQUIRE(ctx, "null ctx");                              \
    B2_REQUIRE(cond, msg);                                    \
    B2_CUDA(cudaSetDevice(ctx->device));                      \
    cudaStream_t stream = (cudaStream_t)stream_;              \
    const size_t n = ctx->n; (void)n;

extern "C" {

static b2_status ew_addsub(int op, b2_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t count, size_t ncomp, size_t table_base, void *stream_) {
    EW_PROLOG(a && out && (op == 2 || b) && table_base + ncomp <= ctx->K && ctx->n >= 2, "bad arguments");
    if (count * ncomp == 0) return B2_OK;
    EwShape S{(u32)ncomp, (u32)(n / 2), (u32)table_base, (u64)count * ncomp * (n / 2)};
    const unsigned grid = ew_grid(S.total2);
    if (op == 0) ew_addsub_kernel<0><<<grid, 256, 0, stream>>>((const ulonglong2 *)a, (const ulonglong2 *)b, (ulonglong2 *)out, ctx->d_mods, S);
    if (op == 1) ew_addsub_kernel<1><<<grid, 256, 0, stream>>>((const ulonglong2 *)a, (const ulonglong2 *)b, (ulonglong2 *)out, ctx->d_mods, S);
    if (op == 2) ew_addsub_kernel<2><<<grid, 256, 0, stream>>>((const ulonglong2 *)a, nullptr, (ulonglong2 *)out, ctx->d_mods, S);
    count_launches(1);
    B2_CUDA(cudaGetLastError());
    return B2_OK;
}
b2_status b2_add(b2_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t count, size_t ncomp, size_t table_base, void *stream) { return ew_addsub(0, ctx, a, b, out, count, ncomp, table_base, stream); }
b2_status b2_sub(b2_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t count, size_t ncomp, size_t table_base, void *stream) { return ew_addsub(1, ctx, a, b, out, count, ncomp, table_base, stream); }
b2_status b2_negate(b2_ctx *ctx, const uint64_t *a, uint64_t *out, size_t count, size_t ncomp, size_t table_base, void *stream) { return ew_addsub(2, ctx, a, nullptr, out, count, ncomp, table_base, stream); }

b2_status b2_dyadic_convolute(b2_ctx *ctx, const uint64_t *a, size_t na, const uint64_t *b, size_t nb, uint64_t *out, size_t batch, size_t ncomp, void *stream_) {
    EW_PROLOG(a && b && out && na >= 1 && nb >= 1 && na <= 3 && nb <= 3 && ncomp <= ctx->K && ctx->n >= 2, "bad arguments (polynomial counts must be 1..3)");
    if (batch * ncomp == 0) return B2_OK;
    const u64 total = (u64)batch * ncomp * (n / 2);
#define B2_CONV(NA, NB) ew_convolute_kernel<NA, NB><<<ew_grid(total), 256, 0, stream>>>((const ulonglong2 *)a, (u32)na, (const ulonglong2 *)b, (u32)nb, (ulonglong2 *)out, ctx->d_mods, (u32)ncomp, (u32)(n / 2), (u32)batch)
    if (na == 2 && nb == 2) B2_CONV(2, 2);
    else if (na == 3 && nb == 2) B2_CONV(3, 2);
    else if (na == 2 && nb == 3) B2_CONV(2, 3);
    else if (na == 3 && nb == 3) B2_CONV(3, 3);
    else B2_CONV(0, 0);
#undef B2_CONV
    count_launches(1);
    B2_CUDA(cudaGetLastError());
    return B2_OK;
}
b2_status b2_dyadic_square(b2_ctx *ctx, const uint64_t *a, uint64_t *out, size_t batch, size_t ncomp, void *stream_) {
    EW_PROLOG(a && out && ncomp <= ctx->K && ctx->n >= 2, "bad arguments");
    if (batch * ncomp == 0) return B2_OK;
    const u64 total = (u64)batch * ncomp * (n / 2);
    ew_square_kernel<<<ew_grid(total), 256, 0, stream>>>((const ulonglong2 *)a, (ulonglong2 *)out, ctx->d_mods, (u32)ncomp, (u32)(n / 2), (u32)batch);
    count_launches(1);
    B2_CUDA(cudaGetLastError());
    return B2_OK;
}
b2_status b2_dyadic_broadcast_batched(b2_ctx *ctx, const uint64_t *a, const uint64_t *plain, uint64_t *out, size_t batch, size_t pcount, size_t ncomp,
                                      int accumulate, void *stream_) {
    EW_PROLOG(a && plain && out && ncomp <= ctx->K && ctx->n >= 2, "bad arguments");
    if (batch * pcount * ncomp == 0) return B2_OK;
    EwShape S{(u32)ncomp, (u32)(n / 2), 0u, (u64)batch * pcount * ncomp * (n / 2)};
    const u64 pstride = (u64)pcount * ncomp * (n / 2);
    if (accumulate) ew_broadcast_kernel<true><<<ew_grid(S.total2), 256, 0, stream>>>((const ulonglong2 *)a, (const ulonglong2 *)plain, (ulonglong2 *)out, ctx->d_mods, S, pstride);
    else ew_broadcast_kernel<false><<<ew_grid(S.total2), 256, 0, stream>>>((const ulonglong2 *)a, (const ulonglong2 *)plain, (ulonglong2 *)out, ctx->d_mods, S, pstride);
    count_launches(1);
    B2_CUDA(cudaGetLastError());
    return B2_OK;
}
b2_status b2_dyadic_broadcast(b2_ctx *ctx, const uint64_t *a, const uint64_t *plain, uint64_t *out, size_t pcount, size_t ncomp, int accumulate, void *stream_) {
    return b2_dyadic_broadcast_batched(ctx, a, plain, out, 1, pcount, ncomp, accumulate, stream_);
}

// Evaluator::multiply_plain_ntt_batched / multiply_plain_ntt_accumulate over scattered operands, one launch.
// ct / pt / dest: HOST arrays of n_terms DEVICE pointers; dest[i] may repeat (every term with the same destination is
// accumulated into it, as the reference's in-thread loop does).  add_to_dest != 0: the destinations' current contents are
// the starting values (set_zero == false in the reference); 0: they are overwritten by the sums.
b2_status b2_multiply_plain_accumulate(b2_ctx *ctx, const uint64_t *const *ct, const uint64_t *const *pt, uint64_t *const *dest, size_t n_terms,
                                       size_t pcount, size_t ncomp, size_t table_base, int add_to_dest, void *stream_) {
    EW_PROLOG((n_terms == 0 || (ct && pt && dest)) && pcount >= 1 && pcount <= 3 && table_base + ncomp <= ctx->K && ctx->n >= 2, "bad arguments (polynomial count must be 1..3)");
    if (n_terms * ncomp == 0) return B2_OK;
    B2_REQUIRE(n_terms < ((size_t)1 << 31), "too many terms");
    // group the terms by destination, keeping first-appearance order of the destinations
    std::vector<uint32_t> order(n_terms);
    std::map<const uint64_t *, uint32_t> slot;
    std::vector<uint64_t *> dests;
    std::vector<uint32_t> cnt;
    for (size_t i = 0; i < n_terms; i++) {
        B2_REQUIRE(ct[i] && pt[i] && dest[i] && ((((uintptr_t)ct[i] | (uintptr_t)pt[i] | (uintptr_t)dest[i]) & 15) == 0), "null or misaligned operand in batch");
        auto it = slot.find(dest[i]);
        if (it == slot.end()) { it = slot.emplace(dest[i], (uint32_t)dests.size()).first; dests.push_back(dest[i]); cnt.push_back(0); }
        order[i] = it->second; cnt[it->second]++;
    }
    const size_t nd = dests.size();
    // one host blob -> one upload: ct[n_terms] | pt[n_terms] | dest[nd] | row_start[nd + 1]
    const size_t words = 2 * n_terms + nd, bytes = words * 8 + (nd + 1) * 4;
    std::vector<uint64_t> blob((bytes + 7) / 8);
    uint32_t *rs = reinterpret_cast<uint32_t *>(blob.data() + words);
    rs[0] = 0;
    for (size_t d = 0; d < nd; d++) rs[d + 1] = rs[d] + cnt[d];
    std::vector<uint32_t> fill(rs, rs + nd);
    for (size_t i = 0; i < n_terms; i++) {
        const uint32_t pos = fill[order[i]]++;
        blob[pos] = (uint64_t)(uintptr_t)ct[i]; blob[n_terms + pos] = (uint64_t)(uintptr_t)pt[i];
    }
    for (size_t d = 0; d < nd; d++) blob[2 * n_terms + d] = (uint64_t)(uintptr_t)dests[d];
    Scratch sc(ctx, stream);
    b2_status st = ensure_ptrtab(sc.s, stream, blob.size() * 8);
    if (st != B2_OK) return st;
    // pageable source: the runtime stages the bytes before returning, so `blob` may go out of scope; the copy is ordered on
    // `stream` after every earlier kernel that still reads the table
    B2_CUDA(cudaMemcpyAsync(sc->ptrtab, blob.data(), blob.size() * 8, cudaMemcpyHostToDevice, stream));
    const u64 *base = (const u64 *)sc->ptrtab;
    TermTables T{(const u64 *const *)base, (const u64 *const *)(base + n_terms), (u64 *const *)(base + 2 * n_terms), (const u32 *)(base + words)};
    const u64 per = (u64)ncomp * (n / 2);
    for (size_t d0 = 0; d0 < nd; d0 += 65535) {        // gridDim.y limit
        const unsigned ny = (unsigned)std::min<size_t>(65535, nd - d0);
        TermTables Td = T; Td.dest = T.dest + d0; Td.row_start = T.row_start + d0;
        // enough CTAs per destination to fill the machine when there are few destinations, grid-stride otherwise
        const unsigned gx = (unsigned)std::max<u64>(1, std::min<u64>((per + 255) / 256, (u64)(148 * 8 + ny - 1) / ny));
        dim3 grid(gx, ny);
        if (pcount == 2) ew_terms_kernel<2><<<grid, 256, 0, stream>>>(Td, ctx->d_mods, (u32)ncomp, (u32)(n / 2), (u32)table_base, 2u, add_to_dest);
        else ew_terms_kernel<0><<<grid, 256, 0, stream>>>(Td, ctx->d_mods, (u32)ncomp, (u32)(n / 2), (u32)table_base, (u32)pcount, add_to_dest);
        count_launches(1);
        B2_CUDA(cudaGetLastError());
    }
    return B2_OK;
}
b2_status b2_scale_inv_degree(b2_ctx *ctx, uint64_t *data, size_t count, size_t ncomp, size_t table_base, uint64_t scalar, void *stream_) {
    EW_PROLOG(data && table_base + ncomp <= ctx->K && ctx->n >= 2, "bad arguments");
    if (count * ncomp == 0) return B2_OK;
    EwShape S{(u32)ncomp, (u32)(n / 2), (u32)table_base, (u64)count * ncomp * (n / 2)};
    ew_scale_inv_degree_kernel<<<ew_grid(S.total2), 256, 0, stream>>>((ulonglong2 *)data, ctx->d_mods, S, scalar);
    count_launches(1);
    B2_CUDA(cudaGetLastError());
    return B2_OK;
}
b2_status b2_negacyclic_shift(b2_ctx *ctx, const uint64_t *a, size_t shift, uint64_t *out, size_t count, size_t ncomp, size_t table_base, void *stream_) {
    EW_PROLOG(a && out && a != out && table_base + ncomp <= ctx->K && shift < 2 * ctx->n, "bad arguments (shift must be below 2N, out of place)");
    if (count * ncomp == 0) return B2_OK;
    const u64 total = (u64)count * ncomp * n;
    ew_negacyclic_shift_kernel<<<ew_grid(total), 256, 0, stream>>>((const u64 *)a, (u64 *)out, ctx->d_mods, (u32)ncomp, (u32)n, (u32)table_base, total, (u32)shift);
    count_launches(1);
    B2_CUDA(cudaGetLastError());
    return B2_OK;
}
b2_status b2_galois(b2_ctx *ctx, const uint64_t *in, uint64_t galois_elt, uint64_t *out, size_t count, size_t ncomp, size_t table_base, int is_ntt, void *stream_) {
    EW_PROLOG(in && out && in != out && table_base + ncomp <= ctx->K, "bad arguments (out of place only)");
    B2_REQUIRE((galois_elt & 1) && galois_elt < 2 * ctx->n, "Galois element must be odd and below 2N");   // reference: src/utils/galois.cu:147-149
    if (count * ncomp == 0) return B2_OK;
    const u64 total = (u64)count * ncomp * n;
    if (is_ntt) ew_galois_ntt_kernel<<<ew_grid(total), 256, 0, stream>>>((const u64 *)in, (u64 *)out, ctx->logn, total, (u32)galois_elt);
    else {
        // g^-1 mod 2N (g odd): Newton iteration
        u32 g = (u32)galois_elt, inv = g;
        for (int it = 0; it < 5; it++) inv *= 2 - g * inv;
        inv &= (u32)(2 * n - 1);
        ew_galois_coeff_kernel<<<ew_grid(total), 256, 0, stream>>>((const u64 *)in, (u64 *)out, ctx->d_mods, (u32)ncomp, ctx->logn, (u32)table_base, total, inv);
    }
    count_launches(1);
    B2_CUDA(cudaGetLastError());
    return B2_OK;
}

}  // extern "C"
