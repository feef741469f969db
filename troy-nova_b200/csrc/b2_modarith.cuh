// b2_modarith.cuh — 64-bit modular arithmetic for sm_100a, on the integer pipes AND on the FP64 pipe.
//
// All hot-path arithmetic is unsigned 64-bit modular integer work with moduli p < 2^61
// (reference: src/modulus.h:8-128, src/utils/uint_small_mod.h:54-148).  B200 has no 64-bit integer
// multiplier.  Measured on the box (tools/microbench*.cu, profiles/r01_microbench.txt):
//     IMAD 32-bit  63 /clk/SM     IMAD.WIDE / IMAD.HI  ~30 /clk/SM (half rate)
//     Shoup lazy modmul (64-bit)  3.9 /clk/SM  (16 IMAD-slots)      DFMA/DADD/DMUL  ~62 /clk/SM
//     integer lazy butterfly      3.6 /clk/SM                       FP64 butterfly   6.6 /clk/SM
// so for moduli that fit (p < 2^41, the 40-bit "scale" primes of every BASELINE config) the butterfly
// runs on the FP64 pipe with an exact double-double modular multiply, and the 60-bit primes use
// Harvey/Shoup lazy butterflies on the integer pipe.  Both produce the canonical residue, so the choice
// is invisible to the reference (bit-exact == equal canonical residues).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b2 {

typedef unsigned long long u64;
typedef unsigned int u32;

// Shoup pair {w, floor(w * 2^64 / p)} — same content as the reference's MultiplyUint64Operand
// (src/utils/uint_small_mod.h:92-125), stored as one 16-byte word so a twiddle is one LDG.128.
typedef ulonglong2 Tw;   // .x = operand w, .y = quotient

__device__ __forceinline__ u64 mulhi64(u64 a, u64 b) { return __umul64hi(a, b); }

// x * w mod p, lazily reduced to [0, 2p) for ANY 64-bit x (src/utils/uint_small_mod.h:141-148).
__device__ __forceinline__ u64 shoup_lazy(u64 x, Tw w, u64 p) {
    u64 q = mulhi64(x, w.y);
    return x * w.x - q * p;
}
__device__ __forceinline__ u64 csub(u64 x, u64 m) { return x >= m ? x - m : x; }   // [0,2m) -> [0,m)
// canonical version (src/utils/uint_small_mod.h:130-139)
__device__ __forceinline__ u64 shoup(u64 x, Tw w, u64 p) { return csub(shoup_lazy(x, w, p), p); }
__device__ __forceinline__ u64 addmod(u64 a, u64 b, u64 p) { return csub(a + b, p); }
__device__ __forceinline__ u64 submod(u64 a, u64 b, u64 p) { return a >= b ? a - b : a + p - b; }
__device__ __forceinline__ u64 negmod(u64 a, u64 p) { return a ? p - a : 0; }

// single-word Barrett: x (any 64-bit) mod p, with ratio = floor(2^64 / p)   (src/modulus.h:20-40)
__device__ __forceinline__ u64 barrett64(u64 x, u64 p, u64 ratio) {
    u64 q = mulhi64(x, ratio);
    return csub(x - q * p, p);
}

// 128-bit value (hi:lo) mod p with the two-word ratio floor(2^128 / p) = (r1:r0)  (src/modulus.h:42-79)
__device__ __forceinline__ u64 barrett128(u64 lo, u64 hi, u64 p, u64 r0, u64 r1) {
    u64 carry = mulhi64(lo, r0);
    u64 t_lo = lo * r1, t_hi = mulhi64(lo, r1);
    u64 tmp1 = t_lo + carry;
    u64 tmp3 = t_hi + (tmp1 < carry);
    u64 u_lo = hi * r0, u_hi = mulhi64(hi, r0);
    u64 s = tmp1 + u_lo;
    u64 c2 = u_hi + (s < tmp1);
    u64 q = hi * r1 + tmp3 + c2;
    return csub(lo - q * p, p);
}
__device__ __forceinline__ u64 mulmod(u64 a, u64 b, u64 p, u64 r0, u64 r1) {
    return barrett128(a * b, mulhi64(a, b), p, r0, r1);
}
// 128-bit accumulate: (hi:lo) += a * b
__device__ __forceinline__ void mac128(u64 &lo, u64 &hi, u64 a, u64 b) {
    u64 pl = a * b, ph = mulhi64(a, b);
    lo += pl;
    hi += ph + (lo < pl);
}

// ---- FP64-pipe arithmetic (p < 2^41) -----------------------------------------------------------
// Values are signed integers held exactly in doubles.  mulmod_dd returns x*w - c*p EXACTLY for an
// integer c within +-1 of x*w/p, so |result| <= (0.5 + 2^-8) p:
//   h + l == x*w exactly (error-free product);  c = rint(h/p) via one FMA against 1.5*2^52;
//   d = h - c*p is an integer below 2^53 -> exact;  d + l exact.
// Preconditions (guaranteed by the growth bounds in b2_ntt.cuh): |x| < 2^10 * p, 0 <= w < p < 2^41.
constexpr double F64_MAGIC = 6755399441055744.0;     // 1.5 * 2^52
constexpr double F64_TWO52 = 4503599627370496.0;     // 2^52
__device__ __forceinline__ double mulmod_dd(double x, double w, double p, double pinv) {
    double h = x * w;
    double l = fma(x, w, -h);
    double c = fma(h, pinv, F64_MAGIC) - F64_MAGIC;
    double d = fma(-c, p, h);
    return d + l;
}
// x -> x - rint(x/p) p, |result| <= (0.5 + eps) p ; |x| < 2^51
__device__ __forceinline__ double reduce_dd(double x, double p, double pinv) {
    double c = fma(x, pinv, F64_MAGIC) - F64_MAGIC;
    return fma(-c, p, x);
}
__device__ __forceinline__ double u64_to_dd(u64 x) {     // exact for x < 2^52
    return __longlong_as_double((long long)(x | 0x4330000000000000ULL)) - F64_TWO52;
}
// g in (-p, p) -> canonical residue as an integer
__device__ __forceinline__ u64 dd_to_canonical(double g, double p) {
    g = g < 0.0 ? g + p : g;
    return (u64)__double_as_longlong(g + F64_TWO52) & 0x000FFFFFFFFFFFFFULL;
}

// Per-modulus constants resident in HBM (one 256-byte record per modulus; read once per CTA).
struct ModInfo {
    u64 p;          // modulus
    u64 ratio;      // floor(2^64 / p)               single-word Barrett
    u64 r0, r1;     // floor(2^128 / p) low, high    two-word Barrett (reference const_ratio)
    Tw  inv_n;      // N^-1 mod p                    (reference NTTTables::inv_degree_modulo)
    Tw  inv_n_w;    // N^-1 * psi^-(N/2) : last inverse-layer twiddle with N^-1 folded in
    const Tw *tw;   // root_powers,     bit-reversed order, N entries (src/utils/ntt.cu:36-46)
    const Tw *itw;  // inv_root_powers, scrambled order,    N entries (src/utils/ntt.cu:48-58)
    u32 fwd_lazy;   // 1: (4 + 2 logN) p < 2^64  -> forward butterflies need no per-layer correction
    u32 inv_lazy;   // 1: p * 2^(logN+2) < 2^64  -> inverse butterflies need no per-layer correction
    // FP64 path (use_f64 != 0 iff p < 2^41): same tables as doubles
    const double *twd, *itwd;
    double pd, pinv, inv_n_d, inv_n_w_d;
    u32 use_f64, bits;
    // TMA kernel (b2_ntt5.cuh; N = 2^16, FP64 path only, else null): 256-entry twiddle tables of the 256-point sub-transforms,
    // one per limb for the column passes (fcol / icol) and one per row for the row passes (frow / irow: [256 rows][256])
    // (frow_u / irow_u: entries [0, 16) of every row table, packed [256 rows][16] — the part that is staged in shared memory)
    // entries are doubles on the FP64 path, Shoup pairs (Tw) on the integer path
    const void *n5_fcol, *n5_frow, *n5_icol, *n5_irow, *n5_frow_u, *n5_irow_u;
    u64 pad[8];
};
static_assert(sizeof(ModInfo) == 256, "ModInfo must be 256 bytes");

}  // namespace b2
