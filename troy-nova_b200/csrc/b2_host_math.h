// b2_host_math.h — host-side number theory used once at context creation (prime selection,
// primitive roots, inverses).  Setup code, not the hot path.  Mirrors the *results* of the
// reference's utils::get_primes / try_minimal_primitive_root / try_invert_uint64_mod
// (src/utils/number_theory.cu:22-87, src/coeff_modulus.cu:64-107); written independently.
#pragma once
#include <cstdint>
#include <vector>

namespace b2 { namespace host {

typedef unsigned __int128 u128;
typedef uint64_t u64;

inline u64 mulmod(u64 a, u64 b, u64 p) { return (u64)(((u128)a * b) % p); }
inline u64 powmod(u64 a, u64 e, u64 p) {
    u64 r = 1 % p; a %= p;
    for (; e; e >>= 1) { if (e & 1) r = mulmod(r, a, p); a = mulmod(a, a, p); }
    return r;
}
inline u64 invmod(u64 a, u64 p) {   // 0 if not invertible
    __int128 t = 0, nt = 1, r = p, nr = a % p;
    while (nr) { __int128 q = r / nr, x = t - q * nt; t = nt; nt = x; x = r - q * nr; r = nr; nr = x; }
    if (r != 1) return 0;
    return (u64)(t < 0 ? t + p : t);
}
inline bool is_prime(u64 n) {       // deterministic Miller-Rabin for 64-bit n
    if (n < 2) return false;
    const u64 bases[12] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (u64 b : bases) { if (n == b) return true; if (n % b == 0) return false; }
    u64 d = n - 1; int s = 0;
    while (!(d & 1)) { d >>= 1; s++; }
    for (u64 b : bases) {
        u64 x = powmod(b, d, n);
        if (x == 1 || x == n - 1) continue;
        bool witness = true;
        for (int i = 1; i < s && witness; i++) { x = mulmod(x, x, n); if (x == n - 1) witness = false; }
        if (witness) return false;
    }
    return true;
}
// `count` largest primes below 2^bits that are 1 mod factor, descending
inline bool primes_below(u64 factor, unsigned bits, size_t count, std::vector<u64> &out) {
    u64 v = (((u64)1 << bits) - 1) / factor * factor + 1, lo = (u64)1 << (bits - 1);
    out.clear();
    for (; out.size() < count && v > lo; v -= factor) if (is_prime(v)) out.push_back(v);
    return out.size() == count;
}
// smallest primitive `degree`-th root of unity mod p (degree a power of two); 0 if none
inline u64 min_primitive_root(u64 degree, u64 p) {
    if ((p - 1) % degree) return 0;
    u64 g = 0;
    for (u64 x = 2; x < p && !g; x++) {
        u64 c = powmod(x, (p - 1) / degree, p);
        if (powmod(c, degree / 2, p) == p - 1) g = c;
    }
    if (!g) return 0;
    u64 best = g, cur = g, sq = mulmod(g, g, p);
    for (u64 i = 0; i < degree / 2; i++) { if (cur < best) best = cur; cur = mulmod(cur, sq, p); }
    return best;
}
inline u64 shoup_quotient(u64 w, u64 p) { return (u64)(((u128)w << 64) / p); }
inline unsigned bit_count(u64 v) { unsigned b = 0; while (v) { b++; v >>= 1; } return b; }
inline uint32_t bitrev(uint32_t x, unsigned bits) { uint32_t r = 0; for (unsigned i = 0; i < bits; i++) r = (r << 1) | ((x >> i) & 1); return r; }

}}  // namespace b2::host
