/*
 * rns_oracle.c — plain-C CPU restatement of the Troy-Nova RNS polynomial hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see rns_oracle.h).  Parity status: PINNED against the reference's
 * known-answer vectors and against the unmodified reference compiled into oracle/_ref.
 *
 * Every function cites the reference file:line it restates.  Where the reference uses a
 * specific reduction algorithm (Barrett with const_ratio, Shoup) whose *stored* result is the
 * canonical residue, the restatement keeps the same algorithm for the primitives that the
 * reference's KATs pin (ro_barrett_*, ro_shoup_*) and is free to use __int128 '%' elsewhere:
 * SURVEY.md fact 5 — all hot-path outputs are canonical residues, so equality of canonical
 * residues is bit-exactness.
 */
#include "rns_oracle.h"
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;

/* ------------------------------------------------------------------ arithmetic */

u64 ro_mulmod(u64 a, u64 b, u64 p) { return (u64)(((u128)a * b) % p); }

u64 ro_powmod(u64 a, u64 e, u64 p) {
    u64 r = 1 % p; a %= p;
    while (e) { if (e & 1) r = ro_mulmod(r, a, p); a = ro_mulmod(a, a, p); e >>= 1; }
    return r;
}

u64 ro_invmod(u64 a, u64 p) { /* p prime or a coprime to p: extended Euclid (uint_small_mod try_invert) */
    __int128 t = 0, nt = 1; __int128 r = p, nr = a % p;
    while (nr != 0) { __int128 q = r / nr, tmp = t - q * nt; t = nt; nt = tmp; tmp = r - q * nr; r = nr; nr = tmp; }
    if (r != 1) return 0;
    if (t < 0) t += p;
    return (u64)t;
}

/* Modulus::reduce, src/modulus.h:20-40: q = hi64(x * floor(2^128/p)[1]); r = x - q p; one correction */
u64 ro_barrett_reduce64(u64 x, u64 p) {
    u128 ratio = (~(u128)0) / p; /* floor((2^128-1)/p) == floor(2^128/p) since p is not a power of two > 1 */
    u64 r1 = (u64)(ratio >> 64);
    u64 q = (u64)(((u128)x * r1) >> 64);
    u64 r = x - q * p;
    return r >= p ? r - p : r;
}

/* Modulus::reduce_uint128_limbs, src/modulus.h:42-79 */
u64 ro_barrett_reduce128(u64 lo, u64 hi, u64 p) {
    u128 ratio = (~(u128)0) / p;
    u64 r0 = (u64)ratio, r1 = (u64)(ratio >> 64);
    u64 carry = (u64)(((u128)lo * r0) >> 64);
    u128 t2 = (u128)lo * r1;
    u64 tmp1 = (u64)t2 + carry;
    u64 tmp3 = (u64)(t2 >> 64) + (tmp1 < carry);
    t2 = (u128)hi * r0;
    u64 s = tmp1 + (u64)t2;
    carry = (u64)(t2 >> 64) + (s < tmp1);
    u64 q = hi * r1 + tmp3 + carry;
    u64 r = lo - q * p;
    return r >= p ? r - p : r;
}

/* MultiplyUint64Operand::set_quotient, src/utils/uint_small_mod.h:97-107: floor(w * 2^64 / p) */
u64 ro_shoup_quotient(u64 w, u64 p) { return (u64)((((u128)w) << 64) / p); }

/* multiply_uint64operand_mod_lazy, src/utils/uint_small_mod.h:141-148: result in [0, 2p) */
u64 ro_shoup_mul_lazy(u64 x, u64 w, u64 wq, u64 p) {
    u64 q = (u64)(((u128)x * wq) >> 64);
    return w * x - q * p;
}

static inline u64 shoup_mul(u64 x, u64 w, u64 wq, u64 p) { /* uint_small_mod.h:130-139, canonical */
    u64 r = ro_shoup_mul_lazy(x, w, wq, p);
    return r >= p ? r - p : r;
}
static inline u64 addmod(u64 a, u64 b, u64 p) { u64 s = a + b; return s >= p ? s - p : s; }   /* :54-61 */
static inline u64 submod(u64 a, u64 b, u64 p) { return a >= b ? a - b : a + p - b; }           /* :63-72 */
static inline u64 negmod(u64 a, u64 p) { return a ? p - a : 0; }                               /* :29-35 */

/* ------------------------------------------------------------------ number theory */

int ro_is_prime(u64 n) { /* accept set == primes (reference uses random-witness Miller-Rabin, number_theory.cu) */
    if (n < 2) return 0;
    static const u64 small[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (int i = 0; i < 12; i++) { if (n == small[i]) return 1; if (n % small[i] == 0) return 0; }
    u64 d = n - 1; int r = 0;
    while (!(d & 1)) { d >>= 1; r++; }
    for (int i = 0; i < 12; i++) { /* deterministic for all 64-bit n */
        u64 x = ro_powmod(small[i], d, n);
        if (x == 1 || x == n - 1) continue;
        int comp = 1;
        for (int k = 1; k < r; k++) { x = ro_mulmod(x, x, n); if (x == n - 1) { comp = 0; break; } }
        if (comp) return 0;
    }
    return 1;
}

/* get_primes, src/utils/number_theory.cu:22-39 */
int ro_get_primes(u64 factor, unsigned bit_size, size_t count, u64 *out) {
    u64 value = (((u64)1 << bit_size) - 1) / factor * factor + 1;
    u64 lower = (u64)1 << (bit_size - 1);
    size_t got = 0;
    while (got < count && value > lower) {
        if (ro_is_prime(value)) out[got++] = value;
        value -= factor;
    }
    return got == count ? 0 : -1;
}

/* CoeffModulus::create, src/coeff_modulus.cu:64-107: for each distinct bit size take `multiplicity`
 * primes descending; hand them out from the BACK of each list in order of appearance. */
int ro_coeff_modulus_create(size_t n, const unsigned *bits, size_t count, u64 *out) {
    if (count > RO_MAX_MODULI) return -1;
    u64 table[64][RO_MAX_MODULI]; size_t have[64]; size_t mult[64];
    memset(have, 0, sizeof have); memset(mult, 0, sizeof mult);
    for (size_t i = 0; i < count; i++) { if (bits[i] >= 64 || bits[i] < 2) return -1; mult[bits[i]]++; }
    for (unsigned b = 0; b < 64; b++) if (mult[b]) {
        if (ro_get_primes(2 * (u64)n, b, mult[b], table[b])) return -1;
        have[b] = mult[b];
    }
    for (size_t i = 0; i < count; i++) out[i] = table[bits[i]][--have[bits[i]]];
    return 0;
}

u64 ro_plain_modulus_batching(size_t n, unsigned bits) { u64 o = 0; ro_coeff_modulus_create(n, &bits, 1, &o); return o; }

/* try_minimal_primitive_root, src/utils/number_theory.cu:68-87. The reference starts from a random
 * primitive root and scans all its odd powers keeping the minimum, i.e. it returns the smallest
 * primitive degree-th root; we do the same scan from a deterministic starting root. */
u64 ro_min_primitive_root(u64 degree, u64 p) {
    if ((p - 1) % degree) return 0;
    u64 e = (p - 1) / degree, g = 0;
    for (u64 x = 2; x < p; x++) {
        g = ro_powmod(x, e, p);
        if (ro_powmod(g, degree >> 1, p) == p - 1) break; /* is_primitive_root for power-of-two degree */
        g = 0;
    }
    if (!g) return 0;
    u64 best = g, cur = g, g2 = ro_mulmod(g, g, p);
    for (u64 i = 0; i < (degree + 1) / 2; i++) { if (cur < best) best = cur; cur = ro_mulmod(cur, g2, p); }
    return best;
}

static inline uint32_t bitrev32(uint32_t x, unsigned bits) {
    uint32_t r = 0; for (unsigned i = 0; i < bits; i++) { r = (r << 1) | ((x >> i) & 1); } return r;
}

/* ------------------------------------------------------------------ context */

struct ro_ctx {
    unsigned log_n; size_t n; size_t nm; int scheme; u64 t;
    u64 q[RO_MAX_MODULI], psi[RO_MAX_MODULI];
    u64 *rp[RO_MAX_MODULI];   /* root_powers [N][2]     (src/utils/ntt.cu:36-46) */
    u64 *irp[RO_MAX_MODULI];  /* inv_root_powers [N][2] (src/utils/ntt.cu:48-58) */
    u64 inv_n[RO_MAX_MODULI], inv_n_q[RO_MAX_MODULI];
};

ro_ctx *ro_ctx_create(unsigned log_n, const u64 *moduli, size_t nm, u64 t, int scheme) {
    if (nm == 0 || nm > RO_MAX_MODULI) return NULL;
    ro_ctx *c = (ro_ctx *)calloc(1, sizeof *c);
    c->log_n = log_n; c->n = (size_t)1 << log_n; c->nm = nm; c->scheme = scheme; c->t = t;
    size_t n = c->n;
    for (size_t i = 0; i < nm; i++) {
        u64 p = moduli[i]; c->q[i] = p;
        u64 psi = ro_min_primitive_root(2 * n, p);
        if (!psi) { ro_ctx_destroy(c); return NULL; }
        c->psi[i] = psi;
        u64 ipsi = ro_invmod(psi, p);
        c->rp[i] = (u64 *)malloc(n * 16); c->irp[i] = (u64 *)malloc(n * 16);
        u64 pw = psi;
        for (size_t k = 1; k < n; k++) {
            size_t idx = bitrev32((uint32_t)k, log_n);
            c->rp[i][2 * idx] = pw; c->rp[i][2 * idx + 1] = ro_shoup_quotient(pw, p);
            pw = ro_mulmod(pw, psi, p);
        }
        c->rp[i][0] = 1; c->rp[i][1] = ro_shoup_quotient(1, p);
        pw = ipsi;
        for (size_t k = 1; k < n; k++) {
            size_t idx = bitrev32((uint32_t)(k - 1), log_n) + 1;
            c->irp[i][2 * idx] = pw; c->irp[i][2 * idx + 1] = ro_shoup_quotient(pw, p);
            pw = ro_mulmod(pw, ipsi, p);
        }
        c->irp[i][0] = 1; c->irp[i][1] = ro_shoup_quotient(1, p);
        c->inv_n[i] = ro_invmod(n % p, p); c->inv_n_q[i] = ro_shoup_quotient(c->inv_n[i], p);
    }
    return c;
}

void ro_ctx_destroy(ro_ctx *c) {
    if (!c) return;
    for (size_t i = 0; i < c->nm; i++) { free(c->rp[i]); free(c->irp[i]); }
    free(c);
}
size_t ro_ctx_n(const ro_ctx *c) { return c->n; }
size_t ro_ctx_n_moduli(const ro_ctx *c) { return c->nm; }
u64 ro_ctx_modulus(const ro_ctx *c, size_t i) { return c->q[i]; }
u64 ro_ctx_root(const ro_ctx *c, size_t i) { return c->psi[i]; }
const u64 *ro_ctx_root_powers(const ro_ctx *c, size_t i) { return c->rp[i]; }
const u64 *ro_ctx_inv_root_powers(const ro_ctx *c, size_t i) { return c->irp[i]; }


/* ------------------------------------------------------------------ tiny pthread parallel-for
 * (no OpenMP runtime in this image). RO_THREADS env var caps the thread count (default: all cores). */
#include <pthread.h>
#include <unistd.h>
typedef void (*par_fn)(long i, void *arg);
typedef struct { par_fn fn; void *arg; long n; long next; pthread_mutex_t mu; } par_job;
static void *par_worker(void *v) {
    par_job *j = (par_job *)v;
    for (;;) {
        pthread_mutex_lock(&j->mu); long i = j->next++; pthread_mutex_unlock(&j->mu);
        if (i >= j->n) return NULL;
        j->fn(i, j->arg);
    }
}
static int ro_threads(void) {
    const char *e = getenv("RO_THREADS"); long t = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    if (t < 1) t = 1; if (t > 256) t = 256; return (int)t;
}
static void par_for(long n, par_fn fn, void *arg) {
    int nt = ro_threads(); if (nt > n) nt = (int)n;
    if (nt <= 1) { for (long i = 0; i < n; i++) fn(i, arg); return; }
    par_job j; j.fn = fn; j.arg = arg; j.n = n; j.next = 0; pthread_mutex_init(&j.mu, NULL);
    pthread_t th[256];
    for (int t = 0; t < nt; t++) pthread_create(&th[t], NULL, par_worker, &j);
    for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
    pthread_mutex_destroy(&j.mu);
}

/* ------------------------------------------------------------------ NTT */

/* host_ntt_transfer_to_rev_layer, src/fgk/ntt_grouped.cu:11-56 (all layers, one limb) */
static void ntt_fwd_limb(const ro_ctx *c, u64 *x, size_t ti) {
    const u64 p = c->q[ti], two_p = p << 1; const u64 *rp = c->rp[ti];
    const unsigned logn = c->log_n; const size_t half = c->n >> 1;
    for (unsigned layer = 0; layer < logn; layer++) {
        size_t m = (size_t)1 << layer; unsigned gp = logn - layer - 1; size_t gap = (size_t)1 << gp;
        for (size_t i = 0; i < half; i++) {
            size_t rid = m + (i >> gp);
            size_t xi = ((i >> gp) << (gp + 1)) + (i & (gap - 1)), yi = xi + gap;
            u64 X = x[xi], Y = x[yi];
            u64 u = X >= two_p ? X - two_p : X;
            u64 v = ro_shoup_mul_lazy(Y, rp[2 * rid], rp[2 * rid + 1], p);
            x[xi] = u + v; x[yi] = u + two_p - v;
        }
    }
    for (size_t i = 0; i < c->n; i++) { u64 v = x[i]; if (v >= two_p) v -= two_p; if (v >= p) v -= p; x[i] = v; }
}

/* host_ntt_transfer_from_rev_layer, src/fgk/ntt_grouped.cu:346-391 */
static void ntt_inv_limb(const ro_ctx *c, u64 *x, size_t ti) {
    const u64 p = c->q[ti], two_p = p << 1; const u64 *irp = c->irp[ti];
    const unsigned logn = c->log_n; const size_t half = c->n >> 1;
    for (unsigned layer = 0; layer < logn; layer++) {
        size_t m = (size_t)1 << (logn - layer - 1); unsigned gp = layer; size_t gap = (size_t)1 << gp;
        for (size_t i = 0; i < half; i++) {
            size_t rid = c->n - (m << 1) + 1 + (i >> gp);
            size_t xi = ((i >> gp) << (gp + 1)) + (i & (gap - 1)), yi = xi + gap;
            u64 u = x[xi], v = x[yi];
            x[xi] = (u + v > two_p) ? (u + v - two_p) : (u + v);
            x[yi] = ro_shoup_mul_lazy(u + two_p - v, irp[2 * rid], irp[2 * rid + 1], p);
        }
    }
    for (size_t i = 0; i < c->n; i++) {
        u64 v = x[i]; if (v >= two_p) v -= two_p; if (v >= p) v -= p;
        x[i] = ro_shoup_mul_lazy(v, c->inv_n[ti], c->inv_n_q[ti], p); /* lazy but exact, SURVEY §8a N2 */
    }
}

/* NTTTableIndexer::get, src/utils/ntt.h:105-123 */
static size_t table_index(const ro_ctx *c, int mode, size_t base, size_t decomp, size_t k, size_t j) {
    switch (mode) {
    case RO_TABLES_KS_SET_PRODUCTS: return k == decomp ? c->nm - 1 : k;
    case RO_TABLES_KS_SKIP_FINALS:  return j == decomp ? c->nm - 1 : j;
    default:                        return base + j;
    }
}

typedef struct { const ro_ctx *c; u64 *d; size_t nc; int mode; size_t base, decomp; int inv; } ntt_job;
static void ntt_job_fn(long e, void *v) {
    ntt_job *j = (ntt_job *)v;
    size_t ti = table_index(j->c, j->mode, j->base, j->decomp, (size_t)e / j->nc, (size_t)e % j->nc);
    if (j->inv) ntt_inv_limb(j->c, j->d + (size_t)e * j->c->n, ti); else ntt_fwd_limb(j->c, j->d + (size_t)e * j->c->n, ti);
}
void ro_ntt_fwd(const ro_ctx *c, u64 *d, size_t pc, size_t nc, int mode, size_t base, size_t decomp) {
    ntt_job j = {c, d, nc, mode, base, decomp, 0}; par_for((long)(pc * nc), ntt_job_fn, &j);
}
void ro_ntt_inv(const ro_ctx *c, u64 *d, size_t pc, size_t nc, int mode, size_t base, size_t decomp) {
    ntt_job j = {c, d, nc, mode, base, decomp, 1}; par_for((long)(pc * nc), ntt_job_fn, &j);
}

/* host_ntt_multiply_inv_degree, src/utils/ntt.cu:78-91: x <- (x * N^-1 lazily) * scalar mod p */
void ro_ntt_multiply_inv_degree(const ro_ctx *c, u64 *d, size_t pc, size_t nc, size_t base, u64 scalar) {
    for (size_t k = 0; k < pc; k++) for (size_t j = 0; j < nc; j++) {
        size_t ti = base + j; u64 p = c->q[ti];
        for (size_t i = 0; i < c->n; i++) {
            u64 *x = d + (k * nc + j) * c->n + i;
            *x = ro_mulmod(ro_shoup_mul_lazy(*x, c->inv_n[ti], c->inv_n_q[ti], p), scalar, p);
        }
    }
}

/* ------------------------------------------------------------------ element-wise */

#define FOR_PLN(pc, nc) for (size_t k = 0; k < (pc); k++) for (size_t j = 0; j < (nc); j++) \
    for (size_t i = 0, o = (k * (nc) + j) * c->n; i < c->n; i++)

void ro_add(const ro_ctx *c, const u64 *a, const u64 *b, u64 *out, size_t pc, size_t nc) {
    FOR_PLN(pc, nc) out[o + i] = addmod(a[o + i], b[o + i], c->q[j]);
}
void ro_sub(const ro_ctx *c, const u64 *a, const u64 *b, u64 *out, size_t pc, size_t nc) {
    FOR_PLN(pc, nc) out[o + i] = submod(a[o + i], b[o + i], c->q[j]);
}
void ro_negate(const ro_ctx *c, const u64 *a, u64 *out, size_t pc, size_t nc) {
    FOR_PLN(pc, nc) out[o + i] = negmod(a[o + i], c->q[j]);
}

/* kernel_dyadic_convolute, src/fgk/dyadic_convolute.cu:6-42 */
void ro_dyadic_convolute(const ro_ctx *c, const u64 *a, size_t na, const u64 *b, size_t nb, u64 *out, size_t nc) {
    size_t n = c->n, no = na + nb - 1;
    for (size_t pidx = 0; pidx < no; pidx++) for (size_t j = 0; j < nc; j++) for (size_t i = 0; i < n; i++) {
        u64 acc = 0, p = c->q[j];
        for (size_t x = 0; x < na; x++) {
            if (pidx < x || pidx - x >= nb) continue;
            size_t y = pidx - x;
            acc = addmod(acc, ro_mulmod(a[(x * nc + j) * n + i], b[(y * nc + j) * n + i], p), p);
        }
        out[(pidx * nc + j) * n + i] = acc;
    }
}

/* kernel_dyadic_square, src/fgk/dyadic_convolute.cu:92-112 */
void ro_dyadic_square(const ro_ctx *c, const u64 *a, u64 *out, size_t nc) {
    size_t n = c->n;
    for (size_t j = 0; j < nc; j++) for (size_t i = 0; i < n; i++) {
        u64 p = c->q[j], c0 = a[j * n + i], c1 = a[(nc + j) * n + i];
        u64 cross = ro_mulmod(c0, c1, p);
        out[j * n + i] = ro_mulmod(c0, c0, p);
        out[(nc + j) * n + i] = addmod(cross, cross, p);
        out[(2 * nc + j) * n + i] = ro_mulmod(c1, c1, p);
    }
}

/* kernel_dyadic_broadcast_product_ps (+accumulate), src/fgk/dyadic_convolute.cu:151-170 */
void ro_dyadic_broadcast(const ro_ctx *c, const u64 *a, const u64 *pl, u64 *out, size_t pc, size_t nc, int acc) {
    FOR_PLN(pc, nc) {
        u64 v = ro_mulmod(a[o + i], pl[j * c->n + i], c->q[j]);
        out[o + i] = acc ? addmod(out[o + i], v, c->q[j]) : v;
    }
}

/* host_negacyclic_shift_ps, src/utils/poly_small_mod.cu:902-925 */
void ro_negacyclic_shift(const ro_ctx *c, const u64 *a, size_t shift, u64 *out, size_t pc, size_t nc) {
    size_t n = c->n;
    FOR_PLN(pc, nc) {
        size_t raw = shift + i, idx = raw & (n - 1);
        u64 v = a[o + i];
        out[o + idx] = (v == 0 || (raw & n) == 0) ? v : c->q[j] - v;
    }
}

/* ------------------------------------------------------------------ Galois */

/* GaloisTool::get_element_from_step, src/utils/galois.cu:44-63 (generator 3) */
u64 ro_galois_element_from_step(size_t n, int step) {
    size_t m = 2 * n;
    if (step == 0) return m - 1;
    size_t pos = step < 0 ? (size_t)(-step) : (size_t)step;
    if (pos >= (n >> 1)) return 0;
    size_t ts = step < 0 ? (n >> 1) - pos : pos;
    u64 g = 1; for (size_t i = 0; i < ts; i++) g = (g * 3) & (m - 1);
    return g;
}

/* GaloisTool::generate_table_ntt, src/utils/galois.cu:25-42 */
void ro_galois_ntt_table(unsigned log_n, u64 g, uint32_t *table) {
    size_t n = (size_t)1 << log_n;
    for (size_t i = 0; i < n; i++) {
        uint32_t rev = bitrev32((uint32_t)(i + n), log_n + 1);
        u64 raw = ((g * (u64)rev) >> 1) & (n - 1);
        table[i] = bitrev32((uint32_t)raw, log_n);
    }
}

/* GaloisTool::apply_ps host, src/utils/galois.cu:141-156 */
void ro_galois_coeff(const ro_ctx *c, const u64 *in, u64 g, u64 *out, size_t pc, size_t nc) {
    size_t n = c->n;
    FOR_PLN(pc, nc) {
        u64 raw = (u64)i * g; size_t idx = raw & (n - 1);
        out[o + idx] = ((raw >> c->log_n) & 1) ? negmod(in[o + i], c->q[j]) : in[o + i];
    }
}

/* GaloisTool::apply_ntt_ps host, src/utils/galois.cu:234-255: out[i] = in[table[i]] */
void ro_galois_ntt(const ro_ctx *c, const u64 *in, u64 g, u64 *out, size_t pc, size_t nc) {
    uint32_t *tab = (uint32_t *)malloc(c->n * sizeof(uint32_t));
    ro_galois_ntt_table(c->log_n, g, tab);
    FOR_PLN(pc, nc) out[o + i] = in[o + tab[i]];
    free(tab);
}

/* ------------------------------------------------------------------ keyswitch */

typedef struct { const ro_ctx *c; const u64 *target; const u64 *tc; const u64 *const *keys; u64 *prod; size_t decomp; int is_ntt; } ks_job;
static void ks_job_fn(long ii, void *v) {
    ks_job *J = (ks_job *)v; const ro_ctx *c = J->c; const u64 *target = J->target, *tc = J->tc; const u64 *const *keys = J->keys;
    u64 *prod = J->prod; size_t decomp = J->decomp; int is_ntt = J->is_ntt;
    const size_t n = c->n, K = c->nm, R = decomp + 1, sp = K - 1;
    size_t i = (size_t)ii, ki = (i == decomp) ? sp : i; u64 m = c->q[ki];
    u64 *tmp = (u64 *)malloc(n * 8);
    u128 *lazy = (u128 *)calloc(2 * n, sizeof(u128));
    for (size_t j = 0; j < decomp; j++) {
        const u64 *op;
        if (is_ntt && i == j) op = target + j * n;                     /* :851-852 */
        else {
            for (size_t x = 0; x < n; x++) tmp[x] = (c->q[j] <= m) ? tc[j * n + x] : tc[j * n + x] % m; /* :854-858 */
            ntt_fwd_limb(c, tmp, ki);
            op = tmp;
        }
        for (size_t k = 0; k < 2; k++) {
            const u64 *kp = keys[j] + (k * K + ki) * n;
            for (size_t x = 0; x < n; x++) lazy[k * n + x] += (u128)op[x] * kp[x];
        }
        if (((j + 1) & 15) == 0) for (size_t x = 0; x < 2 * n; x++) lazy[x] %= m; /* keep the 128-bit sum bounded */
    }
    for (size_t k = 0; k < 2; k++) for (size_t x = 0; x < n; x++) prod[(k * R + i) * n + x] = (u64)(lazy[k * n + x] % m);
    free(tmp); free(lazy);
}

/* Evaluator::switch_key_internal host branch, src/evaluator_keyswitching_core.cu:757-985.
 * Steps follow the reference: (1) INTT target if NTT-form (:815-824); (2) per output modulus i,
 * per digit j: re-reduce c_j mod m_i, NTT, multiply with key_j[k][i] and accumulate (:833-902;
 * the reference accumulates lazily in 128 bits and reduces once — canonical either way);
 * (3) mod-down: BGV ski_util5 (:436-476), BFV/CKKS ski_util6 + ski_util7 (:570-756). */
void ro_keyswitch(const ro_ctx *c, const u64 *target, size_t decomp, int is_ntt,
                  const u64 *const *keys, u64 *dest, int assign) {
    const size_t n = c->n, K = c->nm, R = decomp + 1, sp = K - 1;
    u64 *tc = (u64 *)malloc(decomp * n * 8);
    memcpy(tc, target, decomp * n * 8);
    if (is_ntt) ro_ntt_inv(c, tc, 1, decomp, RO_TABLES_COMPONENTWISE, 0, 0);
    u64 *prod = (u64 *)calloc(2 * R * n, 8);           /* poly_prod [2][R][N] */
    { ks_job kj = {c, target, tc, keys, prod, decomp, is_ntt}; par_for((long)R, ks_job_fn, &kj); }
    const u64 P = c->q[sp], half = P >> 1;
    for (size_t k = 0; k < 2; k++) {
        int add_inplace = (assign == RO_ASSIGN_ADD_INPLACE) || (k == 0 && assign == RO_ASSIGN_OVERWRITE_EXCEPT_FIRST);
        u64 *last = prod + (k * R + decomp) * n;
        ntt_inv_limb(c, last, sp);
        u64 *tn = (u64 *)malloc(decomp * n * 8);
        if (c->scheme == RO_SCHEME_BGV) {
            /* ski_util5, :436-476: kt = -c * P^{-1} mod t; delta_j = (kt mod q_j)*P + c mod q_j */
            u64 t = c->t, pinv_t = ro_invmod(P % t, t);
            for (size_t x = 0; x < n; x++) {
                u64 kt = negmod(last[x] % t, t);
                if (pinv_t != 1) kt = ro_mulmod(kt, pinv_t, t);
                for (size_t j = 0; j < decomp; j++) {
                    u64 q = c->q[j];
                    tn[j * n + x] = addmod(ro_mulmod(kt % q, P % q, q), last[x] % q, q);
                }
            }
            ro_ntt_fwd(c, tn, 1, decomp, RO_TABLES_COMPONENTWISE, 0, 0);
            for (size_t j = 0; j < decomp; j++) {
                u64 q = c->q[j], pinv = ro_invmod(P % q, q);
                for (size_t x = 0; x < n; x++) {
                    u64 v = ro_mulmod(submod(prod[(k * R + j) * n + x], tn[j * n + x], q), pinv, q);
                    u64 *d = dest + (k * decomp + j) * n + x;
                    *d = add_inplace ? addmod(v, *d, q) : v;
                }
            }
        } else {
            /* ski_util6, :570-623 */
            for (size_t x = 0; x < n; x++) {
                u64 r = (last[x] + half) % P;
                for (size_t j = 0; j < decomp; j++) {
                    u64 q = c->q[j];
                    tn[j * n + x] = ((P > q) ? r % q : r) + (q - half % q);
                }
            }
            if (is_ntt) ro_ntt_fwd(c, tn, 1, decomp, RO_TABLES_COMPONENTWISE, 0, 0);                         /* :959-960 */
            else ro_ntt_inv(c, prod + k * R * n, 1, decomp, RO_TABLES_COMPONENTWISE, 0, 0);                  /* :961-969 */
            /* ski_util7, :690-756 */
            for (size_t j = 0; j < decomp; j++) {
                u64 q = c->q[j], pinv = ro_invmod(P % q, q);
                for (size_t x = 0; x < n; x++) {
                    u64 v = prod[(k * R + j) * n + x] + ((c->scheme == RO_SCHEME_CKKS) ? (q << 2) : (q << 1)) - tn[j * n + x];
                    v = ro_mulmod(v % q, pinv, q);
                    u64 *d = dest + (k * decomp + j) * n + x;
                    *d = add_inplace ? addmod(*d, v, q) : v;
                }
            }
        }
        free(tn);
    }
    free(prod); free(tc);
}

/* ------------------------------------------------------------------ rescale / modswitch */

/* RNSTool::divide_and_round_q_last, src/utils/rns_tool.cu:374-466 */
void ro_divide_round_q_last(const ro_ctx *c, const u64 *in, u64 *out, size_t pc, size_t nl) {
    size_t n = c->n; u64 ql = c->q[nl - 1], half = ql >> 1;
    for (size_t p = 0; p < pc; p++) for (size_t x = 0; x < n; x++) {
        u64 v = addmod(in[(p * nl + nl - 1) * n + x], half, ql);
        for (size_t i = 0; i + 1 < nl; i++) {
            u64 q = c->q[i];
            u64 t = submod(v % q, half % q, q);
            u64 d = submod(in[(p * nl + i) * n + x], t, q);
            out[(p * (nl - 1) + i) * n + x] = ro_mulmod(d, ro_invmod(ql % q, q), q);
        }
    }
}

/* RNSTool::divide_and_round_q_last_ntt, src/utils/rns_tool.cu:499-694 (CKKS rescale) */
void ro_divide_round_q_last_ntt(const ro_ctx *c, const u64 *in, u64 *out, size_t pc, size_t nl) {
    size_t n = c->n; u64 ql = c->q[nl - 1], half = ql >> 1;
    u64 *last = (u64 *)malloc(n * 8), *tmp = (u64 *)malloc(n * 8);
    for (size_t p = 0; p < pc; p++) {
        memcpy(last, in + (p * nl + nl - 1) * n, n * 8);
        ntt_inv_limb(c, last, nl - 1);
        for (size_t x = 0; x < n; x++) last[x] = addmod(last[x], half, ql);
        for (size_t i = 0; i + 1 < nl; i++) {
            u64 q = c->q[i], qinv = ro_invmod(ql % q, q);
            for (size_t x = 0; x < n; x++) tmp[x] = submod(last[x] % q, half % q, q);
            ntt_fwd_limb(c, tmp, i);
            for (size_t x = 0; x < n; x++)
                out[(p * (nl - 1) + i) * n + x] = ro_mulmod(submod(in[(p * nl + i) * n + x], tmp[x], q), qinv, q);
        }
    }
    free(last); free(tmp);
}

/* RNSTool::mod_t_and_divide_q_last_ntt, src/utils/rns_tool.cu:1540-1592 (BGV) */
void ro_mod_t_divide_q_last_ntt(const ro_ctx *c, const u64 *in, u64 *out, size_t pc, size_t nl) {
    size_t n = c->n; u64 ql = c->q[nl - 1], t = c->t, qinv_t = ro_invmod(ql % t, t);
    u64 *last = (u64 *)malloc(n * 8), *tmp = (u64 *)malloc(n * 8);
    for (size_t p = 0; p < pc; p++) {
        memcpy(last, in + (p * nl + nl - 1) * n, n * 8);
        ntt_inv_limb(c, last, nl - 1);
        for (size_t i = 0; i + 1 < nl; i++) {
            u64 q = c->q[i], qinv = ro_invmod(ql % q, q);
            for (size_t x = 0; x < n; x++) {
                u64 k = negmod(last[x] % t, t);
                if (qinv_t != 1) k = ro_mulmod(k, qinv_t, t);
                tmp[x] = addmod(ro_mulmod(k % q, ql % q, q), last[x] % q, q);
            }
            ntt_fwd_limb(c, tmp, i);
            for (size_t x = 0; x < n; x++)
                out[(p * (nl - 1) + i) * n + x] = ro_mulmod(submod(in[(p * nl + i) * n + x], tmp[x], q), qinv, q);
        }
    }
    free(last); free(tmp);
}

/* kernel_mod_switch_drop_to, src/evaluator_modswitch.cu:152-171 */
void ro_drop_last_limb(const ro_ctx *c, const u64 *in, u64 *out, size_t pc, size_t nl) {
    for (size_t p = 0; p < pc; p++) memcpy(out + p * (nl - 1) * c->n, in + p * nl * c->n, (nl - 1) * c->n * 8);
}

/* ------------------------------------------------------------------ toy RLWE (semantic tests only) */

static u64 splitmix(u64 *s) { u64 z = (*s += 0x9e3779b97f4a7c15ULL); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31); }

static void sample_small(const ro_ctx *c, u64 *s, size_t nl, int ternary, u64 *out /*[nl][N] coeff form*/) {
    for (size_t x = 0; x < c->n; x++) {
        int v;
        if (ternary) v = (int)(splitmix(s) % 3) - 1;
        else { u64 r = splitmix(s); v = __builtin_popcountll(r & 0x1fffff) - __builtin_popcountll((r >> 21) & 0x1fffff); } /* CBD(21) */
        for (size_t j = 0; j < nl; j++) out[j * c->n + x] = v >= 0 ? (u64)v : c->q[j] - (u64)(-v);
    }
}

void ro_keygen_secret(const ro_ctx *c, u64 seed, u64 *sk) {
    u64 s = seed; sample_small(c, &s, c->nm, 1, sk);
    ro_ntt_fwd(c, sk, 1, c->nm, RO_TABLES_COMPONENTWISE, 0, 0);
}

/* rlwe::symmetric at the key level (src/utils/rlwe.cu:218-322): (c0,c1) = (-(a s + e), a), NTT form */
static void enc_zero_limbs(const ro_ctx *c, const u64 *sk, const size_t *limb_idx, size_t nl, u64 *s, u64 *ct, int is_ntt) {
    size_t n = c->n;
    u64 *e = (u64 *)malloc(c->nm * n * 8), *efull = (u64 *)malloc(c->nm * n * 8);
    sample_small(c, s, c->nm, 0, efull);
    for (size_t j = 0; j < nl; j++) memcpy(e + j * n, efull + limb_idx[j] * n, n * 8);
    for (size_t j = 0; j < nl; j++) { size_t ti = limb_idx[j]; ntt_fwd_limb(c, e + j * n, ti); }
    for (size_t j = 0; j < nl; j++) {
        size_t ti = limb_idx[j]; u64 q = c->q[ti];
        for (size_t x = 0; x < n; x++) {
            u64 a = splitmix(s) % q;
            ct[(nl + j) * n + x] = a;
            ct[j * n + x] = negmod(addmod(ro_mulmod(a, sk[ti * n + x], q), e[j * n + x], q), q);
        }
        if (!is_ntt) { ntt_inv_limb(c, ct + j * n, ti); ntt_inv_limb(c, ct + (nl + j) * n, ti); }
    }
    free(e); free(efull);
}

/* KeyGenerator::generate_one_kswitch_key, src/key_generator.cu:134-156 */
void ro_keygen_kswitch(const ro_ctx *c, const u64 *sk, const u64 *newkey, u64 seed, u64 *out) {
    size_t n = c->n, K = c->nm, D = K - 1; u64 s = seed; size_t idx[RO_MAX_MODULI];
    for (size_t j = 0; j < K; j++) idx[j] = j;
    for (size_t i = 0; i < D; i++) {
        u64 *key = out + i * 2 * K * n;
        enc_zero_limbs(c, sk, idx, K, &s, key, 1);
        u64 q = c->q[i], factor = c->q[K - 1] % q;
        for (size_t x = 0; x < n; x++) key[i * n + x] = addmod(key[i * n + x], ro_mulmod(newkey[i * n + x], factor, q), q);
    }
}

void ro_encrypt_zero_sym(const ro_ctx *c, const u64 *sk, size_t nl, int is_ntt, u64 seed, u64 *ct) {
    u64 s = seed; size_t idx[RO_MAX_MODULI];
    for (size_t j = 0; j < nl; j++) idx[j] = j;
    enc_zero_limbs(c, sk, idx, nl, &s, ct, is_ntt);
}

void ro_phase(const ro_ctx *c, const u64 *sk, const u64 *ct, size_t pc, size_t nl, int is_ntt, u64 *out) {
    size_t n = c->n;
    u64 *tmp = (u64 *)malloc(pc * nl * n * 8);
    memcpy(tmp, ct, pc * nl * n * 8);
    if (!is_ntt) ro_ntt_fwd(c, tmp, pc, nl, RO_TABLES_COMPONENTWISE, 0, 0);
    for (size_t j = 0; j < nl; j++) for (size_t x = 0; x < n; x++) {
        u64 q = c->q[j], acc = 0, spow = 1;
        for (size_t k = 0; k < pc; k++) {
            acc = addmod(acc, ro_mulmod(tmp[(k * nl + j) * n + x], spow, q), q);
            spow = ro_mulmod(spow, sk[j * n + x], q);
        }
        out[j * n + x] = acc;
    }
    ro_ntt_inv(c, out, 1, nl, RO_TABLES_COMPONENTWISE, 0, 0);
    free(tmp);
}

/* NOTE: SURVEY.md §8c quotes the offset basis as 0xcbf29ce484222325, but the probe that produced its
 * fingerprints used the decimal literal 1469598103934665603 (one digit short). The fingerprints only
 * reproduce with the latter, so that is what we use. */
u64 ro_fnv1a64(const u64 *w, size_t n) {
    u64 h = 1469598103934665603ULL;
    for (size_t i = 0; i < n; i++) { h ^= w[i]; h *= 0x100000001b3ULL; }
    return h;
}

#include "rns_oracle_behz.inc"
#include "rns_oracle_decrypt.inc"
#include "rns_oracle_encrypt.inc"
