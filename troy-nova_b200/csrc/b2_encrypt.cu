// b2_encrypt.cu — Encryptor / RLWE sampling and the BFV / BGV plaintext lifts (rank 3).  Replaces the device branch of
//   RandomGenerator (AES-128-CTR)  fill_bytes, sample_poly_ternary / _centered_binomial / _uniform   src/utils/random_generator.cu:12-141,282-505
//   rlwe::asymmetric_with_u_prng, rlwe::symmetric_with_c1_prng                                       src/utils/rlwe.cu:11-92,218-322
//   Encryptor::encrypt_zero_internal / encrypt_internal                                              src/encryptor.cu:12-118,247-338
//   scaling_variant::scale_up (multiply_translate_plain)    src/fgk/translate_plain.cu:6-70, constants src/context_data.cu:223-265
//   scaling_variant::centralize                             src/utils/scaling_variant.cu:326-357
//
// Same generator, same counter arithmetic, same sampling rules: with the same (seed, counter) the stored ciphertext words equal the
// reference's.  AES-128 is implemented from FIPS-197 (S-box computed from the field inverse + affine map at start-up); the key schedule
// is expanded ONCE on the host per call (the reference expands it in every thread) and the S-box lives in shared memory.
// Launches: asymmetric zero encryption 4 (+ the mod switch below the key level), symmetric 4; the reference issues 13-17.
#include "b2_ntt_launch.cuh"
#include "b2_host_math.h"

using namespace b2;

namespace b2 {

// ---- AES-128, host side: S-box, key schedule, single block ----
struct AesTables { uint8_t sbox[256]; };
static uint8_t gmul(uint8_t a, uint8_t b) { uint8_t r = 0; while (b) { if (b & 1) r ^= a; a = (uint8_t)((a << 1) ^ ((a & 0x80) ? 0x1b : 0)); b >>= 1; } return r; }
static const AesTables &aes_tables() {
    static const AesTables T = [] {
        AesTables t;
        for (int x = 0; x < 256; x++) {
            uint8_t inv = 0;
            if (x) for (int y = 1; y < 256; y++) if (gmul((uint8_t)x, (uint8_t)y) == 1) { inv = (uint8_t)y; break; }
            uint8_t s = inv, r = inv;
            for (int k = 0; k < 4; k++) { r = (uint8_t)((r << 1) | (r >> 7)); s ^= r; }
            t.sbox[x] = s ^ 0x63;
        }
        return t;
    }();
    return T;
}
struct AesKey { u32 rk[44]; uint8_t sbox[256]; };      // kernel parameter: expanded key + S-box
static AesKey aes_expand(uint64_t seed_lo, uint64_t seed_hi) {
    const AesTables &T = aes_tables();
    AesKey K;
    memcpy(K.sbox, T.sbox, 256);
    uint8_t rk[176];
    memcpy(rk, &seed_lo, 8); memcpy(rk + 8, &seed_hi, 8);
    uint8_t rcon = 1;
    for (int i = 4; i < 44; i++) {
        uint8_t t[4]; memcpy(t, rk + 4 * (i - 1), 4);
        if (i % 4 == 0) {
            const uint8_t a = t[0];
            t[0] = T.sbox[t[1]] ^ rcon; t[1] = T.sbox[t[2]]; t[2] = T.sbox[t[3]]; t[3] = T.sbox[a];
            rcon = gmul(rcon, 2);
        }
        for (int k = 0; k < 4; k++) rk[4 * i + k] = rk[4 * (i - 4) + k] ^ t[k];
    }
    memcpy(K.rk, rk, 176);
    return K;
}
// columns as little-endian words: byte r of column c is state byte 4c + r
__host__ __device__ __forceinline__ u32 aes_xt(u32 w) { return ((w & 0x7f7f7f7fu) << 1) ^ (((w >> 7) & 0x01010101u) * 0x1bu); }
__host__ __device__ __forceinline__ u32 aes_ror(u32 w, int k) { return (w >> k) | (w << (32 - k)); }
template <class S>
__host__ __device__ __forceinline__ void aes_encrypt(const u32 *rk, const S &sb, u32 s[4]) {
    s[0] ^= rk[0]; s[1] ^= rk[1]; s[2] ^= rk[2]; s[3] ^= rk[3];
#pragma unroll 1
    for (int round = 1; round <= 10; round++) {
        u32 t[4];
#pragma unroll
        for (int c = 0; c < 4; c++)
            t[c] = (u32)sb[s[c] & 0xff] | ((u32)sb[(s[(c + 1) & 3] >> 8) & 0xff] << 8) | ((u32)sb[(s[(c + 2) & 3] >> 16) & 0xff] << 16) |
                   ((u32)sb[(s[(c + 3) & 3] >> 24) & 0xff] << 24);
        if (round < 10) {
#pragma unroll
            for (int c = 0; c < 4; c++) { const u32 a = t[c], x = aes_xt(a); t[c] = x ^ aes_ror(x, 8) ^ aes_ror(a, 8) ^ aes_ror(a, 16) ^ aes_ror(a, 24); }
        }
#pragma unroll
        for (int c = 0; c < 4; c++) s[c] = t[c] ^ rk[4 * round + c];
    }
}
static void host_block(const AesKey &K, uint64_t ctr_lo, uint64_t ctr_hi, uint64_t out[2]) {
    u32 s[4] = {(u32)ctr_lo, (u32)(ctr_lo >> 32), (u32)ctr_hi, (u32)(ctr_hi >> 32)};
    aes_encrypt(K.rk, K.sbox, s);
    out[0] = (uint64_t)s[0] | ((uint64_t)s[1] << 32); out[1] = (uint64_t)s[2] | ((uint64_t)s[3] << 32);
}
static void advance(b2_prng *g, uint64_t blocks) { const uint64_t lo = g->counter_lo + blocks; g->counter_hi += lo < g->counter_lo; g->counter_lo = lo; }

// ---- device side ----
struct Ctr { u64 lo, hi; };
__device__ __forceinline__ void dev_block(const AesKey &K, const uint8_t *sb, Ctr c, u64 idx, u64 &o0, u64 &o1) {
    const u64 lo = c.lo + idx, hi = c.hi + (lo < c.lo);
    u32 s[4] = {(u32)lo, (u32)(lo >> 32), (u32)hi, (u32)(hi >> 32)};
    aes_encrypt(K.rk, sb, s);
    o0 = (u64)s[0] | ((u64)s[1] << 32); o1 = (u64)s[2] | ((u64)s[3] << 32);
}
#define B2_AES_SMEM(K)                                             \
    __shared__ uint8_t sb[256];                                    \
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sb[i] = (K).sbox[i]; \
    __syncthreads();

// out [count][nl][N]: polynomial c uses counter blocks [c * per, (c + 1) * per).  One thread per coefficient (the 16 / 2 threads that share
// a block each encrypt it: ~400 integer instructions, far cheaper than serialising the nl stores of 16 coefficients in one thread), so the
// stores of a warp are 32 consecutive words per limb.
// ternary: one block per 16 coefficients; byte % 3 -> {0, 1, q - 1}
__global__ void __launch_bounds__(256) sample_ternary_kernel(const __grid_constant__ AesKey K, Ctr ctr, u64 *__restrict__ out, const ModInfo *__restrict__ mods,
                                                             u32 nl, u32 n, u32 count) {
    B2_AES_SMEM(K)
    const u32 per = (n + 15) / 16;
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (u64)n * count) return;
    const u32 c = (u32)(idx / n), x = (u32)(idx % n);
    u64 w[2];
    dev_block(K, sb, ctr, (u64)c * per + x / 16, w[0], w[1]);
    const u32 k = x & 15;
    const u32 r = (u32)((w[k >> 3] >> (8 * (k & 7))) & 0xff) % 3;
    for (u32 j = 0; j < nl; j++) out[((u64)c * nl + j) * n + x] = r == 2 ? mods[j].p - 1 : r;
}
__device__ __forceinline__ int cbd_of(u64 w) {      // random_generator.cu:369-380: bytes 0-2 (21 bits) minus bytes 3-5 (21 bits)
    return __popcll(w & 0x1fffffull) - __popcll((w >> 24) & 0x1fffffull);
}
// centred binomial: one block per 2 coefficients
__global__ void __launch_bounds__(256) sample_cbd_kernel(const __grid_constant__ AesKey K, Ctr ctr, u64 *__restrict__ out, const ModInfo *__restrict__ mods,
                                                         u32 nl, u32 n, u32 count) {
    B2_AES_SMEM(K)
    const u32 per = (n + 1) / 2;
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (u64)n * count) return;
    const u32 c = (u32)(idx / n), x = (u32)(idx % n);
    u64 w[2];
    dev_block(K, sb, ctr, (u64)c * per + x / 2, w[0], w[1]);
    const int r = cbd_of(w[x & 1]);
    for (u32 j = 0; j < nl; j++) out[((u64)c * nl + j) * n + x] = r >= 0 ? (u64)r : mods[j].p - (u64)(-r);
}
// uniform: [nl][N] words = nl * N / 2 blocks in memory order, each word reduced modulo its limb's modulus (fill_uint64s + modulo_inplace_p)
__global__ void __launch_bounds__(256) sample_uniform_kernel(const __grid_constant__ AesKey K, Ctr ctr, u64 *__restrict__ out, const ModInfo *__restrict__ mods,
                                                             u32 n, u64 blocks, int reduce) {
    B2_AES_SMEM(K)
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= blocks) return;
    u64 w0, w1;
    dev_block(K, sb, ctr, idx, w0, w1);
    if (reduce) {
        const ModInfo &Q = mods[(2 * idx / n)];       // N is even: both words are in the same limb (one polynomial per launch)
        w0 = barrett64(w0, Q.p, Q.ratio); w1 = barrett64(w1, Q.p, Q.ratio);
    }
    reinterpret_cast<ulonglong2 *>(out)[idx] = make_ulonglong2(w0, w1);
}

// ---- fused epilogues / prologues of the transforms ----
// asymmetric, NTT(u): ct[k][j][x] = NTT(u)[j][x] * pk[k][j][x], k = 0, 1
struct PkMulStore {
    u64 *ct; const u64 *pk; const ModInfo *mods; u32 nl, K, n;
    struct Prep { u64 *o0, *o1; const u64 *p0, *p1; u64 p, r0, r1; };
    __device__ __forceinline__ Prep prep(int g, int, u64) const {
        const ModInfo &Q = mods[g];
        return Prep{ct + (u64)g * n, ct + ((u64)nl + g) * n, pk + (u64)g * n, pk + ((u64)K + g) * n, Q.p, Q.r0, Q.r1};
    }
    __device__ __forceinline__ void operator()(const Prep &q, u32 x, u64 v) const {
        q.o0[x] = mulmod(v, q.p0[x], q.p, q.r0, q.r1); q.o1[x] = mulmod(v, q.p1[x], q.p, q.r0, q.r1);
    }
    __device__ __forceinline__ void st2(const Prep &q, u32 x, u64 v0, u64 v1) const { (*this)(q, x, v0); (*this)(q, x + 1, v1); }
};
// out[c][g][x] = (sign) (base[c][g][x] + s * v): the noise term added to (NTT form: transformed noise; coefficient form: transformed product)
template <bool NEG>
struct AddScaledStore {
    u64 *out; const u64 *base; const ModInfo *mods; u32 n; u64 scale;       // scale: 1, or t (BGV) applied to v when SCALE_V, see below
    bool scale_v;                                                            // true: out = base + scale * v ; false: out = v + scale * base
    struct Prep { u64 *o; const u64 *b; u64 p, r0, r1, s; };
    __device__ __forceinline__ Prep prep(int g, int, u64 off) const {
        const ModInfo &Q = mods[g];
        return Prep{out + off, base + off, Q.p, Q.r0, Q.r1, barrett64(scale, Q.p, Q.ratio)};
    }
    __device__ __forceinline__ void operator()(const Prep &q, u32 x, u64 v) const {
        u64 b = q.b[x];
        if (scale != 1) { if (scale_v) v = mulmod(v, q.s, q.p, q.r0, q.r1); else b = mulmod(b, q.s, q.p, q.r0, q.r1); }
        const u64 r = addmod(v, b, q.p);
        q.o[x] = NEG ? negmod(r, q.p) : r;
    }
    __device__ __forceinline__ void st2(const Prep &q, u32 x, u64 v0, u64 v1) const { (*this)(q, x, v0); (*this)(q, x + 1, v1); }
};
// symmetric, coefficient form: limb group 0 (c == 0) transforms sk o c1 into c0 = -(. + e), group 1 transforms c1 itself
struct SkMulLoad {
    const u64 *c1; const u64 *sk; const ModInfo *mods; u32 nl, n;
    struct Prep { const u64 *s; u64 p, r0, r1; bool mul; };
    __device__ __forceinline__ const u64 *src(int g, int, u64) const { return c1 + (u64)((u32)g % nl) * n; }
    __device__ __forceinline__ Prep prep(int g, int) const {
        const u32 j = (u32)g % nl; const ModInfo &Q = mods[j];
        return Prep{sk + (u64)j * n, Q.p, Q.r0, Q.r1, (u32)g < nl};
    }
    __device__ __forceinline__ u64 xf(const Prep &q, u32 x, u64 raw) const { return q.mul ? mulmod(raw, q.s[x], q.p, q.r0, q.r1) : raw; }
};
struct SymCoeffStore {      // groups [0, nl): c0 = -(v + scale * e) ; groups [nl, 2nl): c1 = v
    u64 *ct; const u64 *e; const ModInfo *mods; u32 nl, n; u64 scale;
    struct Prep { u64 *o; const u64 *e; u64 p, r0, r1, s; bool first; };
    __device__ __forceinline__ Prep prep(int g, int, u64) const {
        const u32 j = (u32)g % nl; const ModInfo &Q = mods[j];
        return Prep{ct + (u64)(u32)g * n, e + (u64)j * n, Q.p, Q.r0, Q.r1, barrett64(scale, Q.p, Q.ratio), (u32)g < nl};
    }
    __device__ __forceinline__ void operator()(const Prep &q, u32 x, u64 v) const {
        if (q.first) {
            u64 ev = q.e[x];
            if (scale != 1) ev = mulmod(ev, q.s, q.p, q.r0, q.r1);
            v = negmod(addmod(v, ev, q.p), q.p);
        }
        q.o[x] = v;
    }
    __device__ __forceinline__ void st2(const Prep &q, u32 x, u64 v0, u64 v1) const { (*this)(q, x, v0); (*this)(q, x + 1, v1); }
};
// symmetric, NTT form: NTT(e) -> c0 = -(sk o c1 + scale * NTT(e))
struct SymNttStore {
    u64 *c0; const u64 *c1; const u64 *sk; const ModInfo *mods; u32 n; u64 scale;
    struct Prep { u64 *o; const u64 *a, *s; u64 p, r0, r1, sc; };
    __device__ __forceinline__ Prep prep(int g, int, u64) const {
        const ModInfo &Q = mods[g];
        return Prep{c0 + (u64)(u32)g * n, c1 + (u64)(u32)g * n, sk + (u64)(u32)g * n, Q.p, Q.r0, Q.r1, barrett64(scale, Q.p, Q.ratio)};
    }
    __device__ __forceinline__ void operator()(const Prep &q, u32 x, u64 v) const {
        if (scale != 1) v = mulmod(v, q.sc, q.p, q.r0, q.r1);
        q.o[x] = negmod(addmod(mulmod(q.a[x], q.s[x], q.p, q.r0, q.r1), v, q.p), q.p);
    }
    __device__ __forceinline__ void st2(const Prep &q, u32 x, u64 v0, u64 v1) const { (*this)(q, x, v0); (*this)(q, x + 1, v1); }
};

// BFV scale-up: dest[j][x] = from[j][x] +/- (Delta_j m_x + floor((m_x (q mod t) + (t+1)/2) / t)) mod q_j     (translate_plain.cu:6-70)
__global__ void __launch_bounds__(256) scale_up_kernel(const u64 *__restrict__ plain, u32 plain_count, const u64 *__restrict__ from, u64 *__restrict__ dest,
                                                       const ModInfo *__restrict__ mods, const u64 *__restrict__ delta, u64 q_mod_t, u64 t, u32 nl, u32 n, int subtract) {
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (u64)nl * n) return;
    const u32 j = (u32)(idx / n), x = (u32)(idx % n);
    const u64 base = from ? from[idx] : 0;
    if (x >= plain_count) { dest[idx] = base; return; }
    const ModInfo &Q = mods[j];
    const u64 m = plain[x];
    // fix = floor((m * q_mod_t + (t + 1) / 2) / t): 128-bit numerator, 64-bit divisor; both factors are below t < 2^61, so the quotient fits a word
    u64 lo = m * q_mod_t, hi = __umul64hi(m, q_mod_t);
    const u64 half = (t + 1) >> 1;
    lo += half; hi += lo < half;
    const unsigned __int128 num = ((unsigned __int128)hi << 64) | lo;
    const u64 fix = (u64)(num / t);
    u64 pl = m * delta[j], ph = __umul64hi(m, delta[j]);
    pl += fix; ph += pl < fix;
    const u64 s = barrett128(pl, ph, Q.p, Q.r0, Q.r1);
    dest[idx] = subtract ? submod(base, s, Q.p) : addmod(base, s, Q.p);
}
// centralize: out[j][x] = m_x (m_x < (t+1)/2) or m_x + q - t, modulo q_j; zero beyond plain_count     (scaling_variant.cu:326-357)
// scalar != 1: the message is first multiplied by it modulo t (BGV add_plain: the ciphertext's correction factor, evaluator_translate_plain.cu:79-82)
__global__ void __launch_bounds__(256) centralize_kernel(const u64 *__restrict__ plain, u32 plain_count, u64 *__restrict__ out, const ModInfo *__restrict__ mods,
                                                         u64 t, u32 nl, u32 n, u64 scalar = 1, u64 t_r0 = 0, u64 t_r1 = 0) {
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (u64)nl * n) return;
    const u32 j = (u32)(idx / n), x = (u32)(idx % n);
    if (x >= plain_count) { out[idx] = 0; return; }
    const ModInfo &Q = mods[j];
    u64 m = plain[x];
    if (scalar != 1) m = mulmod(m, scalar, t, t_r0, t_r1);
    const u64 mq = barrett64(m, Q.p, Q.ratio);
    out[idx] = m >= ((t + 1) >> 1) ? submod(mq, barrett64(t, Q.p, Q.ratio), Q.p) : mq;
}

// key-switching key, digit i: c_0 limb i += (P mod q_i) * new_key limb i     (key_generator.cu:150-155)
__global__ void __launch_bounds__(256) kskey_add_kernel(u64 *__restrict__ c0_limb, const u64 *__restrict__ new_key_limb, Tw factor, u64 p, u32 n) {
    const u32 x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x < n) c0_limb[x] = addmod(c0_limb[x], shoup(new_key_limb[x], factor, p), p);
}

// BatchEncoder (src/batch_encoder.cu:14-60,106-130,345-365): slot i lives at coefficient index_map[i] of the NTT-form polynomial modulo t
__global__ void __launch_bounds__(256) batch_scatter_kernel(const u64 *__restrict__ values, u32 count, const u32 *__restrict__ map, u64 *__restrict__ out, u32 n, u64 t, u64 t_ratio) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[map[i]] = i < count ? barrett64(values[i], t, t_ratio) : 0;
}
__global__ void __launch_bounds__(256) batch_gather_kernel(const u64 *__restrict__ in, const u32 *__restrict__ map, u64 *__restrict__ out, u32 n) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[map[i]];
}
__global__ void __launch_bounds__(256) pad_copy_kernel(const u64 *__restrict__ in, u32 count, u64 *__restrict__ out, u32 n) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i < count ? in[i] : 0;
}
static b2_status ensure_batching(b2_ctx *ctx) {
    std::lock_guard<std::mutex> guard(ctx->mu);
    if (ctx->plain_ctx) return B2_OK;
    const size_t n = ctx->n;
    const uint64_t t = ctx->t;
    B2_REQUIRE(t > 1 && host::is_prime(t) && (t - 1) % (2 * n) == 0, "The parameters does not support vector batching.");      // batch_encoder.cu:191-193
    b2_params prm;
    prm.log_n = ctx->logn; prm.n_moduli = 1; prm.moduli = &t; prm.plain_modulus = 0; prm.scheme = B2_SCHEME_BFV;
    b2_ctx *pc = nullptr;
    b2_status st = b2_ctx_create(ctx->device, &prm, &pc);
    if (st != B2_OK) return st;
    // matrix_reps_index_map (batch_encoder.cu:46-58): rows of N/2 slots, positions generated by 3^i mod 2N
    std::vector<uint32_t> map(n);
    const size_t row = n >> 1, m = n << 1;
    size_t pos = 1;
    for (size_t i = 0; i < row; i++) {
        map[i] = host::bitrev((uint32_t)((pos - 1) >> 1), ctx->logn);
        map[i + row] = host::bitrev((uint32_t)((m - pos - 1) >> 1), ctx->logn);
        pos = (pos * 3) & (m - 1);
    }
    uint32_t *d = nullptr;
    cudaError_t e = cudaMalloc(&d, n * 4);
    if (e == cudaSuccess) e = cudaMemcpy(d, map.data(), n * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(d); b2_ctx_destroy(pc); return cuda_fail(e, "cudaMalloc(slot map)"); }
    ctx->d_slot_map = d;
    ctx->plain_ctx = pc;
    return B2_OK;
}

static uint64_t q_div_t_mod(const std::vector<uint64_t> &q, size_t nl, uint64_t t, uint64_t m) {     // floor(prod q / t) mod m
    std::vector<uint64_t> limbs{1};
    for (size_t i = 0; i < nl; i++) {
        uint64_t carry = 0;
        for (auto &l : limbs) { host::u128 v = (host::u128)l * q[i] + carry; l = (uint64_t)v; carry = (uint64_t)(v >> 64); }
        if (carry) limbs.push_back(carry);
    }
    std::vector<uint64_t> quo(limbs.size());
    host::u128 rem = 0;
    for (size_t k = limbs.size(); k-- > 0;) { host::u128 cur = (rem << 64) | limbs[k]; quo[k] = (uint64_t)(cur / t); rem = cur % t; }
    host::u128 r = 0;
    for (size_t k = quo.size(); k-- > 0;) r = ((r << 64) | quo[k]) % m;
    return (uint64_t)r;
}
static uint64_t prod_mod_t(const std::vector<uint64_t> &q, size_t nl, uint64_t t) {
    host::u128 r = 1 % t;
    for (size_t i = 0; i < nl; i++) r = (r * (q[i] % t)) % t;
    return (uint64_t)r;
}

static b2_status sample_locked(b2_ctx *ctx, int kind, b2_prng *g, u64 *out, size_t count, size_t nl, cudaStream_t stream) {
    const size_t n = ctx->n;
    const AesKey K = aes_expand(g->seed_lo, g->seed_hi);
    const Ctr c{g->counter_lo, g->counter_hi};
    if (kind == B2_SAMPLE_TERNARY) {
        const u64 per = (n + 15) / 16, threads = (u64)n * count;
        sample_ternary_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(K, c, out, ctx->d_mods, (u32)nl, (u32)n, (u32)count);
        advance(g, per * count);
    } else if (kind == B2_SAMPLE_CENTERED_BINOMIAL) {
        const u64 per = (n + 1) / 2, threads = (u64)n * count;
        sample_cbd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(K, c, out, ctx->d_mods, (u32)nl, (u32)n, (u32)count);
        advance(g, per * count);
    } else {
        // consecutive polynomials are consecutive in the stream: nl * N / 2 blocks each; the limb of a word follows from its position
        const u64 per = (u64)nl * n / 2;
        for (size_t i = 0; i < count; i++) {
            const u64 lo = c.lo + i * per;
            sample_uniform_kernel<<<(unsigned)((per + 255) / 256), 256, 0, stream>>>(K, Ctr{lo, c.hi + (lo < c.lo)}, out + i * nl * n, ctx->d_mods, (u32)n, per, 1);
            if (i + 1 < count) count_launches(1);
        }
        advance(g, per * count);
    }
    count_launches(1);
    B2_CUDA(cudaGetLastError());
    return B2_OK;
}

}  // namespace b2

extern "C" {

b2_status b2_prng_fill(b2_ctx *ctx, b2_prng *g, void *out, size_t bytes, void *stream_) {
    B2_REQUIRE(ctx && g && (out || bytes == 0), "null argument");
    B2_REQUIRE((((uintptr_t)out) & 15) == 0, "output must be 16-byte aligned");
    if (bytes == 0) return B2_OK;
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const AesKey K = aes_expand(g->seed_lo, g->seed_hi);
    const u64 main = bytes / 16, tail = bytes % 16;
    if (main) {
        sample_uniform_kernel<<<(unsigned)((main + 255) / 256), 256, 0, stream>>>(K, Ctr{g->counter_lo, g->counter_hi}, (u64 *)out, ctx->d_mods, 2u, main, 0);
        count_launches(1);
        B2_CUDA(cudaGetLastError());
        advance(g, main);
    }
    if (tail) {      // a partial last block consumes one more counter value (random_generator.cu:133-140); generated on the host like the reference
        uint64_t v[2];
        host_block(K, g->counter_lo, g->counter_hi, v);
        B2_CUDA(cudaMemcpyAsync((char *)out + 16 * main, v, tail, cudaMemcpyHostToDevice, stream));
        advance(g, 1);
    }
    return B2_OK;
}

uint64_t b2_prng_sample_uint64(b2_prng *g) {
    if (!g) return 0;
    const AesKey K = aes_expand(g->seed_lo, g->seed_hi);
    uint64_t v[2];
    host_block(K, g->counter_lo, g->counter_hi, v);
    advance(g, 1);
    return v[0];
}

b2_status b2_sample_poly(b2_ctx *ctx, int kind, b2_prng *g, uint64_t *out, size_t count, size_t nl, void *stream_) {
    B2_REQUIRE(ctx && g && out, "null argument");
    B2_REQUIRE(kind >= B2_SAMPLE_TERNARY && kind <= B2_SAMPLE_UNIFORM && nl >= 1 && nl <= ctx->K && ctx->n >= 2, "bad arguments");
    B2_REQUIRE((((uintptr_t)out) & 15) == 0, "output must be 16-byte aligned");
    if (count == 0) return B2_OK;
    B2_CUDA(cudaSetDevice(ctx->device));
    return sample_locked(ctx, kind, g, (u64 *)out, count, nl, (cudaStream_t)stream_);
}

// rlwe::asymmetric_with_u_prng + the level handling of Encryptor::encrypt_zero_internal: pk [2][K][N] (NTT form, key level) -> ct [2][nl][N]
b2_status b2_encrypt_zero_asymmetric(b2_ctx *ctx, const uint64_t *pk, size_t nl, int is_ntt_form, b2_prng *u_prng, b2_prng *e_prng, uint64_t *ct, void *stream_) {
    B2_REQUIRE(ctx && pk && u_prng && e_prng && ct, "null argument");
    B2_REQUIRE(nl >= 1 && nl <= ctx->K && ctx->n >= 2, "bad level");
    B2_REQUIRE((((uintptr_t)pk | (uintptr_t)ct) & 15) == 0, "operands must be 16-byte aligned");
    const bool down = nl < ctx->K;             // below the key level: encrypt one level up, then divide by the extra modulus (encryptor.cu:45-98)
    const bool bgv = ctx->scheme == B2_SCHEME_BGV, ntt = is_ntt_form != 0;
    B2_REQUIRE(!(down && bgv && !ntt), "BGV - Plaintext is not in NTT form.");
    B2_REQUIRE(!bgv || ctx->t > 1, "BGV needs a plain modulus");
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t n = ctx->n, K = ctx->K, L = down ? nl + 1 : nl;
    const u64 escale = bgv ? ctx->t : 1;
    Scratch sc(ctx, stream);
    // workspace: U [L][N] (also the intermediate of its transform) | E [2][L][N] | TMP [2][L][N] | mid [2][L][N] | the mod switch's [2][L][N]
    const size_t front = (size_t)(1 + 2 + 2 + 2) * L * n * 8;
    b2_status st = ensure_ws(sc.s, stream, front + 2 * L * n * 8);
    if (st != B2_OK) return st;
    u64 *U = (u64 *)sc->ws, *E = U + L * n, *TMP = E + 2 * L * n, *mid = TMP + 2 * L * n;
    u64 *C = down ? TMP : (u64 *)ct;
    st = sample_locked(ctx, B2_SAMPLE_TERNARY, u_prng, U, 1, L, stream);
    if (st != B2_OK) return st;
    st = sample_locked(ctx, B2_SAMPLE_CENTERED_BINOMIAL, e_prng, E, 2, L, stream);
    if (st != B2_OK) return st;
    std::vector<LimbGroup> g1, g2;
    for (size_t j = 0; j < L; j++) g1.push_back(LimbGroup{(u32)j, 1u, (u64)(j * n), 0});
    st = run_ntt_dir<true>(ctx, sc.s, g1, U, PlainLoad{U}, PkMulStore{C, (const u64 *)pk, ctx->d_mods, (u32)L, (u32)K, (u32)n}, stream);
    if (st != B2_OK) return st;
    for (size_t j = 0; j < L; j++) g2.push_back(LimbGroup{(u32)j, 2u, (u64)(j * n), (u64)(L * n)});
    if (ntt)      // c_k += scale * NTT(e_k)
        st = run_ntt_dir<true>(ctx, sc.s, g2, mid, PlainLoad{E}, AddScaledStore<false>{C, C, ctx->d_mods, (u32)n, escale, true}, stream);
    else          // c_k = INTT(c_k) + scale * e_k
        st = run_ntt_dir<false>(ctx, sc.s, g2, mid, PlainLoad{C}, AddScaledStore<false>{C, E, ctx->d_mods, (u32)n, escale, false}, stream);
    if (st != B2_OK || !down) return st;
    return mod_switch_locked(ctx, sc.s, TMP, (u64 *)ct, 2, L, bgv ? B2_MODSWITCH_BGV : (ntt ? B2_MODSWITCH_CKKS : B2_MODSWITCH_BFV), front, stream);
}

// rlwe::symmetric_with_c1_prng: sk [K][N] (NTT form) -> ct [2][nl][N] = (-(a s + e), a); *seed_out = the seed of a's generator (always
// returned; the reference stores it in the ciphertext only when save_seed)
b2_status b2_encrypt_zero_symmetric(b2_ctx *ctx, const uint64_t *sk, size_t nl, int is_ntt_form, int save_seed, b2_prng *c1_prng, b2_prng *e_prng, uint64_t *ct,
                                    uint64_t *seed_out, void *stream_) {
    B2_REQUIRE(ctx && sk && c1_prng && e_prng && ct, "null argument");
    B2_REQUIRE(nl >= 1 && nl <= ctx->K && ctx->n >= 2, "bad level");
    B2_REQUIRE((((uintptr_t)sk | (uintptr_t)ct) & 15) == 0, "operands must be 16-byte aligned");
    const bool bgv = ctx->scheme == B2_SCHEME_BGV, ntt = is_ntt_form != 0;
    B2_REQUIRE(!bgv || ctx->t > 1, "BGV needs a plain modulus");
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t n = ctx->n;
    const u64 escale = bgv ? ctx->t : 1;
    uint64_t seed = 0;
    while (seed == 0) seed = b2_prng_sample_uint64(c1_prng);
    if (seed_out) *seed_out = seed;
    b2_prng a_prng{seed, 0, 0, 0};
    Scratch sc(ctx, stream);
    // workspace: E [nl][N] | A [nl][N] | mid [2][nl][N]
    b2_status st = ensure_ws(sc.s, stream, 4 * nl * n * 8);
    if (st != B2_OK) return st;
    u64 *E = (u64 *)sc->ws, *A = E + nl * n, *mid = A + nl * n, *c0 = (u64 *)ct, *c1 = c0 + nl * n;
    const bool both = !ntt && !save_seed;      // coefficient form without a saved seed: the sample is NTT(a) and c1 = INTT(sample)
    st = sample_locked(ctx, B2_SAMPLE_UNIFORM, &a_prng, both ? A : c1, 1, nl, stream);
    if (st != B2_OK) return st;
    st = sample_locked(ctx, B2_SAMPLE_CENTERED_BINOMIAL, e_prng, E, 1, nl, stream);
    if (st != B2_OK) return st;
    std::vector<LimbGroup> g;
    for (size_t j = 0; j < nl; j++) g.push_back(LimbGroup{(u32)j, 1u, (u64)(j * n), 0});
    if (ntt)                 // c0 = -(sk o c1 + scale NTT(e))
        return run_ntt_dir<true>(ctx, sc.s, g, mid, PlainLoad{E}, SymNttStore{c0, c1, (const u64 *)sk, ctx->d_mods, (u32)n, escale}, stream);
    if (save_seed) {         // the sample is a in coefficient form; c1 is left as NTT(a) (rlwe.cu:269-274 never transforms it back)
        st = run_ntt_dir<true>(ctx, sc.s, g, c1, PlainLoad{c1}, PlainStore{c1}, stream);
        if (st != B2_OK) return st;
        // c0 = -(INTT(sk o c1) + scale e): the loader multiplies
        return run_ntt_dir<false>(ctx, sc.s, g, mid, SkMulLoad{c1, (const u64 *)sk, ctx->d_mods, (u32)nl, (u32)n},
                                  SymCoeffStore{c0, E, ctx->d_mods, (u32)nl, (u32)n, escale}, stream);
    }
    // ONE inverse transform over 2 nl limbs of the sample A: groups [0, nl) = sk o A -> c0 = -(. + scale e), groups [nl, 2 nl) = A -> c1
    for (size_t j = 0; j < nl; j++) g.push_back(LimbGroup{(u32)j, 1u, (u64)((nl + j) * n), 0});
    return run_ntt_dir<false>(ctx, sc.s, g, mid, SkMulLoad{A, (const u64 *)sk, ctx->d_mods, (u32)nl, (u32)n}, SymCoeffStore{c0, E, ctx->d_mods, (u32)nl, (u32)n, escale}, stream);
}

// BatchEncoder::encode_slice (src/batch_encoder.cu:169-227): values DEVICE [count <= N] -> plain DEVICE [N] modulo t (coefficient form): the values are
// placed at their slots' positions of the NTT-form polynomial (zero beyond count) and the inverse negacyclic transform modulo t is applied; two launches.
// (values are reduced modulo t on the way in; the reference expects them reduced already)
b2_status b2_batch_encode(b2_ctx *ctx, const uint64_t *values, size_t count, uint64_t *plain, void *stream_) {
    B2_REQUIRE(ctx && (values || count == 0) && plain, "null argument");
    B2_REQUIRE(count <= ctx->n, "Values has size larger than the number of slots.");
    B2_REQUIRE(values != plain, "out of place only");
    b2_status st = ensure_batching(ctx);
    if (st != B2_OK) return st;
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t n = ctx->n;
    batch_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((const u64 *)values, (u32)count, ctx->d_slot_map, (u64 *)plain, (u32)n, ctx->t, ctx->t_ratio);
    count_launches(1);
    B2_CUDA(cudaGetLastError());
    return b2_ntt_inverse(ctx->plain_ctx, plain, plain, 1, 1, 1, B2_TABLES_COMPONENTWISE, 0, 0, stream_);
}

// BatchEncoder::decode_slice (src/batch_encoder.cu:407-460): plain DEVICE [plain_count <= N] modulo t, coefficient form -> values DEVICE [N]; three launches
b2_status b2_batch_decode(b2_ctx *ctx, const uint64_t *plain, size_t plain_count, uint64_t *values, void *stream_) {
    B2_REQUIRE(ctx && plain && values, "null argument");
    B2_REQUIRE(plain != values, "out of place only");
    b2_status st = ensure_batching(ctx);
    if (st != B2_OK) return st;
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t n = ctx->n;
    u64 *tmp;
    {
        Scratch sc(ctx, stream);
        st = ensure_stage(sc.s, stream, n * 8);          // the staging buffer: the transform below takes the plain context's own scratch
        if (st != B2_OK) return st;
        tmp = (u64 *)sc->stage;
        pad_copy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((const u64 *)plain, (u32)std::min(plain_count, n), tmp, (u32)n);
        count_launches(1);
        B2_CUDA(cudaGetLastError());
    }
    st = b2_ntt_forward(ctx->plain_ctx, (const uint64_t *)tmp, (uint64_t *)tmp, 1, 1, 1, B2_TABLES_COMPONENTWISE, 0, 0, stream_);
    if (st != B2_OK) return st;
    batch_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(tmp, ctx->d_slot_map, (u64 *)values, (u32)n);
    count_launches(1);
    B2_CUDA(cudaGetLastError());
    return B2_OK;
}

// KeyGenerator::KeyGenerator (src/key_generator.cu:31-58): sk [K][N] = NTT(ternary sample), two launches
b2_status b2_generate_secret_key(b2_ctx *ctx, b2_prng *g, uint64_t *sk, void *stream_) {
    B2_REQUIRE(ctx && g && sk && ((((uintptr_t)sk) & 15) == 0), "null or misaligned argument");
    B2_CUDA(cudaSetDevice(ctx->device));
    b2_status st = sample_locked(ctx, B2_SAMPLE_TERNARY, g, (u64 *)sk, 1, ctx->K, (cudaStream_t)stream_);
    if (st != B2_OK) return st;
    return b2_ntt_forward(ctx, sk, sk, 1, 1, ctx->K, B2_TABLES_COMPONENTWISE, 0, 0, stream_);
}

// KeyGenerator::generate_one_kswitch_key (src/key_generator.cu:136-157): out [K-1][2][K][N]; digit i is a symmetric encryption of zero at the key level
// (NTT form, generator g for a's seed and for the noise, as rlwe::symmetric does) with (P mod q_i) new_key added into limb i of c_0.
// new_key [K][N] NTT form: s^2 for relinearisation keys (b2_secret_key_powers), the Galois image of s for rotation keys (b2_galois, NTT form).
b2_status b2_generate_kswitch_key(b2_ctx *ctx, const uint64_t *sk, const uint64_t *new_key, b2_prng *g, uint64_t *out, void *stream_) {
    B2_REQUIRE(ctx && sk && new_key && g && out, "null argument");
    B2_REQUIRE(ctx->K >= 2, "key switching needs a special prime");
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t n = ctx->n, K = ctx->K, per = 2 * K * n;
    const uint64_t P = ctx->q[K - 1];
    for (size_t i = 0; i + 1 < K; i++) {
        uint64_t *d = out + i * per;
        b2_status st = b2_encrypt_zero_symmetric(ctx, sk, K, 1, 0, g, g, d, nullptr, stream_);
        if (st != B2_OK) return st;
        const uint64_t qi = ctx->q[i], f = P % qi;
        Tw fw; fw.x = f; fw.y = host::shoup_quotient(f, qi);
        kskey_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((u64 *)d + i * n, (const u64 *)new_key + i * n, fw, qi, (u32)n);
        count_launches(1);
        B2_CUDA(cudaGetLastError());
    }
    return B2_OK;
}

// scaling_variant::scale_up (BFV; Encryptor: multiply_add_plain_inplace on c_0; Evaluator::add_plain / sub_plain): plain = plain_count words
// modulo t; dest / from [nl][N] (from == NULL: zero; from == dest allowed)
b2_status b2_bfv_scale_up(b2_ctx *ctx, const uint64_t *plain, size_t plain_count, size_t nl, const uint64_t *from, int subtract, uint64_t *dest, void *stream_) {
    B2_REQUIRE(ctx && plain && dest, "null argument");
    B2_REQUIRE(ctx->t > 1 && nl >= 1 && nl <= ctx->K && plain_count <= ctx->n, "BFV scale-up needs a plain modulus, 1 <= nl <= K and at most N coefficients");
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t n = ctx->n;
    std::vector<uint64_t> delta(nl);
    for (size_t j = 0; j < nl; j++) delta[j] = q_div_t_mod(ctx->q, nl, ctx->t, ctx->q[j]);
    Scratch sc(ctx, stream);
    b2_status st = ensure_ptrtab(sc.s, stream, nl * 8);
    if (st != B2_OK) return st;
    B2_CUDA(cudaMemcpyAsync(sc->ptrtab, delta.data(), nl * 8, cudaMemcpyHostToDevice, stream));
    const u64 total = (u64)nl * n;
    scale_up_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>((const u64 *)plain, (u32)plain_count, (const u64 *)from, (u64 *)dest, ctx->d_mods, (const u64 *)sc->ptrtab,
                                                                        prod_mod_t(ctx->q, nl, ctx->t), ctx->t, (u32)nl, (u32)n, subtract);
    count_launches(1);
    B2_CUDA(cudaGetLastError());
    return B2_OK;
}

// scaling_variant::centralize: plain (plain_count words modulo t) -> out [nl][N], coefficient form
b2_status b2_centralize(b2_ctx *ctx, const uint64_t *plain, size_t plain_count, size_t nl, uint64_t *out, void *stream_) {
    B2_REQUIRE(ctx && plain && out, "null argument");
    B2_REQUIRE(ctx->t > 1 && nl >= 1 && nl <= ctx->K && plain_count <= ctx->n, "centralize needs a plain modulus, 1 <= nl <= K and at most N coefficients");
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const u64 total = (u64)nl * ctx->n;
    centralize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>((const u64 *)plain, (u32)plain_count, (u64 *)out, ctx->d_mods, ctx->t, (u32)nl, (u32)ctx->n);
    count_launches(1);
    B2_CUDA(cudaGetLastError());
    return B2_OK;
}

// centralize of (scalar * m mod t): BGV Evaluator::add_plain / sub_plain multiply the message by the ciphertext's correction factor first
// (src/evaluator_translate_plain.cu:79-82, utils::multiply_scalar modulo t)
b2_status b2_centralize_scaled(b2_ctx *ctx, const uint64_t *plain, size_t plain_count, size_t nl, uint64_t scalar, uint64_t *out, void *stream_) {
    B2_REQUIRE(ctx && plain && out, "null argument");
    B2_REQUIRE(ctx->t > 1 && nl >= 1 && nl <= ctx->K && plain_count <= ctx->n, "centralize needs a plain modulus, 1 <= nl <= K and at most N coefficients");
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const u64 total = (u64)nl * ctx->n;
    centralize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>((const u64 *)plain, (u32)plain_count, (u64 *)out, ctx->d_mods, ctx->t, (u32)nl, (u32)ctx->n,
                                                                          scalar % ctx->t, ctx->t_r0, ctx->t_r1);
    count_launches(1);
    B2_CUDA(cudaGetLastError());
    return B2_OK;
}

// Encryptor::encrypt_internal (encryptor.cu:247-338).  BFV / BGV: plain = plain_count words modulo t (no parms_id), nl = the first level;
// CKKS: plain = [nl][N] in NTT form (plain_count ignored).  key = pk (asymmetric) or sk.  ct [2][nl][N].
b2_status b2_encrypt(b2_ctx *ctx, const uint64_t *plain, size_t plain_count, const uint64_t *key, int asymmetric, size_t nl, b2_prng *u_prng, b2_prng *e_prng,
                     uint64_t *ct, void *stream_) {
    B2_REQUIRE(ctx && plain && key && ct && u_prng && e_prng, "null argument");
    const int ntt = ctx->scheme != B2_SCHEME_BFV;
    b2_status st = asymmetric ? b2_encrypt_zero_asymmetric(ctx, key, nl, ntt, u_prng, e_prng, ct, stream_)
                              : b2_encrypt_zero_symmetric(ctx, key, nl, ntt, 0, u_prng, e_prng, ct, nullptr, stream_);
    if (st != B2_OK) return st;
    if (ctx->scheme == B2_SCHEME_BFV) return b2_bfv_scale_up(ctx, plain, plain_count, nl, ct, 0, ct, stream_);
    if (ctx->scheme == B2_SCHEME_CKKS) return b2_add(ctx, ct, plain, ct, 1, nl, 0, stream_);
    // BGV: c_0 += NTT(centralize(m))
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t n = ctx->n;
    Scratch sc(ctx, stream);
    st = ensure_ws(sc.s, stream, 2 * nl * n * 8);
    if (st != B2_OK) return st;
    u64 *M = (u64 *)sc->ws, *mid = M + nl * n;
    B2_REQUIRE(ctx->t > 1 && plain_count <= n, "BGV needs a plain modulus and at most N coefficients");
    const u64 total = (u64)nl * n;
    centralize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>((const u64 *)plain, (u32)plain_count, M, ctx->d_mods, ctx->t, (u32)nl, (u32)n);
    count_launches(1);
    B2_CUDA(cudaGetLastError());
    std::vector<LimbGroup> g;
    for (size_t j = 0; j < nl; j++) g.push_back(LimbGroup{(u32)j, 1u, (u64)(j * n), 0});
    return run_ntt_dir<true>(ctx, sc.s, g, mid, PlainLoad{M}, AddScaledStore<false>{(u64 *)ct, (const u64 *)ct, ctx->d_mods, (u32)n, 1, true}, stream);
}

}  // extern "C"
