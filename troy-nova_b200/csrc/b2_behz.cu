// b2_behz.cu — BFV ciphertext multiplication, BEHZ style (M2 in SURVEY.md §8a).  Replaces the device branch of
// Evaluator::bfv_multiply / bfv_square (src/evaluator.cu:29-116,199-282) together with
// RNSTool::fast_b_conv_m_tilde_sm_mrq and RNSTool::fast_floor_fast_b_conv_sk (src/fgk/rns_tool.cu:7-343) and the
// constants of RNSTool::RNSTool (src/utils/rns_tool.cu:29-275).
//
// Six launches per call (any batch size): lift (steps 1-2, all four input polynomials), ONE forward transform over the
// q limbs and the Bsk limbs of both operands, one dyadic convolution over both bases, ONE inverse transform, floor
// (steps 6-8).  The reference launches 4 lift kernels, 4 + 4 NTT passes, 2 convolutions, 4 INTT passes and 1 floor
// kernel with five global-memory scratch arrays; here the per-coefficient scratch lives in registers.
//
// The inexact steps (fast base conversions, Montgomery reduction modulo m_tilde = 2^32, fast floor, Shenoy-Kumaresan)
// are reproduced as specified; every stored value is the canonical residue the reference stores.
#include <map>
#include <algorithm>
#include "b2_ntt_launch.cuh"
#include "b2_host_math.h"

using namespace b2;

// device-resident constants of one level (L = ciphertext limbs, nB = |B|, S = |Bsk| = nB + 1)
struct BehzDev {
    u32 L, nB, S, K;
    const Tw *mt_punc;      // [L]    m_tilde * (q/q_i)^-1 mod q_i
    const Tw *t_punc;       // [L]    t * (q/q_i)^-1 mod q_i
    const u64 *q_to_bsk;    // [S][L] (q/q_i) mod bsk_o
    const u64 *q_to_mt;     // [L]    (q/q_i) mod 2^32
    u64 neg_inv_q_mt;       //        -q^-1 mod 2^32
    const Tw *q_mod_bsk;    // [S]    q mod bsk_o
    const Tw *inv_mt_bsk;   // [S]    m_tilde^-1 mod bsk_o
    const Tw *inv_q_bsk;    // [S]    q^-1 mod bsk_o
    const Tw *t_bsk;        // [S]    t mod bsk_o
    const Tw *punc_B;       // [nB]   (B/B_j)^-1 mod B_j
    const u64 *B_to_q;      // [L][nB] (B/B_j) mod q_i
    const u64 *B_to_msk;    // [nB]   (B/B_j) mod m_sk
    Tw inv_B_msk;           //        B^-1 mod m_sk
    const Tw *B_mod_q;      // [L]    B mod q_i
    const Tw *negB_mod_q;   // [L]    q_i - (B mod q_i)
};
// ModInfo slot of Bsk element o at this level: B_o for o < nB, m_sk last
__device__ __host__ __forceinline__ u32 bsk_slot(u32 K, u32 nB, u32 o) { return o < nB ? K + o : 2 * K + 1; }

struct b2_behz {
    std::vector<uint64_t> aux;            // B_0 .. B_K, m_sk  (host copy)
    uint64_t gamma = 0;
    Tw *d_tw = nullptr, *d_itw = nullptr; // [K+2][N] twiddles of the auxiliary primes (integer pipe: they are 61-bit)
    struct Level { BehzDev dev; void *blob; };
    std::map<size_t, Level> levels;
};

namespace b2 {

void behz_release(b2_ctx *ctx) {
    if (!ctx->behz) return;
    cudaFree(ctx->behz->d_tw); cudaFree(ctx->behz->d_itw);
    for (auto &kv : ctx->behz->levels) cudaFree(kv.second.blob);
    delete ctx->behz; ctx->behz = nullptr;
}

static uint64_t prod_except(const std::vector<uint64_t> &base, size_t count, size_t skip, uint64_t m) {
    host::u128 r = 1 % m;
    for (size_t i = 0; i < count; i++) if (i != skip) r = (r * (base[i] % m)) % m;
    return (uint64_t)r;
}
static Tw shoup_pair(uint64_t w, uint64_t p) { Tw t; t.x = w; t.y = host::shoup_quotient(w, p); return t; }

// auxiliary primes + their NTT tables (once per context)
static b2_status ensure_behz(b2_ctx *ctx) {
    if (ctx->behz) return B2_OK;
    const size_t n = ctx->n, K = ctx->K, A = K + 2;
    std::vector<uint64_t> pr;
    // rns_tool.cu:68-80: m_sk, gamma, B_0, B_1, ... = successive 61-bit primes = 1 mod 2N, descending from 2^61
    if (!host::primes_below(2 * (uint64_t)n, 61, K + 3, pr)) { set_error("b2_bfv_multiply: not enough 61-bit primes = 1 mod 2N"); return B2_ERR_INVALID_ARG; }
    b2_behz *bz = new b2_behz();
    bz->gamma = pr[1];
    for (size_t j = 0; j <= K; j++) bz->aux.push_back(pr[2 + j]);
    bz->aux.push_back(pr[0]);
    std::vector<Tw> tw(A * n), itw(A * n);
    std::vector<ModInfo> mods(A);
    for (size_t a = 0; a < A; a++) {
        uint64_t psi;
        if (build_modulus_tables(bz->aux[a], ctx->logn, &tw[a * n], &itw[a * n], nullptr, nullptr, mods[a], psi)) { delete bz; set_error("b2_bfv_multiply: auxiliary prime unusable"); return B2_ERR_INVALID_ARG; }
    }
    cudaError_t e;
    if ((e = cudaMalloc(&bz->d_tw, A * n * sizeof(Tw))) != cudaSuccess || (e = cudaMalloc(&bz->d_itw, A * n * sizeof(Tw))) != cudaSuccess) {
        cudaFree(bz->d_tw); delete bz; return cuda_fail(e, "cudaMalloc(BEHZ tables)");
    }
    for (size_t a = 0; a < A; a++) { mods[a].tw = bz->d_tw + a * n; mods[a].itw = bz->d_itw + a * n; }
    // upload everything first, publish last: a failed upload must not leave a half-initialised BEHZ state behind
    if ((e = cudaMemcpy(bz->d_tw, tw.data(), A * n * sizeof(Tw), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(bz->d_itw, itw.data(), A * n * sizeof(Tw), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(ctx->d_mods + K, mods.data(), A * sizeof(ModInfo), cudaMemcpyHostToDevice)) != cudaSuccess) {
        cudaFree(bz->d_tw); cudaFree(bz->d_itw); delete bz; return cuda_fail(e, "cudaMemcpy(BEHZ tables)");
    }
    // host copies: run_ntt reads the arithmetic path of every table from h_mods.  The capacity for these 2K+2 records was
    // reserved by b2_ctx_create, so appending never moves the first K records that other threads may be reading.
    ctx->h_mods.resize(K);
    ctx->h_mods.insert(ctx->h_mods.end(), mods.begin(), mods.end());
    ctx->behz = bz;
    return B2_OK;
}

// constants of the level with L ciphertext limbs (RNSTool::RNSTool for base_q = q_0..q_{L-1})
static b2_status ensure_behz_level(b2_ctx *ctx, size_t L, const BehzDev **out) {
    std::lock_guard<std::mutex> guard(ctx->mu);      // lazy, once per (context, level): concurrent first calls serialise here
    b2_status st = ensure_behz(ctx);
    if (st != B2_OK) return st;
    b2_behz *bz = ctx->behz;
    auto it = bz->levels.find(L);
    if (it != bz->levels.end()) { *out = &it->second.dev; return B2_OK; }
    const size_t K = ctx->K;
    const uint64_t t = ctx->t, MT = (uint64_t)1 << 32;
    std::vector<uint64_t> q(ctx->q.begin(), ctx->q.begin() + L);
    // |B| = |q| (+1 if 32 + bits(t) + bits(prod q) >= 61 |q| + 61)            rns_tool.cu:52-62
    size_t qbits;
    {
        std::vector<uint64_t> limbs{1};
        for (uint64_t p : q) {
            uint64_t carry = 0;
            for (auto &l : limbs) { host::u128 v = (host::u128)l * p + carry; l = (uint64_t)v; carry = (uint64_t)(v >> 64); }
            if (carry) limbs.push_back(carry);
        }
        qbits = 64 * (limbs.size() - 1) + host::bit_count(limbs.back());
    }
    size_t nB = L;
    if (32 + host::bit_count(t) + qbits >= 61 * L + 61) nB++;
    const size_t S = nB + 1;
    std::vector<uint64_t> B(bz->aux.begin(), bz->aux.begin() + nB), bsk(B);
    const uint64_t msk = bz->aux[K + 1];
    bsk.push_back(msk);
    // one blob: Tw arrays first (16-byte aligned), then u64 arrays
    const size_t n_tw = 2 * L + 4 * S + nB + 2 * L, n_u64 = S * L + L + L * nB + nB;
    std::vector<Tw> tws; tws.reserve(n_tw);
    std::vector<uint64_t> u64s; u64s.reserve(n_u64);
    for (size_t i = 0; i < L; i++) { uint64_t ip = host::invmod(prod_except(q, L, i, q[i]), q[i]); tws.push_back(shoup_pair(host::mulmod(MT % q[i], ip, q[i]), q[i])); }
    for (size_t i = 0; i < L; i++) { uint64_t ip = host::invmod(prod_except(q, L, i, q[i]), q[i]); tws.push_back(shoup_pair(host::mulmod(t % q[i], ip, q[i]), q[i])); }
    for (size_t o = 0; o < S; o++) tws.push_back(shoup_pair(prod_except(q, L, (size_t)-1, bsk[o]), bsk[o]));
    for (size_t o = 0; o < S; o++) tws.push_back(shoup_pair(host::invmod(MT % bsk[o], bsk[o]), bsk[o]));
    for (size_t o = 0; o < S; o++) tws.push_back(shoup_pair(host::invmod(prod_except(q, L, (size_t)-1, bsk[o]), bsk[o]), bsk[o]));
    for (size_t o = 0; o < S; o++) tws.push_back(shoup_pair(t % bsk[o], bsk[o]));
    for (size_t j = 0; j < nB; j++) tws.push_back(shoup_pair(host::invmod(prod_except(B, nB, j, B[j]), B[j]), B[j]));
    for (size_t i = 0; i < L; i++) tws.push_back(shoup_pair(prod_except(B, nB, (size_t)-1, q[i]), q[i]));
    for (size_t i = 0; i < L; i++) tws.push_back(shoup_pair(q[i] - prod_except(B, nB, (size_t)-1, q[i]), q[i]));
    for (size_t o = 0; o < S; o++) for (size_t i = 0; i < L; i++) u64s.push_back(prod_except(q, L, i, bsk[o]));
    for (size_t i = 0; i < L; i++) u64s.push_back(prod_except(q, L, i, MT));
    for (size_t i = 0; i < L; i++) for (size_t j = 0; j < nB; j++) u64s.push_back(prod_except(B, nB, j, q[i]));
    for (size_t j = 0; j < nB; j++) u64s.push_back(prod_except(B, nB, j, msk));
    void *blob = nullptr;
    B2_CUDA(cudaMalloc(&blob, tws.size() * sizeof(Tw) + u64s.size() * 8));
    B2_CUDA(cudaMemcpy(blob, tws.data(), tws.size() * sizeof(Tw), cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy((char *)blob + tws.size() * sizeof(Tw), u64s.data(), u64s.size() * 8, cudaMemcpyHostToDevice));
    BehzDev d;
    d.L = (u32)L; d.nB = (u32)nB; d.S = (u32)S; d.K = (u32)K;
    const Tw *tp = (const Tw *)blob;
    d.mt_punc = tp; tp += L; d.t_punc = tp; tp += L; d.q_mod_bsk = tp; tp += S; d.inv_mt_bsk = tp; tp += S; d.inv_q_bsk = tp; tp += S;
    d.t_bsk = tp; tp += S; d.punc_B = tp; tp += nB; d.B_mod_q = tp; tp += L; d.negB_mod_q = tp; tp += L;
    const u64 *up = (const u64 *)tp;
    d.q_to_bsk = up; up += S * L; d.q_to_mt = up; up += L; d.B_to_q = up; up += L * nB; d.B_to_msk = up;
    // -q^-1 mod 2^32 (q odd): Newton iteration
    {
        uint64_t a = prod_except(q, L, (size_t)-1, MT), x = a;
        for (int i = 0; i < 6; i++) x *= 2 - a * x;
        d.neg_inv_q_mt = (MT - (x & (MT - 1))) & (MT - 1);
    }
    d.inv_B_msk = shoup_pair(host::invmod(prod_except(B, nB, (size_t)-1, msk), msk), msk);
    auto &lv = bz->levels[L];
    lv.dev = d; lv.blob = blob;
    *out = &lv.dev;
    return B2_OK;
}

// exact sum_i a_i * m_i modulo M.p (dot_product_mod): 128-bit accumulation, at most 64 terms of < 2^122 each
template <int LM>
__device__ __forceinline__ u64 dot_mod(const u64 (&a)[LM], const u64 *__restrict__ m, u32 cnt, const ModInfo &M) {
    u64 lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < LM; i++) if ((u32)i < cnt) mac128(lo, hi, a[i], __ldg(m + i));
    return barrett128(lo, hi, M.p, M.r0, M.r1);
}

// ---- steps (1)-(2): x (base q) -> Bsk, Montgomery-reduced modulo m_tilde      (src/fgk/rns_tool.cu:29-98)
// grid.y = polynomial index c in [0, 4 batch): batch item c>>2, (operand a/b, poly 0/1) = c&3
template <int LM>
__global__ void __launch_bounds__(128) behz_lift_kernel(const u64 *__restrict__ a, const u64 *__restrict__ b, u64 *__restrict__ lifted,
                                                          const ModInfo *__restrict__ mods, const BehzDev D, u32 n) {
    const u32 x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    const u32 c = blockIdx.y, bi = c >> 2, w = c & 3;
    const u64 *src = ((w < 2) ? a : b) + ((u64)(bi * 2 + (w & 1)) * D.L) * n + x;
    u64 tmp[LM];
    u64 c2 = 0;
#pragma unroll
    for (int i = 0; i < LM; i++) if ((u32)i < D.L) {
        tmp[i] = shoup(src[(u64)i * n], __ldg(D.mt_punc + i), mods[i].p);
        c2 += tmp[i] * __ldg(D.q_to_mt + i);
    }
    const u64 MT = 1ull << 32;
    const u64 r = ((c2 & (MT - 1)) * D.neg_inv_q_mt) & (MT - 1);
    u64 *dst = lifted + ((u64)c * D.S) * n + x;
    for (u32 o = 0; o < D.S; o++) {
        const ModInfo &M = mods[bsk_slot(D.K, D.nB, o)];
        const u64 m = M.p;
        const u64 c1 = dot_mod<LM>(tmp, D.q_to_bsk + (u64)o * D.L, D.L, M);
        const u64 tt = r >= (MT >> 1) ? r + m - MT : r;
        const u64 v = addmod(shoup(tt, __ldg(D.q_mod_bsk + o), m), c1, m);
        dst[(u64)o * n] = shoup(v, __ldg(D.inv_mt_bsk + o), m);
    }
}

// ---- step (4): out[bi][p] = sum_{x+y=p} A[bi][x] o B[bi][y], over the q limbs and the Bsk limbs in one launch.
// W layout [4 batch][nl][N] (nl = L or S), polynomials 4bi, 4bi+1 = operand a, 4bi+2, 4bi+3 = operand b; out [3 batch][nl][N]
__global__ void __launch_bounds__(256) behz_convolute_kernel(const ulonglong2 *__restrict__ wq, const ulonglong2 *__restrict__ wb, ulonglong2 *__restrict__ oq,
                                                               ulonglong2 *__restrict__ ob, const ModInfo *__restrict__ mods, const BehzDev D, u32 n2, u32 batch) {
    const u64 per = (u64)(D.L + D.S) * n2, total = (u64)batch * per;
    for (u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (u64)gridDim.x * blockDim.x) {
        const u64 bi = v / per, r = v % per;
        const u32 limb = (u32)(r / n2), x = (u32)(r % n2);
        const bool isq = limb < D.L;
        const u32 l = isq ? limb : limb - D.L, nl = isq ? D.L : D.S;
        const ModInfo &M = mods[isq ? l : bsk_slot(D.K, D.nB, l)];
        const u64 p = M.p, r0 = M.r0, r1 = M.r1;
        const ulonglong2 *w = (isq ? wq : wb) + ((u64)(4 * bi) * nl + l) * n2 + x;
        const u64 ps = (u64)nl * n2;   // polynomial stride
        const ulonglong2 a0 = w[0], a1 = w[ps], b0 = w[2 * ps], b1 = w[3 * ps];
        ulonglong2 *o = (isq ? oq : ob) + ((u64)(3 * bi) * nl + l) * n2 + x;
        ulonglong2 res;
        res.x = mulmod(a0.x, b0.x, p, r0, r1); res.y = mulmod(a0.y, b0.y, p, r0, r1);
        o[0] = res;
        u64 lo = 0, hi = 0, lo2 = 0, hi2 = 0;
        mac128(lo, hi, a0.x, b1.x); mac128(lo, hi, a1.x, b0.x); mac128(lo2, hi2, a0.y, b1.y); mac128(lo2, hi2, a1.y, b0.y);
        res.x = barrett128(lo, hi, p, r0, r1); res.y = barrett128(lo2, hi2, p, r0, r1);
        o[ps] = res;
        res.x = mulmod(a1.x, b1.x, p, r0, r1); res.y = mulmod(a1.y, b1.y, p, r0, r1);
        o[2 * ps] = res;
    }
}

// ---- steps (6)-(8): x t, fast floor to Bsk, Shenoy-Kumaresan back to q        (src/fgk/rns_tool.cu:186-285)
// grid.y = output polynomial index in [0, 3 batch)
template <int LM>
__global__ void __launch_bounds__(128) behz_floor_kernel(const u64 *__restrict__ yq, const u64 *__restrict__ yb, u64 *__restrict__ dest,
                                                           const ModInfo *__restrict__ mods, const BehzDev D, u32 n) {
    const u32 x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    const u32 k = blockIdx.y;
    const u64 *sq = yq + ((u64)k * D.L) * n + x, *sb = yb + ((u64)k * D.S) * n + x;
    u64 tq[LM], h[LM + 1];
#pragma unroll
    for (int i = 0; i < LM; i++) if ((u32)i < D.L) tq[i] = shoup(sq[(u64)i * n], __ldg(D.t_punc + i), mods[i].p);
    u64 g_msk = 0;
#pragma unroll
    for (int o = 0; o < LM + 2; o++) if ((u32)o < D.S) {
        const ModInfo &M = mods[bsk_slot(D.K, D.nB, o)];
        const u64 m = M.p;
        const u64 zz = shoup(sb[(u64)o * n], __ldg(D.t_bsk + o), m);
        const u64 f = dot_mod<LM>(tq, D.q_to_bsk + (u64)o * D.L, D.L, M);
        const u64 g = shoup(zz + m - f, __ldg(D.inv_q_bsk + o), m);
        if ((u32)o < D.nB) { if (o < LM + 1) h[o < LM + 1 ? o : 0] = shoup(g, __ldg(D.punc_B + o), m); }
        else g_msk = g;
    }
    const ModInfo &MS = mods[2 * D.K + 1];
    const u64 msk = MS.p;
    u64 alpha;
    {
        u64 lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < LM + 1; j++) if ((u32)j < D.nB) mac128(lo, hi, h[j], __ldg(D.B_to_msk + j));
        const u64 ms = barrett128(lo, hi, msk, MS.r0, MS.r1);
        alpha = shoup(ms + (msk - g_msk), D.inv_B_msk, msk);
    }
    const bool neg = alpha > (msk >> 1);
    const u64 av = neg ? msk - alpha : alpha;
    u64 *dp = dest + ((u64)k * D.L) * n + x;
    for (u32 i = 0; i < D.L; i++) {
        const ModInfo &M = mods[i];
        const u64 q = M.p;
        u64 lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < LM + 1; j++) if ((u32)j < D.nB) mac128(lo, hi, h[j], __ldg(D.B_to_q + (u64)i * D.nB + j));
        const u64 d = barrett128(lo, hi, q, M.r0, M.r1);
        const u64 corr = shoup(av, __ldg((neg ? D.B_mod_q : D.negB_mod_q) + i), q);
        dp[(u64)i * n] = addmod(corr, d, q);
    }
}

// forward-transform loader: q groups read the caller's operands in place, Bsk groups read the lifted buffer
struct BehzLoad {
    const u64 *a, *b, *lifted;
    u32 L, S, n;
    __device__ __forceinline__ const u64 *src(int g, int c, u64) const {
        if ((u32)g < L) { const u32 bi = (u32)c >> 2, w = (u32)c & 3; return ((w < 2) ? a : b) + ((u64)(bi * 2 + (w & 1)) * L + g) * n; }
        return lifted + ((u64)c * S + ((u32)g - L)) * n;
    }
    B2_IDENTITY_XF
};

}  // namespace b2

extern "C" b2_status b2_bfv_multiply(b2_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t batch, size_t nl, void *stream_) {
    B2_REQUIRE(ctx && a && b && out, "null argument");
    B2_REQUIRE(ctx->scheme == B2_SCHEME_BFV, "context scheme is not BFV");                       // evaluator.cu:175-197 dispatch
    B2_REQUIRE(ctx->t > 1, "BFV needs a plain modulus");
    B2_REQUIRE(nl >= 1 && nl <= ctx->K, "limb count out of range");
    B2_REQUIRE(ctx->logn >= 1, "degree too small");
    B2_REQUIRE((((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) == 0, "operands must be 16-byte aligned");
    if (batch == 0) return B2_OK;
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const BehzDev *Dp = nullptr;
    b2_status st = ensure_behz_level(ctx, nl, &Dp);
    if (st != B2_OK) return st;
    const BehzDev D = *Dp;
    const size_t n = ctx->n, K = ctx->K, L = nl, S = D.S;
    // workspace: lifted [4B][S][N] | wq [4B][L][N] | wb [4B][S][N] | pq [3B][L][N] | pb [3B][S][N]
    const size_t w_l = 4 * batch * S * n, w_q = 4 * batch * L * n, w_b = w_l, w_pq = 3 * batch * L * n, w_pb = 3 * batch * S * n;
    Scratch scl(ctx, stream);
    b2_scratch *sc = scl.s;
    st = ensure_ws(sc, stream, (w_l + w_q + w_b + w_pq + w_pb) * 8);
    if (st != B2_OK) return st;
    u64 *lifted = (u64 *)sc->ws, *wq = lifted + w_l, *wb = wq + w_q, *pq = wb + w_b, *pb = pq + w_pq;
    const bool small = L <= 8;
    {   // (1)-(2)
        dim3 grid((unsigned)((n + 127) / 128), (unsigned)(4 * batch));
        if (small) behz_lift_kernel<8><<<grid, 128, 0, stream>>>((const u64 *)a, (const u64 *)b, lifted, ctx->d_mods, D, (u32)n);
        else behz_lift_kernel<64><<<grid, 128, 0, stream>>>((const u64 *)a, (const u64 *)b, lifted, ctx->d_mods, D, (u32)n);
        count_launches(1);
        B2_CUDA(cudaGetLastError());
    }
    {   // (3) forward transforms: L groups in base q + S groups in base Bsk, 4 batch limbs each.  The intermediate / output
        // buffer is addressed relative to wq; wb follows it contiguously, so the Bsk groups use base offsets beyond w_q.
        std::vector<LimbGroup> groups;
        for (size_t j = 0; j < L; j++) groups.push_back(LimbGroup{(u32)j, (u32)(4 * batch), (u64)(j * n), (u64)(L * n)});
        for (size_t o = 0; o < S; o++) groups.push_back(LimbGroup{bsk_slot((u32)K, D.nB, (u32)o), (u32)(4 * batch), (u64)(w_q + o * n), (u64)(S * n)});
        BehzLoad ld{(const u64 *)a, (const u64 *)b, lifted, (u32)L, (u32)S, (u32)n};
        st = run_ntt_dir<true>(ctx, sc, groups, wq, ld, PlainStore{wq}, stream);
        if (st != B2_OK) return st;
    }
    {   // (4)
        const u64 total = (u64)batch * (L + S) * (n / 2);
        behz_convolute_kernel<<<(unsigned)std::min<u64>((total + 255) / 256, 148u * 32u), 256, 0, stream>>>((const ulonglong2 *)wq, (const ulonglong2 *)wb, (ulonglong2 *)pq,
                                                                                                            (ulonglong2 *)pb, ctx->d_mods, D, (u32)(n / 2), (u32)batch);
        count_launches(1);
        B2_CUDA(cudaGetLastError());
    }
    {   // (5) inverse transforms in place (pq and pb are contiguous)
        std::vector<LimbGroup> groups;
        for (size_t j = 0; j < L; j++) groups.push_back(LimbGroup{(u32)j, (u32)(3 * batch), (u64)(j * n), (u64)(L * n)});
        for (size_t o = 0; o < S; o++) groups.push_back(LimbGroup{bsk_slot((u32)K, D.nB, (u32)o), (u32)(3 * batch), (u64)(w_pq + o * n), (u64)(S * n)});
        st = run_ntt_dir<false>(ctx, sc, groups, pq, PlainLoad{pq}, PlainStore{pq}, stream);
        if (st != B2_OK) return st;
    }
    {   // (6)-(8)
        dim3 grid((unsigned)((n + 127) / 128), (unsigned)(3 * batch));
        if (small) behz_floor_kernel<8><<<grid, 128, 0, stream>>>(pq, pb, (u64 *)out, ctx->d_mods, D, (u32)n);
        else behz_floor_kernel<64><<<grid, 128, 0, stream>>>(pq, pb, (u64 *)out, ctx->d_mods, D, (u32)n);
        count_launches(1);
        B2_CUDA(cudaGetLastError());
    }
    return B2_OK;
}

// introspection for tests: the auxiliary base chosen for level nl (B..., m_sk) and gamma — must equal the reference's
extern "C" b2_status b2_behz_bases(b2_ctx *ctx, size_t nl, uint64_t *bsk, size_t *bsk_count, uint64_t *gamma) {
    B2_REQUIRE(ctx && bsk && bsk_count && gamma && nl >= 1 && nl <= ctx->K && ctx->t > 1, "bad arguments");
    B2_CUDA(cudaSetDevice(ctx->device));
    const BehzDev *D = nullptr;
    b2_status st = ensure_behz_level(ctx, nl, &D);
    if (st != B2_OK) return st;
    for (u32 o = 0; o < D->nB; o++) bsk[o] = ctx->behz->aux[o];
    bsk[D->nB] = ctx->behz->aux[ctx->K + 1];
    *bsk_count = D->S; *gamma = ctx->behz->gamma;
    return B2_OK;
}
