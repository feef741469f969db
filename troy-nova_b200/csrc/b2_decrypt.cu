// b2_decrypt.cu — Decryptor.  Replaces the device branch of
//   Decryptor::dot_product_ct_sk_array[_batched]      src/decryptor.cu:27-231      <ct, (1, s, s^2, ...)>
//   RNSTool::decrypt_scale_and_round[_batched]        src/utils/rns_tool.cu:1189-1430 (constants :168-214)      BFV
//   RNSTool::decrypt_mod_t = BaseConverter::exact_convey_array   src/utils/rns_base.cu:445-583
//     + scaling_variant::decentralize                 src/utils/scaling_variant.cu:415-431                      BGV
//   KeyGenerator::compute_secret_key_powers           src/key_generator.cu:84-110
//   Decryptor::bfv_decrypt / ckks_decrypt / bgv_decrypt (+ _batched)   src/decryptor.cu:333-535
//
// Launches per call, for ANY batch size (the reference: 6-8 per ciphertext, or the same number of *_bp launches per batch):
//   CKKS  1   the phase, element-wise (NTT form in, NTT form out)
//   BGV   2   ONE inverse transform whose loader computes the phase on the fly, then the reduction modulo t
//   BFV   3   forward transform of c_1.. with the product by s^k fused into its store, inverse transform with the sum over k
//             fused into its loader and "+ c_0" into its store, then the t/q scaling (all constants in registers / L1; the
//             reference's kernel round-trips two global scratch arrays per coefficient)
// Every stored value is the canonical residue the reference stores; the BGV path's double-precision quotient estimate uses the
// same operations in the same order (IEEE division, sum in limb order, round half away from zero).
#include <map>
#include "b2_ntt_launch.cuh"
#include "b2_host_math.h"

using namespace b2;

namespace b2 {

// ---- constants of one level (nl ciphertext limbs) ----
struct DecDev {
    u32 nl;
    const Tw *ptg;       // [nl]  t * gamma mod q_i
    const Tw *punc;      // [nl]  (q/q_i)^-1 mod q_i
    const u64 *to_t;     // [nl]  (q/q_i) mod t
    const u64 *to_g;     // [nl]  (q/q_i) mod gamma
    u64 t, t_ratio, t_r0, t_r1, gamma, g_r0, g_r1, q_mod_t;
    Tw neg_iq_t, neg_iq_g, inv_gamma_t;
};
}  // namespace b2

struct b2_decrypt_state {
    struct Level { DecDev dev; void *blob; };
    std::map<size_t, Level> levels;
};

namespace b2 {

void decrypt_release(b2_ctx *ctx) {
    if (!ctx->dec) return;
    for (auto &kv : ctx->dec->levels) cudaFree(kv.second.blob);
    delete ctx->dec; ctx->dec = nullptr;
}

static uint64_t prod_mod(const std::vector<uint64_t> &q, size_t count, size_t skip, uint64_t m) {
    host::u128 r = 1 % m;
    for (size_t i = 0; i < count; i++) if (i != skip) r = (r * (q[i] % m)) % m;
    return (uint64_t)r;
}
static Tw pair_of(uint64_t w, uint64_t p) { Tw t; t.x = w; t.y = host::shoup_quotient(w, p); return t; }

static b2_status ensure_decrypt_level(b2_ctx *ctx, size_t nl, const DecDev **out) {
    std::lock_guard<std::mutex> guard(ctx->mu);
    if (!ctx->dec) ctx->dec = new b2_decrypt_state();
    auto it = ctx->dec->levels.find(nl);
    if (it != ctx->dec->levels.end()) { *out = &it->second.dev; return B2_OK; }
    const uint64_t t = ctx->t;
    std::vector<uint64_t> pr;
    // rns_tool.cu:68-76: m_sk, gamma = the first two 61-bit primes = 1 mod 2N, descending
    if (!host::primes_below(2 * (uint64_t)ctx->n, 61, 2, pr)) { set_error("b2_decrypt: no 61-bit primes = 1 mod 2N"); return B2_ERR_INVALID_ARG; }
    const uint64_t gamma = pr[1];
    std::vector<Tw> tws(2 * nl);
    std::vector<uint64_t> w(2 * nl);
    for (size_t i = 0; i < nl; i++) {
        const uint64_t qi = ctx->q[i];
        tws[i] = pair_of(host::mulmod(t % qi, gamma % qi, qi), qi);
        tws[nl + i] = pair_of(host::invmod(prod_mod(ctx->q, nl, i, qi), qi), qi);
        w[i] = prod_mod(ctx->q, nl, i, t);
        w[nl + i] = prod_mod(ctx->q, nl, i, gamma);
    }
    DecDev D;
    D.nl = (u32)nl; D.t = t; D.t_ratio = ctx->t_ratio; D.t_r0 = ctx->t_r0; D.t_r1 = ctx->t_r1; D.gamma = gamma;
    const host::u128 gr = (~(host::u128)0) / gamma;
    D.g_r0 = (uint64_t)gr; D.g_r1 = (uint64_t)(gr >> 64);
    D.q_mod_t = prod_mod(ctx->q, nl, (size_t)-1, t);
    const uint64_t iq_t = host::invmod(D.q_mod_t, t), iq_g = host::invmod(prod_mod(ctx->q, nl, (size_t)-1, gamma), gamma), ig_t = host::invmod(gamma % t, t);
    // BGV needs none of the inverses (a plain modulus that shares a factor with q or gamma only rules out the BFV scaling)
    D.neg_iq_t = pair_of(iq_t ? t - iq_t : 0, t); D.neg_iq_g = pair_of(iq_g ? gamma - iq_g : 0, gamma); D.inv_gamma_t = pair_of(ig_t, t);
    if (!iq_t || !ig_t || !iq_g) D.inv_gamma_t.x = 0, D.inv_gamma_t.y = 0, D.neg_iq_t.x = 0;      // marks "BFV scaling unavailable"
    void *blob = nullptr;
    const size_t bytes = 2 * nl * sizeof(Tw) + 2 * nl * 8;
    B2_CUDA(cudaMalloc(&blob, bytes));
    cudaError_t e = cudaMemcpy(blob, tws.data(), 2 * nl * sizeof(Tw), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy((char *)blob + 2 * nl * sizeof(Tw), w.data(), 2 * nl * 8, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(blob); return cuda_fail(e, "cudaMemcpy(decrypt constants)"); }
    D.ptg = (const Tw *)blob; D.punc = D.ptg + nl; D.to_t = (const u64 *)(D.punc + nl); D.to_g = D.to_t + nl;
    auto &L = ctx->dec->levels[nl];
    L.dev = D; L.blob = blob;
    *out = &L.dev;
    return B2_OK;
}

// ---- kernels ----
// s^k = s^(k-1) o s, one power per launch (the powers are sequentially dependent)
__global__ void __launch_bounds__(256) sk_power_kernel(const u64 *__restrict__ prev, const u64 *__restrict__ s, u64 *__restrict__ next,
                                                       const ModInfo *__restrict__ mods, u32 n, u64 total) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (u64)gridDim.x * blockDim.x) {
        const ModInfo &Q = mods[i / n];
        next[i] = mulmod(prev[i], s[i], Q.p, Q.r0, Q.r1);
    }
}

// NTT-form phase: out[b][j][x] = c_0 + sum_k c_k o s^k.  One thread per two coefficients.
__global__ void __launch_bounds__(256) phase_ntt_kernel(const u64 *const *__restrict__ ct, u64 *const *__restrict__ out, const u64 *__restrict__ skp,
                                                        const ModInfo *__restrict__ mods, u32 pc, u32 nl, u32 K, u32 n2, u64 total2) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < total2; i += (u64)gridDim.x * blockDim.x) {
        const u32 x = (u32)(i % n2);
        const u64 r = i / n2;
        const u32 j = (u32)(r % nl), b = (u32)(r / nl);
        const ModInfo &Q = mods[j];
        const u64 p = Q.p, r0 = Q.r0, r1 = Q.r1;
        const ulonglong2 *c = reinterpret_cast<const ulonglong2 *>(ct[b]);
        ulonglong2 acc = c[(u64)j * n2 + x];
        for (u32 k = 1; k < pc; k++) {
            const ulonglong2 v = c[((u64)k * nl + j) * n2 + x];
            const ulonglong2 s = reinterpret_cast<const ulonglong2 *>(skp)[((u64)(k - 1) * K + j) * n2 + x];
            acc.x = addmod(acc.x, mulmod(v.x, s.x, p, r0, r1), p);
            acc.y = addmod(acc.y, mulmod(v.y, s.y, p, r0, r1), p);
        }
        reinterpret_cast<ulonglong2 *>(out[b])[(u64)j * n2 + x] = acc;
    }
}

// ---- functors of the fused transforms ----
// forward transform of c_1 .. c_{pc-1}: limb c of group g is polynomial 1 + c % (pc-1) of ciphertext c / (pc-1)
struct CtTailLoad {
    const u64 *const *ct; u32 pm1, nl, n;
    __device__ __forceinline__ const u64 *src(int g, int c, u64) const {
        const u32 b = (u32)c / pm1, k = (u32)c - b * pm1;
        return ct[b] + ((u64)(k + 1) * nl + (u32)g) * n;
    }
    B2_IDENTITY_XF
};
// ... whose store multiplies by s^(k+1) (NTT form): T[c][g][x] = NTT(c_{k+1}) o s^{k+1}
struct SkMulStore {
    u64 *T; const u64 *skp; const ModInfo *mods; u32 pm1, K, n;
    struct Prep { u64 *o; const u64 *s; u64 p, r0, r1; };
    __device__ __forceinline__ Prep prep(int g, int c, u64 off) const {
        const ModInfo &Q = mods[g];
        return Prep{T + off, skp + ((u64)((u32)c % pm1) * K + (u32)g) * n, Q.p, Q.r0, Q.r1};
    }
    __device__ __forceinline__ void operator()(const Prep &pc, u32 x, u64 v) const { pc.o[x] = mulmod(v, pc.s[x], pc.p, pc.r0, pc.r1); }
    __device__ __forceinline__ void st2(const Prep &pc, u32 x, u64 v0, u64 v1) const {
        const ulonglong2 s = *reinterpret_cast<const ulonglong2 *>(pc.s + x);
        *reinterpret_cast<ulonglong2 *>(pc.o + x) = make_ulonglong2(mulmod(v0, s.x, pc.p, pc.r0, pc.r1), mulmod(v1, s.y, pc.p, pc.r0, pc.r1));
    }
};
// inverse transform of sum_k T[b][k]: the staged tile is k = 0, the loader adds the others
struct SumLoad {
    const u64 *T; const ModInfo *mods; u32 pm1, nl, n;
    struct Prep { const u64 *t; u64 p; };
    __device__ __forceinline__ const u64 *src(int g, int c, u64) const { return T + ((u64)c * pm1 * nl + (u32)g) * n; }
    __device__ __forceinline__ Prep prep(int g, int c) const { return Prep{T + ((u64)c * pm1 * nl + (u32)g) * n, mods[g].p}; }
    __device__ __forceinline__ u64 xf(const Prep &pc, u32 x, u64 raw) const {
        for (u32 k = 1; k < pm1; k++) raw = addmod(raw, pc.t[(u64)k * nl * n + x], pc.p);
        return raw;
    }
};
// ... whose store adds c_0 (coefficient form) and writes the phase of ciphertext c
struct AddC0Store {
    const u64 *const *ct; u64 *const *out; const ModInfo *mods; u32 n;
    struct Prep { u64 *o; const u64 *c0; u64 p; };
    __device__ __forceinline__ Prep prep(int g, int c, u64) const { return Prep{out[c] + (u64)(u32)g * n, ct[c] + (u64)(u32)g * n, mods[g].p}; }
    __device__ __forceinline__ void operator()(const Prep &pc, u32 x, u64 v) const { pc.o[x] = addmod(v, pc.c0[x], pc.p); }
    __device__ __forceinline__ void st2(const Prep &pc, u32 x, u64 v0, u64 v1) const {
        const ulonglong2 c = *reinterpret_cast<const ulonglong2 *>(pc.c0 + x);
        *reinterpret_cast<ulonglong2 *>(pc.o + x) = make_ulonglong2(addmod(v0, c.x, pc.p), addmod(v1, c.y, pc.p));
    }
};
// BGV: inverse transform of the NTT-form phase, computed by the loader (staged tile = c_0)
struct PhaseLoad {
    const u64 *const *ct; const u64 *skp; const ModInfo *mods; u32 pc, nl, K, n;
    struct Prep { const u64 *c; const u64 *s; u64 p, r0, r1; };
    __device__ __forceinline__ const u64 *src(int g, int c, u64) const { return ct[c] + (u64)(u32)g * n; }
    __device__ __forceinline__ Prep prep(int g, int c) const {
        const ModInfo &Q = mods[g];
        return Prep{ct[c] + (u64)(u32)g * n, skp + (u64)(u32)g * n, Q.p, Q.r0, Q.r1};
    }
    __device__ __forceinline__ u64 xf(const Prep &q, u32 x, u64 raw) const {
        for (u32 k = 1; k < pc; k++) raw = addmod(raw, mulmod(q.c[(u64)k * nl * n + x], q.s[(u64)(k - 1) * K * n + x], q.p, q.r0, q.r1), q.p);
        return raw;
    }
};

// BFV: round(t/q * phase) mod t for one coefficient per thread (rns_tool.cu:1189-1268, all scratch in registers)
__global__ void __launch_bounds__(256) scale_round_kernel(const u64 *const *__restrict__ phase, u64 *const *__restrict__ out, const ModInfo *__restrict__ mods,
                                                          const DecDev D, u32 n, u64 total) {
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const u32 b = (u32)(idx / n), x = (u32)(idx % n);
    const u64 *ph = phase[b];
    u64 tl = 0, th = 0, gl = 0, gh = 0;
    for (u32 i = 0; i < D.nl; i++) {
        const u64 q = mods[i].p;
        const u64 v = shoup(shoup(ph[(u64)i * n + x], D.ptg[i], q), D.punc[i], q);
        mac128(tl, th, v, D.to_t[i]);
        mac128(gl, gh, v, D.to_g[i]);
    }
    const u64 vt = shoup(barrett128(tl, th, D.t, D.t_r0, D.t_r1), D.neg_iq_t, D.t);
    const u64 vg = shoup(barrett128(gl, gh, D.gamma, D.g_r0, D.g_r1), D.neg_iq_g, D.gamma);
    const u64 d = vg > (D.gamma >> 1) ? addmod(vt, barrett64(D.gamma - vg, D.t, D.t_ratio), D.t) : submod(vt, barrett64(vg, D.t, D.t_ratio), D.t);
    out[b][x] = d ? shoup(d, D.inv_gamma_t, D.t) : 0;
}

// BGV: exact_convey_array to the base {t}, then the correction factor (rns_base.cu:445-583, scaling_variant.cu:424-430)
__global__ void __launch_bounds__(256) mod_t_kernel(const u64 *const *__restrict__ phase, u64 *const *__restrict__ out, const ModInfo *__restrict__ mods,
                                                    const DecDev D, const u64 *__restrict__ fix, u32 n, u64 total) {
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const u32 b = (u32)(idx / n), x = (u32)(idx % n);
    const u64 *ph = phase[b];
    u64 lo = 0, hi = 0;
    double v = 0.0;
    for (u32 i = 0; i < D.nl; i++) {
        const u64 q = mods[i].p;
        const u64 y = shoup(ph[(u64)i * n + x], D.punc[i], q);
        v = __dadd_rn(v, __ddiv_rn(__ull2double_rn(y), __ull2double_rn(q)));
        mac128(lo, hi, y, D.to_t[i]);
    }
    const u64 rounded = (u64)round(v);
    const u64 corr = mulmod(barrett64(rounded, D.t, D.t_ratio), D.q_mod_t, D.t, D.t_r0, D.t_r1);
    u64 r = submod(barrett128(lo, hi, D.t, D.t_r0, D.t_r1), corr, D.t);
    const u64 f = fix[b];
    out[b][x] = f != 1 ? mulmod(r, f, D.t, D.t_r0, D.t_r1) : r;
}

static inline unsigned grid_for(u64 total) { return (unsigned)std::min<u64>((total + 255) / 256, 148u * 32u); }

// uploads host pointer arrays (and trailing words) into the scratch record's table; returns the device base
static b2_status upload_words(b2_scratch *sc, cudaStream_t stream, const std::vector<uint64_t> &blob, const u64 **base) {
    b2_status st = ensure_ptrtab(sc, stream, blob.size() * 8);
    if (st != B2_OK) return st;
    B2_CUDA(cudaMemcpyAsync(sc->ptrtab, blob.data(), blob.size() * 8, cudaMemcpyHostToDevice, stream));   // pageable source: staged before return
    *base = (const u64 *)sc->ptrtab;
    return B2_OK;
}

// phase of `batch` ciphertexts into out[b] ([nl][N] each), with the scratch record already held.  d_ct / d_out: device tables.
static b2_status phase_locked(b2_ctx *ctx, b2_scratch *sc, const u64 *const *d_ct, u64 *const *d_out, size_t batch, size_t pc, size_t nl, bool is_ntt,
                              const u64 *skp, size_t ws_offset_bytes, cudaStream_t stream) {
    const size_t n = ctx->n, K = ctx->K;
    if (is_ntt) {
        const u64 total2 = (u64)batch * nl * (n / 2);
        phase_ntt_kernel<<<grid_for(total2), 256, 0, stream>>>(d_ct, d_out, skp, ctx->d_mods, (u32)pc, (u32)nl, (u32)K, (u32)(n / 2), total2);
        count_launches(1);
        B2_CUDA(cudaGetLastError());
        return B2_OK;
    }
    // workspace (behind ws_offset_bytes, which the caller owns): T [batch][pc-1][nl][N] | mid [batch][nl][N]
    const size_t pm1 = pc - 1;
    b2_status st = ensure_ws(sc, stream, ws_offset_bytes + (batch * pm1 * nl * n + batch * nl * n) * 8);
    if (st != B2_OK) return st;
    u64 *T = (u64 *)((char *)sc->ws + ws_offset_bytes), *mid = T + batch * pm1 * nl * n;
    std::vector<LimbGroup> groups;
    for (size_t j = 0; j < nl; j++) groups.push_back(LimbGroup{(u32)j, (u32)(batch * pm1), (u64)(j * n), (u64)(nl * n)});
    st = run_ntt_dir<true>(ctx, sc, groups, T, CtTailLoad{d_ct, (u32)pm1, (u32)nl, (u32)n}, SkMulStore{T, skp, ctx->d_mods, (u32)pm1, (u32)K, (u32)n}, stream);
    if (st != B2_OK) return st;
    groups.clear();
    for (size_t j = 0; j < nl; j++) groups.push_back(LimbGroup{(u32)j, (u32)batch, (u64)(j * n), (u64)(nl * n)});
    return run_ntt_dir<false>(ctx, sc, groups, mid, SumLoad{T, ctx->d_mods, (u32)pm1, (u32)nl, (u32)n}, AddC0Store{d_ct, d_out, ctx->d_mods, (u32)n}, stream);
}

static b2_status check_batch(const uint64_t *const *a, uint64_t *const *b, size_t batch) {
    for (size_t i = 0; i < batch; i++) {
        if (!a[i] || !b[i]) { set_error("null operand in batch"); return B2_ERR_INVALID_ARG; }
        if ((((uintptr_t)a[i] | (uintptr_t)b[i]) & 15) != 0) { set_error("operands must be 16-byte aligned"); return B2_ERR_INVALID_ARG; }
    }
    return B2_OK;
}

}  // namespace b2

extern "C" {

b2_status b2_secret_key_powers(b2_ctx *ctx, const uint64_t *sk, size_t max_power, uint64_t *out, void *stream_) {
    B2_REQUIRE(ctx && sk && out && max_power >= 1, "bad arguments");
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t w = ctx->K * ctx->n;
    if (out != sk) B2_CUDA(cudaMemcpyAsync(out, sk, w * 8, cudaMemcpyDeviceToDevice, stream));
    for (size_t k = 1; k < max_power; k++) {
        sk_power_kernel<<<grid_for(w), 256, 0, stream>>>((const u64 *)out + (k - 1) * w, (const u64 *)out, (u64 *)out + k * w, ctx->d_mods, (u32)ctx->n, (u64)w);
        count_launches(1);
        B2_CUDA(cudaGetLastError());
    }
    return B2_OK;
}

b2_status b2_decrypt_phase(b2_ctx *ctx, const uint64_t *const *ct, uint64_t *const *phase, size_t batch, size_t pcount, size_t nl, int is_ntt_form,
                           const uint64_t *sk_powers, void *stream_) {
    B2_REQUIRE(ctx && (batch == 0 || (ct && phase)) && sk_powers, "null argument");
    B2_REQUIRE(pcount >= 2 && nl >= 1 && nl <= ctx->K && ctx->n >= 2, "bad shape (a ciphertext has at least two polynomials)");
    if (batch == 0) return B2_OK;
    b2_status st = check_batch(ct, phase, batch);
    if (st != B2_OK) return st;
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    Scratch sc(ctx, stream);
    std::vector<uint64_t> blob(2 * batch);
    for (size_t b = 0; b < batch; b++) { blob[b] = (uint64_t)(uintptr_t)ct[b]; blob[batch + b] = (uint64_t)(uintptr_t)phase[b]; }
    const u64 *base;
    st = upload_words(sc.s, stream, blob, &base);
    if (st != B2_OK) return st;
    return phase_locked(ctx, sc.s, (const u64 *const *)base, (u64 *const *)(base + batch), batch, pcount, nl, is_ntt_form != 0, (const u64 *)sk_powers, 0, stream);
}

b2_status b2_decrypt_scale_and_round(b2_ctx *ctx, const uint64_t *const *phase, uint64_t *const *plain, size_t batch, size_t nl, void *stream_) {
    B2_REQUIRE(ctx && (batch == 0 || (phase && plain)), "null argument");
    B2_REQUIRE(ctx->t > 1 && nl >= 1 && nl <= ctx->K, "BFV scaling needs a plain modulus and 1 <= nl <= K");
    if (batch == 0) return B2_OK;
    b2_status st = check_batch(phase, plain, batch);
    if (st != B2_OK) return st;
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const DecDev *D;
    st = ensure_decrypt_level(ctx, nl, &D);
    if (st != B2_OK) return st;
    B2_REQUIRE(D->inv_gamma_t.x != 0, "plain modulus shares a factor with q or gamma (the reference's RNSTool rejects it too)");
    Scratch sc(ctx, stream);
    std::vector<uint64_t> blob(2 * batch);
    for (size_t b = 0; b < batch; b++) { blob[b] = (uint64_t)(uintptr_t)phase[b]; blob[batch + b] = (uint64_t)(uintptr_t)plain[b]; }
    const u64 *base;
    st = upload_words(sc.s, stream, blob, &base);
    if (st != B2_OK) return st;
    const u64 total = (u64)batch * ctx->n;
    scale_round_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>((const u64 *const *)base, (u64 *const *)(base + batch), ctx->d_mods, *D, (u32)ctx->n, total);
    count_launches(1);
    B2_CUDA(cudaGetLastError());
    return B2_OK;
}

b2_status b2_decrypt_mod_t(b2_ctx *ctx, const uint64_t *const *phase, uint64_t *const *plain, size_t batch, size_t nl, const uint64_t *correction_factors,
                           void *stream_) {
    B2_REQUIRE(ctx && (batch == 0 || (phase && plain)), "null argument");
    B2_REQUIRE(ctx->t > 1 && nl >= 1 && nl <= ctx->K, "BGV decryption needs a plain modulus and 1 <= nl <= K");
    if (batch == 0) return B2_OK;
    b2_status st = check_batch(phase, plain, batch);
    if (st != B2_OK) return st;
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const DecDev *D;
    st = ensure_decrypt_level(ctx, nl, &D);
    if (st != B2_OK) return st;
    std::vector<uint64_t> blob(3 * batch);
    for (size_t b = 0; b < batch; b++) {
        blob[b] = (uint64_t)(uintptr_t)phase[b]; blob[batch + b] = (uint64_t)(uintptr_t)plain[b];
        uint64_t fix = 1;
        if (correction_factors && correction_factors[b] != 1) {
            fix = host::invmod(correction_factors[b] % ctx->t, ctx->t);
            B2_REQUIRE(fix != 0, "correction factor is not invertible modulo t");     // scaling_variant.cu:426-428
        }
        blob[2 * batch + b] = fix;
    }
    Scratch sc(ctx, stream);
    const u64 *base;
    st = upload_words(sc.s, stream, blob, &base);
    if (st != B2_OK) return st;
    const u64 total = (u64)batch * ctx->n;
    mod_t_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>((const u64 *const *)base, (u64 *const *)(base + batch), ctx->d_mods, *D, base + 2 * batch, (u32)ctx->n, total);
    count_launches(1);
    B2_CUDA(cudaGetLastError());
    return B2_OK;
}

// Decryptor::decrypt[_batched] for the scheme of the context.
//   BFV  (coefficient-form ciphertexts)  plain[b] = [N] words modulo t
//   BGV  (NTT-form ciphertexts)          plain[b] = [N] words modulo t (correction_factors[b], NULL = all 1)
//   CKKS (NTT-form ciphertexts)          plain[b] = [nl][N], NTT form
b2_status b2_decrypt(b2_ctx *ctx, const uint64_t *const *ct, uint64_t *const *plain, size_t batch, size_t pcount, size_t nl, const uint64_t *sk_powers,
                     const uint64_t *correction_factors, void *stream_) {
    B2_REQUIRE(ctx && (batch == 0 || (ct && plain)) && sk_powers, "null argument");
    B2_REQUIRE(pcount >= 2 && nl >= 1 && nl <= ctx->K && ctx->n >= 2, "bad shape (a ciphertext has at least two polynomials)");
    if (batch == 0) return B2_OK;
    if (ctx->scheme == B2_SCHEME_CKKS) return b2_decrypt_phase(ctx, ct, plain, batch, pcount, nl, 1, sk_powers, stream_);
    B2_REQUIRE(ctx->t > 1, "BFV / BGV decryption needs a plain modulus");
    b2_status st = check_batch(ct, plain, batch);
    if (st != B2_OK) return st;
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t n = ctx->n, K = ctx->K;
    const DecDev *D;
    st = ensure_decrypt_level(ctx, nl, &D);
    if (st != B2_OK) return st;
    const bool bfv = ctx->scheme == B2_SCHEME_BFV;
    if (bfv) B2_REQUIRE(D->inv_gamma_t.x != 0, "plain modulus shares a factor with q or gamma (the reference's RNSTool rejects it too)");
    Scratch sc(ctx, stream);
    // workspace: PH [batch][nl][N] (the phase, coefficient form) | whatever phase_locked / the transform needs behind it
    const size_t ph_bytes = batch * nl * n * 8;
    st = ensure_ws(sc.s, stream, ph_bytes + (bfv ? (batch * (pcount - 1) * nl * n + batch * nl * n) * 8 : batch * nl * n * 8));
    if (st != B2_OK) return st;
    u64 *PH = (u64 *)sc->ws;
    // table: ct[batch] | ph[batch] | plain[batch] | fix[batch]
    std::vector<uint64_t> blob(4 * batch);
    for (size_t b = 0; b < batch; b++) {
        blob[b] = (uint64_t)(uintptr_t)ct[b]; blob[batch + b] = (uint64_t)(uintptr_t)(PH + b * nl * n); blob[2 * batch + b] = (uint64_t)(uintptr_t)plain[b];
        uint64_t fix = 1;
        if (!bfv && correction_factors && correction_factors[b] != 1) {
            fix = host::invmod(correction_factors[b] % ctx->t, ctx->t);
            B2_REQUIRE(fix != 0, "correction factor is not invertible modulo t");
        }
        blob[3 * batch + b] = fix;
    }
    const u64 *base;
    st = upload_words(sc.s, stream, blob, &base);
    if (st != B2_OK) return st;
    const u64 *const *d_ct = (const u64 *const *)base;
    u64 *const *d_ph = (u64 *const *)(base + batch), *const *d_plain = (u64 *const *)(base + 2 * batch);
    const u64 total = (u64)batch * n;
    if (bfv) {
        st = phase_locked(ctx, sc.s, d_ct, d_ph, batch, pcount, nl, false, (const u64 *)sk_powers, ph_bytes, stream);
        if (st != B2_OK) return st;
        scale_round_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>((const u64 *const *)d_ph, d_plain, ctx->d_mods, *D, (u32)n, total);
    } else {
        u64 *mid = PH + batch * nl * n;
        std::vector<LimbGroup> groups;
        for (size_t j = 0; j < nl; j++) groups.push_back(LimbGroup{(u32)j, (u32)batch, (u64)(j * n), (u64)(nl * n)});
        st = run_ntt_dir<false>(ctx, sc.s, groups, mid, PhaseLoad{d_ct, (const u64 *)sk_powers, ctx->d_mods, (u32)pcount, (u32)nl, (u32)K, (u32)n}, PlainStore{PH}, stream);
        if (st != B2_OK) return st;
        mod_t_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>((const u64 *const *)d_ph, d_plain, ctx->d_mods, *D, base + 3 * batch, (u32)n, total);
    }
    count_launches(1);
    B2_CUDA(cudaGetLastError());
    return B2_OK;
}

}  // extern "C"
