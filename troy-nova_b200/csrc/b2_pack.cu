// b2_pack.cu — RLWE packing: the device branch of Evaluator::pack_rlwe_ciphertexts_new
// (src/evaluator_lwes.cu:305-491, which pack_lwe_ciphertexts_new :201-229 and the BumbleBee / Cheetah output packing
// src/app/matmul.cu:572-619 call) and Evaluator::field_trace_inplace (:96-105).
//
// A layer of the reference is, for every (even, odd) pair: temp = x^shift * odd (negacyclic_shift_bps); odd = even - temp (sub_batched);
// even += temp (add_inplace_batched); [NTT]; odd = apply_galois(odd) = permute both polynomials + keyswitch c_1 (OverwriteExceptFirst);
// [INTT]; even += odd — six element-wise launches over the batch, two transforms for the NTT-form schemes, and the keyswitch.  Here a
// layer is ONE element-wise launch and ONE batched keyswitch:
//   pack_layer_kernel writes   out = (even + temp) + (galois(even - temp).c_0, 0)   and   target = galois(even - temp).c_1
//   b2_keyswitch_batched(target, AddInplace, out)                                     adds the key-switched pair
// The tree runs in coefficient form (for CKKS too: the automorphism commutes with the transform and the keyswitch computes the same ring
// element in either form, so the final transform gives the reference's NTT-form words, which removes the two transforms per layer of the
// reference's CKKS branch; BGV transforms the keyswitch target and result per layer as the reference does, its mod-down being defined on
// NTT-form operands).  Live ciphertexts are kept compact: layer l pairs (cur[2i], cur[2i+1]) -> next[i].
#include <vector>
#include "b2_ctx.h"
#include "b2_host_math.h"

using namespace b2;

namespace b2 {

// one thread per (pair, limb, coefficient y).  in: [M][2][nl][N] (pair b = slots 2b, 2b+1); out: [M/2][2][nl][N]; target: [M/2][nl][N]
__global__ void __launch_bounds__(256) pack_layer_kernel(const u64 *__restrict__ in, u64 *__restrict__ out, u64 *__restrict__ target, const ModInfo *__restrict__ mods,
                                                         u32 nl, u32 logn, u32 shift, u32 ginv, u64 total) {
    const u32 n = 1u << logn;
    for (u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (u64)gridDim.x * blockDim.x) {
        const u32 y = (u32)(v & (n - 1));
        const u64 r = v >> logn;
        const u32 j = (u32)(r % nl);
        const u64 b = r / nl;
        const u64 p = mods[j].p;
        const u64 ct = (u64)2 * nl * n;                                  // words per ciphertext
        const u64 *even = in + 2 * b * ct, *odd = even + ct;
        // galois gather: destination y takes source i with sign       (ew_galois_coeff_kernel)
        const u32 raw = (y * ginv) & (2 * n - 1), i = raw & (n - 1);
        const bool gneg = raw >= n;
        // x^shift * odd at positions y and i                           (ew_negacyclic_shift_kernel)
        const u32 sy = (y - shift) & (n - 1), si = (i - shift) & (n - 1);
        const bool ny = ((sy + shift) & n) != 0, ni = ((si + shift) & n) != 0;
        u64 res[2], tgt = 0;
#pragma unroll
        for (u32 k = 0; k < 2; k++) {
            const u64 off = ((u64)k * nl + j) * n;
            u64 ty = odd[off + sy]; ty = (ny && ty) ? p - ty : ty;
            u64 ti = odd[off + si]; ti = (ni && ti) ? p - ti : ti;
            const u64 sum = addmod(even[off + y], ty, p);
            u64 d = submod(even[off + i], ti, p);
            d = (gneg && d) ? p - d : d;
            if (k == 0) res[0] = addmod(sum, d, p);
            else { res[1] = sum; tgt = d; }
        }
        u64 *o = out + b * ct;
        o[(u64)j * n + y] = res[0];
        o[((u64)nl + j) * n + y] = res[1];
        target[(b * nl + j) * n + y] = tgt;
    }
}
// slot i of the arena = scale_j * x^shift * cipher[src[i]] (coefficient form), or zero when src[i] < 0
__global__ void __launch_bounds__(256) pack_init_kernel(const u64 *const *__restrict__ ciphers, const int *__restrict__ src, u64 *__restrict__ out,
                                                        const ModInfo *__restrict__ mods, const Tw *__restrict__ scale, u32 nl, u32 logn, u32 shift, u64 total) {
    const u32 n = 1u << logn;
    for (u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (u64)gridDim.x * blockDim.x) {
        const u32 x = (u32)(v & (n - 1));
        const u64 limb = v >> logn;                                      // (slot, k, j)
        const u32 j = (u32)(limb % nl);
        const u64 slot = limb / (2 * nl);
        const int s = src[slot];
        if (s < 0) { out[v] = 0; continue; }
        const u64 p = mods[j].p;
        const u32 sx = (x - shift) & (n - 1);
        const bool neg = ((sx + shift) & n) != 0;
        u64 val = shoup(ciphers[s][(limb % (2 * nl)) * n + sx], scale[j], p);
        out[v] = (neg && val) ? p - val : val;
    }
}
b2_status arena_pool(b2_ctx *ctx, cudaMemPool_t *out) {
    std::lock_guard<std::mutex> guard(ctx->mu);
    if (!ctx->mempool) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = ctx->device;
        cudaMemPool_t pool = nullptr;
        B2_CUDA(cudaMemPoolCreate(&pool, &props));
        unsigned long long keep = ~0ull;
        cudaError_t e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        if (e != cudaSuccess) { cudaMemPoolDestroy(pool); return cuda_fail(e, "cudaMemPoolSetAttribute"); }
        ctx->mempool = pool;
    }
    *out = ctx->mempool;
    return B2_OK;
}
static inline unsigned pk_grid(u64 total) { return (unsigned)std::min<u64>((total + 255) / 256, 148u * 32u); }
static u32 inv_mod_2n(u32 g, size_t n) { u32 inv = g; for (int it = 0; it < 5; it++) inv *= 2 - g * inv; return inv & (u32)(2 * n - 1); }
static const b2_kswitch_key *find_key(const uint64_t *elts, const b2_kswitch_key *const *keys, size_t n_keys, uint64_t g) {
    for (size_t i = 0; i < n_keys; i++) if (elts[i] == g) return keys[i];
    return nullptr;
}

}  // namespace b2

extern "C" {

// one layer over `pairs` pairs, all operands contiguous device arrays in coefficient form: in [2 pairs][2][nl][N], out [pairs][2][nl][N],
// target [pairs][nl][N] (scratch)
b2_status b2_pack_layer(b2_ctx *ctx, const uint64_t *in, uint64_t *out, uint64_t *target, size_t pairs, size_t nl, size_t shift, uint64_t galois_element,
                        const b2_kswitch_key *key, void *stream_) {
    B2_REQUIRE(ctx && in && out && target && key, "null argument");
    B2_REQUIRE(nl >= 1 && nl + 1 <= ctx->K && shift < ctx->n, "bad shape (packing needs a special prime; shift must be below N)");
    B2_REQUIRE((galois_element & 1) && galois_element < 2 * ctx->n, "Galois element must be odd and below 2N");
    B2_REQUIRE(in != out, "out of place only");
    if (pairs == 0) return B2_OK;
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t n = ctx->n;
    const u64 total = (u64)pairs * nl * n;
    pack_layer_kernel<<<pk_grid(total), 256, 0, stream>>>((const u64 *)in, (u64 *)out, (u64 *)target, ctx->d_mods, (u32)nl, ctx->logn, (u32)shift,
                                                          inv_mod_2n((u32)galois_element, n), total);
    count_launches(1);
    B2_CUDA(cudaGetLastError());
    std::vector<const uint64_t *> t(pairs);
    std::vector<uint64_t *> d(pairs);
    for (size_t b = 0; b < pairs; b++) { t[b] = target + b * nl * n; d[b] = out + b * 2 * nl * n; }
    if (ctx->scheme != B2_SCHEME_BGV) return b2_keyswitch_batched(ctx, t.data(), d.data(), pairs, nl, 0, key, B2_ASSIGN_ADD_INPLACE, stream_);
    // BGV: the mod-down of its keyswitch is defined on NTT-form operands (the t-correction), so the key-switched pair is computed as the
    // reference does — NTT(target), keyswitch, INTT — and added: three more launches per layer
    uint64_t *ks = nullptr;
    cudaMemPool_t pool;
    b2_status pst = arena_pool(ctx, &pool);
    if (pst != B2_OK) return pst;
    B2_CUDA(cudaMallocFromPoolAsync((void **)&ks, pairs * 2 * nl * n * 8, pool, stream));
    for (size_t b = 0; b < pairs; b++) d[b] = ks + b * 2 * nl * n;
    b2_status st = b2_ntt_forward(ctx, target, target, pairs, 1, nl, B2_TABLES_COMPONENTWISE, 0, 0, stream_);
    if (st == B2_OK) st = b2_keyswitch_batched(ctx, t.data(), d.data(), pairs, nl, 1, key, B2_ASSIGN_OVERWRITE, stream_);
    if (st == B2_OK) st = b2_ntt_inverse(ctx, ks, ks, pairs, 2, nl, B2_TABLES_COMPONENTWISE, 0, 0, stream_);
    if (st == B2_OK) st = b2_add(ctx, out, ks, out, pairs * 2, nl, 0, stream_);
    cudaFreeAsync(ks, stream);
    return st;
}

}  // extern "C"

namespace b2 {
// the packing tree over coefficient-form inputs (arguments validated by the callers)
static b2_status pack_run(b2_ctx *ctx, const uint64_t *const *ciphers, const size_t *counts, size_t groups, size_t count, size_t nl, size_t shift, size_t input_interval,
                          size_t output_interval, const uint64_t *galois_elements, const b2_kswitch_key *const *keys, size_t n_keys, bool trace, int output_ntt_form,
                          uint64_t *const *outs, void *stream_) {
    const size_t n = ctx->n, M = input_interval / output_interval;
    size_t layers = 0;
    while (((size_t)1 << layers) < M) layers++;
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t ctw = 2 * nl * n, slots = groups * M;
    // arena: A [slots][ctw] | B [max(slots/2, groups)][ctw] | T [max(slots/2, groups)][nl][N] | table: pointers[count] + src[slots]
    const size_t half = std::max<size_t>(slots / 2, groups), tab_bytes = ((count * 8 + slots * 4 + 15) / 16) * 16;
    const size_t words = slots * ctw + half * ctw + half * nl * n;
    void *arena = nullptr;
    cudaMemPool_t pool;
    b2_status pst = arena_pool(ctx, &pool);
    if (pst != B2_OK) return pst;
    B2_CUDA(cudaMallocFromPoolAsync(&arena, words * 8 + tab_bytes, pool, stream));
    u64 *A = (u64 *)arena, *Bf = A + slots * ctw, *T = Bf + half * ctw;
    char *tab = (char *)(T + half * nl * n);
    b2_status st = B2_OK;
    Tw *d_scale = nullptr;
    auto fail = [&](b2_status s) {      // every early exit returns the stream-ordered allocations
        if (d_scale) cudaFreeAsync(d_scale, stream);
        cudaFreeAsync(arena, stream);
        return s;
    };
    // slot (g, i) <- cipher[g][bitrev(i)] (evaluator_lwes.cu:347-358), or zero
    std::vector<uint64_t> blob((tab_bytes + 7) / 8);
    int *src = reinterpret_cast<int *>(blob.data() + count);
    for (size_t i = 0; i < count; i++) blob[i] = (uint64_t)(uintptr_t)ciphers[i];
    for (size_t g = 0, first = 0; g < groups; first += counts[g], g++)
        for (size_t i = 0; i < M; i++) {
            const size_t idx = layers ? host::bitrev((uint32_t)i, (unsigned)layers) : 0;
            src[g * M + i] = idx < counts[g] ? (int)(first + idx) : -1;
        }
    if (cudaMemcpyAsync(tab, blob.data(), tab_bytes, cudaMemcpyHostToDevice, stream) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMemcpyAsync(pack table)"));
    const u64 *const *d_ciphers = (const u64 *const *)tab;
    const int *d_src = (const int *)(tab + count * 8);
    std::vector<Tw> scale(nl);
    const uint64_t mul = n / input_interval;
    for (size_t j = 0; j < nl; j++) {            // N^-1 * mul mod q_j          (divide_by_poly_modulus_degree_inplace, :132-141)
        const uint64_t q = ctx->q[j], sc = host::mulmod(host::invmod(n % q, q), mul % q, q);
        scale[j].x = sc; scale[j].y = host::shoup_quotient(sc, q);
    }
    if (cudaMallocFromPoolAsync((void **)&d_scale, nl * sizeof(Tw), pool, stream) != cudaSuccess) { d_scale = nullptr; return fail(cuda_fail(cudaGetLastError(), "cudaMallocFromPoolAsync(pack scale)")); }
    if (cudaMemcpyAsync(d_scale, scale.data(), nl * sizeof(Tw), cudaMemcpyHostToDevice, stream) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMemcpyAsync(pack scale)"));
    const u64 *const *init_src = d_ciphers;
    if (st == B2_OK) {
        const u64 total = (u64)slots * ctw;
        pack_init_kernel<<<pk_grid(total), 256, 0, stream>>>(init_src, d_src, A, ctx->d_mods, d_scale, (u32)nl, ctx->logn, (u32)shift, total);
        count_launches(1);
        if (cudaGetLastError() != cudaSuccess) st = B2_ERR_CUDA;
    }
    u64 *cur = A, *nxt = Bf;
    size_t live = slots;         // pairs never straddle groups: M is a power of two >= 2 whenever there is a layer
    for (size_t l = 0; l < layers && st == B2_OK; l++) {
        const size_t sh = input_interval >> (l + 1);
        const uint64_t g = (n / input_interval) * ((uint64_t)1 << (l + 1)) + 1;
        st = b2_pack_layer(ctx, (const uint64_t *)cur, (uint64_t *)nxt, (uint64_t *)T, live / 2, nl, sh, g, find_key(galois_elements, keys, n_keys, g), stream_);
        std::swap(cur, nxt);
        live /= 2;
    }
    // results: cur[g], g < groups; transform / trace them together, in the form of the result (field trace: evaluator_lwes.cu:96-130)
    if (st == B2_OK && output_ntt_form) st = b2_ntt_forward(ctx, (const uint64_t *)cur, (uint64_t *)cur, groups, 2, nl, B2_TABLES_COMPONENTWISE, 0, 0, stream_);
    if (st == B2_OK && trace) {
        uint64_t *res = (uint64_t *)cur, *tmp = (uint64_t *)nxt;      // [groups][2][nl][N] fits in either buffer
        std::vector<const uint64_t *> t(groups);
        std::vector<uint64_t *> d(groups);
        for (size_t g = 0; g < groups; g++) { t[g] = tmp + g * ctw + nl * n; d[g] = tmp + g * ctw; }
        for (size_t deg = n; deg > n / output_interval && st == B2_OK; deg >>= 1) {
            st = b2_galois(ctx, res, deg + 1, tmp, 2 * groups, nl, 0, output_ntt_form, stream_);
            if (st == B2_OK) st = b2_keyswitch_batched(ctx, t.data(), d.data(), groups, nl, output_ntt_form, find_key(galois_elements, keys, n_keys, deg + 1), B2_ASSIGN_OVERWRITE_EXCEPT_FIRST, stream_);
            if (st == B2_OK) st = b2_add(ctx, res, tmp, res, 2 * groups, nl, 0, stream_);
        }
    }
    for (size_t g = 0; g < groups && st == B2_OK; g++)
        if (cudaMemcpyAsync(outs[g], cur + g * ctw, ctw * 8, cudaMemcpyDeviceToDevice, stream) != cudaSuccess) st = cuda_fail(cudaGetLastError(), "cudaMemcpyAsync(pack result)");
    return fail(st);      // releases the arena and the scale table (stream-ordered: after the work enqueued above)
}

// argument checks shared by the two packing entry points (evaluator_lwes.cu:324-339 and the key look-ups of GaloisKeys::key)
static b2_status pack_check(b2_ctx *ctx, const size_t *counts, size_t groups, size_t nl, size_t shift, size_t input_interval, size_t output_interval,
                            const uint64_t *galois_elements, const b2_kswitch_key *const *keys, size_t n_keys, bool trace, uint64_t *const *outs, size_t *count_out) {
    const size_t n = ctx->n;
    B2_REQUIRE(nl >= 1 && nl + 1 <= ctx->K, "bad level (packing needs a special prime)");
    B2_REQUIRE(input_interval >= 1 && input_interval <= n, "input_interval must be less than poly_modulus_degree.");
    B2_REQUIRE(output_interval <= input_interval && output_interval >= 1, "output_interval must be less than input_interval.");
    B2_REQUIRE((input_interval & (input_interval - 1)) == 0, "input_interval must be power of two.");
    B2_REQUIRE((output_interval & (output_interval - 1)) == 0, "output_interval must be power of two.");
    const size_t M = input_interval / output_interval;
    size_t count = 0;
    for (size_t g = 0; g < groups; g++) {
        B2_REQUIRE(counts[g] >= 1 && counts[g] <= M, "ciphers count must be less than input_interval / output_interval.");
        B2_REQUIRE(outs[g] && (((uintptr_t)outs[g]) & 15) == 0, "null or misaligned output");
        count += counts[g];
    }
    B2_REQUIRE(shift < 2 * n, "shift must be below 2N");
    B2_REQUIRE(ctx->scheme != B2_SCHEME_BGV || ctx->t > 1, "BGV needs a plain modulus");
    size_t layers = 0;
    while (((size_t)1 << layers) < M) layers++;
    // every key the tree and the trace will use must be present before anything is enqueued
    for (size_t l = 0; l < layers; l++)
        B2_REQUIRE(find_key(galois_elements, keys, n_keys, (n / input_interval) * ((uint64_t)1 << (l + 1)) + 1), "Galois key not present");
    if (trace)
        for (size_t deg = n; deg > n / output_interval; deg >>= 1) B2_REQUIRE(find_key(galois_elements, keys, n_keys, deg + 1), "Galois key not present");
    *count_out = count;
    return B2_OK;
}

// LWE (c0 [nl] scalars, c1 [nl][N]) -> RLWE [2][nl][N]: c_0 = the constant polynomial c0, c_1 = c1   (src/lwe_ciphertext.cu:12-22)
__global__ void __launch_bounds__(256) lwe_assemble_kernel(const u64 *const *__restrict__ c0s, const u64 *const *__restrict__ c1s, u64 *const *__restrict__ outs,
                                                           u32 nl, u32 n, u64 total) {
    for (u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (u64)gridDim.x * blockDim.x) {
        const u64 per = (u64)nl * n, b = v / per, r = v % per;
        const u32 x = (u32)(r % n), j = (u32)(r / n);
        u64 *o = outs[b];
        o[r] = x == 0 ? c0s[b][j] : 0;
        o[per + r] = c1s[b][r];
    }
}
// RLWE (coefficient form) -> LWE of coefficient `term`: c1 = x^(2N - term) c_1 (term > 0), c0[j] = c_0[j][term]   (src/evaluator_lwes.cu:51-94)
__global__ void __launch_bounds__(256) lwe_extract_kernel(const u64 *__restrict__ ct, u64 *__restrict__ c0, u64 *__restrict__ c1, const ModInfo *__restrict__ mods,
                                                          u32 nl, u32 n, u32 term, u32 shift, u64 total) {
    for (u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (u64)gridDim.x * blockDim.x) {
        const u32 x = (u32)(v % n), j = (u32)(v / n);
        const u64 p = mods[j].p;
        const u32 src = (x - shift) & (n - 1);
        const bool neg = ((src + shift) & n) != 0;
        const u64 val = ct[((u64)nl + j) * n + src];
        c1[v] = (neg && val) ? p - val : val;
        if (x == 0) c0[j] = ct[(u64)j * n + term];
    }
}
}  // namespace b2

extern "C" {

// Evaluator::pack_rlwe_ciphertexts_new / pack_rlwe_ciphertexts_batched for `groups` groups packed side by side (every layer is one fused
// launch + one batched keyswitch over ALL groups).  ciphers: HOST array of sum(counts) DEVICE pointers (group after group), each
// [2][nl][N] in the form given by input_ntt_form; counts: HOST [groups]; galois_elements / keys: the automorphism keys (GaloisKeys) as
// parallel HOST arrays; outs: HOST array of `groups` DEVICE pointers, each [2][nl][N], NTT form iff output_ntt_form (the reference: CKKS /
// BGV).  Everything is enqueued on `stream`; the arena is stream-ordered memory.
b2_status b2_pack_rlwes(b2_ctx *ctx, const uint64_t *const *ciphers, const size_t *counts, size_t groups, int input_ntt_form, size_t nl, size_t shift,
                        size_t input_interval, size_t output_interval, const uint64_t *galois_elements, const b2_kswitch_key *const *keys, size_t n_keys,
                        int apply_field_trace, int output_ntt_form, uint64_t *const *outs, void *stream_) {
    B2_REQUIRE(ctx && ciphers && counts && outs && (n_keys == 0 || (galois_elements && keys)), "null argument");
    if (groups == 0) return B2_OK;
    const bool trace = output_interval != 1 && apply_field_trace;
    size_t count = 0;
    b2_status st = pack_check(ctx, counts, groups, nl, shift, input_interval, output_interval, galois_elements, keys, n_keys, trace, outs, &count);
    if (st != B2_OK) return st;
    for (size_t i = 0; i < count; i++) B2_REQUIRE(ciphers[i] && (((uintptr_t)ciphers[i]) & 15) == 0, "null or misaligned ciphertext");
    B2_CUDA(cudaSetDevice(ctx->device));
    if (!input_ntt_form) return pack_run(ctx, ciphers, counts, groups, count, nl, shift, input_interval, output_interval, galois_elements, keys, n_keys, trace, output_ntt_form, outs, stream_);
    // NTT-form inputs: one batched inverse transform into a staging area, which the tree then reads
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t ctw = 2 * nl * ctx->n;
    cudaMemPool_t pool;
    st = arena_pool(ctx, &pool);
    if (st != B2_OK) return st;
    uint64_t *staged = nullptr;
    B2_CUDA(cudaMallocFromPoolAsync((void **)&staged, count * ctw * 8, pool, stream));
    std::vector<uint64_t *> souts(count);
    for (size_t i = 0; i < count; i++) souts[i] = staged + i * ctw;
    st = b2_ntt_inverse_ptrs(ctx, ciphers, souts.data(), count, 2, nl, B2_TABLES_COMPONENTWISE, 0, 0, stream_);
    if (st == B2_OK) st = pack_run(ctx, souts.data(), counts, groups, count, nl, shift, input_interval, output_interval, galois_elements, keys, n_keys, trace, output_ntt_form, outs, stream_);
    cudaFreeAsync(staged, stream);
    return st;
}

// LWECiphertext::assemble_lwe_batched (src/lwe_ciphertext.cu:74-140): c0s[b] DEVICE [nl], c1s[b] DEVICE [nl][N] -> outs[b] DEVICE [2][nl][N] (coefficient form), one launch
b2_status b2_assemble_lwes(b2_ctx *ctx, const uint64_t *const *c0s, const uint64_t *const *c1s, uint64_t *const *outs, size_t count, size_t nl, void *stream_) {
    B2_REQUIRE(ctx && (count == 0 || (c0s && c1s && outs)) && nl >= 1 && nl <= ctx->K, "bad arguments");
    if (count == 0) return B2_OK;
    for (size_t i = 0; i < count; i++) B2_REQUIRE(c0s[i] && c1s[i] && outs[i], "null operand in batch");
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    Scratch sc(ctx, stream);
    b2_status st = ensure_ptrtab(sc.s, stream, 3 * count * 8);
    if (st != B2_OK) return st;
    std::vector<uint64_t> blob(3 * count);
    for (size_t i = 0; i < count; i++) { blob[i] = (uint64_t)(uintptr_t)c0s[i]; blob[count + i] = (uint64_t)(uintptr_t)c1s[i]; blob[2 * count + i] = (uint64_t)(uintptr_t)outs[i]; }
    B2_CUDA(cudaMemcpyAsync(sc->ptrtab, blob.data(), blob.size() * 8, cudaMemcpyHostToDevice, stream));
    const u64 *const *t = (const u64 *const *)sc->ptrtab;
    const u64 total = (u64)count * nl * ctx->n;
    lwe_assemble_kernel<<<pk_grid(total), 256, 0, stream>>>(t, t + count, (u64 *const *)(t + 2 * count), (u32)nl, (u32)ctx->n, total);
    count_launches(1);
    B2_CUDA(cudaGetLastError());
    return B2_OK;
}

// Evaluator::extract_lwe_new (src/evaluator_lwes.cu:51-94) on a coefficient-form ciphertext ct [2][nl][N]: c0 DEVICE [nl], c1 DEVICE [nl][N], one launch
b2_status b2_extract_lwe(b2_ctx *ctx, const uint64_t *ct, size_t term, size_t nl, uint64_t *c0, uint64_t *c1, void *stream_) {
    B2_REQUIRE(ctx && ct && c0 && c1 && nl >= 1 && nl <= ctx->K && term < ctx->n, "bad arguments (term must be below N)");
    B2_REQUIRE(ct != c1, "out of place only");
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t n = ctx->n;
    const u64 total = (u64)nl * n;
    lwe_extract_kernel<<<pk_grid(total), 256, 0, stream>>>((const u64 *)ct, (u64 *)c0, (u64 *)c1, ctx->d_mods, (u32)nl, (u32)n, (u32)term, (u32)(term == 0 ? 0 : 2 * n - term), total);
    count_launches(1);
    B2_CUDA(cudaGetLastError());
    return B2_OK;
}

// Evaluator::pack_lwe_ciphertexts_new / _batched (src/evaluator_lwes.cu:201-303): assemble every LWE into an RLWE (one launch for all groups) and pack each
// group of counts[g] LWEs into one ciphertext: input_interval = N, output_interval = N / 2^ceil(log2 counts[g]) — all groups must share it (the reference's
// batched form requires equal group sizes up to the power of two as well).
b2_status b2_pack_lwes(b2_ctx *ctx, const uint64_t *const *c0s, const uint64_t *const *c1s, const size_t *counts, size_t groups, size_t nl, const uint64_t *galois_elements,
                       const b2_kswitch_key *const *keys, size_t n_keys, int apply_field_trace, int output_ntt_form, uint64_t *const *outs, void *stream_) {
    B2_REQUIRE(ctx && c0s && c1s && counts && outs && (n_keys == 0 || (galois_elements && keys)), "null argument");
    if (groups == 0) return B2_OK;
    const size_t n = ctx->n;
    size_t l = 0, maxc = 0;
    for (size_t g = 0; g < groups; g++) { B2_REQUIRE(counts[g] >= 1, "LWE ciphertexts must not be empty."); maxc = std::max(maxc, counts[g]); }
    B2_REQUIRE(maxc <= n, "LWE ciphertexts count must be less than poly_modulus_degree.");
    while (((size_t)1 << l) < maxc) l++;
    const size_t output_interval = n >> l;
    const bool trace = output_interval != 1 && apply_field_trace;
    size_t count = 0;
    b2_status st = pack_check(ctx, counts, groups, nl, 0, n, output_interval, galois_elements, keys, n_keys, trace, outs, &count);
    if (st != B2_OK) return st;
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t ctw = 2 * nl * n;
    cudaMemPool_t pool;
    st = arena_pool(ctx, &pool);
    if (st != B2_OK) return st;
    uint64_t *staged = nullptr;
    B2_CUDA(cudaMallocFromPoolAsync((void **)&staged, count * ctw * 8, pool, stream));
    std::vector<uint64_t *> souts(count);
    for (size_t i = 0; i < count; i++) souts[i] = staged + i * ctw;
    st = b2_assemble_lwes(ctx, c0s, c1s, souts.data(), count, nl, stream_);
    if (st == B2_OK) st = pack_run(ctx, souts.data(), counts, groups, count, nl, 0, n, output_interval, galois_elements, keys, n_keys, trace, output_ntt_form, outs, stream_);
    cudaFreeAsync(staged, stream);
    return st;
}

}  // extern "C"
