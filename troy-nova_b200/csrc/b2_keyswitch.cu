// b2_keyswitch.cu — the keyswitching core (K1-K5 in SURVEY.md §8a).  Replaces
// Evaluator::switch_key_internal(_batched) (src/evaluator_keyswitching_core.cu:757-1267) and its
// helpers fgk::switch_key::set_accumulate / accumulate_products (src/fgk/switch_key.cu:6-176) and
// kernel_ski_util5/6/7_merged (src/evaluator_keyswitching_core.cu:436-688).
//
// Pipeline (R = decomp + 1 output moduli m_i = q_0..q_{decomp-1}, P; `temp` and `prod` live in the
// context's workspace arena, nothing is allocated on the hot path):
//   1. c = INTT(target)                                            (skipped for coefficient-form input)
//   2. temp[i][j] = NTT_{m_i}(c_j mod m_i)   — the mod-up re-reduction (the reference's set_accumulate
//      kernel and its (L+1)·L·N-word buffer) is fused into the NTT's first-pass loader
//   3. prod[k][i] = sum_j temp[i][j] o key_j[k][i]  mod m_i        — 128-bit lazy accumulation, one
//      Barrett reduction per output (the reference reduces every term; result is canonical either way)
//   4. u_k = INTT_P(prod[k][decomp])
//   5. mod-down, fused into ONE forward transform: its loader builds t_{k,j} from u_k on the fly
//      (ski_util6 / ski_util5_step1) and its last-pass store computes (prod[k][j] - t)·P^-1 mod q_j and
//      assigns / accumulates into the destination (ski_util7 / ski_util5_step2).
//      For coefficient-form (BFV) ciphertexts the same happens around INTT(prod[k][j]).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "b2_ntt5_launch.cuh"
#include "b2_ks5.cuh"
#include "b2_fused.cuh"

using namespace b2;

struct b2_kswitch_key {
    const uint64_t **d_ptrs = nullptr;   // device array of D digit base pointers, each [2][rows][N] NTT form
    std::vector<const uint64_t *> h_ptrs; // host copy of the same pointers
    size_t digits = 0;
    int device = 0;
    bool borrowed = false;               // b2_kswitch_key_wrap: d_ptrs belongs to the caller
    // limb shard (b2_kswitch_key_shard): rows first .. first+count-1 of the output moduli 0..decomp only, owned storage
    uint64_t *own = nullptr;
    size_t rows = 0, first = 0, decomp = 0;   // rows == 0: a full key ([2][K][N] per digit)
};

namespace b2 {

// ---- step 1 loader: limb j of the target of ciphertext b (NTT form) -------------------------------------
struct TargetLoad {
    PtrTab t; u32 n;
    __device__ __forceinline__ const u64 *src(int j, int b, u64) const { return t.p[b] + (u64)j * n; }
    B2_IDENTITY_XF
};

// ---- step 2 loader: c_j re-reduced modulo the group's modulus (src/fgk/switch_key.cu:6-27) ----------
// limb index c = b * decomp + j; coefficient-form source of ciphertext b at tc.p[b] ([decomp][N])
struct ModUpLoad {
    PtrTab tc;
    const ModInfo *mods;
    u32 decomp, K, n;
    u32 g0;          // output modulus of group 0 (limb-sharded calls transform a sub-range of the output moduli)
    u32 f64_path;    // 1 iff FP64-pipe moduli really run the FP64 back end in this launch (N >= 4096); the tiny-N kernel is integer-only
    u32 special_first = 0;   // 1: group 0 is the special prime (output modulus `decomp`) and group g > 0 is output modulus g - 1: the slow
                             // integer-pipe limbs of every digit are claimed first, so the queue does not end on one
    __device__ __forceinline__ const u64 *src(int, int c, u64) const {
        const u32 b = (u32)c / decomp, j = (u32)c - b * decomp;
        return tc.p[b] + (u64)j * n;
    }
    struct Prep { u64 m, ratio; bool reduce; };
    __device__ __forceinline__ Prep prep(int g, int c) const {
        const u32 i = special_first ? ((u32)g == 0 ? decomp : (u32)g - 1) : (u32)g + g0;
        const u32 ki = i == decomp ? K - 1 : i;
        const u32 j = (u32)c % decomp;
        const ModInfo &Mi = mods[ki];
        Prep r;
        r.m = Mi.p; r.ratio = Mi.ratio;
        // q_j <= m_i: already canonical modulo m_i.  FP64-pipe moduli take lazy inputs (|x| < 2^10 m_i, b2_modarith.cuh), so a
        // digit modulus at most 4 bits longer needs no reduction here either: the transform's first multiplication does it
        r.reduce = !(mods[j].p <= r.m || (f64_path && Mi.use_f64 && mods[j].bits <= Mi.bits + 4));
        return r;
    }
    __device__ __forceinline__ u64 xf(const Prep &pc, u32, u64 v) const { return pc.reduce ? barrett64(v, pc.m, pc.ratio) : v; }
};

// ---- step 3: key inner product (src/fgk/switch_key.cu:77-132) --------------------------------------
// one thread per (ciphertext b, output modulus i, coefficient pair); loops over the decomp digits with 128-bit lazy
// accumulators.  temp [R][B][decomp][N], prod [B][2][R][N].  blockIdx.x = b so that the CTAs that read the same key
// rows are co-scheduled and all but the first hit in L2.
__global__ void __launch_bounds__(256) ks_inner_product_kernel(const u64 *__restrict__ temp, const u64 *const *__restrict__ keys,
                                                                 u64 *__restrict__ prod, const ModInfo *__restrict__ mods,
                                                                 u32 decomp, u32 K, u32 n, u32 i0, u32 key_rows, u32 prod_rows, u32 prod_row0) {
    const u32 b = blockIdx.x, nb = gridDim.x;
    const u32 il = blockIdx.z;                      // local output modulus index; the modulus is i0 + il in 0..decomp
    const u32 i = il;
    const u32 ki = i0 + il == decomp ? K - 1 : i0 + il;
    const u32 krow = key_rows == K ? ki : il;       // full key: row of the key-level modulus; limb shard: local row
    const u32 x = (blockIdx.y * blockDim.x + threadIdx.x) * 2;
    if (x >= n) return;
    const ModInfo &M = mods[ki];
    const u64 p = M.p, r0 = M.r0, r1 = M.r1;
    u64 a0l = 0, a0h = 0, a1l = 0, a1h = 0, b0l = 0, b0h = 0, b1l = 0, b1h = 0;
    const u64 *tp = temp + (((u64)i * nb + b) * decomp) * n + x;
    for (u32 j = 0; j < decomp; j++) {
        const ulonglong2 d = __ldcs(reinterpret_cast<const ulonglong2 *>(tp + (u64)j * n));   // streamed once: evict-first
        const u64 *kp = keys[j] + (u64)krow * n + x;
        ulonglong2 k0, k1;
        if (nb == 1) { k0 = __ldcs(reinterpret_cast<const ulonglong2 *>(kp)); k1 = __ldcs(reinterpret_cast<const ulonglong2 *>(kp + (u64)key_rows * n)); }
        else { k0 = __ldg(reinterpret_cast<const ulonglong2 *>(kp)); k1 = __ldg(reinterpret_cast<const ulonglong2 *>(kp + (u64)key_rows * n)); }
        mac128(a0l, a0h, d.x, k0.x); mac128(a1l, a1h, d.y, k0.y);
        mac128(b0l, b0h, d.x, k1.x); mac128(b1l, b1h, d.y, k1.y);
        if ((j & 7) == 7) {     // keep the lazy 128-bit sums below 2^125 (8 products of 61-bit residues)
            a0l = barrett128(a0l, a0h, p, r0, r1); a0h = 0; a1l = barrett128(a1l, a1h, p, r0, r1); a1h = 0;
            b0l = barrett128(b0l, b0h, p, r0, r1); b0h = 0; b1l = barrett128(b1l, b1h, p, r0, r1); b1h = 0;
        }
    }
    const u32 R = prod_rows, io = prod_row0 + il;       // row of this modulus in prod [B][2][prod_rows][N]
    ulonglong2 o0, o1;
    o0.x = barrett128(a0l, a0h, p, r0, r1); o0.y = barrett128(a1l, a1h, p, r0, r1);
    o1.x = barrett128(b0l, b0h, p, r0, r1); o1.y = barrett128(b1l, b1h, p, r0, r1);
    *reinterpret_cast<ulonglong2 *>(prod + (((u64)b * 2 + 0) * R + io) * n + x) = o0;
    *reinterpret_cast<ulonglong2 *>(prod + (((u64)b * 2 + 1) * R + io) * n + x) = o1;
}

// limb-sharded path: the all-gather delivers [world][pad][N] (rank r's owned ciphertext limbs in slot r); one launch
// compacts it to tc[decomp][N].  Rank ranges as in b2_limb_partition: the first `rem` ranks own base + 1 output moduli.
__global__ void __launch_bounds__(256) ks_compact_gathered_kernel(const u64 *__restrict__ gathered, u64 *__restrict__ tc, u32 n, u32 pad,
                                                                    u32 base, u32 rem) {
    const u32 j = blockIdx.y;                      // ciphertext limb
    const u32 split = rem * (base + 1);
    const u32 r = j < split ? j / (base + 1) : rem + (j - split) / base;
    const u32 first = r * base + min(r, rem);
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(gathered + ((u64)r * pad + (j - first)) * n);
    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(tc + (u64)j * n);
    for (u32 x = blockIdx.x * blockDim.x + threadIdx.x; x < n / 2; x += gridDim.x * blockDim.x) dst[x] = src[x];
}

// ---- P2P data plane of the limb-sharded keyswitch -------------------------------------------------------------------------
// The peers' buffers are mapped into this process (symmetric memory / CUDA IPC): a rank PUSHES its part of an exchange into
// every peer's buffer with plain stores over NVLink and then raises its flag in every peer's flag array; the consumer side
// spins on its local flags.  No collective call, no staging copy, no compaction: the exchange is two small kernels ordered on
// the compute stream (reference has no counterpart: it replicates whole contexts per device).
constexpr int SHARD_MAX_WORLD = 16;
struct PeerTab { u64 *p[SHARD_MAX_WORLD]; };
__global__ void __launch_bounds__(256) shard_push_kernel(const ulonglong2 *__restrict__ src, PeerTab dst, PeerTab flags, u32 world, u64 chunks, u32 flag_slot, u64 epoch,
                                                         u32 *__restrict__ counter) {
    for (u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < chunks; v += (u64)gridDim.x * blockDim.x) {
        const ulonglong2 d = src[v];
        for (u32 r = 0; r < world; r++) reinterpret_cast<ulonglong2 *>(dst.p[r])[v] = d;
    }
    __threadfence_system();                  // this thread's peer stores are performed system-wide ...
    __shared__ u32 last;
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (last) {                              // ... and the last block to get here publishes the epoch to every peer
        __threadfence_system();
        if (threadIdx.x < world) {
            u64 *f = flags.p[threadIdx.x] + flag_slot;
            asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(f), "l"(epoch) : "memory");
        }
        if (threadIdx.x == 0) *counter = 0;
    }
}
struct SlotTab { int s[SHARD_MAX_WORLD]; };
__global__ void shard_wait_kernel(const u64 *__restrict__ flags, SlotTab slots, u32 n_slots, u64 epoch) {
    if (threadIdx.x < n_slots) {
        const u64 *f = flags + slots.s[threadIdx.x];
        u64 v = 0;
        for (u32 spins = 0;; spins++) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(f) : "memory");
            if (v >= epoch) break;
            __nanosleep(100);
            if (spins == (1u << 25)) { printf("b2 shard wait: flag slot %d stuck at %llu, waiting for %llu\n", slots.s[threadIdx.x], v, epoch); __trap(); }
        }
    }
}

struct SpecialLoad {      // step 4: the special-prime limb of prod[b][k], c = 2 b + k (row `decomp` of R rows)
    const u64 *prod; u32 R, decomp, n;
    __device__ __forceinline__ const u64 *src(int, int c, u64) const { return prod + ((u64)c * R + decomp) * n; }
    B2_IDENTITY_XF
};
struct ProdLoad {         // coefficient-form path: transform prod[b][k][j] itself (inverse NTT); `rows` rows per polynomial
    const u64 *prod; u32 rows, n;
    __device__ __forceinline__ const u64 *src(int j, int c, u64) const { return prod + ((u64)c * rows + j) * n; }
    B2_IDENTITY_XF
};


// step 5 launcher: the scheme (BGV or not) and the ciphertext form select one kernel instantiation each
template <bool BGV>
static b2_status mod_down_t(b2_ctx *ctx, b2_scratch *sc, const std::vector<LimbGroup> &groups, u64 *mid, const u64 *u, const u64 *prod, const PtrTabW &dp, size_t decomp,
                            int assign, int is_ntt, size_t jbase, size_t prod_rows, size_t dest_rows, cudaStream_t stream) {
    const size_t K = ctx->K, n = ctx->n;
    ModDownTerm<BGV> T{u, ctx->d_mods, (u32)(K - 1), (u32)n, ctx->t, ctx->t_ratio, ctx->t_r0, ctx->t_r1, BGV ? ctx->h_inv_mod_t[K - 1] : 0};
    if (is_ntt) {
        ModDownStore<BGV, true> sst{T, prod, ctx->d_inv + (K - 1) * K, dp, (u32)decomp, (u32)assign, (u32)jbase, (u32)prod_rows, (u32)dest_rows};
        return run_ntt_dir<true>(ctx, sc, groups, mid, ModDownLoad<BGV>{T, (u32)jbase}, sst, stream);
    }
    ModDownStore<BGV, false> sst{T, prod, ctx->d_inv + (K - 1) * K, dp, (u32)decomp, (u32)assign, (u32)jbase, (u32)prod_rows, (u32)dest_rows};
    return run_ntt_dir<false>(ctx, sc, groups, mid, ProdLoad{prod, (u32)prod_rows, (u32)n}, sst, stream);
}
static b2_status mod_down(b2_ctx *ctx, b2_scratch *sc, const std::vector<LimbGroup> &groups, u64 *mid, const u64 *u, const u64 *prod, const PtrTabW &dp, size_t decomp,
                          int assign, int is_ntt, size_t jbase, size_t prod_rows, size_t dest_rows, cudaStream_t stream) {
    // BGV ciphertexts are always NTT-form (checked by the callers): the coefficient-form BGV kernel is never instantiated
    if (ctx->scheme == B2_SCHEME_BGV) {
        ModDownTerm<true> T{u, ctx->d_mods, (u32)(ctx->K - 1), (u32)ctx->n, ctx->t, ctx->t_ratio, ctx->t_r0, ctx->t_r1, ctx->h_inv_mod_t[ctx->K - 1]};
        ModDownStore<true, true> sst{T, prod, ctx->d_inv + (ctx->K - 1) * ctx->K, dp, (u32)decomp, (u32)assign, (u32)jbase, (u32)prod_rows, (u32)dest_rows};
        return run_ntt_dir<true>(ctx, sc, groups, mid, ModDownLoad<true>{T, (u32)jbase}, sst, stream);
    }
    return mod_down_t<false>(ctx, sc, groups, mid, u, prod, dp, decomp, assign, is_ntt, jbase, prod_rows, dest_rows, stream);
}
}  // namespace b2

extern "C" {

b2_status b2_kswitch_key_create(b2_ctx *ctx, const uint64_t *const *digit_ptrs, size_t digits, b2_kswitch_key **out) {
    B2_REQUIRE(ctx && digit_ptrs && out && digits >= 1 && digits <= 64, "bad arguments");
    B2_CUDA(cudaSetDevice(ctx->device));
    b2_kswitch_key *k = new b2_kswitch_key();
    k->digits = digits; k->device = ctx->device;
    k->h_ptrs.assign(digit_ptrs, digit_ptrs + digits);
    cudaError_t e = cudaMalloc((void **)&k->d_ptrs, digits * sizeof(uint64_t *));
    if (e != cudaSuccess) { delete k; return cuda_fail(e, "cudaMalloc(key ptrs)"); }
    e = cudaMemcpy((void *)k->d_ptrs, digit_ptrs, digits * sizeof(uint64_t *), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree((void *)k->d_ptrs); delete k; return cuda_fail(e, "cudaMemcpy(key ptrs)"); }
    *out = k;
    return B2_OK;
}
b2_status b2_kswitch_key_create_contiguous(b2_ctx *ctx, const uint64_t *key, size_t digits, b2_kswitch_key **out) {
    B2_REQUIRE(ctx && key && digits >= 1 && digits <= 64, "bad arguments");
    const uint64_t *ptrs[64];
    for (size_t j = 0; j < digits; j++) ptrs[j] = key + j * 2 * ctx->K * ctx->n;
    return b2_kswitch_key_create(ctx, ptrs, digits, out);
}
// zero-copy handle over a pointer table the caller already keeps on the device (the reference's KSwitchKeys::get_data_ptrs,
// src/kswitch_keys.h:36-57): no allocation, no copy, no synchronisation — cheap enough to build per keyswitch call
b2_status b2_kswitch_key_wrap(b2_ctx *ctx, const uint64_t *const *device_ptr_table, const uint64_t *const *digit_ptrs, size_t digits, b2_kswitch_key **out) {
    B2_REQUIRE(ctx && device_ptr_table && digit_ptrs && out && digits >= 1 && digits <= 64, "bad arguments");
    b2_kswitch_key *k = new b2_kswitch_key();
    k->digits = digits; k->device = ctx->device; k->borrowed = true;
    k->h_ptrs.assign(digit_ptrs, digit_ptrs + digits);
    k->d_ptrs = (const uint64_t **)device_ptr_table;
    *out = k;
    return B2_OK;
}
void b2_kswitch_key_destroy(b2_kswitch_key *k) {
    if (!k) return;
    if (k->borrowed) { delete k; return; }
    cudaSetDevice(k->device);
    cudaFree((void *)k->d_ptrs);
    cudaFree(k->own);
    delete k;
}

// one chunk of nb <= KS_MAXB ciphertexts sharing `key`
static b2_status keyswitch_chunk(b2_ctx *ctx, const uint64_t *const *targets, uint64_t *const *dests, size_t nb, size_t decomp, int is_ntt,
                                 const b2_kswitch_key *key, int assign, cudaStream_t stream) {
    const size_t n = ctx->n, K = ctx->K, R = decomp + 1;
    // workspace: tc [B][decomp][N] | temp [blk][B][decomp][N] | prod [B][2][R][N] | u [B][2][N] | mid [B][2][decomp][N]
    static const size_t blk_env = [] { const char *e = getenv("B2_KS_BLOCK"); return e && atoi(e) > 0 ? (size_t)atoi(e) : (size_t)0; }();
    // output moduli per mod-up / inner-product block.  Default: all of them in one block.  Smaller blocks keep the block's temp
    // slice L2-resident between the transform and the inner product, but the shorter transforms lose more to wave
    // quantisation than the saved HBM traffic buys (C3 on B200, r01h: 21 / 11 / 7 / 5 / 3 moduli per block -> 545 / 570 / 593 /
    // 646 / 672 us per keyswitch); B2_KS_BLOCK is the developer override the sweep used
    const size_t blk = std::min(R, blk_env ? blk_env : R);
    const size_t w_tc = nb * decomp * n, w_temp = blk * nb * decomp * n, w_prod = nb * 2 * R * n, w_u = nb * 2 * n, w_mid = nb * 2 * decomp * n;
    Scratch scl(ctx, stream);
    b2_scratch *sc = scl.s;
    b2_status st = ensure_ws(sc, stream, (w_tc + w_temp + w_prod + w_u + w_mid) * 8);
    if (st != B2_OK) return st;
    u64 *tc = (u64 *)sc->ws, *temp = tc + w_tc, *prod = temp + w_temp, *u = prod + w_prod, *mid = u + w_u;

    // developer aid: B2_KS_TIMING=1 prints per-stage device times (synchronises; never set in production)
    static const bool timing = getenv("B2_KS_TIMING") != nullptr;
    cudaEvent_t ev[6];
    int nev = 0;
    auto mark = [&]() { if (timing) { cudaEventCreate(&ev[nev]); cudaEventRecord(ev[nev], stream); nev++; } };
    mark();
    // 1. coefficient form of the targets
    PtrTab coef;
    for (size_t b = 0; b < nb; b++) coef.p[b] = (const u64 *)targets[b];
    for (size_t b = nb; b < (size_t)KS_MAXB; b++) coef.p[b] = nullptr;
    if (is_ntt) {
        std::vector<LimbGroup> groups;
        for (size_t j = 0; j < decomp; j++) groups.push_back(LimbGroup{(u32)j, (u32)nb, (u64)(j * n), (u64)(decomp * n)});
        TargetLoad ld{coef, (u32)n};
        st = run_ntt_dir<false>(ctx, sc, groups, tc, ld, PlainStore{tc}, stream);
        if (st != B2_OK) return st;
        for (size_t b = 0; b < nb; b++) coef.p[b] = tc + b * decomp * n;
    }
    mark();
    // 2 + 3. mod-up NTT and key inner product, per block of `blk` output moduli (one block by default, see above)
    //   group i = output modulus, limbs c = (b, j); temp[il][b][j], il = i - i0
    // 4. u_{b,k} = INTT_P(prod[b][k][decomp]).  In the one-block case the special prime's inner product is launched first and its
    //   inverse transform (a few limbs on the cluster kernel: latency-bound, ~30 us) runs on a side stream UNDER the inner
    //   product of the other moduli instead of after it.
    static const bool no_overlap = getenv("B2_KS_NO_OVERLAP") != nullptr;      // developer switch for A/B timing
    // N = 2^16, one ciphertext: steps 2 + 3 as one persistent launch whose temp is an L2-resident ring (b2_ks5.cuh).  Bit-exact, but
    // measured SLOWER than the two launches on B200 (585 us vs 460 us per C3 keyswitch; the in-order queue stalls on the group
    // dependencies and L2 still writes 239 MB of the ring back: profiles/r03b_ks5_*), so it is opt-in (B2_KS5=1) until that is solved.
    static const bool ks5_off = [] { const char *e = getenv("B2_KS5"); return !(e && e[0] == '1'); }();
    bool fused = false;
    if (ctx->logn == N5_LOGN && nb == 1 && blk == R && R * decomp >= (size_t)KS5_RING * decomp && R <= (size_t)KS5_MAX_GROUPS && !ks5_off && !timing && n5_enabled() &&
        (((uintptr_t)coef.p[0]) & 127) == 0) {
        Ks5Job J;
        J.mods = ctx->d_mods; J.keys = (const u64 *const *)key->d_ptrs; J.ring = temp; J.prod = prod;
        J.G = (u32)R; J.C = (u32)decomp; J.K = (u32)K; J.key_rows = (u32)K;
        J.total_items = (u32)(R * decomp * 32 + R * KS5_IPI);
        J.f64_lazy = 1;
        for (size_t i = 0; i < R; i++) {
            const u32 ti = (u32)(i == decomp ? K - 1 : i);
            const ModInfo &M = ctx->h_mods[ti];
            J.table[i] = tag_table(ti, M.use_f64 ? PATH_F64 : (M.fwd_lazy ? PATH_INT_LAZY : PATH_INT), M.use_f64 ? PATH_F64 : (M.inv_lazy ? PATH_INT_LAZY : PATH_INT));
        }
        Ks5Maps MP;
        if (n5_make_map(&MP.tc, N5_NATURAL, coef.p[0], decomp) && n5_make_map(&MP.ring_perm, N5_PERMUTED, temp, KS5_RING * decomp) &&
            n5_make_map(&MP.ring_rows, N5_ROWS, temp, KS5_RING * decomp)) {
            st = ensure_sync(sc, R * decomp + 2 * R, stream);
            if (st != B2_OK) return st;
            J.sync = sc->d_sync;
            static bool configured[B2_MAX_DEVICES] = {};
            if (!configured[ctx->device & (B2_MAX_DEVICES - 1)]) {
                B2_CUDA(cudaFuncSetAttribute(ks5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, N5_SMEM_BYTES));
                configured[ctx->device & (B2_MAX_DEVICES - 1)] = true;
            }
            const unsigned grid = std::min<unsigned>(J.total_items, 3u * (unsigned)ctx->sm_count);
            ks5_kernel<<<grid, N5_THREADS, N5_SMEM_BYTES, stream>>>(J, MP);
            count_launches(1);
            B2_CUDA(cudaGetLastError());
            fused = true;
        }
    }
    const bool overlap_special = blk == R && 2 * nb <= cluster_max_limbs() && ctx->logn == 16 && !timing && !no_overlap;
    bool special_done = false;
    for (size_t i0 = 0; i0 < R && !fused; i0 += blk) {
        const size_t cnt = std::min(blk, R - i0);
        std::vector<LimbGroup> groups;
        // one block covering every output modulus: the special prime's group goes first (its place in `temp` stays row `decomp`)
        static const bool sp_first_env = [] { const char *e = getenv("B2_KS_SPECIAL_FIRST"); return !(e && e[0] == '0'); }();
        const bool sp_first = sp_first_env && i0 == 0 && cnt == R;
        if (sp_first) groups.push_back(LimbGroup{(u32)(K - 1), (u32)(nb * decomp), (u64)(decomp * nb * decomp * n), (u64)n});
        for (size_t il = 0; il < cnt; il++) {
            const size_t i = i0 + il;
            if (sp_first && i == decomp) continue;
            groups.push_back(LimbGroup{(u32)(i == decomp ? K - 1 : i), (u32)(nb * decomp), (u64)(il * nb * decomp * n), (u64)n});
        }
        ModUpLoad ld{coef, ctx->d_mods, (u32)decomp, (u32)K, (u32)n, (u32)i0, ctx->logn > 11 ? 1u : 0u, sp_first ? 1u : 0u};
        st = run_ntt_dir<true>(ctx, sc, groups, temp, ld, PlainStore{temp}, stream);
        if (st != B2_OK) return st;
        if (i0 == 0) mark();
        const unsigned gy = (unsigned)((n / 2 + 255) / 256);
        if (overlap_special) {
            st = ensure_side(sc, 0, stream);
            if (st != B2_OK) return st;
            // special prime first (row `decomp` of temp and of prod) ...
            ks_inner_product_kernel<<<dim3((unsigned)nb, gy, 1), 256, 0, stream>>>(temp + decomp * nb * decomp * n, (const u64 *const *)key->d_ptrs, prod, ctx->d_mods,
                                                                                 (u32)decomp, (u32)K, (u32)n, (u32)decomp, (u32)K, (u32)R, (u32)decomp);
            B2_CUDA(cudaGetLastError());
            B2_CUDA(cudaEventRecord(sc->ev_fork, stream));
            B2_CUDA(cudaStreamWaitEvent(sc->s_side, sc->ev_fork, 0));
            {   // ... its inverse transform on the side stream ...
                std::vector<LimbGroup> sg{LimbGroup{(u32)(K - 1), (u32)(2 * nb), 0, (u64)n}};
                SpecialLoad sld{prod, (u32)R, (u32)decomp, (u32)n};
                st = run_ntt_dir<false>(ctx, sc, sg, u, sld, PlainStore{u}, sc->s_side);
                if (st != B2_OK) return st;
            }
            B2_CUDA(cudaEventRecord(sc->ev_join, sc->s_side));
            // ... under the inner product of the ciphertext moduli
            ks_inner_product_kernel<<<dim3((unsigned)nb, gy, (unsigned)decomp), 256, 0, stream>>>(temp, (const u64 *const *)key->d_ptrs, prod, ctx->d_mods, (u32)decomp, (u32)K, (u32)n,
                                                                                             0u, (u32)K, (u32)R, 0u);
            count_launches(2);
            B2_CUDA(cudaGetLastError());
            B2_CUDA(cudaStreamWaitEvent(stream, sc->ev_join, 0));
            special_done = true;
            continue;
        }
        dim3 grid((unsigned)nb, gy, (unsigned)cnt);
        ks_inner_product_kernel<<<grid, 256, 0, stream>>>(temp, (const u64 *const *)key->d_ptrs, prod, ctx->d_mods, (u32)decomp, (u32)K, (u32)n,
                                                          (u32)i0, (u32)K, (u32)R, (u32)i0);
        count_launches(1);
        B2_CUDA(cudaGetLastError());
    }
    mark();
    if (!special_done) {
        std::vector<LimbGroup> groups{LimbGroup{(u32)(K - 1), (u32)(2 * nb), 0, (u64)n}};
        SpecialLoad ld{prod, (u32)R, (u32)decomp, (u32)n};
        st = run_ntt_dir<false>(ctx, sc, groups, u, ld, PlainStore{u}, stream);
        if (st != B2_OK) return st;
    }
    mark();
    // 5. mod-down fused around one transform of [B][2][decomp] limbs (group j = modulus, limb c = 2 b + k)
    {
        std::vector<LimbGroup> groups;
        for (size_t j = 0; j < decomp; j++) groups.push_back(LimbGroup{(u32)j, (u32)(2 * nb), (u64)(j * n), (u64)(decomp * n)});
        PtrTabW dp;
        for (size_t b = 0; b < (size_t)KS_MAXB; b++) dp.p[b] = b < nb ? (u64 *)dests[b] : nullptr;
        st = mod_down(ctx, sc, groups, mid, u, prod, dp, decomp, assign, is_ntt, 0, R, decomp, stream);
        if (st != B2_OK) return st;
    }
    mark();
    if (timing) {
        cudaEventSynchronize(ev[nev - 1]);
        static const char *names[] = {"intt(target)", "modup+ntt (block 0)", "other blocks + inner products", "intt(special)", "moddown+ntt"};
        float tot = 0; cudaEventElapsedTime(&tot, ev[0], ev[nev - 1]);
        fprintf(stderr, "[b2 keyswitch x%zu] total %.1f us:", nb, tot * 1e3f);
        for (int i = 0; i + 1 < nev; i++) { float ms = 0; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]); fprintf(stderr, " %s %.1f", names[i], ms * 1e3f); }
        fprintf(stderr, "\n");
        for (int i = 0; i < nev; i++) cudaEventDestroy(ev[i]);
    }
    return B2_OK;
}

b2_status b2_keyswitch_batched(b2_ctx *ctx, const uint64_t *const *targets, uint64_t *const *dests, size_t batch, size_t decomp, int is_ntt,
                               const b2_kswitch_key *key, int assign, void *stream_) {
    B2_REQUIRE(ctx && key && (batch == 0 || (targets && dests)), "null argument");
    B2_REQUIRE(decomp >= 1 && decomp + 1 <= ctx->K, "decomp must be in [1, K-1] (keyswitching needs a special prime)");
    B2_REQUIRE(key->digits >= decomp, "keyswitching key has fewer digits than the ciphertext has limbs");
    B2_REQUIRE(key->rows == 0, "a limb-sharded key can only be used with the b2_keyswitch_shard_* entry points");
    B2_REQUIRE(assign >= 0 && assign <= 2, "unknown assign method");
    B2_REQUIRE(ctx->scheme != B2_SCHEME_BGV || is_ntt, "BGV ciphertexts are NTT-form");
    B2_REQUIRE(ctx->scheme != B2_SCHEME_BGV || ctx->t > 1, "BGV needs a plain modulus");
    for (size_t b = 0; b < batch; b++) {
        B2_REQUIRE(targets[b] && dests[b], "null operand in batch");
        B2_REQUIRE((((uintptr_t)targets[b] | (uintptr_t)dests[b]) & 15) == 0, "operands must be 16-byte aligned");
    }
    B2_CUDA(cudaSetDevice(ctx->device));
    // chunk so that the workspace of one chunk stays below ~4 GiB
    const size_t per_ct = (decomp + (decomp + 1) * decomp + 2 * (decomp + 1) + 2 + 2 * decomp) * ctx->n * 8;
    size_t chunk = std::max<size_t>(1, std::min<size_t>((size_t)KS_MAXB, ((size_t)4 << 30) / per_ct));
    for (size_t b0 = 0; b0 < batch; b0 += chunk) {
        b2_status st = keyswitch_chunk(ctx, targets + b0, dests + b0, std::min(chunk, batch - b0), decomp, is_ntt, key, assign, (cudaStream_t)stream_);
        if (st != B2_OK) return st;
    }
    return B2_OK;
}

// Host-buffer keyswitch (the e2e leg of bench.py; what a host-memory caller of the reference does with to_device / to_host
// around Evaluator::relinearize / apply_galois): the key stays resident on the device, target and destination live in host
// memory (pinned for full-rate DMA).  H2D of the target (and of the destination when it is accumulated into), the five
// launches, D2H of the destination; synchronised on return.
b2_status b2_keyswitch_host(b2_ctx *ctx, const uint64_t *target, size_t decomp, int is_ntt, const b2_kswitch_key *key,
                            uint64_t *dest, int assign, void *stream_) {
    B2_REQUIRE(ctx && target && key && dest, "null argument");
    B2_REQUIRE(decomp >= 1 && decomp + 1 <= ctx->K, "decomp must be in [1, K-1]");
    B2_REQUIRE(assign >= 0 && assign <= 2, "unknown assign method");
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t n = ctx->n, tb = decomp * n * 8, db = 2 * decomp * n * 8;
    uint64_t *d_t = nullptr, *d_d = nullptr, *d_o = nullptr;
    cudaStream_t s_in = nullptr;
    cudaEvent_t ev_start = nullptr, ev_dest = nullptr;
    {
        Scratch sc(ctx, stream);
        b2_status st = ensure_stage(sc.s, stream, tb + 2 * db);
        if (st != B2_OK) return st;
        st = ensure_copy_pipeline(sc.s, 2);
        if (st != B2_OK) return st;
        d_t = (uint64_t *)sc->stage; d_d = d_t + decomp * n; d_o = d_d + 2 * decomp * n;
        s_in = sc->s_in; ev_start = sc->events[0]; ev_dest = sc->events[1];
    }   // (the scratch lock is released: b2_keyswitch takes it again; the staging buffer and the copy stream are private to this stream)
    // The destination's old contents are only needed by the final addition: they travel on the copy stream UNDER the keyswitch, which
    // writes its result (Overwrite) into a third buffer; one element-wise launch then adds them.  (Before: target and destination were
    // uploaded back to back in front of the keyswitch: 31 MB instead of 10 MB on the critical path of a C3 relinearize.)
    B2_CUDA(cudaMemcpyAsync(d_t, target, tb, cudaMemcpyHostToDevice, stream));
    const size_t old_polys = assign == B2_ASSIGN_ADD_INPLACE ? 2 : (assign == B2_ASSIGN_OVERWRITE_EXCEPT_FIRST ? 1 : 0);
    if (old_polys) {
        B2_CUDA(cudaEventRecord(ev_start, stream));                 // after the caller's earlier work on `stream` (the staging buffer may still be read by it)
        B2_CUDA(cudaStreamWaitEvent(s_in, ev_start, 0));
        B2_CUDA(cudaMemcpyAsync(d_d, dest, old_polys * decomp * n * 8, cudaMemcpyHostToDevice, s_in));
        B2_CUDA(cudaEventRecord(ev_dest, s_in));
    }
    const uint64_t *tp = d_t; uint64_t *op = d_o;
    b2_status st = b2_keyswitch_batched(ctx, &tp, &op, 1, decomp, is_ntt, key, B2_ASSIGN_OVERWRITE, stream_);
    if (st != B2_OK) return st;
    if (old_polys) {
        B2_CUDA(cudaStreamWaitEvent(stream, ev_dest, 0));
        st = b2_add(ctx, d_o, d_d, d_o, old_polys, decomp, 0, stream_);
        if (st != B2_OK) return st;
    }
    B2_CUDA(cudaMemcpyAsync(dest, d_o, db, cudaMemcpyDeviceToHost, stream));
    B2_CUDA(cudaStreamSynchronize(stream));
    return B2_OK;
}

b2_status b2_keyswitch(b2_ctx *ctx, const uint64_t *target, size_t decomp, int is_ntt, const b2_kswitch_key *key,
                       uint64_t *dest, int assign, void *stream_) {
    B2_REQUIRE(ctx && target && key && dest, "null argument");
    return b2_keyswitch_batched(ctx, &target, &dest, 1, decomp, is_ntt, key, assign, stream_);
}


// ---- limb-sharded keyswitch: the decomp + 1 output moduli are partitioned over the ranks ----------------
// Rank r owns the contiguous range [first, first + count) of the output moduli 0 .. decomp (decomp = the special prime)
// and holds the ciphertext limbs and the key rows of those moduli only.  One keyswitch is three local stages around
// two collectives that the caller (one process per GPU) issues between them:
//   prepare   c_j = INTT(target_j) for the owned ciphertext limbs        -> ALL-GATHER of the coefficient-form target
//   products  temp[i][j] = NTT_{m_i}(c_j mod m_i), prod[k][i] = sum_j temp[i][j] o key_j[k][i] for the owned i;
//             the owner of the special prime also computes u_k = INTT_P(prod[k][decomp])   -> BROADCAST of u (2 N words)
//   finish    mod-down of the owned ciphertext limbs into the local destination
// The per-rank work and key traffic are 1 / world of the single-GPU keyswitch; results are bit-identical to it.

b2_status b2_limb_partition(size_t n_out_moduli, int world, int rank, size_t *first, size_t *count) {
    B2_REQUIRE(first && count && world >= 1 && rank >= 0 && rank < world && n_out_moduli >= 1, "bad arguments");
    const size_t base = n_out_moduli / (size_t)world, rem = n_out_moduli % (size_t)world, r = (size_t)rank;
    *count = base + (r < rem ? 1 : 0);
    *first = r * base + std::min(r, rem);
    return B2_OK;
}

b2_status b2_kswitch_key_shard(b2_ctx *ctx, const b2_kswitch_key *full, size_t decomp, size_t first, size_t count, b2_kswitch_key **out) {
    B2_REQUIRE(ctx && full && out, "null argument");
    B2_REQUIRE(full->rows == 0, "the source must be a full key");
    B2_REQUIRE(decomp >= 1 && decomp + 1 <= ctx->K && full->digits >= decomp, "bad decomp");
    B2_REQUIRE(count >= 1 && first + count <= decomp + 1, "modulus range out of bounds");
    B2_CUDA(cudaSetDevice(ctx->device));
    const size_t n = ctx->n, K = ctx->K;
    b2_kswitch_key *k = new b2_kswitch_key();
    k->digits = decomp; k->device = ctx->device; k->rows = count; k->first = first; k->decomp = decomp;
    cudaError_t e = cudaMalloc((void **)&k->own, decomp * 2 * count * n * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&k->d_ptrs, decomp * sizeof(uint64_t *));
    if (e != cudaSuccess) { cudaFree(k->own); delete k; return cuda_fail(e, "cudaMalloc(key shard)"); }
    for (size_t j = 0; j < decomp && e == cudaSuccess; j++) {
        k->h_ptrs.push_back(k->own + j * 2 * count * n);
        for (size_t p = 0; p < 2 && e == cudaSuccess; p++)
            for (size_t il = 0; il < count && e == cudaSuccess; il++) {
                const size_t i = first + il, ki = i == decomp ? K - 1 : i;
                e = cudaMemcpy(k->own + ((j * 2 + p) * count + il) * n, full->h_ptrs[j] + (p * K + ki) * n, n * 8, cudaMemcpyDeviceToDevice);
            }
    }
    if (e == cudaSuccess) e = cudaMemcpy((void *)k->d_ptrs, k->h_ptrs.data(), decomp * sizeof(uint64_t *), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { b2_kswitch_key_destroy(k); return cuda_fail(e, "cudaMemcpy(key shard)"); }
    *out = k;
    return B2_OK;
}

// Partition weighted by what a modulus costs its owner: a modulus above 41 bits runs on the integer pipe (~2 x a 40-bit one), and
// the special prime's owner carries the serial tail (inverse transform of the special limb + its push), so it counts once more.
// Contiguous ranges, maximal load minimised (dynamic programme over the decomp + 1 output moduli).
b2_status b2_limb_partition_weighted(const b2_ctx *ctx, size_t decomp, int world, int rank, size_t *first, size_t *count) {
    B2_REQUIRE(ctx && first && count && world >= 1 && rank >= 0 && rank < world && decomp >= 1 && decomp + 1 <= ctx->K, "bad arguments");
    const size_t R = decomp + 1;
    std::vector<unsigned> w(R);
    for (size_t i = 0; i < R; i++) { const size_t ki = i == decomp ? ctx->K - 1 : i; w[i] = (ctx->h_mods[ki].use_f64 ? 1u : 2u) + (i == decomp ? 1u : 0u); }
    std::vector<unsigned> pre(R + 1, 0);
    for (size_t i = 0; i < R; i++) pre[i + 1] = pre[i] + w[i];
    const size_t W = (size_t)world;
    // best[k][i]: minimal maximal load of the first i moduli over k ranks (ranks may be empty when world > R)
    std::vector<std::vector<unsigned>> best(W + 1, std::vector<unsigned>(R + 1, ~0u));
    std::vector<std::vector<size_t>> cut(W + 1, std::vector<size_t>(R + 1, 0));
    best[0][0] = 0;
    for (size_t k = 1; k <= W; k++)
        for (size_t i = 0; i <= R; i++)
            for (size_t j = 0; j <= i; j++) {
                if (best[k - 1][j] == ~0u) continue;
                const unsigned load = std::max(best[k - 1][j], pre[i] - pre[j]);
                if (load < best[k][i] || (load == best[k][i] && j > cut[k][i])) { best[k][i] = load; cut[k][i] = j; }
            }
    size_t hi = R;
    for (size_t k = W; k >= 1; k--) {
        const size_t lo = cut[k][hi];
        if ((int)(k - 1) == rank) { *first = lo; *count = hi - lo; return B2_OK; }
        hi = lo;
    }
    *first = 0; *count = 0;
    return B2_OK;
}

b2_status b2_shard_push(b2_ctx *ctx, const uint64_t *src, uint64_t *const *peer_dst, int world, size_t words, uint64_t *const *peer_flags, size_t flag_slot,
                        uint64_t epoch, void *stream_) {
    B2_REQUIRE(ctx && peer_dst && peer_flags && world >= 1 && world <= SHARD_MAX_WORLD && (words == 0 || src) && (words & 1) == 0, "bad arguments");
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    PeerTab d, f;
    for (int r = 0; r < SHARD_MAX_WORLD; r++) { d.p[r] = r < world ? (u64 *)peer_dst[r] : nullptr; f.p[r] = r < world ? (u64 *)peer_flags[r] : nullptr; }
    for (int r = 0; r < world; r++) B2_REQUIRE(f.p[r] && (words == 0 || (d.p[r] && (((uintptr_t)d.p[r]) & 15) == 0)), "null or misaligned peer pointer");
    Scratch sc(ctx, stream);
    b2_status st = ensure_sync(sc.s, 0, stream);
    if (st != B2_OK) return st;
    // the push counter lives behind the transform queue's counters (zero between launches)
    u32 *counter = sc->d_sync + sc->sync_limbs + 3;
    const u64 chunks = words / 2;
    const unsigned grid = (unsigned)std::max<u64>(1, std::min<u64>((chunks + 255) / 256, (u64)ctx->sm_count * 2));
    shard_push_kernel<<<grid, 256, 0, stream>>>((const ulonglong2 *)src, d, f, (u32)world, chunks, (u32)flag_slot, epoch, counter);
    count_launches(1);
    B2_CUDA(cudaGetLastError());
    return B2_OK;
}
b2_status b2_shard_wait(b2_ctx *ctx, const uint64_t *flags, const int *slots, int n_slots, uint64_t epoch, void *stream_) {
    B2_REQUIRE(ctx && flags && (n_slots == 0 || slots) && n_slots >= 0 && n_slots <= SHARD_MAX_WORLD, "bad arguments");
    if (n_slots == 0) return B2_OK;
    B2_CUDA(cudaSetDevice(ctx->device));
    SlotTab t;
    for (int i = 0; i < SHARD_MAX_WORLD; i++) t.s[i] = i < n_slots ? slots[i] : 0;
    shard_wait_kernel<<<1, 32, 0, (cudaStream_t)stream_>>>((const u64 *)flags, t, (u32)n_slots, epoch);
    count_launches(1);
    B2_CUDA(cudaGetLastError());
    return B2_OK;
}

static inline size_t owned_ct_limbs(size_t decomp, size_t first, size_t count) {
    return first >= decomp ? 0 : std::min(first + count, decomp) - first;
}

b2_status b2_keyswitch_shard_prepare(b2_ctx *ctx, const uint64_t *target_local, size_t decomp, int is_ntt, size_t first, size_t count,
                                     uint64_t *send, void *stream_) {
    B2_REQUIRE(ctx && decomp >= 1 && decomp + 1 <= ctx->K && count >= 1 && first + count <= decomp + 1, "bad arguments");
    const size_t n = ctx->n, nct = owned_ct_limbs(decomp, first, count);
    if (nct == 0) return B2_OK;
    B2_REQUIRE(target_local && send && (((uintptr_t)target_local | (uintptr_t)send) & 15) == 0, "operands must be non-null and 16-byte aligned");
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!is_ntt) { B2_CUDA(cudaMemcpyAsync(send, target_local, nct * n * 8, cudaMemcpyDeviceToDevice, stream)); return B2_OK; }
    std::vector<LimbGroup> groups;
    for (size_t jl = 0; jl < nct; jl++) groups.push_back(LimbGroup{(u32)(first + jl), 1u, (u64)(jl * n), 0});
    PtrTab coef;
    for (int b = 0; b < KS_MAXB; b++) coef.p[b] = b == 0 ? (const u64 *)target_local : nullptr;
    Scratch scl(ctx, stream);
    b2_scratch *sc = scl.s;
    return run_ntt_dir<false>(ctx, sc, groups, (u64 *)send, TargetLoad{coef, (u32)n}, PlainStore{(u64 *)send}, stream);
}

// workspace of the sharded stages: prod [2][count][N] | mid [2][nct][N] | tc [decomp][N] | temp [count][decomp][N];
// prod survives from `products` to `finish` (no other call on this context in between)
b2_status b2_keyswitch_shard_products(b2_ctx *ctx, const uint64_t *gathered, int world, size_t pad, size_t decomp, size_t first, size_t count,
                                      const b2_kswitch_key *shard, uint64_t *u, void *stream_) {
    B2_REQUIRE(ctx && gathered && shard && world >= 0, "null argument");
    B2_REQUIRE(decomp >= 1 && decomp + 1 <= ctx->K && count >= 1 && first + count <= decomp + 1, "bad modulus range");
    B2_REQUIRE(shard->rows == count && shard->first == first && shard->decomp == decomp, "key shard does not match the modulus range");
    B2_REQUIRE((((uintptr_t)gathered | (uintptr_t)u) & 15) == 0, "operands must be 16-byte aligned");
    const bool owns_special = first + count == decomp + 1;
    B2_REQUIRE(!owns_special || u, "the owner of the special prime needs the u buffer");
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t n = ctx->n, K = ctx->K, nct = owned_ct_limbs(decomp, first, count);
    Scratch scl(ctx, stream);
    b2_scratch *sc = scl.s;
    b2_status st = ensure_ws(sc, stream, (2 * count + 2 * nct + decomp + count * decomp) * n * 8);
    if (st != B2_OK) return st;
    u64 *prod = (u64 *)sc->ws, *tc = prod + (2 * count + 2 * nct) * n, *temp = tc + decomp * n;
    // world == 0: `gathered` already IS the coefficient-form target [decomp][N] (P2P path: the peers pushed their limbs in place)
    if (world == 0) tc = (u64 *)gathered;
    // the all-gather delivered [world][pad][N]: rank r's owned ciphertext limbs first_r .. ; compact to tc[decomp][N]
    else {
        const size_t R = decomp + 1, base = R / (size_t)world, rem = R % (size_t)world;
        B2_REQUIRE(pad >= std::min<size_t>(base + (rem ? 1 : 0), decomp), "pad is smaller than a rank's share of the ciphertext limbs");
        if (base == 0) {    // more ranks than moduli: rank r < rem owns modulus r
            for (size_t j = 0; j < decomp; j++)
                B2_CUDA(cudaMemcpyAsync(tc + j * n, gathered + j * pad * n, n * 8, cudaMemcpyDeviceToDevice, stream));
        } else {
            dim3 grid((unsigned)std::min<size_t>(64, (n / 2 + 255) / 256), (unsigned)decomp);
            ks_compact_gathered_kernel<<<grid, 256, 0, stream>>>((const u64 *)gathered, tc, (u32)n, (u32)pad, (u32)base, (u32)rem);
            count_launches(1);
            B2_CUDA(cudaGetLastError());
        }
    }
    PtrTab coef;
    for (int b = 0; b < KS_MAXB; b++) coef.p[b] = b == 0 ? tc : nullptr;
    {
        std::vector<LimbGroup> groups;
        for (size_t il = 0; il < count; il++) {
            const size_t i = first + il;
            groups.push_back(LimbGroup{(u32)(i == decomp ? K - 1 : i), (u32)decomp, (u64)(il * decomp * n), (u64)n});
        }
        ModUpLoad ld{coef, ctx->d_mods, (u32)decomp, (u32)K, (u32)n, (u32)first, ctx->logn > 11 ? 1u : 0u};
        st = run_ntt_dir<true>(ctx, sc, groups, temp, ld, PlainStore{temp}, stream);
        if (st != B2_OK) return st;
    }
    {
        dim3 grid(1, (unsigned)((n / 2 + 255) / 256), (unsigned)count);
        ks_inner_product_kernel<<<grid, 256, 0, stream>>>(temp, (const u64 *const *)shard->d_ptrs, prod, ctx->d_mods, (u32)decomp, (u32)K, (u32)n,
                                                          (u32)first, (u32)count, (u32)count, 0u);
        count_launches(1);
        B2_CUDA(cudaGetLastError());
    }
    if (owns_special) {
        std::vector<LimbGroup> groups{LimbGroup{(u32)(K - 1), 2u, 0, (u64)n}};
        SpecialLoad ld{prod, (u32)count, (u32)(count - 1), (u32)n};
        st = run_ntt_dir<false>(ctx, sc, groups, (u64 *)u, ld, PlainStore{(u64 *)u}, stream);
        if (st != B2_OK) return st;
    }
    return B2_OK;
}

b2_status b2_keyswitch_shard_finish(b2_ctx *ctx, const uint64_t *u, size_t decomp, int is_ntt, size_t first, size_t count,
                                    uint64_t *dest_local, int assign, void *stream_) {
    B2_REQUIRE(ctx && u, "null argument");
    B2_REQUIRE(decomp >= 1 && decomp + 1 <= ctx->K && count >= 1 && first + count <= decomp + 1, "bad modulus range");
    B2_REQUIRE(assign >= 0 && assign <= 2, "unknown assign method");
    B2_REQUIRE(ctx->scheme != B2_SCHEME_BGV || (is_ntt && ctx->t > 1), "BGV ciphertexts are NTT-form and need a plain modulus");
    const size_t n = ctx->n, K = ctx->K, nct = owned_ct_limbs(decomp, first, count);
    if (nct == 0) return B2_OK;
    B2_REQUIRE(dest_local && (((uintptr_t)dest_local | (uintptr_t)u) & 15) == 0, "operands must be non-null and 16-byte aligned");
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    Scratch scl(ctx, stream);
    b2_scratch *sc = scl.s;
    B2_REQUIRE(sc->ws && sc->ws_bytes >= (2 * count + 2 * nct) * n * 8, "b2_keyswitch_shard_products must run first (same stream, same thread)");
    u64 *prod = (u64 *)sc->ws, *mid = prod + 2 * count * n;
    std::vector<LimbGroup> groups;
    for (size_t jl = 0; jl < nct; jl++) groups.push_back(LimbGroup{(u32)(first + jl), 2u, (u64)(jl * n), (u64)(nct * n)});
    PtrTabW dp;
    for (int b = 0; b < KS_MAXB; b++) dp.p[b] = b == 0 ? (u64 *)dest_local : nullptr;
    return mod_down(ctx, sc, groups, mid, (const u64 *)u, prod, dp, decomp, assign, is_ntt, first, count, nct, stream);
}

}  // extern "C"
