// b2_ntt5_launch.cuh — host side of the TMA transform kernel (b2_ntt5.cuh): tensor maps + launch.
#pragma once
#include "b2_ntt_launch.cuh"
#include "b2_ntt5.cuh"

namespace b2 {

// cuTensorMapEncodeTiled through the runtime's driver entry point: libb2rns.so keeps no link-time dependency on libcuda
// (the library must load on a box without a driver: tests/test_abi.py)
typedef CUresult (*n5_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline n5_encode_fn n5_encoder() {
    static const n5_encode_fn fn = [] {
        void *f = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return (n5_encode_fn)f;
    }();
    return fn;
}
enum N5View { N5_NATURAL, N5_PERMUTED, N5_ROWS };
// views of `limbs` consecutive N = 2^16 limbs at `base` (each a 256 x 256 matrix of 8-byte words), all SWIZZLE_128B, 32 KiB boxes:
//   NATURAL   (col, row, limb)             box 16 x 256 x 1        shared-memory line index = row
//   PERMUTED  (col, row_hi, row_lo, limb)  box 16 x 16 x 16 x 1    shared-memory line index = 16 row_lo + row_hi
//   ROWS      (word, line)                 box 16 x 256            16 whole rows = 256 lines of 128 bytes
static inline bool n5_make_map(CUtensorMap *m, N5View view, const void *base, size_t limbs) {
    n5_encode_fn enc = n5_encoder();
    if (!enc) return false;
    cuuint64_t dims[4], strides[3];
    cuuint32_t box[4], es[4] = {1, 1, 1, 1};
    cuuint32_t rank;
    if (view == N5_NATURAL) {
        rank = 3; dims[0] = 256; dims[1] = 256; dims[2] = limbs; strides[0] = 2048; strides[1] = 524288; box[0] = 16; box[1] = 256; box[2] = 1;
    } else if (view == N5_PERMUTED) {
        rank = 4; dims[0] = 256; dims[1] = 16; dims[2] = 16; dims[3] = limbs; strides[0] = 32768; strides[1] = 2048; strides[2] = 524288;
        box[0] = 16; box[1] = 16; box[2] = 16; box[3] = 1;
    } else {
        rank = 2; dims[0] = 16; dims[1] = 4096 * limbs; strides[0] = 128; box[0] = 16; box[1] = 256;
    }
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, rank, const_cast<void *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static inline bool n5_enabled() {
    static const bool on = [] { const char *e = getenv("B2_NTT5"); return !(e && e[0] == '0'); }();      // developer switch for A/B timing
    return on && n5_encoder() != nullptr;
}

// All `groups` must be moduli of an N = 2^16 context (either arithmetic path), with element offsets that are multiples of N; `dst` holds the
// intermediate and the result, `src` the operand (`src_limbs` limbs; the prologue functor names the source limb of each limb).
template <bool FWD, class Pro>
static b2_status run_ntt5(b2_ctx *ctx, b2_scratch *sc, const std::vector<LimbGroup> &groups, const u64 *src, size_t src_limbs, u64 *dst, Pro pro, cudaStream_t stream) {
    if (groups.empty() || groups[0].count == 0) return B2_OK;
    B2_REQUIRE(ctx->logn == N5_LOGN, "internal: the TMA transform serves N = 2^16 only");
    B2_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 127) == 0, "internal: TMA operands must be 128-byte aligned");
    const size_t n = ctx->n;
    size_t dst_limbs = 0;
    for (const LimbGroup &g : groups) {
        B2_REQUIRE(g.count == groups[0].count, "internal: limb groups of one transform must have equal counts");
        B2_REQUIRE(g.base % n == 0 && g.stride % n == 0 && ctx->h_mods[g.table].n5_frow != nullptr, "internal: group not eligible for the TMA transform");
        dst_limbs = std::max<size_t>(dst_limbs, (g.base + (size_t)(g.count - 1) * g.stride) / n + 1);
    }
    Ntt5Maps MP;
    bool ok;
    if (FWD) ok = n5_make_map(&MP.first, N5_NATURAL, src, src_limbs) && n5_make_map(&MP.mid_st, N5_PERMUTED, dst, dst_limbs) && n5_make_map(&MP.mid_ld, N5_ROWS, dst, dst_limbs);
    else ok = n5_make_map(&MP.first, N5_ROWS, src, src_limbs) && n5_make_map(&MP.mid_st, N5_ROWS, dst, dst_limbs) && n5_make_map(&MP.mid_ld, N5_PERMUTED, dst, dst_limbs);
    ok = ok && n5_make_map(&MP.last, FWD ? N5_ROWS : N5_NATURAL, dst, dst_limbs);
    if (!ok) { set_error("cuTensorMapEncodeTiled failed"); return B2_ERR_CUDA; }
    static bool configured[B2_MAX_DEVICES] = {};
    if (!configured[ctx->device & (B2_MAX_DEVICES - 1)]) {
        B2_CUDA(cudaFuncSetAttribute(ntt5_kernel<FWD, Pro>, cudaFuncAttributeMaxDynamicSharedMemorySize, N5_SMEM_BYTES));
        configured[ctx->device & (B2_MAX_DEVICES - 1)] = true;
    }
    NttJob P;
    P.mods = ctx->d_mods; P.mid = dst; P.logn = N5_LOGN; P.count = groups[0].count;
    P.out_tab = nullptr; P.tab_pc = 0; P.tab_pstride = 0;
    static const u32 dbg = [] { const char *e = getenv("B2_NTT5_DBG"); return e ? (u32)atoi(e) : 0u; }();
    P.dbg = dbg;
    P.tile_shift = N5_LOGN - NTT_TB;
    for (size_t g0 = 0; g0 < groups.size(); g0 += NTT_MAX_GROUPS) {
        const size_t ng = std::min(groups.size() - g0, (size_t)NTT_MAX_GROUPS);
        P.g_base = (int)g0; P.n_groups = (int)ng;
        for (size_t g = 0; g < ng; g++) {
            P.groups[g] = groups[g0 + g];
            const ModInfo &M = ctx->h_mods[groups[g0 + g].table];
            P.groups[g].table = tag_table(groups[g0 + g].table, M.use_f64 ? PATH_F64 : (M.fwd_lazy ? PATH_INT_LAZY : PATH_INT),
                                          M.use_f64 ? PATH_F64 : (M.inv_lazy ? PATH_INT_LAZY : PATH_INT));
        }
        P.n_limbs = (u32)(ng * P.count);
        // claimed-but-unfinished items: 3 CTAs x 2 stages + 1 claim ahead per CTA, ~1350 items = 42 limbs; a lag of 96 limbs keeps a
        // second-pass item from being claimed while first-pass tiles of its limb are still in that window (48 MiB of intermediate
        // in flight, inside the 126 MB L2)
        P.lag = std::min<u32>(96u, P.n_limbs);
        P.total_items = (P.n_limbs << P.tile_shift) * 2;
        b2_status s = ensure_sync(sc, P.n_limbs, stream);
        if (s != B2_OK) return s;
        P.sync = sc->d_sync;
        const unsigned grid = std::min<unsigned>(P.total_items, 3u * (unsigned)ctx->sm_count);
        ntt5_kernel<FWD, Pro><<<grid, N5_THREADS, N5_SMEM_BYTES, stream>>>(P, MP, pro);
        count_launches(1);
        B2_CUDA(cudaGetLastError());
    }
    return B2_OK;
}

// Split a transform's groups by arithmetic path: FP64-pipe moduli -> TMA kernel, the 60-bit moduli -> integer-pipe kernel
// (b2_ntt.cuh).  Plain operands only (src + off / dst + off addressing).
template <bool FWD>
static b2_status run_ntt_split(b2_ctx *ctx, b2_scratch *sc, const std::vector<LimbGroup> &groups, const u64 *src, u64 *dst, size_t total_limbs, cudaStream_t stream) {
    // developer switch for the mixed-path experiments (profiles/r02_ntt5_experiments.txt): 0 everything on the TMA kernel,
    // 1 / 2 integer-path limbs on the cp.async kernel next to it (capped at SMs/2 / SMs CTAs), 3 the same, one after the other
    static const int mix = [] { const char *e = getenv("B2_NTT5_MIX"); return e ? atoi(e) : -1; }();
    if (mix < 0) {
        // Default, by measurement on B200 (profiles/r02_ntt5_experiments.txt): launches whose limbs are all FP64-path take the TMA
        // kernel (380 limbs 168 us, 3800 limbs 1484 us vs 1542 us); a launch that also has integer-path (60-bit) limbs is still
        // served fastest by the cp.async kernel, whose CTAs overlap the two pipes (420 limbs: 208 us vs 218 us all-TMA, 224-231 us split)
        bool pure = true;
        for (const LimbGroup &g : groups) pure = pure && ctx->h_mods[g.table].use_f64 && ctx->h_mods[g.table].n5_frow && g.base % ctx->n == 0 && g.stride % ctx->n == 0;
        if (pure) return run_ntt5<FWD>(ctx, sc, groups, src, total_limbs, dst, N5Identity{}, stream);
        return run_ntt_dir<FWD>(ctx, sc, groups, dst, PlainLoad{src}, PlainStore{dst}, stream);
    }
    std::vector<LimbGroup> f64, rest;
    for (const LimbGroup &g : groups) {
        const bool ok = ctx->h_mods[g.table].n5_frow && g.base % ctx->n == 0 && g.stride % ctx->n == 0 && (mix == 0 || ctx->h_mods[g.table].use_f64);
        (ok ? f64 : rest).push_back(g);
    }
    b2_status st = B2_OK;
    if (rest.empty()) return run_ntt5<FWD>(ctx, sc, f64, src, total_limbs, dst, N5Identity{}, stream);
    if (f64.empty()) return run_ntt_dir<FWD>(ctx, sc, rest, dst, PlainLoad{src}, PlainStore{dst}, stream);
    // Both kinds: the few integer-pipe limbs run NEXT TO the TMA kernel instead of after it — a capped grid of the persistent
    // kernel on a high-priority side stream (one CTA on every second SM: 20 k registers + 64 KiB, which still leaves room for two
    // CTAs of the TMA kernel there and three elsewhere).  The two kernels use different pipes (integer vs FP64) and touch
    // disjoint limbs; each has its own work queue.
    size_t rest_limbs = 0;
    for (const LimbGroup &g : rest) rest_limbs += g.count;
    st = ensure_side(sc, rest_limbs, stream);
    if (st != B2_OK) return st;
    B2_CUDA(cudaEventRecord(sc->ev_fork, stream));
    B2_CUDA(cudaStreamWaitEvent(sc->s_side, sc->ev_fork, 0));
    if (mix == 3) {
        st = run_ntt_dir<FWD>(ctx, sc, rest, dst, PlainLoad{src}, PlainStore{dst}, stream);
        return st != B2_OK ? st : run_ntt5<FWD>(ctx, sc, f64, src, total_limbs, dst, N5Identity{}, stream);
    }
    NttScatter side; side.max_ctas = (unsigned)std::max(1, mix == 2 ? ctx->sm_count : ctx->sm_count / 2); side.sync = sc->d_sync2;
    st = run_ntt_dir<FWD>(ctx, sc, rest, dst, PlainLoad{src}, PlainStore{dst}, sc->s_side, side);
    if (st != B2_OK) return st;
    B2_CUDA(cudaEventRecord(sc->ev_join, sc->s_side));
    st = run_ntt5<FWD>(ctx, sc, f64, src, total_limbs, dst, N5Identity{}, stream);
    B2_CUDA(cudaStreamWaitEvent(stream, sc->ev_join, 0));
    return st;
}

}  // namespace b2
