// b2_ntt.cu — C-ABI entry points b2_ntt_forward / b2_ntt_inverse (+ _ptrs, _host).  Reference seam:
// utils::ntt_transfer_to_rev / ntt_transfer_from_rev and their _batched forms (src/utils/ntt.h:132-144,
// src/fgk/ntt_grouped.cu:258-344,597-685).
#include "b2_ntt5_launch.cuh"

namespace b2 {

static b2_status check_ntt_args(b2_ctx *ctx, size_t batch, size_t pcount, size_t ncomp, int mode, size_t table_base, size_t decomp) {
    B2_REQUIRE(ctx, "null ctx");
    B2_REQUIRE(mode >= 0 && mode <= 2, "unknown table mode");
    B2_REQUIRE(batch * pcount * ncomp > 0, "empty operand");   // callers filter empties; reference would loop zero times
    if (mode == B2_TABLES_COMPONENTWISE) B2_REQUIRE(table_base + ncomp <= ctx->K, "table_base + ncomp exceeds modulus count");
    if (mode == B2_TABLES_KS_SET_PRODUCTS) B2_REQUIRE(decomp < ctx->K && pcount <= decomp + 1, "bad decomp for KS_SET_PRODUCTS");
    if (mode == B2_TABLES_KS_SKIP_FINALS) B2_REQUIRE(decomp < ctx->K && ncomp <= decomp + 1, "bad decomp for KS_SKIP_FINALS");
    return B2_OK;
}

b2_status ntt_launch(b2_ctx *ctx, bool forward, const uint64_t *in, uint64_t *out, size_t batch, size_t pcount,
                     size_t ncomp, int mode, size_t table_base, size_t decomp, cudaStream_t stream) {
    if (batch * pcount * ncomp == 0) return B2_OK;
    b2_status st = check_ntt_args(ctx, batch, pcount, ncomp, mode, table_base, decomp);
    if (st != B2_OK) return st;
    B2_REQUIRE(in && out, "null data pointer");
    B2_REQUIRE((((uintptr_t)in | (uintptr_t)out) & 15) == 0, "operands must be 16-byte aligned");
    B2_CUDA(cudaSetDevice(ctx->device));
    const size_t n = ctx->n, K = ctx->K;
    std::vector<LimbGroup> groups;
    if (mode == B2_TABLES_KS_SET_PRODUCTS) {
        for (size_t b = 0; b < batch; b++)
            for (size_t k = 0; k < pcount; k++)
                groups.push_back(LimbGroup{(u32)(k == decomp ? K - 1 : k), (u32)ncomp, (u64)((b * pcount + k) * ncomp * n), (u64)n});
    } else {
        for (size_t j = 0; j < ncomp; j++) {
            u32 table = mode == B2_TABLES_COMPONENTWISE ? (u32)(table_base + j) : (u32)(j == decomp ? K - 1 : j);
            groups.push_back(LimbGroup{table, (u32)(batch * pcount), (u64)(j * n), (u64)(ncomp * n)});
        }
    }
    // the work queue is in group order within every batch item: the slow integer-pipe moduli go first so that the queue does not end on one
    // (the groups carry their own offsets, so their order is free)
    static const bool int_first = [] { const char *e = getenv("B2_NTT_INT_FIRST"); return !(e && e[0] == '0'); }();
    if (int_first) std::stable_partition(groups.begin(), groups.end(), [&](const LimbGroup &g) { return !ctx->h_mods[g.table].use_f64; });
    Scratch sc(ctx, stream);
    // throughput regime at N = 2^16: the FP64-pipe moduli take the TMA kernel (b2_ntt5.cuh); few-limb calls stay on the
    // cluster kernel (latency regime)
    const size_t limbs = batch * pcount * ncomp;
    if (ctx->logn == N5_LOGN && limbs > cluster_max_limbs() && ((((uintptr_t)in | (uintptr_t)out) & 127) == 0) && n5_enabled())
        return forward ? run_ntt_split<true>(ctx, sc.s, groups, (const u64 *)in, (u64 *)out, limbs, stream)
                       : run_ntt_split<false>(ctx, sc.s, groups, (const u64 *)in, (u64 *)out, limbs, stream);
    return run_ntt(ctx, sc.s, forward, groups, (u64 *)out, PlainLoad{(const u64 *)in}, PlainStore{(u64 *)out}, stream);
}

}  // namespace b2

using namespace b2;

extern "C" {

b2_status b2_ntt_forward(b2_ctx *ctx, const uint64_t *in, uint64_t *out, size_t batch, size_t pcount, size_t ncomp,
                         int mode, size_t table_base, size_t decomp, void *stream) {
    return ntt_launch(ctx, true, in, out, batch, pcount, ncomp, mode, table_base, decomp, (cudaStream_t)stream);
}
b2_status b2_ntt_inverse(b2_ctx *ctx, const uint64_t *in, uint64_t *out, size_t batch, size_t pcount, size_t ncomp,
                         int mode, size_t table_base, size_t decomp, void *stream) {
    return ntt_launch(ctx, false, in, out, batch, pcount, ncomp, mode, table_base, decomp, (cudaStream_t)stream);
}

// Scattered batch (the reference's ntt_transfer_*_batched over a SliceVec, src/fgk/ntt_grouped.cu:297-344,642-685: one
// launch with blockIdx.y = batch item).  Here too ONE launch: the pointer tables are uploaded once into the stream's scratch
// record and the batch is a dimension of the persistent kernel's work queue (limb c of group j = (item c / pcount, poly
// c % pcount)); the intermediate between the two passes lives in each item's own output, exactly as for contiguous operands.
static b2_status ntt_ptrs(b2_ctx *ctx, bool fwd, const uint64_t *const *in, uint64_t *const *out, size_t batch, size_t pcount,
                          size_t ncomp, int mode, size_t table_base, size_t decomp, void *stream_) {
    B2_REQUIRE(ctx && (batch == 0 || (in && out)), "null argument");
    if (batch * pcount * ncomp == 0) return B2_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (batch == 1 || mode == B2_TABLES_KS_SET_PRODUCTS) {      // KS_SET_PRODUCTS groups by polynomial: not a batched shape the reference uses
        for (size_t b = 0; b < batch; b++) {
            b2_status st = ntt_launch(ctx, fwd, in[b], out[b], 1, pcount, ncomp, mode, table_base, decomp, stream);
            if (st != B2_OK) return st;
        }
        return B2_OK;
    }
    b2_status st = check_ntt_args(ctx, batch, pcount, ncomp, mode, table_base, decomp);
    if (st != B2_OK) return st;
    for (size_t b = 0; b < batch; b++) {
        B2_REQUIRE(in[b] && out[b], "null operand in batch");
        B2_REQUIRE((((uintptr_t)in[b] | (uintptr_t)out[b]) & 15) == 0, "operands must be 16-byte aligned");
    }
    B2_CUDA(cudaSetDevice(ctx->device));
    const size_t n = ctx->n, K = ctx->K;
    Scratch sc(ctx, stream);
    st = ensure_ptrtab(sc.s, stream, 2 * batch * 8);
    if (st != B2_OK) return st;
    std::vector<uint64_t> blob(2 * batch);
    for (size_t b = 0; b < batch; b++) { blob[b] = (uint64_t)(uintptr_t)in[b]; blob[batch + b] = (uint64_t)(uintptr_t)out[b]; }
    B2_CUDA(cudaMemcpyAsync(sc->ptrtab, blob.data(), blob.size() * 8, cudaMemcpyHostToDevice, stream));
    const u64 *const *d_in = (const u64 *const *)sc->ptrtab;
    const u64 *const *d_out = d_in + batch;
    std::vector<LimbGroup> groups;
    for (size_t j = 0; j < ncomp; j++) {
        const u32 table = mode == B2_TABLES_COMPONENTWISE ? (u32)(table_base + j) : (u32)(j == decomp ? K - 1 : j);
        groups.push_back(LimbGroup{table, (u32)(batch * pcount), (u64)(j * n), 0});      // stride unused: offsets come from the table
    }
    NttScatter sca; sca.out_tab = d_out; sca.pc = (u32)pcount; sca.pstride = (u64)ncomp * n;
    u64 *mid = (u64 *)out[0];
    return run_ntt(ctx, sc.s, fwd, groups, mid, PtrTabLoad{d_in, (u32)pcount, (u32)n, (u64)ncomp * n}, PlainStore{mid}, stream, sca);
}
b2_status b2_ntt_forward_ptrs(b2_ctx *ctx, const uint64_t *const *in, uint64_t *const *out, size_t batch, size_t pcount,
                              size_t ncomp, int mode, size_t table_base, size_t decomp, void *stream) {
    return ntt_ptrs(ctx, true, in, out, batch, pcount, ncomp, mode, table_base, decomp, stream);
}
b2_status b2_ntt_inverse_ptrs(b2_ctx *ctx, const uint64_t *const *in, uint64_t *const *out, size_t batch, size_t pcount,
                              size_t ncomp, int mode, size_t table_base, size_t decomp, void *stream) {
    return ntt_ptrs(ctx, false, in, out, batch, pcount, ncomp, mode, table_base, decomp, stream);
}

// Host-buffer transform, pipelined: the operand is cut into chunks of whole limbs (~16 MiB: 1 / 2 / 4 / 8 / 16 / 32 MiB chunks measured 51 / 59 / 70 / 78 / 81 / 79 k limbs/s on B200, r01h); chunk c+1 travels host->device
// on a copy stream while chunk c is transformed on `stream` and chunk c-1 travels device->host on a second copy stream, so
// both PCIe directions and the SMs are busy at the same time.  Synchronised on return (the caller's buffers are plain
// host memory; pinned memory gives full-rate DMA).
static b2_status ntt_host(b2_ctx *ctx, bool fwd, const uint64_t *in, uint64_t *out, size_t batch, size_t pcount, size_t ncomp,
                          int mode, size_t table_base, size_t decomp, void *stream) {
    B2_REQUIRE(ctx && in && out, "null argument");
    const size_t n = ctx->n, K = ctx->K, limbs = batch * pcount * ncomp, bytes = limbs * n * 8;
    if (bytes == 0) return B2_OK;
    b2_status st = check_ntt_args(ctx, batch, pcount, ncomp, mode, table_base, decomp);
    if (st != B2_OK) return st;
    B2_CUDA(cudaSetDevice(ctx->device));
    Scratch sc(ctx, (cudaStream_t)stream);
    st = ensure_stage(sc.s, (cudaStream_t)stream, bytes);
    if (st != B2_OK) return st;
    // chunk size: large enough that a copy runs at link rate, small enough that filling / draining the three-stage pipeline
    // stays a small part of the call (B2_HOST_CHUNK_MB: developer override for tuning)
    static const size_t chunk_mb = [] { const char *e = getenv("B2_HOST_CHUNK_MB"); return e && atoi(e) > 0 ? (size_t)atoi(e) : (size_t)16; }();
    const size_t chunk = std::max<size_t>(1, std::min<size_t>((size_t)NTT_MAX_GROUPS, (chunk_mb << 20) / (n * 8)));
    const size_t nchunks = (limbs + chunk - 1) / chunk;
    st = ensure_copy_pipeline(sc.s, 2 * nchunks + 2);
    if (st != B2_OK) return st;
    cudaStream_t s = (cudaStream_t)stream, s_in = sc->s_in, s_out = sc->s_out;
    cudaEvent_t *ev = sc->events.data();
    u64 *dev = (u64 *)sc->stage;
    B2_CUDA(cudaEventRecord(ev[0], s));                 // copies start after the caller's earlier work on `stream`
    B2_CUDA(cudaStreamWaitEvent(s_in, ev[0], 0));
    std::vector<LimbGroup> groups;
    for (size_t c = 0; c < nchunks; c++) {
        const size_t l0 = c * chunk, l1 = std::min(limbs, l0 + chunk), cb = (l1 - l0) * n * 8;
        B2_CUDA(cudaMemcpyAsync(dev + l0 * n, in + l0 * n, cb, cudaMemcpyHostToDevice, s_in));
        B2_CUDA(cudaEventRecord(ev[2 + 2 * c], s_in));
        B2_CUDA(cudaStreamWaitEvent(s, ev[2 + 2 * c], 0));
        groups.clear();
        for (size_t l = l0; l < l1; l++) {      // one single-limb group per limb; table as NTTTableIndexer::get picks it
            const size_t j = l % ncomp, k = (l / ncomp) % pcount;
            const u32 table = mode == B2_TABLES_COMPONENTWISE ? (u32)(table_base + j)
                            : mode == B2_TABLES_KS_SET_PRODUCTS ? (u32)(k == decomp ? K - 1 : k) : (u32)(j == decomp ? K - 1 : j);
            groups.push_back(LimbGroup{table, 1u, (u64)(l * n), 0});
        }
        st = run_ntt(ctx, sc.s, fwd, groups, dev, PlainLoad{dev}, PlainStore{dev}, s);
        if (st != B2_OK) return st;
        B2_CUDA(cudaEventRecord(ev[3 + 2 * c], s));
        B2_CUDA(cudaStreamWaitEvent(s_out, ev[3 + 2 * c], 0));
        B2_CUDA(cudaMemcpyAsync(out + l0 * n, dev + l0 * n, cb, cudaMemcpyDeviceToHost, s_out));
    }
    B2_CUDA(cudaEventRecord(ev[1], s_out));
    B2_CUDA(cudaStreamWaitEvent(s, ev[1], 0));
    B2_CUDA(cudaStreamSynchronize(s));
    return B2_OK;
}
b2_status b2_ntt_forward_host(b2_ctx *ctx, const uint64_t *in, uint64_t *out, size_t batch, size_t pcount, size_t ncomp,
                              int mode, size_t table_base, size_t decomp, void *stream) {
    return ntt_host(ctx, true, in, out, batch, pcount, ncomp, mode, table_base, decomp, stream);
}
b2_status b2_ntt_inverse_host(b2_ctx *ctx, const uint64_t *in, uint64_t *out, size_t batch, size_t pcount, size_t ncomp,
                              int mode, size_t table_base, size_t decomp, void *stream) {
    return ntt_host(ctx, false, in, out, batch, pcount, ncomp, mode, table_base, decomp, stream);
}

}  // extern "C"
