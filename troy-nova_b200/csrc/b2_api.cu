// b2_api.cu — error handling, device memory helpers and misc C-ABI entry points of libb2rns.so.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include "b2_ctx.h"
#include "b2_host_math.h"

#include <atomic>
#include <functional>
#include <thread>
namespace b2 {
static std::atomic<unsigned long long> g_launches{0};
void count_launches(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}
b2_status cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return B2_ERR_NO_DEVICE;
    if (e == cudaErrorMemoryAllocation) return B2_ERR_OUT_OF_MEMORY;
    return B2_ERR_CUDA;
}
// Scratch records are keyed by the stream handle and — for the two handles that name a different stream in every host
// thread (cudaStreamPerThread) or no particular one — by the calling thread as well.
Scratch::Scratch(b2_ctx *ctx, cudaStream_t stream) : s(nullptr) {
    const uintptr_t h = (uintptr_t)stream;
    const size_t tid = (stream == cudaStreamPerThread) ? std::hash<std::thread::id>()(std::this_thread::get_id()) : 0;
    {
        std::lock_guard<std::mutex> g(ctx->mu);
        b2_scratch *&slot = ctx->scratch[std::make_pair(h, tid)];
        if (!slot) slot = new b2_scratch();
        s = slot;
    }
    lock = std::unique_lock<std::mutex>(s->mu);
}
void scratch_release_all(b2_ctx *ctx) {
    for (auto &kv : ctx->scratch) {
        b2_scratch *sc = kv.second;
        cudaFree(sc->ws); cudaFree(sc->stage); cudaFree(sc->d_sync); cudaFree(sc->ptrtab); cudaFree(sc->d_sync2);
        if (sc->s_side) cudaStreamDestroy(sc->s_side);
        if (sc->ev_fork) cudaEventDestroy(sc->ev_fork);
        if (sc->ev_join) cudaEventDestroy(sc->ev_join);
        for (cudaEvent_t e : sc->events) cudaEventDestroy(e);
        if (sc->s_in) cudaStreamDestroy(sc->s_in);
        if (sc->s_out) cudaStreamDestroy(sc->s_out);
        delete sc;
    }
    ctx->scratch.clear();
}
// Growth happens outside the steady state (first use of a bigger shape on this stream).  Only this stream can have work
// in flight on the old buffer (the record is private to it), so synchronising the stream is enough before freeing.
b2_status ensure_ws(b2_scratch *sc, cudaStream_t stream, size_t bytes) {
    if (sc->ws_bytes >= bytes) return B2_OK;
    if (sc->ws) { B2_CUDA(cudaStreamSynchronize(stream)); B2_CUDA(cudaFree(sc->ws)); sc->ws = nullptr; sc->ws_bytes = 0; }
    size_t want = bytes + bytes / 8;
    B2_CUDA(cudaMalloc(&sc->ws, want));
    sc->ws_bytes = want;
    return B2_OK;
}
b2_status ensure_stage(b2_scratch *sc, cudaStream_t stream, size_t bytes) {
    if (sc->stage_bytes >= bytes) return B2_OK;
    if (sc->stage) { B2_CUDA(cudaStreamSynchronize(stream)); B2_CUDA(cudaFree(sc->stage)); sc->stage = nullptr; sc->stage_bytes = 0; }
    B2_CUDA(cudaMalloc(&sc->stage, bytes));
    sc->stage_bytes = bytes;
    return B2_OK;
}
b2_status ensure_side(b2_scratch *sc, size_t limbs, cudaStream_t stream) {
    if (!sc->s_side) {
        int lo = 0, hi = 0;
        B2_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        B2_CUDA(cudaStreamCreateWithPriority(&sc->s_side, cudaStreamNonBlocking, hi));      // its few CTAs should be placed first
        B2_CUDA(cudaEventCreateWithFlags(&sc->ev_fork, cudaEventDisableTiming));
        B2_CUDA(cudaEventCreateWithFlags(&sc->ev_join, cudaEventDisableTiming));
    }
    if (sc->d_sync2 && sc->sync2_limbs >= limbs) return B2_OK;
    if (sc->d_sync2) { B2_CUDA(cudaStreamSynchronize(stream)); B2_CUDA(cudaStreamSynchronize(sc->s_side)); B2_CUDA(cudaFree(sc->d_sync2)); sc->d_sync2 = nullptr; sc->sync2_limbs = 0; }
    const size_t want = std::max<size_t>(limbs * 2, 1024);
    B2_CUDA(cudaMalloc((void **)&sc->d_sync2, (want + 2) * sizeof(unsigned)));
    B2_CUDA(cudaMemsetAsync(sc->d_sync2, 0, (want + 2) * sizeof(unsigned), stream));
    sc->sync2_limbs = want;
    return B2_OK;
}
b2_status ensure_ptrtab(b2_scratch *sc, cudaStream_t stream, size_t bytes) {
    if (sc->ptrtab_bytes >= bytes) return B2_OK;
    if (sc->ptrtab) { B2_CUDA(cudaStreamSynchronize(stream)); B2_CUDA(cudaFree(sc->ptrtab)); sc->ptrtab = nullptr; sc->ptrtab_bytes = 0; }
    const size_t want = std::max<size_t>(bytes * 2, 1 << 16);
    B2_CUDA(cudaMalloc(&sc->ptrtab, want));
    sc->ptrtab_bytes = want;
    return B2_OK;
}
b2_status ensure_copy_pipeline(b2_scratch *sc, size_t n_events) {
    if (!sc->s_in) B2_CUDA(cudaStreamCreateWithFlags(&sc->s_in, cudaStreamNonBlocking));
    if (!sc->s_out) B2_CUDA(cudaStreamCreateWithFlags(&sc->s_out, cudaStreamNonBlocking));
    while (sc->events.size() < n_events) {
        cudaEvent_t e;
        B2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        sc->events.push_back(e);
    }
    return B2_OK;
}
b2_status ensure_sync(b2_scratch *sc, size_t limbs, cudaStream_t stream) {
    if (sc->d_sync && sc->sync_limbs >= limbs) return B2_OK;
    if (sc->d_sync) { B2_CUDA(cudaStreamSynchronize(stream)); B2_CUDA(cudaFree(sc->d_sync)); sc->d_sync = nullptr; sc->sync_limbs = 0; }
    const size_t want = std::max<size_t>(limbs * 2, 4096);
    // + 2 queue words in front, + 2 spare words behind (the last one is the block counter of b2_shard_push)
    B2_CUDA(cudaMalloc((void **)&sc->d_sync, (want + 4) * sizeof(unsigned)));
    B2_CUDA(cudaMemsetAsync(sc->d_sync, 0, (want + 4) * sizeof(unsigned), stream));
    sc->sync_limbs = want;
    return B2_OK;
}
}  // namespace b2

extern "C" {

const char *b2_last_error(void) { return b2::g_err; }
const char *b2_version(void) { return "b2rns 0.1 (sm_100a)"; }
unsigned long long b2_launch_count(void) { return b2::g_launches.load(); }
int b2_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

b2_status b2_coeff_modulus_create(size_t n, const uint32_t *bits, size_t count, uint64_t *out) {
    B2_REQUIRE(n >= 2 && (n & (n - 1)) == 0 && bits && out && count >= 1 && count <= 64, "bad arguments");
    // CoeffModulus::create: per distinct bit size take as many primes as it occurs (descending from
    // 2^bits), then hand them out from the smallest upwards in order of appearance.
    std::vector<uint64_t> lists[64]; size_t used[64] = {0}, mult[64] = {0};
    for (size_t i = 0; i < count; i++) { B2_REQUIRE(bits[i] >= 2 && bits[i] <= 61, "bit size out of range"); mult[bits[i]]++; }
    for (unsigned b = 0; b < 64; b++)
        if (mult[b] && !b2::host::primes_below(2 * (uint64_t)n, b, mult[b], lists[b])) {
            b2::set_error("b2_coeff_modulus_create: not enough %u-bit primes = 1 mod %zu", b, 2 * n);
            return B2_ERR_INVALID_ARG;
        }
    for (size_t i = 0; i < count; i++) { unsigned b = bits[i]; out[i] = lists[b][mult[b] - 1 - used[b]++]; }
    return B2_OK;
}

b2_status b2_ctx_root(const b2_ctx *ctx, size_t i, uint64_t *psi) {
    B2_REQUIRE(ctx && psi && i < ctx->K, "bad arguments");
    *psi = ctx->psi[i];
    return B2_OK;
}

b2_status b2_malloc(b2_ctx *ctx, size_t bytes, void **dptr) {
    B2_REQUIRE(ctx && dptr, "null argument");
    B2_CUDA(cudaSetDevice(ctx->device));
    B2_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
    return B2_OK;
}
b2_status b2_free(b2_ctx *ctx, void *dptr) {
    B2_REQUIRE(ctx, "null ctx");
    B2_CUDA(cudaSetDevice(ctx->device));
    B2_CUDA(cudaFree(dptr));
    return B2_OK;
}
b2_status b2_malloc_host(size_t bytes, void **hptr) { B2_REQUIRE(hptr, "null argument"); B2_CUDA(cudaMallocHost(hptr, bytes ? bytes : 1)); return B2_OK; }
b2_status b2_free_host(void *hptr) { B2_CUDA(cudaFreeHost(hptr)); return B2_OK; }
b2_status b2_memcpy_h2d(b2_ctx *ctx, void *dst, const void *src, size_t bytes, void *stream) {
    B2_REQUIRE(ctx, "null ctx");
    B2_CUDA(cudaSetDevice(ctx->device));
    B2_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return B2_OK;
}
b2_status b2_memcpy_d2h(b2_ctx *ctx, void *dst, const void *src, size_t bytes, void *stream) {
    B2_REQUIRE(ctx, "null ctx");
    B2_CUDA(cudaSetDevice(ctx->device));
    B2_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return B2_OK;
}
b2_status b2_memcpy_d2d(b2_ctx *ctx, void *dst, const void *src, size_t bytes, void *stream) {
    B2_REQUIRE(ctx, "null ctx");
    B2_CUDA(cudaSetDevice(ctx->device));
    B2_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return B2_OK;
}
b2_status b2_stream_sync(b2_ctx *ctx, void *stream) {
    B2_REQUIRE(ctx, "null ctx");
    B2_CUDA(cudaSetDevice(ctx->device));
    B2_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return B2_OK;
}

}  // extern "C"
