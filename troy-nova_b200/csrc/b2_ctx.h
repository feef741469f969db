// b2_ctx.h — internal definition of the opaque b2_ctx handle.
#pragma once
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/b2rns.h"
#include "b2_modarith.cuh"

struct b2_ctx {
    int device = 0;
    unsigned logn = 0;
    size_t n = 0, K = 0;
    uint64_t t = 0;
    int scheme = 0;
    std::vector<uint64_t> q;      // key-level moduli (host copy)
    std::vector<uint64_t> psi;    // minimal primitive 2N-th roots
    b2::ModInfo *d_mods = nullptr;   // [K] on device
    b2::Tw *d_tw = nullptr;          // [K][N] forward twiddles
    b2::Tw *d_itw = nullptr;         // [K][N] inverse twiddles
    double *d_twd = nullptr, *d_itwd = nullptr;   // same tables as doubles (FP64-pipe path)
    void *d_n5 = nullptr;                         // per-row / per-limb tables of the TMA kernel (N = 2^16 only), see ModInfo::n5_*
    std::vector<b2::ModInfo> h_mods;
    b2::Tw *d_inv = nullptr;         // [K][K]: d_inv[a*K+b] = q_a^-1 mod q_b as a Shoup pair (a != b)
    uint64_t *d_inv_mod_t = nullptr; // [K]: q_a^-1 mod t (BFV/BGV), 0 when t == 0
    uint64_t t_ratio = 0, t_r0 = 0, t_r1 = 0;   // floor(2^64 / t), floor(2^128 / t)
    std::vector<uint64_t> h_inv_mod_t;
    int sm_count = 148;
    // Everything above is immutable after b2_ctx_create.  The MUTABLE per-call state (workspace arena, persistent-NTT work
    // queue, host staging) lives in b2_scratch records, one per (stream, host thread) that ever used the context: calls on
    // one stream are ordered by the stream, calls on different streams / threads never share a record, so a context is
    // re-entrant the way the reference is per host thread on cudaStreamPerThread (SURVEY §8b "Threading / streams").
    std::mutex mu;                                   // guards `scratch` and the lazy BEHZ tables
    std::map<std::pair<uintptr_t, size_t>, struct b2_scratch *> scratch;
    // BEHZ (BFV multiply) auxiliary base, built lazily by the first b2_bfv_multiply: ModInfo slots K .. 2K+1 hold
    // B_0 .. B_K (K+1 primes) and m_sk; per-level constant blobs keyed by the number of ciphertext limbs
    struct b2_behz *behz = nullptr;
    // Decryptor constants (gamma, base conversions q -> {t, gamma}), built lazily per level by b2_decrypt.cu
    struct b2_decrypt_state *dec = nullptr;
    // stream-ordered arena of the multi-launch entry points (b2_pack_rlwes): a private pool with an unbounded release threshold, so that the
    // memory stays mapped between calls (the device's default pool hands it back to the OS at every synchronisation)
    cudaMemPool_t mempool = nullptr;
    // BatchEncoder: a one-modulus context over the plain modulus t (its negacyclic NTT tables) and the slot -> coefficient index map, built lazily
    b2_ctx *plain_ctx = nullptr;
    uint32_t *d_slot_map = nullptr;
};

// Mutable state of one (stream, host thread).  `mu` is held by an entry point for the whole time it enqueues its launches,
// so two host threads that share one explicit stream cannot interleave their launch sequences on the shared workspace.
struct b2_scratch {
    std::mutex mu;
    void *ws = nullptr; size_t ws_bytes = 0;          // workspace arena (grown on demand; never shrinks)
    void *stage = nullptr; size_t stage_bytes = 0;    // device staging of the *_host entry points
    cudaStream_t s_in = nullptr, s_out = nullptr;     // copy streams of the pipelined *_host entry points
    std::vector<cudaEvent_t> events;
    // persistent-NTT work queue + per-limb completion counters (zero between launches; the kernel resets them)
    unsigned *d_sync = nullptr; size_t sync_limbs = 0;
    // device copy of the pointer tables of the *_ptrs / batched entry points (one upload per call, like the reference's
    // construct_batch, src/utils/box_batch.h:222-247, but into a buffer that is reused)
    void *ptrtab = nullptr; size_t ptrtab_bytes = 0;
    // mixed transforms at N = 2^16: the integer-pipe limbs run on a side stream next to the TMA kernel (b2_ntt5_launch.cuh), with
    // their own work queue
    cudaStream_t s_side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    unsigned *d_sync2 = nullptr; size_t sync2_limbs = 0;
};

namespace b2 {
// RAII: the scratch record of (ctx, stream, calling thread), locked for the lifetime of the object
struct Scratch {
    b2_scratch *s;
    std::unique_lock<std::mutex> lock;
    Scratch(b2_ctx *ctx, cudaStream_t stream);
    b2_scratch *operator->() const { return s; }
};
// number of kernels this library launched in this process (bench.py reports it as gpu_launches)
void count_launches(unsigned n);
void set_error(const char *fmt, ...);
b2_status cuda_fail(cudaError_t e, const char *what);
b2_status ensure_ws(b2_scratch *sc, cudaStream_t stream, size_t bytes);
b2_status ensure_stage(b2_scratch *sc, cudaStream_t stream, size_t bytes);
b2_status ensure_copy_pipeline(b2_scratch *sc, size_t n_events);
b2_status ensure_sync(b2_scratch *sc, size_t limbs, cudaStream_t stream);
b2_status ensure_ptrtab(b2_scratch *sc, cudaStream_t stream, size_t bytes);
b2_status ensure_side(b2_scratch *sc, size_t limbs, cudaStream_t stream);
void scratch_release_all(b2_ctx *ctx);
int build_modulus_tables(uint64_t p, unsigned logn, Tw *T, Tw *IT, double *TD, double *ITD, ModInfo &M, uint64_t &psi_out);
void behz_release(b2_ctx *ctx);
void decrypt_release(b2_ctx *ctx);
b2_status arena_pool(b2_ctx *ctx, cudaMemPool_t *out);
b2_status mod_switch_locked(b2_ctx *ctx, b2_scratch *sc, const u64 *in, u64 *out, size_t count, size_t nl, int kind, size_t ws_off, cudaStream_t stream);
#define B2_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return b2::cuda_fail(e__, #call); } while (0)
#define B2_REQUIRE(cond, msg) do { if (!(cond)) { b2::set_error("%s: %s", __func__, msg); return B2_ERR_INVALID_ARG; } } while (0)

// launchers (b2_ntt.cu)
struct LimbGroup;
b2_status ntt_launch(b2_ctx *ctx, bool forward, const uint64_t *in, uint64_t *out, size_t batch, size_t pcount,
                     size_t ncomp, int mode, size_t table_base, size_t decomp, cudaStream_t stream);
}
