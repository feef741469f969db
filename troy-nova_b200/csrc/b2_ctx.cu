// b2_ctx.cu — context creation: per-modulus constants and twiddle tables, built on the host once and
// made device-resident.  Replaces the device copies the reference makes in
// HeContext::to_device_inplace (src/he_context.cu:134-145) of NTTTables (src/utils/ntt.cu:14-76).
#include "b2_ctx.h"
#include "b2_host_math.h"
#include <type_traits>

using namespace b2;

namespace b2 {
// position of twiddle jj of the layer with 2^m twiddles inside a 256-entry table of the TMA kernel (layout: b2_ntt5.cuh)
static size_t n5_table_slot(int m, size_t jj, bool f64) {
    if (m <= 4) return ((size_t)1 << m) + jj;
    const int per = 1 << (m - 4);                       // twiddles per thread y: 2 (m = 5), 4 (m = 6) or 8 (m = 7)
    const size_t y = jj / per, i = jj % per;
    if (f64) return ((size_t)1 << m) + (i >> 1) * 32 + 2 * y + (i & 1);      // 8-byte entries: two per 16-byte chunk
    return ((size_t)1 << m) + 16 * i + y;                                    // 16-byte entries: one per chunk
}
// Per-modulus constants and twiddle tables (host side).  T / IT: N Shoup pairs each; TD / ITD: the same values as
// doubles (may be null).  Returns 0, 1 (not a prime below 2^61) or 2 (not 1 mod 2N).
// root_powers[bitrev(k)] = psi^k ; inv_root_powers[bitrev(k-1)+1] = psi^-k  (src/utils/ntt.cu:36-58)
int build_modulus_tables(uint64_t p, unsigned logn, Tw *T, Tw *IT, double *TD, double *ITD, ModInfo &M, uint64_t &psi_out) {
    const size_t n = (size_t)1 << logn;
    if (p < 2 || (p >> 61) != 0 || !host::is_prime(p)) return 1;
    const uint64_t psi = host::min_primitive_root(2 * n, p);
    if (!psi) return 2;
    psi_out = psi;
    const uint64_t ipsi = host::invmod(psi, p);
    uint64_t pw = psi, ipw = ipsi;
    T[0].x = 1; T[0].y = host::shoup_quotient(1, p); IT[n - 1] = T[0];
    for (size_t k = 1; k < n; k++) {
        size_t a = host::bitrev((uint32_t)k, logn), b = host::bitrev((uint32_t)(k - 1), logn) + 1;
        T[a].x = pw; T[a].y = host::shoup_quotient(pw, p);
        IT[b - 1].x = ipw; IT[b - 1].y = host::shoup_quotient(ipw, p);   // device table = inv_root_powers shifted down by one (16-byte aligned runs)
        pw = host::mulmod(pw, psi, p); ipw = host::mulmod(ipw, ipsi, p);
    }
    memset(&M, 0, sizeof M);
    M.p = p;
    host::u128 ratio = (~(host::u128)0) / p;
    M.r0 = (uint64_t)ratio; M.r1 = (uint64_t)(ratio >> 64); M.ratio = M.r1;
    const uint64_t inv_n = host::invmod(n % p, p);
    M.inv_n.x = inv_n; M.inv_n.y = host::shoup_quotient(inv_n, p);
    const uint64_t last_w = n > 1 ? IT[n - 2].x : 1;    // twiddle of the last inverse layer
    const uint64_t inw = host::mulmod(inv_n, last_w, p);
    M.inv_n_w.x = inw; M.inv_n_w.y = host::shoup_quotient(inw, p);
    const unsigned bits = host::bit_count(p);
    M.fwd_lazy = ((host::u128)(4 + 2 * logn) * p) < ((host::u128)1 << 64) ? 1u : 0u;
    M.inv_lazy = (bits + logn + 2 <= 64) ? 1u : 0u;
    // FP64-pipe path: exact double-double arithmetic needs |x| w / p < 2^51 with |x| < 2^10 p
    M.bits = bits;
    M.use_f64 = (bits <= 41 && TD && ITD) ? 1u : 0u;
    M.pd = (double)p; M.pinv = 1.0 / (double)p;
    M.inv_n_d = (double)inv_n; M.inv_n_w_d = (double)inw;
    if (TD && ITD) for (size_t k = 0; k < n; k++) { TD[k] = (double)T[k].x; ITD[k] = (double)IT[k].x; }
    return 0;
}
}  // namespace b2

extern "C" void b2_ctx_destroy(b2_ctx *c);
extern "C" b2_status b2_ctx_create(int device, const b2_params *prm, b2_ctx **out) {
    B2_REQUIRE(prm && out && prm->moduli, "null argument");
    B2_REQUIRE(prm->log_n >= 1 && prm->log_n <= 17, "log_n must be in [1, 17]");
    B2_REQUIRE(prm->n_moduli >= 1 && prm->n_moduli <= 64, "n_moduli must be in [1, 64]");
    B2_REQUIRE(prm->scheme >= B2_SCHEME_BFV && prm->scheme <= B2_SCHEME_BGV, "unknown scheme");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("b2_ctx_create: no CUDA device (libb2rns has no CPU fallback)");
        return B2_ERR_NO_DEVICE;
    }
    B2_REQUIRE(device >= 0 && device < ndev, "device index out of range");
    B2_CUDA(cudaSetDevice(device));
    int sms = 0;
    B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));

    // owned by a guard until the last upload has succeeded: every early return below releases the context and its tables
    struct Guard { b2_ctx *c; ~Guard() { if (c) b2_ctx_destroy(c); } } guard{new b2_ctx()};
    b2_ctx *c = guard.c;
    c->sm_count = sms > 0 ? sms : 148;
    c->device = device; c->logn = prm->log_n; c->n = (size_t)1 << prm->log_n; c->K = prm->n_moduli;
    c->t = prm->plain_modulus; c->scheme = (int)prm->scheme;
    const size_t n = c->n, K = c->K; const unsigned logn = c->logn;
    c->q.assign(prm->moduli, prm->moduli + K);
    c->psi.resize(K);
    std::vector<Tw> tw(K * n), itw(K * n);
    std::vector<double> twd(K * n), itwd(K * n);
    c->h_mods.reserve(2 * K + 2);      // room for the lazily added BEHZ records: appending them never reallocates
    c->h_mods.resize(K);
    for (size_t i = 0; i < K; i++) {
        const int rc = build_modulus_tables(c->q[i], logn, &tw[i * n], &itw[i * n], &twd[i * n], &itwd[i * n], c->h_mods[i], c->psi[i]);
        if (rc == 1) { set_error("b2_ctx_create: modulus %zu is not a prime below 2^61", i); return B2_ERR_INVALID_ARG; }
        if (rc == 2) { set_error("b2_ctx_create: modulus %zu is not 1 mod 2N", i); return B2_ERR_INVALID_ARG; }
    }
    cudaError_t ce;
    if ((ce = cudaMalloc(&c->d_tw, K * n * sizeof(Tw))) != cudaSuccess || (ce = cudaMalloc(&c->d_itw, K * n * sizeof(Tw))) != cudaSuccess ||
        (ce = cudaMalloc(&c->d_twd, K * n * sizeof(double))) != cudaSuccess || (ce = cudaMalloc(&c->d_itwd, K * n * sizeof(double))) != cudaSuccess ||
        (ce = cudaMalloc(&c->d_mods, (2 * K + 2) * sizeof(ModInfo))) != cudaSuccess) {
        return cuda_fail(ce, "cudaMalloc(tables)");
    }
    for (size_t i = 0; i < K; i++) { c->h_mods[i].tw = c->d_tw + i * n; c->h_mods[i].itw = c->d_itw + i * n;
        c->h_mods[i].twd = c->d_twd + i * n; c->h_mods[i].itwd = c->d_itwd + i * n; }
    B2_CUDA(cudaMemcpy(c->d_tw, tw.data(), K * n * sizeof(Tw), cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy(c->d_itw, itw.data(), K * n * sizeof(Tw), cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy(c->d_twd, twd.data(), K * n * sizeof(double), cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy(c->d_itwd, itwd.data(), K * n * sizeof(double), cudaMemcpyHostToDevice));
    // tables of the TMA kernel (b2_ntt5.cuh): the same twiddles, regrouped per 256-point sub-transform — doubles for the
    // FP64-path moduli, Shoup pairs for the integer-path moduli
    if (logn == 16) {
        const size_t ent = 2 * 256 + 2 * 65536 + 2 * 4096;          // fcol | icol | frow | irow | frow_u | irow_u  (entries)
        size_t bytes = 0;
        for (size_t i = 0; i < K; i++) bytes += ent * (c->h_mods[i].use_f64 ? sizeof(double) : sizeof(Tw));
        std::vector<unsigned char> blob(bytes);
        B2_CUDA(cudaMalloc(&c->d_n5, bytes));
        size_t off = 0;
        auto fill = [&](auto *base, const auto *ft, const auto *it, bool f64, ModInfo &M) {
            typedef typename std::remove_pointer<decltype(base)>::type E;
            E *fcol = base, *icol = fcol + 256, *frow = icol + 256, *irow = frow + 65536, *fu = irow + 65536, *iu = fu + 4096;
            for (int m = 0; m < 8; m++)
                for (size_t jj = 0; jj < ((size_t)1 << m); jj++) {
                    const size_t s = n5_table_slot(m, jj, f64);
                    fcol[s] = ft[((size_t)1 << m) + jj];
                    icol[s] = it[n - (n >> (8 + (7 - m))) + jj];          // layer 8 + u, u = 7 - m, 128 >> u = 2^m entries
                    for (size_t r = 0; r < 256; r++) {
                        frow[r * 256 + s] = ft[((size_t)1 << (8 + m)) + (r << m) + jj];
                        irow[r * 256 + s] = it[n - (n >> (7 - m)) + (r << m) + jj];   // layer l = 7 - m, 128 >> l = 2^m entries per row
                    }
                }
            for (size_t r = 0; r < 256; r++)
                for (size_t h = 0; h < 16; h++) { fu[r * 16 + h] = frow[r * 256 + h]; iu[r * 16 + h] = irow[r * 256 + h]; }
            const E *d = reinterpret_cast<const E *>(reinterpret_cast<const unsigned char *>(c->d_n5) + off);
            M.n5_fcol = d; M.n5_icol = d + 256; M.n5_frow = d + 512; M.n5_irow = d + 512 + 65536;
            M.n5_frow_u = d + 512 + 2 * 65536; M.n5_irow_u = d + 512 + 2 * 65536 + 4096;
        };
        for (size_t i = 0; i < K; i++) {
            if (c->h_mods[i].use_f64) { fill(reinterpret_cast<double *>(blob.data() + off), &twd[i * n], &itwd[i * n], true, c->h_mods[i]); off += ent * sizeof(double); }
            else { fill(reinterpret_cast<Tw *>(blob.data() + off), &tw[i * n], &itw[i * n], false, c->h_mods[i]); off += ent * sizeof(Tw); }
        }
        B2_CUDA(cudaMemcpy(c->d_n5, blob.data(), bytes, cudaMemcpyHostToDevice));
    }
    B2_CUDA(cudaMemcpy(c->d_mods, c->h_mods.data(), K * sizeof(ModInfo), cudaMemcpyHostToDevice));
    // cross-modulus inverses used by mod-down / rescale (reference: RNSTool::inv_q_last_mod_q, inv_q_last_mod_t,
    // src/utils/rns_tool.cu:214-231 — one RNSTool per level there, one K x K table here)
    {
        std::vector<Tw> inv(K * K); std::vector<uint64_t> invt(K, 0);
        for (size_t a = 0; a < K; a++) {
            for (size_t b = 0; b < K; b++) {
                if (a == b) { inv[a * K + b].x = 0; inv[a * K + b].y = 0; continue; }
                uint64_t v = host::invmod(c->q[a] % c->q[b], c->q[b]);
                inv[a * K + b].x = v; inv[a * K + b].y = host::shoup_quotient(v, c->q[b]);
            }
            if (c->t > 1) invt[a] = host::invmod(c->q[a] % c->t, c->t);
        }
        if (c->t > 1) { host::u128 r = (~(host::u128)0) / c->t; if ((~(host::u128)0) % c->t == c->t - 1) r += 1;   /* floor(2^128 / t), exact for powers of two too */
            c->t_r0 = (uint64_t)r; c->t_r1 = (uint64_t)(r >> 64); c->t_ratio = c->t_r1; }
        c->h_inv_mod_t = invt;
        B2_CUDA(cudaMalloc(&c->d_inv, K * K * sizeof(Tw)));
        B2_CUDA(cudaMalloc(&c->d_inv_mod_t, K * sizeof(uint64_t)));
        B2_CUDA(cudaMemcpy(c->d_inv, inv.data(), K * K * sizeof(Tw), cudaMemcpyHostToDevice));
        B2_CUDA(cudaMemcpy(c->d_inv_mod_t, invt.data(), K * sizeof(uint64_t), cudaMemcpyHostToDevice));
    }
    guard.c = nullptr;
    *out = c;
    return B2_OK;
}

extern "C" void b2_ctx_destroy(b2_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaFree(c->d_tw); cudaFree(c->d_itw); cudaFree(c->d_twd); cudaFree(c->d_itwd); cudaFree(c->d_n5); cudaFree(c->d_mods); b2::behz_release(c); b2::decrypt_release(c); if (c->mempool) cudaMemPoolDestroy(c->mempool); if (c->plain_ctx) b2_ctx_destroy(c->plain_ctx); cudaFree(c->d_slot_map); cudaFree(c->d_inv); cudaFree(c->d_inv_mod_t); b2::scratch_release_all(c);
    delete c;
}
