// b2_rescale.cu — rounding division by the last modulus (R1-R4 in SURVEY.md §8a).  Replaces
// RNSTool::divide_and_round_q_last (BFV mod_switch_to_next, coefficient form, src/utils/rns_tool.cu:374-497),
// RNSTool::divide_and_round_q_last_ntt (CKKS rescale, :499-760), RNSTool::mod_t_and_divide_q_last_ntt (BGV
// mod_switch_to_next, :1540-1833) and kernel_mod_switch_drop_to (CKKS mod_switch_to_next,
// src/evaluator_modswitch.cu:152-171).
//
// NTT-form variants: only the LAST limb is inverse-transformed (the reference's device branch inverse-
// transforms all nl limbs, rns_tool.cu:674-676), and the correction term / final scaling are fused into
// the loader / last-pass store of ONE forward transform over the remaining nl-1 limbs.
#include "b2_ntt_launch.cuh"
#include "b2_fused.cuh"

using namespace b2;

namespace b2 {
// BFV coefficient form: out_i = (c_i - ((c_L + h) mod q_L mod q_i - h mod q_i)) q_L^-1   (rns_tool.cu:374-403)
__global__ void __launch_bounds__(256) divide_round_kernel(const u64 *__restrict__ in, u64 *__restrict__ out, const ModInfo *__restrict__ mods,
                                                             const Tw *__restrict__ inv, u32 nl, u32 n, u32 count) {
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (u64)count * n) return;
    const u32 p = (u32)(idx / n), x = (u32)(idx % n);
    const u64 qL = mods[nl - 1].p, half = qL >> 1;
    u64 v = in[((u64)p * nl + nl - 1) * n + x] + half;
    v = v >= qL ? v - qL : v;
    for (u32 i = 0; i + 1 < nl; i++) {
        const ModInfo &Q = mods[i];
        const u64 q = Q.p;
        const u64 t = submod(barrett64(v, q, Q.ratio), barrett64(half, q, Q.ratio), q);
        const u64 s = in[((u64)p * nl + i) * n + x];
        out[((u64)p * (nl - 1) + i) * n + x] = shoup(submod(s, t, q), inv[i], q);
    }
}
__global__ void __launch_bounds__(256) drop_limb_kernel(const ulonglong2 *__restrict__ in, ulonglong2 *__restrict__ out, u32 nl, u32 n2, u32 count) {
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;     // over count * (nl-1) * n/2
    const u64 per = (u64)(nl - 1) * n2;
    if (idx >= (u64)count * per) return;
    const u64 p = idx / per, r = idx % per;
    out[idx] = in[p * (u64)nl * n2 + r];
}
}  // namespace b2

namespace b2 {
// the work of b2_mod_switch_to_next with the scratch record already held by the caller; the workspace is used from ws_off on
// (callers that keep their own data in front of it: b2_encrypt.cu)
b2_status mod_switch_locked(b2_ctx *ctx, b2_scratch *sc, const u64 *in, u64 *out, size_t count, size_t nl, int kind, size_t ws_off, cudaStream_t stream) {
    const size_t n = ctx->n, K = ctx->K, last = nl - 1;
    if (kind == B2_MODSWITCH_DROP) {
        B2_REQUIRE(n >= 2, "degree too small");
        const u64 total = (u64)count * (nl - 1) * (n / 2);
        drop_limb_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>((const ulonglong2 *)in, (ulonglong2 *)out, (u32)nl, (u32)(n / 2), (u32)count);
        count_launches(1);
        B2_CUDA(cudaGetLastError());
        return B2_OK;
    }
    if (kind == B2_MODSWITCH_BFV) {
        const u64 total = (u64)count * n;
        divide_round_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, out, ctx->d_mods, ctx->d_inv + last * K, (u32)nl, (u32)n, (u32)count);
        count_launches(1);
        B2_CUDA(cudaGetLastError());
        return B2_OK;
    }
    B2_REQUIRE(kind != B2_MODSWITCH_BGV || ctx->t > 1, "BGV needs a plain modulus");
    // workspace: u [count][N] | mid [count][nl-1][N]
    b2_status st = ensure_ws(sc, stream, ws_off + (count * n + count * (nl - 1) * n) * 8);
    if (st != B2_OK) return st;
    u64 *u = (u64 *)((char *)sc->ws + ws_off), *mid = u + count * n;
    {   // u_p = INTT_{q_last}(in[p][last])
        std::vector<LimbGroup> groups{LimbGroup{(u32)last, (u32)count, 0, (u64)n}};
        st = run_ntt_dir<false>(ctx, sc, groups, u, LastLimbLoad{in, (u32)nl, (u32)n}, PlainStore{u}, stream);
        if (st != B2_OK) return st;
    }
    {   // NTT of the correction term, final store scales by q_last^-1
        std::vector<LimbGroup> groups;
        for (size_t i = 0; i + 1 < nl; i++) groups.push_back(LimbGroup{(u32)i, (u32)count, (u64)(i * n), (u64)((nl - 1) * n)});
        RescaleStore sst{in, out, ctx->d_mods, ctx->d_inv + last * K, (u32)nl, (u32)n};
        if (kind == B2_MODSWITCH_BGV) {
            ModDownTerm<true> T{u, ctx->d_mods, (u32)last, (u32)n, ctx->t, ctx->t_ratio, ctx->t_r0, ctx->t_r1, ctx->h_inv_mod_t[last]};
            st = run_ntt_dir<true>(ctx, sc, groups, mid, ModDownLoad<true>{T}, sst, stream);
        } else {
            ModDownTerm<false> T{u, ctx->d_mods, (u32)last, (u32)n, ctx->t, ctx->t_ratio, ctx->t_r0, ctx->t_r1, 0};
            st = run_ntt_dir<true>(ctx, sc, groups, mid, ModDownLoad<false>{T}, sst, stream);
        }
        if (st != B2_OK) return st;
    }
    return B2_OK;
}
}  // namespace b2

extern "C" b2_status b2_mod_switch_to_next(b2_ctx *ctx, const uint64_t *in, uint64_t *out, size_t count, size_t nl, int kind,
                                           void *stream_) {
    B2_REQUIRE(ctx && in && out, "null argument");
    B2_REQUIRE((((uintptr_t)in | (uintptr_t)out) & 15) == 0, "operands must be 16-byte aligned");
    B2_REQUIRE(nl >= 2 && nl <= ctx->K, "need at least two limbs to drop one");
    B2_REQUIRE(kind >= 0 && kind <= 3, "unknown kind");
    B2_REQUIRE(in != out, "in-place mod switch is not supported: the output has one limb less per polynomial (the reference writes a new ciphertext too)");
    if (count == 0) return B2_OK;
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    Scratch scl(ctx, stream);
    return mod_switch_locked(ctx, scl.s, (const u64 *)in, (u64 *)out, count, nl, kind, 0, stream);
}
