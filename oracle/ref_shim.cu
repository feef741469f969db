/*
 * ref_shim.cu — C-ABI shim over the UNMODIFIED reference library (oracle/_ref/libtroy_ref.so).
 *
 * TEST INFRASTRUCTURE ONLY.  This file is ours; it contains no reference code, it only CALLS the
 * reference's public C++ API (HeContext / KeyGenerator / Encryptor / Evaluator / utils::ntt_*)
 * so that python (ctypes) can (a) pin oracle/rns_oracle.c against the real implementation,
 * (b) generate the golden vectors under tests/golden/, (c) time the reference host branch as the
 * CPU baseline (bench.py --impl reference, cpu_baseline.kind == "reference") and (d) time the
 * reference *device* branch on the B200 as the "recompiled generic kernels" GPU baseline.
 *
 * All arrays are HOST pointers in the reference layout [poly][limb][N] (src/ciphertext.h:211-252).
 * device != 0 makes the context device-resident and routes every op through the reference's
 * device branch (data is copied in and out around the op).
 */
#include "troy.h"
#include <chrono>
#include <cstring>
#include <memory>
#include <vector>

using namespace troy;
using troy::utils::ConstSlice;
using troy::utils::Slice;
using troy::utils::DynamicArray;

struct RefShim {
    HeContextPointer ctx;
    std::unique_ptr<KeyGenerator> keygen;
    std::unique_ptr<Evaluator> eval;
    std::unique_ptr<Encryptor> enc;
    std::unique_ptr<Decryptor> dec;
    RelinKeys rlk; bool has_rlk = false;
    GaloisKeys glk; bool has_glk = false;
    SchemeType scheme; size_t n; bool device; double ckks_scale = 1099511627776.0;
    std::string err;
};

static ContextDataPointer level_with(RefShim *h, size_t nl) {
    auto kd = h->ctx->key_context_data().value();
    if (kd->parms().coeff_modulus().size() == nl) return kd;
    auto cd = h->ctx->first_context_data().value();
    while (cd->parms().coeff_modulus().size() != nl) {
        auto nx = cd->next_context_data();
        if (!nx.has_value()) throw std::invalid_argument("no level with that many limbs");
        cd = nx.value();
    }
    return cd;
}

static Ciphertext make_ct(RefShim *h, const uint64_t *data, size_t pc, size_t nl, bool ntt_form) {
    auto cd = level_with(h, nl);
    size_t total = pc * nl * h->n;
    DynamicArray<uint64_t> arr = DynamicArray<uint64_t>::create_uninitialized(total, false);
    std::memcpy(arr.raw_pointer(), data, total * 8);
    Ciphertext ct = Ciphertext::from_members(pc, nl, h->n, cd->parms_id(), h->scheme == SchemeType::CKKS ? h->ckks_scale : 1.0,
                                             ntt_form, 1, 0, std::move(arr));
    if (h->device) ct.to_device_inplace();
    return ct;
}

static void export_ct(const Ciphertext &ct, uint64_t *out) {
    Ciphertext hct = ct.clone(); if (hct.on_device()) hct.to_host_inplace();
    std::memcpy(out, hct.data().raw_pointer(), hct.data().size() * 8);
}

static bool default_ntt_form(RefShim *h) { return h->scheme != SchemeType::BFV; } /* SURVEY §9.J */

#define RS_TRY(...) try { __VA_ARGS__; return 0; } catch (const std::exception &e) { h->err = e.what(); return -1; }

extern "C" {

void *rs_create(int scheme, size_t n, const unsigned *bits, size_t nbits, unsigned t_bits, uint64_t seed, int device) {
    RefShim *h = new RefShim();
    try {
        h->scheme = (SchemeType)scheme; h->n = n; h->device = device != 0;
        EncryptionParameters parms(h->scheme);
        parms.set_poly_modulus_degree(n);
        std::vector<size_t> b(bits, bits + nbits);
        parms.set_coeff_modulus(CoeffModulus::create(n, b));
        if (h->scheme != SchemeType::CKKS) parms.set_plain_modulus(PlainModulus::batching(n, t_bits));
        h->ctx = HeContext::create(parms, true, SecurityLevel::Nil, seed);
        if (h->device) h->ctx->to_device_inplace();
        h->keygen.reset(new KeyGenerator(h->ctx));
        h->eval.reset(new Evaluator(h->ctx));
        h->enc.reset(new Encryptor(h->ctx));
        h->enc->set_public_key(h->keygen->create_public_key(false));
        h->enc->set_secret_key(h->keygen->secret_key());
        h->dec.reset(new Decryptor(h->ctx, h->keygen->secret_key()));
    } catch (const std::exception &e) { h->err = e.what(); }
    return h;
}
const char *rs_error(void *hh) { return ((RefShim *)hh)->err.c_str(); }
void rs_destroy(void *hh) { delete (RefShim *)hh; }
int rs_device_count() { return (int)utils::device_count(); }

size_t rs_n_key_moduli(void *hh) { return ((RefShim *)hh)->ctx->key_context_data().value()->parms().coeff_modulus().size(); }
void rs_key_moduli(void *hh, uint64_t *out) {
    RefShim *h = (RefShim *)hh;
    auto cd = h->ctx->key_context_data().value();
    auto parms = cd->parms().on_device() ? cd->parms().to_host() : cd->parms().clone();
    auto m = parms.coeff_modulus();
    for (size_t i = 0; i < m.size(); i++) out[i] = m[i].value();
}
uint64_t rs_plain_modulus(void *hh) {
    RefShim *h = (RefShim *)hh;
    if (h->scheme == SchemeType::CKKS) return 0;
    auto cd = h->ctx->key_context_data().value();
    auto parms = cd->parms().on_device() ? cd->parms().to_host() : cd->parms().clone();
    return parms.plain_modulus_host().value();
}
/* psi_i and the first `count` root_powers operands/quotients of key-level table i (host contexts only) */
int rs_ntt_table(void *hh, size_t i, uint64_t *root, uint64_t *root_powers, uint64_t *inv_root_powers, size_t count) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({
        if (h->device) throw std::invalid_argument("rs_ntt_table: host context only");
        auto tabs = h->ctx->key_context_data().value()->small_ntt_tables();
        const utils::NTTTables &t = tabs[i];
        *root = t.root();
        for (size_t k = 0; k < count; k++) {
            root_powers[2 * k] = t.root_powers()[k].operand; root_powers[2 * k + 1] = t.root_powers()[k].quotient;
            inv_root_powers[2 * k] = t.inv_root_powers()[k].operand; inv_root_powers[2 * k + 1] = t.inv_root_powers()[k].quotient;
        }
    })
}

int rs_gen_relin(void *hh) { RefShim *h = (RefShim *)hh; RS_TRY({ h->rlk = h->keygen->create_relin_keys(false, 2); h->has_rlk = true; }) }
int rs_gen_galois(void *hh, const uint64_t *elts, size_t count) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({ std::vector<size_t> e(elts, elts + count); h->glk = h->keygen->create_galois_keys_from_elements(e, false); h->has_glk = true; })
}
static int export_key(RefShim *h, const std::vector<PublicKey> &kv, uint64_t *out) {
    size_t per = 0;
    for (size_t j = 0; j < kv.size(); j++) {
        const Ciphertext &c = kv[j].as_ciphertext();
        per = c.data().size();
        export_ct(c, out + j * per);
    }
    return 0;
}
/* relin key: [decomp][2][K][N] */
int rs_relin_key(void *hh, uint64_t *out) { RefShim *h = (RefShim *)hh; RS_TRY(export_key(h, h->rlk.key(2), out)) }
int rs_galois_key(void *hh, uint64_t elt, uint64_t *out) { RefShim *h = (RefShim *)hh; RS_TRY(export_key(h, h->glk.key(elt), out)) }
int rs_secret_key(void *hh, uint64_t *out) { /* NTT form, [K][N] */
    RefShim *h = (RefShim *)hh;
    RS_TRY({
        const Plaintext &p = h->keygen->secret_key().as_plaintext();
        Plaintext hp = p.clone(); if (hp.on_device()) hp.to_host_inplace();
        std::memcpy(out, hp.data().raw_pointer(), hp.data().size() * 8);
    })
}
uint64_t rs_galois_element_from_step(void *hh, int step) {
    RefShim *h = (RefShim *)hh;
    return h->ctx->key_context_data().value()->galois_tool().get_element_from_step(step);
}

int rs_encrypt_zero(void *hh, uint64_t *out) { /* asymmetric, first level, [2][L][N]; CKKS scale = 2^40 */
    RefShim *h = (RefShim *)hh;
    RS_TRY({ Ciphertext c = h->enc->encrypt_zero_asymmetric_new(); export_ct(c, out); })
}

/* ---- raw-limb NTT through the L2 seam (src/utils/ntt.h:132-144,164-190) ---- */
static int ntt_common(RefShim *h, uint64_t *data, size_t pc, size_t nc, int mode, size_t base, size_t decomp, bool inverse) {
    RS_TRY({
        auto tabs = h->ctx->key_context_data().value()->small_ntt_tables();
        size_t total = pc * nc * h->n;
        utils::Array<uint64_t> arr(total, false, nullptr);
        std::memcpy(arr.raw_pointer(), data, total * 8);
        if (h->device) arr.to_device_inplace();
        utils::NTTTableIndexer idx = (mode == 1) ? utils::NTTTableIndexer::key_switching_set_products(tabs, decomp)
                                   : (mode == 2) ? utils::NTTTableIndexer::key_switching_skip_finals(tabs, decomp)
                                   : utils::NTTTableIndexer(tabs.const_slice(base, base + nc));
        if (inverse) utils::intt_inplace_ps(arr.reference(), pc, nc, h->n, idx);
        else utils::ntt_inplace_ps(arr.reference(), pc, nc, h->n, idx);
        if (h->device) { cudaDeviceSynchronize(); arr.to_host_inplace(); }
        std::memcpy(data, arr.raw_pointer(), total * 8);
    })
}
int rs_ntt_fwd(void *hh, uint64_t *d, size_t pc, size_t nc, int mode, size_t base, size_t decomp) { return ntt_common((RefShim *)hh, d, pc, nc, mode, base, decomp, false); }
int rs_ntt_inv(void *hh, uint64_t *d, size_t pc, size_t nc, int mode, size_t base, size_t decomp) { return ntt_common((RefShim *)hh, d, pc, nc, mode, base, decomp, true); }

/* ---- Evaluator ops on raw ciphertext arrays ---- */
int rs_multiply(void *hh, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t nl) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({ Ciphertext ca = make_ct(h, a, 2, nl, default_ntt_form(h)); Ciphertext cb = make_ct(h, b, 2, nl, default_ntt_form(h));
             export_ct(h->eval->multiply_new(ca, cb), out); })
}
int rs_square(void *hh, const uint64_t *a, uint64_t *out, size_t nl) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({ Ciphertext ca = make_ct(h, a, 2, nl, default_ntt_form(h)); export_ct(h->eval->square_new(ca), out); })
}
int rs_relinearize(void *hh, const uint64_t *a, uint64_t *out, size_t nl) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({ Ciphertext ca = make_ct(h, a, 3, nl, default_ntt_form(h)); export_ct(h->eval->relinearize_new(ca, h->rlk), out); })
}
int rs_apply_galois(void *hh, const uint64_t *a, uint64_t elt, uint64_t *out, size_t nl) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({ Ciphertext ca = make_ct(h, a, 2, nl, default_ntt_form(h)); export_ct(h->eval->apply_galois_new(ca, elt, h->glk), out); })
}
int rs_rotate(void *hh, const uint64_t *a, int steps, uint64_t *out, size_t nl) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({ Ciphertext ca = make_ct(h, a, 2, nl, default_ntt_form(h));
             export_ct(h->scheme == SchemeType::CKKS ? h->eval->rotate_vector_new(ca, steps, h->glk) : h->eval->rotate_rows_new(ca, steps, h->glk), out); })
}
/* CKKS: rescale_to_next; BFV/BGV: mod_switch_to_next. in [pc][nl][N] -> out [pc][nl-1][N] */
int rs_mod_switch_scale_to_next(void *hh, const uint64_t *a, size_t pc, uint64_t *out, size_t nl) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({ Ciphertext ca = make_ct(h, a, pc, nl, default_ntt_form(h));
             export_ct(h->scheme == SchemeType::CKKS ? h->eval->rescale_to_next_new(ca) : h->eval->mod_switch_to_next_new(ca), out); })
}
/* CKKS mod_switch_to_next (drop limb) */
int rs_drop_to_next(void *hh, const uint64_t *a, size_t pc, uint64_t *out, size_t nl) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({ Ciphertext ca = make_ct(h, a, pc, nl, default_ntt_form(h)); export_ct(h->eval->mod_switch_to_next_new(ca), out); })
}
int rs_transform_to_ntt(void *hh, const uint64_t *a, size_t pc, uint64_t *out, size_t nl) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({ Ciphertext ca = make_ct(h, a, pc, nl, false); export_ct(h->eval->transform_to_ntt_new(ca), out); })
}
int rs_transform_from_ntt(void *hh, const uint64_t *a, size_t pc, uint64_t *out, size_t nl) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({ Ciphertext ca = make_ct(h, a, pc, nl, true); export_ct(h->eval->transform_from_ntt_new(ca), out); })
}
/* ct (NTT form or not per ct_ntt) x plaintext given as raw limbs [nl][N] in the same form */
int rs_multiply_plain_raw(void *hh, const uint64_t *a, size_t pc, const uint64_t *plain, uint64_t *out, size_t nl, int ntt_form) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({
        Ciphertext ca = make_ct(h, a, pc, nl, ntt_form != 0);
        auto cd = level_with(h, nl);
        Plaintext p; p.resize_rns(*h->ctx, cd->parms_id());
        std::memcpy(p.data().raw_pointer(), plain, nl * h->n * 8);
        p.is_ntt_form() = ntt_form != 0;
        if (h->device) p.to_device_inplace();
        export_ct(h->eval->multiply_plain_new(ca, p), out);
    })
}
int rs_galois_raw(void *hh, const uint64_t *a, size_t pc, size_t nl, uint64_t elt, uint64_t *out, int ntt_form) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({
        auto cd = level_with(h, nl);
        size_t total = pc * nl * h->n;
        utils::Array<uint64_t> in(total, false, nullptr), res(total, false, nullptr);
        std::memcpy(in.raw_pointer(), a, total * 8);
        if (h->device) { in.to_device_inplace(); res.to_device_inplace(); }
        if (ntt_form) cd->galois_tool().apply_ntt_ps(in.const_reference(), pc, nl, elt, res.reference());
        else cd->galois_tool().apply_ps(in.const_reference(), pc, elt, cd->parms().coeff_modulus(), res.reference());
        if (h->device) { cudaDeviceSynchronize(); res.to_host_inplace(); }
        std::memcpy(out, res.raw_pointer(), total * 8);
    })
}

/* ---- Decryptor::decrypt on a raw ciphertext (src/decryptor.cu:233-250): BFV / BGV -> out [N], CKKS -> out [nl][N] ---- */
int rs_decrypt_raw(void *hh, const uint64_t *a, size_t pc, size_t nl, uint64_t correction_factor, uint64_t *out) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({
        Ciphertext ct = make_ct(h, a, pc, nl, default_ntt_form(h));
        ct.correction_factor() = correction_factor;
        Plaintext p = h->dec->decrypt_new(ct);
        if (p.on_device()) p.to_host_inplace();
        size_t words = h->scheme == SchemeType::CKKS ? nl * h->n : h->n;
        if (p.data().size() < words) throw std::logic_error("plaintext smaller than expected");
        std::memcpy(out, p.data().raw_pointer(), words * 8);
    })
}
/* the phase alone, in the ciphertext's form: out [nl][N]  (Decryptor::dot_product_ct_sk_array is private; CKKS decrypt returns
 * exactly it, and for BFV the reference exposes bfv_decrypt_without_scaling_down) */
int rs_decrypt_phase_raw(void *hh, const uint64_t *a, size_t pc, size_t nl, uint64_t *out) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({
        Ciphertext ct = make_ct(h, a, pc, nl, default_ntt_form(h));
        Plaintext p;
        if (h->scheme == SchemeType::BFV) h->dec->bfv_decrypt_without_scaling_down(ct, p);
        else if (h->scheme == SchemeType::CKKS) p = h->dec->decrypt_new(ct);
        else throw std::invalid_argument("phase export: BFV / CKKS only");
        if (p.on_device()) p.to_host_inplace();
        std::memcpy(out, p.data().raw_pointer(), nl * h->n * 8);
    })
}

/* ---- RandomGenerator / Encryptor with an explicit generator state (src/utils/random_generator.cu, src/encryptor.cu) ---- */
static __uint128_t u128_of(uint64_t lo, uint64_t hi) { return ((__uint128_t)hi << 64) | lo; }
/* kind: 0 fill_bytes(words * 8 bytes; `words` may be fractional via bytes), 1 ternary, 2 centered binomial, 3 uniform; nl limbs of the key-level base.
 * st = {seed_lo, seed_hi, counter_lo, counter_hi}, updated */
int rs_prng(void *hh, int kind, uint64_t *st, size_t nl, size_t bytes, uint64_t *out) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({
        utils::RandomGenerator g(u128_of(st[0], st[1]));
        g.set_counter(u128_of(st[2], st[3]));
        auto mods = h->ctx->key_context_data().value()->parms().coeff_modulus();
        utils::Array<Modulus> hm = utils::Array<Modulus>::create_and_copy_from_slice(mods, false, nullptr);
        if (kind == 0) g.fill_bytes(Slice<uint8_t>((uint8_t *)out, bytes, false, nullptr));
        else {
            Slice<uint64_t> d(out, nl * h->n, false, nullptr);
            if (kind == 1) g.sample_poly_ternary(d, h->n, hm.const_slice(0, nl));
            else if (kind == 2) g.sample_poly_centered_binomial(d, h->n, hm.const_slice(0, nl));
            else g.sample_poly_uniform(d, h->n, hm.const_slice(0, nl));
        }
        __uint128_t c = g.get_counter();
        st[2] = (uint64_t)c; st[3] = (uint64_t)(c >> 64);
    })
}
int rs_public_key(void *hh, uint64_t *out) { /* [2][K][N], NTT form */
    RefShim *h = (RefShim *)hh;
    RS_TRY(export_ct(h->enc->public_key().as_ciphertext(), out))
}
/* the context's generator is reset to (seed, counter 0) first, so the call is a deterministic function of its arguments.
 * plain == NULL: encryption of zero at the level with nl limbs.  Otherwise BFV / BGV: plain = plain_count words modulo t (no parms_id);
 * CKKS: plain = [nl][N] NTT form at the level with nl limbs.  out [2][nl][N]; *seed_out = the saved seed (symmetric + save_seed) */
int rs_encrypt_seeded(void *hh, const uint64_t *plain, size_t plain_count, int asymmetric, int save_seed, size_t nl, uint64_t seed_lo, uint64_t seed_hi,
                      uint64_t *out, uint64_t *seed_out) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({
        h->ctx->random_generator().reset_seed(u128_of(seed_lo, seed_hi));
        auto cd = level_with(h, nl);
        Ciphertext c;
        if (!plain) {
            c = asymmetric ? h->enc->encrypt_zero_asymmetric_new(cd->parms_id()) : h->enc->encrypt_zero_symmetric_new(save_seed != 0, cd->parms_id());
        } else {
            Plaintext p;
            if (h->scheme == SchemeType::CKKS) {
                p.resize_rns(*h->ctx, cd->parms_id(), false, false);
                std::memcpy(p.data().raw_pointer(), plain, nl * h->n * 8);
                p.is_ntt_form() = true; p.scale() = h->ckks_scale;
            } else {
                p.resize(plain_count);
                std::memcpy(p.data().raw_pointer(), plain, plain_count * 8);
            }
            if (h->device) p.to_device_inplace();
            c = asymmetric ? h->enc->encrypt_asymmetric_new(p) : h->enc->encrypt_symmetric_new(p, save_seed != 0);
        }
        if (seed_out) *seed_out = c.seed();
        export_ct(c, out);
    })
}

/* ---- Evaluator::pack_rlwe_ciphertexts_new (src/evaluator_lwes.cu:305-491) on raw ciphertexts [count][2][nl][N] with the Galois keys
 * generated by rs_gen_galois; out [2][nl][N] ---- */
int rs_pack_rlwes(void *hh, const uint64_t *cts, size_t count, size_t nl, int input_ntt_form, size_t shift, size_t input_interval, size_t output_interval,
                  int apply_field_trace, uint64_t *out) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({
        std::vector<Ciphertext> v; v.reserve(count);
        for (size_t i = 0; i < count; i++) v.push_back(make_ct(h, cts + i * 2 * nl * h->n, 2, nl, input_ntt_form != 0));
        std::vector<const Ciphertext *> ptrs;
        for (auto &c : v) ptrs.push_back(&c);
        Ciphertext r = h->eval->pack_rlwe_ciphertexts_new(ptrs, h->glk, shift, input_interval, output_interval, MemoryPool::GlobalPool(), apply_field_trace != 0);
        if (h->device) cudaDeviceSynchronize();
        export_ct(r, out);
    })
}

/* ---- Evaluator::extract_lwe_new (src/evaluator_lwes.cu:51-94) and pack_lwe_ciphertexts_new (:201-229) on raw data ---- */
int rs_extract_lwe(void *hh, const uint64_t *ct, size_t term, size_t nl, int ntt_form, uint64_t *c0, uint64_t *c1) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({
        Ciphertext c = make_ct(h, ct, 2, nl, ntt_form != 0);
        LWECiphertext l = h->eval->extract_lwe_new(c, term);
        if (l.on_device()) l.to_host_inplace();
        std::memcpy(c0, l.const_c0().raw_pointer(), nl * 8);
        std::memcpy(c1, l.const_c1().raw_pointer(), nl * h->n * 8);
    })
}
int rs_pack_lwes(void *hh, const uint64_t *c0s, const uint64_t *c1s, size_t count, size_t nl, int apply_field_trace, uint64_t *out) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({
        /* LWEs are built by extracting term 0 of RLWEs assembled from the raw (c0, c1): assemble o extract is the identity on them */
        std::vector<LWECiphertext> lwes; lwes.reserve(count);
        std::vector<uint64_t> raw(2 * nl * h->n);
        for (size_t i = 0; i < count; i++) {
            std::fill(raw.begin(), raw.end(), 0);
            for (size_t j = 0; j < nl; j++) raw[j * h->n] = c0s[i * nl + j];
            std::memcpy(raw.data() + nl * h->n, c1s + i * nl * h->n, nl * h->n * 8);
            Ciphertext c = make_ct(h, raw.data(), 2, nl, false);
            lwes.push_back(h->eval->extract_lwe_new(c, 0));
        }
        std::vector<const LWECiphertext *> ptrs;
        for (auto &l : lwes) ptrs.push_back(&l);
        Ciphertext r = h->eval->pack_lwe_ciphertexts_new(ptrs, h->glk, MemoryPool::GlobalPool(), apply_field_trace != 0);
        if (h->device) cudaDeviceSynchronize();
        export_ct(r, out);
    })
}

/* ---- Evaluator::multiply_plain / add_plain / sub_plain with a plaintext modulo t (no parms_id), on a raw ciphertext [2][nl][N] in the scheme's
 * default form (src/evaluator_multiply_plain.cu:13-68,309-328, src/evaluator_translate_plain.cu:95-170); op 0 multiply, 1 add, 2 sub ---- */
int rs_plain_op(void *hh, int op, const uint64_t *ct, size_t nl, const uint64_t *plain, size_t plain_count, uint64_t correction_factor, uint64_t *out) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({
        Ciphertext c = make_ct(h, ct, 2, nl, default_ntt_form(h));
        c.correction_factor() = correction_factor;
        Plaintext p; p.resize(plain_count);
        std::memcpy(p.data().raw_pointer(), plain, plain_count * 8);
        if (h->device) p.to_device_inplace();
        Ciphertext r = op == 0 ? h->eval->multiply_plain_new(c, p) : op == 1 ? h->eval->add_plain_new(c, p) : h->eval->sub_plain_new(c, p);
        if (h->device) cudaDeviceSynchronize();
        export_ct(r, out);
    })
}

/* ---- BatchEncoder::encode / decode (src/batch_encoder.cu:169-227,407-460), host contexts only (the reference builds the encoder from a host context) ---- */
int rs_batch_encode(void *hh, const uint64_t *values, size_t count, uint64_t *out) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({
        if (h->device) throw std::invalid_argument("batch encoder shim: host context only");
        BatchEncoder be(h->ctx);
        std::vector<uint64_t> v(values, values + count);
        Plaintext p = be.encode_new(v);
        std::memcpy(out, p.data().raw_pointer(), h->n * 8);
    })
}
int rs_batch_decode(void *hh, const uint64_t *plain, size_t count, uint64_t *out) {
    RefShim *h = (RefShim *)hh;
    RS_TRY({
        if (h->device) throw std::invalid_argument("batch encoder shim: host context only");
        BatchEncoder be(h->ctx);
        Plaintext p; p.resize(count);
        std::memcpy(p.data().raw_pointer(), plain, count * 8);
        std::vector<uint64_t> v = be.decode_new(p);
        std::memcpy(out, v.data(), h->n * 8);
    })
}

/* ---- timing helper: seconds per op, averaged over iters, data resident where the context is ----
 * op: 0 ntt_fwd(pc*nc limbs, componentwise key tables base 0) 1 ntt_inv 2 multiply 3 relinearize
 *     4 rotate(1) 5 rescale/modswitch_to_next 6 square 7 encrypt_zero_asymmetric (level nl) 8 decrypt (2 polynomials, level nl)
 *     9 pack_rlwe_ciphertexts_new of pc ciphertexts, input_interval = N, output_interval = N / pc, with the field trace (needs the automorphism keys) */
double rs_time_op(void *hh, int op, size_t nl, size_t pc, int warmup, int iters) {
    RefShim *h = (RefShim *)hh;
    try {
        std::vector<uint64_t> raw(3 * std::max(nl, (size_t)1) * h->n * std::max(pc, (size_t)1));
        uint64_t s = 1; for (auto &w : raw) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; w = (s >> 24) & 0xffffffffULL; }
        bool nf = default_ntt_form(h);
        Ciphertext c2, c3; utils::Array<uint64_t> arr;
        auto tabs = h->ctx->key_context_data().value()->small_ntt_tables();
        if (op <= 1) { arr = utils::Array<uint64_t>(pc * nl * h->n, false, nullptr); std::memcpy(arr.raw_pointer(), raw.data(), pc * nl * h->n * 8); if (h->device) arr.to_device_inplace(); }
        else { c2 = make_ct(h, raw.data(), 2, nl, nf); c3 = make_ct(h, raw.data(), 3, nl, nf); }
        double total = 0;
        for (int it = 0; it < warmup + iters; it++) {
            if (h->device) cudaDeviceSynchronize();
            auto t0 = std::chrono::steady_clock::now();
            switch (op) {
            case 0: utils::ntt_inplace_ps(arr.reference(), pc, nl, h->n, utils::NTTTableIndexer(tabs.const_slice(0, nl))); break;
            case 1: utils::intt_inplace_ps(arr.reference(), pc, nl, h->n, utils::NTTTableIndexer(tabs.const_slice(0, nl))); break;
            case 2: { Ciphertext r = h->eval->multiply_new(c2, c2); } break;
            case 3: { Ciphertext r = h->eval->relinearize_new(c3, h->rlk); } break;
            case 4: { Ciphertext r = (h->scheme == SchemeType::CKKS) ? h->eval->rotate_vector_new(c2, 1, h->glk) : h->eval->rotate_rows_new(c2, 1, h->glk); } break;
            case 5: { Ciphertext r = (h->scheme == SchemeType::CKKS) ? h->eval->rescale_to_next_new(c2) : h->eval->mod_switch_to_next_new(c2); } break;
            case 6: { Ciphertext r = h->eval->square_new(c2); } break;
            case 7: { Ciphertext r = h->enc->encrypt_zero_asymmetric_new(level_with(h, nl)->parms_id()); } break;
            case 8: { Plaintext r = h->dec->decrypt_new(c2); } break;
            case 9: {
                std::vector<const Ciphertext *> ptrs(pc, &c2);
                Ciphertext r = h->eval->pack_rlwe_ciphertexts_new(ptrs, h->glk, 0, h->n, h->n / pc);
            } break;
            default: throw std::invalid_argument("bad op");
            }
            if (h->device) cudaDeviceSynchronize();
            auto t1 = std::chrono::steady_clock::now();
            if (it >= warmup) total += std::chrono::duration<double>(t1 - t0).count();
        }
        return total / iters;
    } catch (const std::exception &e) { h->err = e.what(); return -1.0; }
}

} /* extern "C" */
