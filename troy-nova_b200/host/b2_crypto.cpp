// b2_crypto.cpp — see b2_crypto.h.  Orchestration only: every step is a C-ABI call into libb2rns.
#include "b2_crypto.h"

namespace b2troy {

SecretKey SecretKey::from_host(const HeContextPointer &ctx, const uint64_t *words) {
    SecretKey k; k.ctx_ = ctx;
    k.buf_ = DeviceArray(ctx, ctx->key_modulus_size() * ctx->poly_modulus_degree());
    k.buf_.upload(words, k.buf_.size());
    return k;
}
SecretKey SecretKey::uninitialised(const HeContextPointer &ctx) {
    SecretKey k; k.ctx_ = ctx;
    k.buf_ = DeviceArray(ctx, ctx->key_modulus_size() * ctx->poly_modulus_degree());
    return k;
}
SecretKey SecretKey::clone() const { SecretKey k; k.ctx_ = ctx_; k.buf_ = buf_.clone(); return k; }
std::vector<uint64_t> SecretKey::to_host() const {
    std::vector<uint64_t> v(buf_.size());
    if (!v.empty()) buf_.download(v.data(), v.size());
    return v;
}
PublicKey PublicKey::from_host(const HeContextPointer &ctx, const uint64_t *words) {
    PublicKey k; k.ctx_ = ctx;
    k.buf_ = DeviceArray(ctx, 2 * ctx->key_modulus_size() * ctx->poly_modulus_degree());
    k.buf_.upload(words, k.buf_.size());
    return k;
}
PublicKey PublicKey::create(const HeContextPointer &ctx, const SecretKey &sk, RandomGenerator &prng) {
    if (sk.empty()) throw std::invalid_argument("[KeyGenerator::create_public_key] Secret key is empty.");
    PublicKey k; k.ctx_ = ctx;
    k.buf_ = DeviceArray(ctx, 2 * ctx->key_modulus_size() * ctx->poly_modulus_degree());
    check(b2_encrypt_zero_symmetric(ctx->raw(), sk.data(), ctx->key_modulus_size(), 1, 0, prng.raw(), prng.raw(), k.buf_.data(), nullptr, ctx->stream),
          "[KeyGenerator::create_public_key]");
    return k;
}
std::vector<uint64_t> PublicKey::to_host() const {
    std::vector<uint64_t> v(buf_.size());
    if (!v.empty()) buf_.download(v.data(), v.size());
    return v;
}

// ---- BatchEncoder (src/batch_encoder.cu) ----------------------------------------------------------------------------------------------------
BatchEncoder::BatchEncoder(HeContextPointer context) : context_(std::move(context)) {
    if (context_->scheme() != SchemeType::BFV && context_->scheme() != SchemeType::BGV) throw std::invalid_argument("[BatchEncoder::BatchEncoder] Unsupported scheme.");
}
void BatchEncoder::encode(const std::vector<uint64_t> &values, Plaintext &destination) const {
    const size_t n = slot_count();
    if (values.size() > n) throw std::invalid_argument("[BatchEncoder::encode_slice] Values has size larger than the number of slots.");
    DeviceArray v(context_, (values.size() + 2) & ~(size_t)1);
    if (!values.empty()) v.upload(values.data(), values.size());
    Plaintext p = Plaintext::mod_t(context_, n);
    const b2_status st = b2_batch_encode(context_->raw(), v.data(), values.size(), p.data(), context_->stream);
    if (st == B2_ERR_INVALID_ARG) throw std::logic_error(std::string("[BatchEncoder::encode_slice] ") + b2_last_error());      // "does not support vector batching"
    check(st, "[BatchEncoder::encode_slice]");
    check(b2_stream_sync(context_->raw(), context_->stream), "[BatchEncoder::encode_slice]");
    destination = std::move(p);
}
std::vector<uint64_t> BatchEncoder::decode_new(const Plaintext &plain) const {
    if (plain.empty() || !plain.is_mod_t()) throw std::invalid_argument("[BatchEncoder::decode] Plaintext must be modulo t (no parms_id).");
    if (plain.is_ntt_form()) throw std::invalid_argument("[BatchEncoder::decode] Plaintext is in NTT form.");
    const size_t n = slot_count();
    DeviceArray out(context_, n);
    const b2_status st = b2_batch_decode(context_->raw(), plain.data(), plain.coeff_count(), out.data(), context_->stream);
    if (st == B2_ERR_INVALID_ARG) throw std::logic_error(std::string("[BatchEncoder::encode] ") + b2_last_error());
    check(st, "[BatchEncoder::decode]");
    std::vector<uint64_t> v(n);
    out.download(v.data(), n);
    return v;
}
Plaintext BatchEncoder::encode_polynomial_new(const std::vector<uint64_t> &values) const {
    if (values.size() > slot_count()) throw std::invalid_argument("[BatchEncoder::encode_polynomial_slice] Values has size larger than the number of slots.");
    if (values.empty()) throw std::invalid_argument("[BatchEncoder::encode_polynomial_slice] Values is empty.");
    std::vector<uint64_t> r(values);
    const uint64_t t = context_->plain_modulus();
    if (t < 2) throw std::invalid_argument("[BatchEncoder::encode_polynomial_slice] The context has no plain modulus.");
    for (uint64_t &w : r) w %= t;
    return Plaintext::from_host_mod_t(context_, r.data(), r.size());
}

// ---- KeyGenerator (src/key_generator.cu:31-78, 136-256) ---------------------------------------------------------------------------------
KeyGenerator::KeyGenerator(HeContextPointer context, RandomGenerator &prng) : context_(std::move(context)), prng_(&prng) {
    sk_ = SecretKey::uninitialised(context_);
    check(b2_generate_secret_key(context_->raw(), prng_->raw(), sk_.data(), context_->stream), "[KeyGenerator::KeyGenerator]");
}
KeyGenerator::KeyGenerator(HeContextPointer context, RandomGenerator &prng, const SecretKey &secret_key) : context_(std::move(context)), prng_(&prng) {
    if (secret_key.empty()) throw std::invalid_argument("[KeyGenerator::KeyGenerator] Secret key is empty.");
    sk_ = secret_key.clone();
}
void KeyGenerator::one_kswitch_key(const uint64_t *new_key, KSwitchKeys &dest, size_t index) {
    const size_t K = context_->key_modulus_size(), n = context_->poly_modulus_degree();
    if (K < 2) throw std::logic_error("[KeyGenerator::generate_one_kswitch_key] Keyswitching is not enabled.");
    DeviceArray key(context_, (K - 1) * 2 * K * n);
    check(b2_generate_kswitch_key(context_->raw(), sk_.data(), new_key, prng_->raw(), key.data(), context_->stream), "[KeyGenerator::generate_one_kswitch_key]");
    dest.adopt_key(index, std::move(key), K - 1);
}
RelinKeys KeyGenerator::create_relin_keys() {
    const size_t w = context_->key_modulus_size() * context_->poly_modulus_degree();
    DeviceArray powers(context_, 2 * w);
    check(b2_secret_key_powers(context_->raw(), sk_.data(), 2, powers.data(), context_->stream), "[KeyGenerator::create_relin_keys]");
    RelinKeys rk(context_);
    one_kswitch_key(powers.data() + w, rk, RelinKeys::get_index(2));
    check(b2_stream_sync(context_->raw(), context_->stream), "[KeyGenerator::create_relin_keys]");      // `powers` is freed on return
    return rk;
}
GaloisKeys KeyGenerator::create_galois_keys(const std::vector<size_t> &galois_elements) {
    const size_t K = context_->key_modulus_size(), n = context_->poly_modulus_degree();
    GaloisKeys gk(context_);
    DeviceArray rotated(context_, K * n);
    for (size_t e : galois_elements) {
        if (e % 2 == 0 || e >= 2 * n) throw std::invalid_argument("[KeyGenerator::generate_galois_keys] Galois element is not valid.");
        if (gk.has_key(e)) continue;
        check(b2_galois(context_->raw(), sk_.data(), e, rotated.data(), 1, K, 0, 1, context_->stream), "[KeyGenerator::generate_galois_keys]");
        one_kswitch_key(rotated.data(), gk, GaloisKeys::get_index(e));
    }
    check(b2_stream_sync(context_->raw(), context_->stream), "[KeyGenerator::generate_galois_keys]");
    return gk;
}
GaloisKeys KeyGenerator::create_automorphism_keys() {
    std::vector<size_t> elts;
    for (size_t e = 2; e <= context_->poly_modulus_degree(); e <<= 1) elts.push_back(e + 1);
    return create_galois_keys(elts);
}

// ---- Encryptor (src/encryptor.cu:12-118, 247-338) ---------------------------------------------------------------------------------
void Encryptor::encrypt_zero_internal(size_t nl, bool is_ntt_form, bool is_asymmetric, bool save_seed, RandomGenerator *u_prng, Ciphertext &destination) {
    if (is_asymmetric && public_key_.empty()) throw std::invalid_argument("[Encryptor::encrypt_zero_internal] Public key not set for asymmetric encryption.");
    if (!is_asymmetric && secret_key_.empty()) throw std::invalid_argument("[Encryptor::encrypt_zero_internal] Secret key not set for symmetric encryption.");
    if (save_seed && is_asymmetric) throw std::invalid_argument("[Encryptor::encrypt_zero_internal] Cannot save seed when using asymmetric encryption.");
    if (nl == 0) nl = context_->first_modulus_size();
    if (nl > context_->key_modulus_size()) throw std::invalid_argument("[Encryptor::encrypt_zero_internal] parms_id is not valid for encryption parameters.");
    if (is_asymmetric && context_->scheme() == SchemeType::BGV && !is_ntt_form && nl < context_->key_modulus_size())
        throw std::invalid_argument("[Encryptor::encrypt_zero_internal] BGV - Plaintext is not in NTT form.");
    Ciphertext r(context_, 2, nl, is_ntt_form);
    b2_prng *first = u_prng ? u_prng->raw() : prng_.raw();
    if (is_asymmetric) {
        check(b2_encrypt_zero_asymmetric(context_->raw(), public_key_.data(), nl, is_ntt_form ? 1 : 0, first, prng_.raw(), r.data(), context_->stream),
              "[Encryptor::encrypt_zero_internal]");
    } else {
        uint64_t seed = 0;
        check(b2_encrypt_zero_symmetric(context_->raw(), secret_key_.data(), nl, is_ntt_form ? 1 : 0, save_seed ? 1 : 0, first, prng_.raw(), r.data(), &seed, context_->stream),
              "[Encryptor::encrypt_zero_internal]");
        r.seed() = save_seed ? seed : 0;
    }
    destination = std::move(r);
}
void Encryptor::encrypt_zero_asymmetric(Ciphertext &d, size_t nl, RandomGenerator *u_prng) {
    encrypt_zero_internal(nl, context_->scheme() != SchemeType::BFV, true, false, u_prng, d);
}
void Encryptor::encrypt_zero_symmetric(bool save_seed, Ciphertext &d, size_t nl, RandomGenerator *u_prng) {
    encrypt_zero_internal(nl, context_->scheme() != SchemeType::BFV, false, save_seed, u_prng, d);
}
void Encryptor::encrypt_internal(const Plaintext &plain, bool is_asymmetric, bool save_seed, RandomGenerator *u_prng, Ciphertext &destination) {
    if (plain.empty() || plain.context().get() != context_.get()) throw std::invalid_argument("[Encryptor::encrypt_internal] The arguments are not on the same device.");
    b2_ctx *c = context_->raw();
    const size_t first = context_->first_modulus_size();
    Ciphertext r;
    switch (context_->scheme()) {
    case SchemeType::BFV:
        if (plain.is_mod_t()) {
            encrypt_zero_internal(first, false, is_asymmetric, save_seed, u_prng, r);
            check(b2_bfv_scale_up(c, plain.data(), plain.coeff_count(), first, r.poly(0), 0, r.poly(0), context_->stream), "[Encryptor::encrypt_internal]");
        } else {      // an RNS plaintext at some level, in its own form: added as is (encryptor.cu:268-281)
            encrypt_zero_internal(plain.coeff_modulus_size(), plain.is_ntt_form(), is_asymmetric, save_seed, u_prng, r);
            check(b2_add(c, r.poly(0), plain.data(), r.poly(0), 1, plain.coeff_modulus_size(), 0, context_->stream), "[Encryptor::encrypt_internal]");
        }
        break;
    case SchemeType::CKKS:
        if (plain.is_mod_t()) throw std::invalid_argument("[Encryptor::encrypt_internal] CKKS - Plaintext parms_id is not valid.");
        encrypt_zero_internal(plain.coeff_modulus_size(), plain.is_ntt_form(), is_asymmetric, save_seed, u_prng, r);
        check(b2_add(c, r.poly(0), plain.data(), r.poly(0), 1, plain.coeff_modulus_size(), 0, context_->stream), "[Encryptor::encrypt_internal]");
        r.scale() = plain.scale();
        break;
    case SchemeType::BGV:
        if (plain.is_mod_t()) {      // c_0 += NTT(centralize(m)): b2_encrypt does the zero encryption and the lift in one call
            if (is_asymmetric && public_key_.empty()) throw std::invalid_argument("[Encryptor::encrypt_zero_internal] Public key not set for asymmetric encryption.");
            if (!is_asymmetric && secret_key_.empty()) throw std::invalid_argument("[Encryptor::encrypt_zero_internal] Secret key not set for symmetric encryption.");
            if (save_seed) throw std::logic_error("[Encryptor::encrypt_internal] BGV with a saved seed is not mirrored.");
            r = Ciphertext(context_, 2, first, true);
            b2_prng *fp = u_prng ? u_prng->raw() : prng_.raw();
            check(b2_encrypt(c, plain.data(), plain.coeff_count(), is_asymmetric ? public_key_.data() : secret_key_.data(), is_asymmetric ? 1 : 0, first, fp, prng_.raw(), r.data(),
                             context_->stream), "[Encryptor::encrypt_internal]");
        } else {
            if (!plain.is_ntt_form()) throw std::invalid_argument("[Encryptor::encrypt_internal] BGV - an RNS plaintext must be in NTT form.");
            encrypt_zero_internal(first, true, is_asymmetric, save_seed, u_prng, r);
            check(b2_add(c, r.poly(0), plain.data(), r.poly(0), 1, first, 0, context_->stream), "[Encryptor::encrypt_internal]");
        }
        break;
    default: throw std::invalid_argument("[Encryptor::encrypt_internal] Unsupported scheme.");
    }
    destination = std::move(r);
}
void Encryptor::encrypt_asymmetric(const Plaintext &plain, Ciphertext &d, RandomGenerator *u_prng) { encrypt_internal(plain, true, false, u_prng, d); }
void Encryptor::encrypt_symmetric(const Plaintext &plain, bool save_seed, Ciphertext &d, RandomGenerator *u_prng) { encrypt_internal(plain, false, save_seed, u_prng, d); }

// ---- Decryptor (src/decryptor.cu) ---------------------------------------------------------------------------------------------------
Decryptor::Decryptor(HeContextPointer context, const SecretKey &secret_key) : context_(std::move(context)) {
    if (secret_key.empty() || secret_key.context().get() != context_.get())
        throw std::invalid_argument("[Decryptor::Decryptor] secret_key is not valid for encryption parameters");
    const size_t w = context_->key_modulus_size() * context_->poly_modulus_degree();
    powers_ = DeviceArray(context_, w);
    check(b2_secret_key_powers(context_->raw(), secret_key.data(), 1, powers_.data(), context_->stream), "[Decryptor::Decryptor]");
    n_powers_ = 1;
}
const uint64_t *Decryptor::powers(size_t max_power) const {
    std::lock_guard<std::mutex> g(mu_);
    if (max_power > n_powers_) {      // KeyGenerator::compute_secret_key_powers: extend the array, keep the prefix
        const size_t w = context_->key_modulus_size() * context_->poly_modulus_degree();
        DeviceArray bigger(context_, max_power * w);
        check(b2_secret_key_powers(context_->raw(), powers_.data(), max_power, bigger.data(), context_->stream), "[Decryptor::dot_product_ct_sk_array]");
        check(b2_stream_sync(context_->raw(), context_->stream), "[Decryptor::dot_product_ct_sk_array]");      // the old array is freed below
        powers_ = std::move(bigger); n_powers_ = max_power;
    }
    return powers_.data();
}
static void check_decrypt_input(const HeContextPointer &ctx, const Ciphertext &e) {
    if (e.contains_seed()) throw std::invalid_argument("[Decryptor::decrypt] Seed should be expanded first.");
    if (e.empty() || e.polynomial_count() < 2) throw std::invalid_argument("[Decryptor::decrypt] Ciphertext is empty.");
    if (e.context().get() != ctx.get()) throw std::invalid_argument("[Decryptor::decrypt] Ciphertext belongs to another context.");
    const SchemeType s = ctx->scheme();
    if (s == SchemeType::BFV && e.is_ntt_form()) throw std::invalid_argument("[Decryptor::bfv_decrypt] Ciphertext is in NTT form.");
    if (s == SchemeType::CKKS && !e.is_ntt_form()) throw std::invalid_argument("[Decryptor::ckks_decrypt] Ciphertext is not in NTT form.");
    if (s == SchemeType::BGV && !e.is_ntt_form()) throw std::invalid_argument("[Decryptor::bgv_decrypt] Ciphertext is not in NTT form.");
}
void Decryptor::decrypt_batched(const std::vector<const Ciphertext *> &es, const std::vector<Plaintext *> &ds) const {
    if (es.size() != ds.size()) throw std::invalid_argument("[Decryptor::decrypt_batched] encrypted and destination are not the same size.");
    if (es.empty()) return;
    for (const Ciphertext *e : es) {
        if (!e) throw std::invalid_argument("[Decryptor::decrypt] Ciphertext is empty.");
        check_decrypt_input(context_, *e);
        if (e->coeff_modulus_size() != es[0]->coeff_modulus_size()) throw std::invalid_argument("[Decryptor::bfv_decrypt_batched] Ciphertexts are not in the same context.");
        if (e->polynomial_count() != es[0]->polynomial_count()) throw std::invalid_argument("[Decryptor::dot_product_ct_sk_array_batched] Ciphertexts are not in the same size.");
    }
    const size_t pc = es[0]->polynomial_count(), nl = es[0]->coeff_modulus_size(), n = context_->poly_modulus_degree();
    const bool ckks = context_->scheme() == SchemeType::CKKS;
    std::vector<Plaintext> out(es.size());
    std::vector<const uint64_t *> in(es.size());
    std::vector<uint64_t *> op(es.size());
    std::vector<uint64_t> cf(es.size());
    for (size_t i = 0; i < es.size(); i++) {
        out[i] = ckks ? Plaintext::rns(context_, nl, true) : Plaintext::mod_t(context_, n);
        if (ckks) out[i].scale() = es[i]->scale();
        in[i] = es[i]->data(); op[i] = out[i].data(); cf[i] = es[i]->correction_factor();
    }
    check(b2_decrypt(context_->raw(), in.data(), op.data(), es.size(), pc, nl, powers(pc - 1), cf.data(), context_->stream), "[Decryptor::decrypt]");
    for (size_t i = 0; i < es.size(); i++) *ds[i] = std::move(out[i]);
}
void Decryptor::decrypt(const Ciphertext &e, Plaintext &d) const { decrypt_batched({&e}, {&d}); }
void Decryptor::phase(const Ciphertext &e, Plaintext &d) const {
    if (e.empty() || e.polynomial_count() < 2) throw std::invalid_argument("[Decryptor::decrypt] Ciphertext is empty.");
    Plaintext r = Plaintext::rns(context_, e.coeff_modulus_size(), e.is_ntt_form());
    const uint64_t *in = e.data(); uint64_t *o = r.data();
    check(b2_decrypt_phase(context_->raw(), &in, &o, 1, e.polynomial_count(), e.coeff_modulus_size(), e.is_ntt_form() ? 1 : 0, powers(e.polynomial_count() - 1), context_->stream),
          "[Decryptor::dot_product_ct_sk_array]");
    d = std::move(r);
}

}  // namespace b2troy
