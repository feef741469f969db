// b2_evaluator.cpp — see b2_evaluator.h.  Orchestration only: every arithmetic step is one C-ABI call into libb2rns.
#include "b2_evaluator.h"

#include <cmath>
#include <cstdlib>

namespace b2troy {

void check(b2_status st, const char *where) {
    if (st == B2_OK) return;
    const std::string msg = std::string(where) + ": " + b2_last_error();
    if (st == B2_ERR_INVALID_ARG) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}

// ---- HeContext ---------------------------------------------------------------------------------------------------------
HeContextPointer HeContext::create(SchemeType scheme, size_t poly_modulus_degree, const std::vector<uint64_t> &coeff_modulus,
                                   uint64_t plain_modulus, int device) {
    if (poly_modulus_degree < 2 || (poly_modulus_degree & (poly_modulus_degree - 1)) != 0)
        throw std::invalid_argument("[HeContext::create] poly_modulus_degree must be a power of two.");
    if (coeff_modulus.empty()) throw std::invalid_argument("[HeContext::create] coeff_modulus is empty.");
    unsigned log_n = 0;
    while (((size_t)1 << log_n) < poly_modulus_degree) log_n++;
    b2_params prm;
    prm.log_n = log_n; prm.n_moduli = (uint32_t)coeff_modulus.size(); prm.moduli = coeff_modulus.data();
    prm.plain_modulus = plain_modulus; prm.scheme = (uint32_t)scheme;
    HeContextPointer c(new HeContext());
    check(b2_ctx_create(device, &prm, &c->ctx_), "[HeContext::create]");
    c->scheme_ = scheme; c->n_ = poly_modulus_degree; c->moduli_ = coeff_modulus; c->t_ = plain_modulus;
    return c;
}
HeContext::~HeContext() { b2_ctx_destroy(ctx_); }
void HeContext::synchronize() const { check(b2_stream_sync(ctx_, stream), "[HeContext::synchronize]"); }

// ---- DeviceArray -------------------------------------------------------------------------------------------------------
DeviceArray::DeviceArray(const HeContextPointer &ctx, size_t words) : ctx_(ctx), words_(words) {
    if (words) { void *p = nullptr; check(b2_malloc(ctx->raw(), words * 8, &p), "[DeviceArray]"); p_ = (uint64_t *)p; }
}
DeviceArray::~DeviceArray() { if (p_ && ctx_) b2_free(ctx_->raw(), p_); }
DeviceArray &DeviceArray::operator=(DeviceArray &&o) noexcept {
    if (this != &o) {
        if (p_ && ctx_) b2_free(ctx_->raw(), p_);
        ctx_ = std::move(o.ctx_); p_ = o.p_; words_ = o.words_; o.p_ = nullptr; o.words_ = 0;
    }
    return *this;
}
DeviceArray DeviceArray::clone() const {
    DeviceArray r(ctx_, words_);
    if (words_) check(b2_memcpy_d2d(ctx_->raw(), r.p_, p_, words_ * 8, ctx_->stream), "[DeviceArray::clone]");
    return r;
}
void DeviceArray::upload(const uint64_t *host, size_t words) {
    if (words > words_) throw std::invalid_argument("[DeviceArray::upload] too many words.");
    check(b2_memcpy_h2d(ctx_->raw(), p_, host, words * 8, ctx_->stream), "[DeviceArray::upload]");
    check(b2_stream_sync(ctx_->raw(), ctx_->stream), "[DeviceArray::upload]");
}
void DeviceArray::download(uint64_t *host, size_t words) const {
    if (words > words_) throw std::invalid_argument("[DeviceArray::download] too many words.");
    check(b2_memcpy_d2h(ctx_->raw(), host, p_, words * 8, ctx_->stream), "[DeviceArray::download]");
    check(b2_stream_sync(ctx_->raw(), ctx_->stream), "[DeviceArray::download]");
}

// ---- Ciphertext / Plaintext -----------------------------------------------------------------------------------------------
Ciphertext::Ciphertext(const HeContextPointer &ctx, size_t polynomial_count, size_t coeff_modulus_size, bool is_ntt_form)
    : ctx_(ctx), polys_(polynomial_count), limbs_(coeff_modulus_size), ntt_(is_ntt_form),
      buf_(ctx, polynomial_count * coeff_modulus_size * ctx->poly_modulus_degree()) {}
Ciphertext Ciphertext::like(const Ciphertext &o, size_t polynomial_count) {
    Ciphertext r(o.ctx_, polynomial_count, o.limbs_, o.ntt_);
    r.scale_ = o.scale_; r.cf_ = o.cf_;
    return r;
}
Ciphertext Ciphertext::from_host(const HeContextPointer &ctx, const uint64_t *words, size_t polynomial_count, size_t coeff_modulus_size,
                                 bool is_ntt_form, double scale, uint64_t correction_factor) {
    if (coeff_modulus_size == 0 || coeff_modulus_size > ctx->first_modulus_size())
        throw std::invalid_argument("[Ciphertext::from_host] coeff_modulus_size out of range for this context.");
    Ciphertext r(ctx, polynomial_count, coeff_modulus_size, is_ntt_form);
    r.scale_ = scale; r.cf_ = correction_factor;
    r.buf_.upload(words, r.buf_.size());
    return r;
}
Ciphertext Ciphertext::clone() const {
    Ciphertext r;
    r.ctx_ = ctx_; r.polys_ = polys_; r.limbs_ = limbs_; r.ntt_ = ntt_; r.scale_ = scale_; r.cf_ = cf_; r.seed_ = seed_;
    r.buf_ = buf_.clone();
    return r;
}
std::vector<uint64_t> Ciphertext::to_host() const {
    std::vector<uint64_t> v(buf_.size());
    if (!v.empty()) buf_.download(v.data(), v.size());
    return v;
}
Plaintext Plaintext::rns(const HeContextPointer &ctx, size_t coeff_modulus_size, bool is_ntt_form) {
    if (coeff_modulus_size == 0 || coeff_modulus_size > ctx->key_modulus_size()) throw std::invalid_argument("[Plaintext::rns] coeff_modulus_size out of range for this context.");
    Plaintext p;
    p.ctx_ = ctx; p.limbs_ = coeff_modulus_size; p.coeff_count_ = ctx->poly_modulus_degree(); p.ntt_ = is_ntt_form;
    p.buf_ = DeviceArray(ctx, coeff_modulus_size * ctx->poly_modulus_degree());
    return p;
}
Plaintext Plaintext::mod_t(const HeContextPointer &ctx, size_t coeff_count) {
    if (coeff_count == 0 || coeff_count > ctx->poly_modulus_degree()) throw std::invalid_argument("[Plaintext::mod_t] coeff_count must be in [1, N].");
    Plaintext p;
    p.ctx_ = ctx; p.limbs_ = 0; p.coeff_count_ = coeff_count; p.ntt_ = false;
    p.buf_ = DeviceArray(ctx, (coeff_count + 1) & ~(size_t)1);
    return p;
}
Plaintext Plaintext::from_host(const HeContextPointer &ctx, const uint64_t *words, size_t coeff_modulus_size, bool is_ntt_form) {
    Plaintext p = rns(ctx, coeff_modulus_size, is_ntt_form);
    p.buf_.upload(words, p.buf_.size());
    return p;
}
Plaintext Plaintext::from_host_mod_t(const HeContextPointer &ctx, const uint64_t *words, size_t coeff_count) {
    Plaintext p = mod_t(ctx, coeff_count);
    p.buf_.upload(words, coeff_count);
    return p;
}
std::vector<uint64_t> Plaintext::to_host() const {
    std::vector<uint64_t> v(is_mod_t() ? coeff_count_ : buf_.size());
    if (!v.empty()) buf_.download(v.data(), v.size());
    return v;
}

// ---- keys -------------------------------------------------------------------------------------------------------------------
KSwitchKeys::~KSwitchKeys() { for (b2_kswitch_key *k : keys_) b2_kswitch_key_destroy(k); }
KSwitchKeys::KSwitchKeys(KSwitchKeys &&o) noexcept { *this = std::move(o); }
KSwitchKeys &KSwitchKeys::operator=(KSwitchKeys &&o) noexcept {
    if (this != &o) {
        for (b2_kswitch_key *k : keys_) b2_kswitch_key_destroy(k);
        ctx_ = std::move(o.ctx_); keys_ = std::move(o.keys_); digits_ = std::move(o.digits_); storage_ = std::move(o.storage_);
        o.keys_.clear();
    }
    return *this;
}
void KSwitchKeys::set_key(size_t index, const uint64_t *words, size_t digits) {
    if (!ctx_) throw std::invalid_argument("[KSwitchKeys::set_key] no context.");
    const size_t per = 2 * ctx_->key_modulus_size() * ctx_->poly_modulus_degree();
    if (index >= keys_.size()) { keys_.resize(index + 1, nullptr); digits_.resize(index + 1, 0); storage_.resize(index + 1); }
    if (keys_[index]) { b2_kswitch_key_destroy(keys_[index]); keys_[index] = nullptr; }
    storage_[index] = DeviceArray(ctx_, digits * per);
    storage_[index].upload(words, digits * per);
    check(b2_kswitch_key_create_contiguous(ctx_->raw(), storage_[index].data(), digits, &keys_[index]), "[KSwitchKeys::set_key]");
    digits_[index] = digits;
}

void KSwitchKeys::adopt_key(size_t index, DeviceArray &&storage, size_t digits) {
    if (!ctx_) throw std::invalid_argument("[KSwitchKeys::adopt_key] no context.");
    if (index >= keys_.size()) { keys_.resize(index + 1, nullptr); digits_.resize(index + 1, 0); storage_.resize(index + 1); }
    if (keys_[index]) { b2_kswitch_key_destroy(keys_[index]); keys_[index] = nullptr; }
    storage_[index] = std::move(storage);
    check(b2_kswitch_key_create_contiguous(ctx_->raw(), storage_[index].data(), digits, &keys_[index]), "[KSwitchKeys::adopt_key]");
    digits_[index] = digits;
}
std::vector<uint64_t> KSwitchKeys::key_to_host(size_t index) const {
    if (!has_index(index)) throw std::invalid_argument("[KSwitchKeys::key_to_host] key not present.");
    std::vector<uint64_t> v(storage_[index].size());
    storage_[index].download(v.data(), v.size());
    return v;
}

std::vector<int> naf(int value) {
    std::vector<int> res;
    const bool negative = value < 0;
    unsigned v = (unsigned)std::abs(value);
    for (int i = 0; v > 0; i++) {
        const int zi = (v & 1u) ? 2 - (int)(v & 3u) : 0;
        v = (unsigned)((int)v - zi) >> 1;
        if (zi != 0) res.push_back((negative ? -zi : zi) * (1 << i));
    }
    return res;
}
size_t galois_element_from_step(size_t n, int step) {
    const size_t m = 2 * n;
    if (step == 0) return m - 1;
    const size_t pos = (size_t)(step < 0 ? -step : step);
    if (pos >= (n >> 1)) throw std::invalid_argument("[GaloisTool::get_element_from_step] Step count too large");
    const size_t true_step = step < 0 ? (n >> 1) - pos : pos;
    size_t e = 1;
    for (size_t i = 0; i < true_step; i++) e = (e * 3) & (m - 1);
    return e;
}

// ---- Evaluator --------------------------------------------------------------------------------------------------------------
void Evaluator::check_ct(const char *who, const Ciphertext &c) const {
    if (c.empty() || c.context().get() != context_.get())
        throw std::invalid_argument(std::string(who) + " Ciphertext is empty or belongs to another context.");
}

void Evaluator::negate(const Ciphertext &e, Ciphertext &d) const {
    check_ct("[Evaluator::negate]", e);
    Ciphertext r = Ciphertext::like(e, e.polynomial_count());
    check(b2_negate(context_->raw(), e.data(), r.data(), e.polynomial_count(), e.coeff_modulus_size(), 0, context_->stream), "[Evaluator::negate]");
    d = std::move(r);
}
void Evaluator::negate_inplace(Ciphertext &e) const {
    check_ct("[Evaluator::negate_inplace]", e);
    check(b2_negate(context_->raw(), e.data(), e.data(), e.polynomial_count(), e.coeff_modulus_size(), 0, context_->stream), "[Evaluator::negate_inplace]");
}

// src/evaluator.cu translate: same level, same form, CKKS scales must match; the result has max(size) polynomials
static void check_translate(const Ciphertext &a, const Ciphertext &b, SchemeType s) {
    if (a.coeff_modulus_size() != b.coeff_modulus_size()) throw std::invalid_argument("[Evaluator::translate] Ciphertexts are at different levels.");
    if (a.is_ntt_form() != b.is_ntt_form()) throw std::invalid_argument("[Evaluator::translate] NTT form mismatch.");
    if (s == SchemeType::CKKS && std::fabs(a.scale() - b.scale()) > 1e-9 * std::fabs(a.scale())) throw std::invalid_argument("[Evaluator::translate] Scale mismatch.");
    if (s == SchemeType::BGV && a.correction_factor() != b.correction_factor())
        throw std::invalid_argument("[Evaluator::translate] BGV correction factors differ (not supported by this mirror).");
}
void Evaluator::translate(const Ciphertext &a, const Ciphertext &b, Ciphertext &d, bool subtract) const {
    check_ct("[Evaluator::translate]", a); check_ct("[Evaluator::translate]", b);
    check_translate(a, b, context_->scheme());
    const size_t pa = a.polynomial_count(), pb = b.polynomial_count(), pmin = pa < pb ? pa : pb, pmax = pa < pb ? pb : pa, L = a.coeff_modulus_size();
    const size_t n = context_->poly_modulus_degree();
    Ciphertext r = Ciphertext::like(a, pmax);
    b2_ctx *c = context_->raw(); void *s = context_->stream;
    check((subtract ? b2_sub : b2_add)(c, a.data(), b.data(), r.data(), pmin, L, 0, s), "[Evaluator::translate]");
    if (pa > pb) {          // the surplus polynomials of a are copied
        check(b2_memcpy_d2d(c, r.data() + pmin * L * n, a.data() + pmin * L * n, (pa - pmin) * L * n * 8, s), "[Evaluator::translate]");
    } else if (pb > pa) {   // +-b's surplus polynomials
        if (subtract) check(b2_negate(c, b.data() + pmin * L * n, r.data() + pmin * L * n, pb - pmin, L, 0, s), "[Evaluator::translate]");
        else check(b2_memcpy_d2d(c, r.data() + pmin * L * n, b.data() + pmin * L * n, (pb - pmin) * L * n * 8, s), "[Evaluator::translate]");
    }
    d = std::move(r);
}
void Evaluator::translate_inplace(Ciphertext &a, const Ciphertext &b, bool subtract) const {
    if (a.polynomial_count() < b.polynomial_count()) { Ciphertext d; translate(a, b, d, subtract); a = std::move(d); return; }
    check_ct("[Evaluator::translate_inplace]", a); check_ct("[Evaluator::translate_inplace]", b);
    check_translate(a, b, context_->scheme());
    check((subtract ? b2_sub : b2_add)(context_->raw(), a.data(), b.data(), a.data(), b.polynomial_count(), a.coeff_modulus_size(), 0, context_->stream),
          "[Evaluator::translate_inplace]");
}

static uint64_t mulmod_u64(uint64_t a, uint64_t b, uint64_t m) { return (uint64_t)((unsigned __int128)a * b % m); }

void Evaluator::multiply(const Ciphertext &a, const Ciphertext &b, Ciphertext &d) const {
    check_ct("[Evaluator::multiply]", a); check_ct("[Evaluator::multiply]", b);
    if (a.coeff_modulus_size() != b.coeff_modulus_size()) throw std::invalid_argument("[Evaluator::multiply] Ciphertexts are at different levels.");
    const SchemeType s = context_->scheme();
    const size_t L = a.coeff_modulus_size();
    if (s == SchemeType::BFV) {      // Evaluator::bfv_multiply (src/evaluator.cu:29-116): coefficient form, size-2 operands
        if (a.is_ntt_form() || b.is_ntt_form()) throw std::invalid_argument("[Evaluator::bfv_multiply] Ciphertexts must not be in NTT form.");
        if (a.polynomial_count() != 2 || b.polynomial_count() != 2) throw std::invalid_argument("[Evaluator::bfv_multiply] Ciphertexts must have 2 polynomials.");
        Ciphertext r = Ciphertext::like(a, 3);
        check(b2_bfv_multiply(context_->raw(), a.data(), b.data(), r.data(), 1, L, context_->stream), "[Evaluator::bfv_multiply]");
        d = std::move(r);
        return;
    }
    // ckks_multiply / bgv_multiply (src/evaluator.cu:118-197): NTT form, dyadic convolution
    if (!a.is_ntt_form() || !b.is_ntt_form()) throw std::invalid_argument("[Evaluator::multiply] Ciphertexts must be in NTT form.");
    const size_t pa = a.polynomial_count(), pb = b.polynomial_count();
    if (pa < 2 || pb < 2 || pa > 3 || pb > 3) throw std::invalid_argument("[Evaluator::multiply] Operand sizes must be 2 or 3.");
    Ciphertext r = Ciphertext::like(a, pa + pb - 1);
    check(b2_dyadic_convolute(context_->raw(), a.data(), pa, b.data(), pb, r.data(), 1, L, context_->stream), "[Evaluator::multiply]");
    if (s == SchemeType::CKKS) r.scale() = a.scale() * b.scale();
    if (s == SchemeType::BGV) r.correction_factor() = mulmod_u64(a.correction_factor(), b.correction_factor(), context_->plain_modulus());
    d = std::move(r);
}
void Evaluator::square(const Ciphertext &a, Ciphertext &d) const {
    check_ct("[Evaluator::square]", a);
    if (context_->scheme() == SchemeType::BFV) { multiply(a, a, d); return; }      // bfv_square = the same BEHZ pipeline on one operand
    if (!a.is_ntt_form()) throw std::invalid_argument("[Evaluator::square] Ciphertext must be in NTT form.");
    if (a.polynomial_count() != 2) throw std::invalid_argument("[Evaluator::square] Ciphertext must have 2 polynomials.");
    Ciphertext r = Ciphertext::like(a, 3);
    check(b2_dyadic_square(context_->raw(), a.data(), r.data(), 1, a.coeff_modulus_size(), context_->stream), "[Evaluator::square]");
    if (context_->scheme() == SchemeType::CKKS) r.scale() = a.scale() * a.scale();
    if (context_->scheme() == SchemeType::BGV) r.correction_factor() = mulmod_u64(a.correction_factor(), a.correction_factor(), context_->plain_modulus());
    d = std::move(r);
}

// Evaluator::switch_key_internal (src/evaluator_keyswitching_core.cu:757-1052)
void Evaluator::switch_key_internal(const Ciphertext &e, const uint64_t *target, const KSwitchKeys &keys, size_t index,
                                    SwitchKeyDestinationAssignMethod assign, Ciphertext &destination) const {
    if (!keys.has_index(index)) throw std::invalid_argument("[Evaluator::switch_key_internal] Key switch keys do not contain the requested key.");
    if (context_->key_modulus_size() < 2) throw std::invalid_argument("[Evaluator::switch_key_internal] Keyswitching needs a special prime.");
    if (destination.polynomial_count() < 2 || destination.coeff_modulus_size() != e.coeff_modulus_size())
        throw std::invalid_argument("[Evaluator::switch_key_internal] Destination has the wrong shape.");
    check(b2_keyswitch(context_->raw(), target, e.coeff_modulus_size(), e.is_ntt_form() ? 1 : 0, keys.key(index), destination.data(), (int)assign,
                       context_->stream), "[Evaluator::switch_key_internal]");
}

void Evaluator::apply_keyswitching(const Ciphertext &e, const KSwitchKeys &keys, Ciphertext &d) const {
    check_ct("[Evaluator::apply_keyswitching]", e);
    if (keys.size() != 1) throw std::invalid_argument("[Evaluator::apply_keyswitching_inplace] Key switch keys size must be 1.");
    if (e.polynomial_count() != 2) throw std::invalid_argument("[Evaluator::apply_keyswitching_inplace] Ciphertext polynomial count must be 2.");
    Ciphertext r = e.clone();         // c0 kept; (c0', c1') from the keyswitch of c1: c0 += c0', c1 = c1'
    switch_key_internal(e, e.poly(1), keys, 0, OverwriteExceptFirst, r);
    d = std::move(r);
}
void Evaluator::apply_keyswitching_inplace(Ciphertext &e, const KSwitchKeys &keys) const {
    Ciphertext d; apply_keyswitching(e, keys, d); e = std::move(d);
}

void Evaluator::relinearize(const Ciphertext &e, const RelinKeys &keys, Ciphertext &d) const {
    check_ct("[Evaluator::relinearize]", e);
    size_t size = e.polynomial_count();
    if (size < 2) throw std::invalid_argument("[Evaluator::relinearize_inplace_internal] Destination size must be at least 2 and less/equal to the size of the encrypted polynomial.");
    if (size == 2) { d = e.clone(); return; }
    // destination = (c0, c1) + sum over the higher polynomials of their keyswitch (relinearize_internal, :119-144)
    Ciphertext r = Ciphertext::like(e, 2);
    check(b2_memcpy_d2d(context_->raw(), r.data(), e.data(), 2 * e.coeff_modulus_size() * context_->poly_modulus_degree() * 8, context_->stream), "[Evaluator::relinearize]");
    while (size > 2) {
        switch_key_internal(e, e.poly(size - 1), keys, RelinKeys::get_index(size - 1), AddInplace, r);
        size--;
    }
    d = std::move(r);
}

void Evaluator::relinearize_batched(const std::vector<const Ciphertext *> &es, const RelinKeys &keys, const std::vector<Ciphertext *> &ds) const {
    if (es.size() != ds.size()) throw std::invalid_argument("[Evaluator::relinearize_batched] Encrypted and destination size must be the same.");
    if (es.empty()) return;
    if (!keys.has_index(0)) throw std::invalid_argument("[Evaluator::relinearize_batched] Relin keys do not contain the s^2 key.");
    const size_t L = es[0]->coeff_modulus_size();
    std::vector<Ciphertext> out(es.size());
    std::vector<const uint64_t *> targets(es.size());
    std::vector<uint64_t *> dests(es.size());
    for (size_t i = 0; i < es.size(); i++) {
        check_ct("[Evaluator::relinearize_batched]", *es[i]);
        if (es[i]->polynomial_count() != 3 || es[i]->coeff_modulus_size() != L || es[i]->is_ntt_form() != es[0]->is_ntt_form())
            throw std::invalid_argument("[Evaluator::relinearize_batched] All ciphertexts must have 3 polynomials and share level and form.");
        out[i] = Ciphertext::like(*es[i], 2);
        check(b2_memcpy_d2d(context_->raw(), out[i].data(), es[i]->data(), 2 * L * context_->poly_modulus_degree() * 8, context_->stream), "[Evaluator::relinearize_batched]");
        targets[i] = es[i]->poly(2); dests[i] = out[i].data();
    }
    check(b2_keyswitch_batched(context_->raw(), targets.data(), dests.data(), es.size(), L, es[0]->is_ntt_form() ? 1 : 0, keys.key(0), (int)AddInplace,
                               context_->stream), "[Evaluator::relinearize_batched]");
    for (size_t i = 0; i < es.size(); i++) *ds[i] = std::move(out[i]);
}

// Evaluator::apply_galois (src/evaluator_keyswitching.cu:147-197): sigma(c0), then c1' = sigma(c1) switched back to s
void Evaluator::apply_galois(const Ciphertext &e, size_t galois_element, const GaloisKeys &keys, Ciphertext &d) const {
    check_ct("[Evaluator::apply_galois]", e);
    if (e.polynomial_count() > 2) throw std::invalid_argument("[Evaluator::apply_galois_inplace] Ciphertext size must be 2.");
    if (!keys.has_key(galois_element)) throw std::invalid_argument("[Evaluator::apply_galois_inplace] Galois key not present.");
    Ciphertext r = Ciphertext::like(e, 2);
    check(b2_galois(context_->raw(), e.data(), galois_element, r.data(), 2, e.coeff_modulus_size(), 0, e.is_ntt_form() ? 1 : 0, context_->stream),
          "[Evaluator::apply_galois]");
    // the keyswitch reads sigma(c1) from r.poly(1) and overwrites it: it stages the target before writing the destination
    switch_key_internal(e, r.poly(1), keys, GaloisKeys::get_index(galois_element), OverwriteExceptFirst, r);
    d = std::move(r);
}
void Evaluator::apply_galois_batched(const std::vector<const Ciphertext *> &es, size_t galois_element, const GaloisKeys &keys,
                                     const std::vector<Ciphertext *> &ds) const {
    if (es.size() != ds.size()) throw std::invalid_argument("[Evaluator::apply_galois_batched] Encrypted and destination size must be the same.");
    if (es.empty()) return;
    if (!keys.has_key(galois_element)) throw std::invalid_argument("[Evaluator::apply_galois_inplace] Galois key not present.");
    const size_t L = es[0]->coeff_modulus_size();
    std::vector<Ciphertext> out(es.size());
    std::vector<const uint64_t *> targets(es.size());
    std::vector<uint64_t *> dests(es.size());
    for (size_t i = 0; i < es.size(); i++) {
        check_ct("[Evaluator::apply_galois_batched]", *es[i]);
        if (es[i]->polynomial_count() != 2 || es[i]->coeff_modulus_size() != L || es[i]->is_ntt_form() != es[0]->is_ntt_form())
            throw std::invalid_argument("[Evaluator::apply_galois_batched] All ciphertexts must have 2 polynomials and share level and form.");
        out[i] = Ciphertext::like(*es[i], 2);
        check(b2_galois(context_->raw(), es[i]->data(), galois_element, out[i].data(), 2, L, 0, es[i]->is_ntt_form() ? 1 : 0, context_->stream),
              "[Evaluator::apply_galois_batched]");
        targets[i] = out[i].poly(1); dests[i] = out[i].data();
    }
    check(b2_keyswitch_batched(context_->raw(), targets.data(), dests.data(), es.size(), L, es[0]->is_ntt_form() ? 1 : 0,
                               keys.key(GaloisKeys::get_index(galois_element)), (int)OverwriteExceptFirst, context_->stream), "[Evaluator::apply_galois_batched]");
    for (size_t i = 0; i < es.size(); i++) *ds[i] = std::move(out[i]);
}

// Evaluator::rotate_internal (src/evaluator_keyswitching.cu:263-293): direct key if present, else the NAF decomposition
void Evaluator::rotate_internal(const Ciphertext &e, int steps, const GaloisKeys &keys, Ciphertext &d) const {
    check_ct("[Evaluator::rotate_internal]", e);
    if (steps == 0) { d = e.clone(); return; }
    const size_t n = context_->poly_modulus_degree();
    const size_t element = galois_element_from_step(n, steps);
    if (keys.has_key(element)) { apply_galois(e, element, keys, d); return; }
    const std::vector<int> parts = naf(steps);
    if (parts.size() == 1) throw std::invalid_argument("[Evaluator::rotate_inplace_internal] Galois key not present.");
    Ciphertext acc;
    bool first = true;
    for (int part : parts) {
        Ciphertext tmp;
        rotate_internal(first ? e : acc, part, keys, tmp);
        acc = std::move(tmp);
        first = false;
    }
    d = std::move(acc);
}
void Evaluator::rotate_internal_batched(const std::vector<const Ciphertext *> &es, int steps, const GaloisKeys &keys, const std::vector<Ciphertext *> &ds) const {
    if (es.size() != ds.size()) throw std::invalid_argument("[Evaluator::rotate_inplace_internal] Encrypted and destination size must be the same.");
    if (es.empty()) return;
    if (steps == 0) { for (size_t i = 0; i < es.size(); i++) *ds[i] = es[i]->clone(); return; }
    const size_t element = galois_element_from_step(context_->poly_modulus_degree(), steps);
    if (keys.has_key(element)) { apply_galois_batched(es, element, keys, ds); return; }
    const std::vector<int> parts = naf(steps);
    if (parts.size() == 1) throw std::invalid_argument("[Evaluator::rotate_inplace_internal] Galois key not present.");
    std::vector<Ciphertext> acc(es.size()), tmp(es.size());
    std::vector<const Ciphertext *> src = es;
    for (int part : parts) {
        std::vector<Ciphertext *> dst(es.size());
        for (size_t i = 0; i < es.size(); i++) dst[i] = &tmp[i];
        rotate_internal_batched(src, part, keys, dst);
        for (size_t i = 0; i < es.size(); i++) { acc[i] = std::move(tmp[i]); src[i] = &acc[i]; }
    }
    for (size_t i = 0; i < es.size(); i++) *ds[i] = std::move(acc[i]);
}
void Evaluator::rotate_rows(const Ciphertext &e, int steps, const GaloisKeys &k, Ciphertext &d) const {
    if (context_->scheme() != SchemeType::BFV && context_->scheme() != SchemeType::BGV) throw std::invalid_argument("[Evaluator::rotate_rows] Rotate rows only applies for BFV or BGV");
    rotate_internal(e, steps, k, d);
}
void Evaluator::rotate_vector(const Ciphertext &e, int steps, const GaloisKeys &k, Ciphertext &d) const {
    if (context_->scheme() != SchemeType::CKKS) throw std::invalid_argument("[Evaluator::rotate_vector] Rotate vector only applies for CKKS");
    rotate_internal(e, steps, k, d);
}
void Evaluator::rotate_columns(const Ciphertext &e, const GaloisKeys &k, Ciphertext &d) const {
    if (context_->scheme() != SchemeType::BFV && context_->scheme() != SchemeType::BGV) throw std::invalid_argument("[Evaluator::rotate_columns] Rotate columns only applies for BFV or BGV");
    apply_galois(e, galois_element_from_step(context_->poly_modulus_degree(), 0), k, d);
}
void Evaluator::complex_conjugate(const Ciphertext &e, const GaloisKeys &k, Ciphertext &d) const {
    if (context_->scheme() != SchemeType::CKKS) throw std::invalid_argument("[Evaluator::complex_conjugate] Complex conjugate only applies for CKKS");
    apply_galois(e, galois_element_from_step(context_->poly_modulus_degree(), 0), k, d);
}
void Evaluator::rotate_rows_batched(const std::vector<const Ciphertext *> &es, int steps, const GaloisKeys &k, const std::vector<Ciphertext *> &ds) const {
    if (context_->scheme() != SchemeType::BFV && context_->scheme() != SchemeType::BGV) throw std::invalid_argument("[Evaluator::rotate_rows] Rotate rows only applies for BFV or BGV");
    rotate_internal_batched(es, steps, k, ds);
}
void Evaluator::rotate_vector_batched(const std::vector<const Ciphertext *> &es, int steps, const GaloisKeys &k, const std::vector<Ciphertext *> &ds) const {
    if (context_->scheme() != SchemeType::CKKS) throw std::invalid_argument("[Evaluator::rotate_vector] Rotate vector only applies for CKKS");
    rotate_internal_batched(es, steps, k, ds);
}

// Evaluator::mod_switch_to_next (src/evaluator_modswitch.cu:279-343): BFV / BGV divide-and-round, CKKS drops the limb
void Evaluator::mod_switch_to_next(const Ciphertext &e, Ciphertext &d) const {
    check_ct("[Evaluator::mod_switch_to_next]", e);
    const size_t L = e.coeff_modulus_size();
    if (L < 2) throw std::invalid_argument("[Evaluator::mod_switch_to_next] End of modulus switching chain reached.");
    const SchemeType s = context_->scheme();
    int kind;
    if (s == SchemeType::BFV) { if (e.is_ntt_form()) throw std::invalid_argument("[Evaluator::mod_switch_to_next] BFV ciphertext cannot be in NTT form."); kind = B2_MODSWITCH_BFV; }
    else if (s == SchemeType::BGV) { if (!e.is_ntt_form()) throw std::invalid_argument("[Evaluator::mod_switch_to_next] BGV ciphertext must be in NTT form."); kind = B2_MODSWITCH_BGV; }
    else { if (!e.is_ntt_form()) throw std::invalid_argument("[Evaluator::mod_switch_to_next] CKKS ciphertext must be in NTT form."); kind = B2_MODSWITCH_DROP; }
    Ciphertext r(context_, e.polynomial_count(), L - 1, e.is_ntt_form());
    r.scale() = e.scale(); r.correction_factor() = e.correction_factor();
    check(b2_mod_switch_to_next(context_->raw(), e.data(), r.data(), e.polynomial_count(), L, kind, context_->stream), "[Evaluator::mod_switch_to_next]");
    if (s == SchemeType::BGV) {      // correction_factor *= q_last^-1 mod t (src/evaluator_modswitch.cu:69-73)
        const uint64_t t = context_->plain_modulus(), ql = context_->coeff_modulus()[L - 1] % t;
        // general modular inverse (extended Euclid): the plain modulus need not be prime (the reference uses
        // utils::try_invert_uint64_mod here); a non-invertible q_last is a parameter error
        __int128 r0 = (__int128)t, r1 = (__int128)ql, s0 = 0, s1 = 1;
        while (r1 != 0) { const __int128 qq = r0 / r1, r2 = r0 - qq * r1, s2 = s0 - qq * s1; r0 = r1; r1 = r2; s0 = s1; s1 = s2; }
        if (r0 != 1) throw std::invalid_argument("[Evaluator::mod_switch_to_next] q_last is not invertible modulo the plain modulus.");
        const uint64_t inv = (uint64_t)((s0 % (__int128)t + (__int128)t) % (__int128)t);
        r.correction_factor() = mulmod_u64(e.correction_factor(), inv, t);
    }
    d = std::move(r);
}
// Evaluator::rescale_to_next (src/evaluator_modswitch.cu:445-461): CKKS only, scale /= q_last
void Evaluator::rescale_to_next(const Ciphertext &e, Ciphertext &d) const {
    check_ct("[Evaluator::rescale_to_next]", e);
    if (context_->scheme() != SchemeType::CKKS) throw std::invalid_argument("[Evaluator::rescale_to_next] unsupported scheme");
    if (!e.is_ntt_form()) throw std::invalid_argument("[Evaluator::rescale_to_next] CKKS ciphertext must be in NTT form.");
    const size_t L = e.coeff_modulus_size();
    if (L < 2) throw std::invalid_argument("[Evaluator::rescale_to_next] End of modulus switching chain reached.");
    Ciphertext r(context_, e.polynomial_count(), L - 1, true);
    check(b2_mod_switch_to_next(context_->raw(), e.data(), r.data(), e.polynomial_count(), L, B2_MODSWITCH_CKKS, context_->stream), "[Evaluator::rescale_to_next]");
    r.scale() = e.scale() / (double)context_->coeff_modulus()[L - 1];
    d = std::move(r);
}

void Evaluator::transform_to_ntt_inplace(Ciphertext &e) const {
    check_ct("[Evaluator::transform_to_ntt_inplace]", e);
    if (e.is_ntt_form()) throw std::invalid_argument("[Evaluator::transform_to_ntt_inplace] Ciphertext is already in NTT form.");
    check(b2_ntt_forward(context_->raw(), e.data(), e.data(), 1, e.polynomial_count(), e.coeff_modulus_size(), B2_TABLES_COMPONENTWISE, 0, 0, context_->stream),
          "[Evaluator::transform_to_ntt_inplace]");
    e.is_ntt_form() = true;
}
void Evaluator::transform_from_ntt_inplace(Ciphertext &e) const {
    check_ct("[Evaluator::transform_from_ntt_inplace]", e);
    if (!e.is_ntt_form()) throw std::invalid_argument("[Evaluator::transform_from_ntt_inplace] Ciphertext is not in NTT form.");
    check(b2_ntt_inverse(context_->raw(), e.data(), e.data(), 1, e.polynomial_count(), e.coeff_modulus_size(), B2_TABLES_COMPONENTWISE, 0, 0, context_->stream),
          "[Evaluator::transform_from_ntt_inplace]");
    e.is_ntt_form() = false;
}
static void transform_batched(const HeContextPointer &ctx, const std::vector<Ciphertext *> &es, bool to_ntt) {
    if (es.empty()) return;
    std::vector<const uint64_t *> in(es.size());
    std::vector<uint64_t *> out(es.size());
    for (size_t i = 0; i < es.size(); i++) {
        if (es[i]->empty() || es[i]->is_ntt_form() == to_ntt || es[i]->polynomial_count() != es[0]->polynomial_count() ||
            es[i]->coeff_modulus_size() != es[0]->coeff_modulus_size())
            throw std::invalid_argument("[Evaluator::transform_batched] Ciphertexts must share shape and be in the opposite form.");
        in[i] = es[i]->data(); out[i] = es[i]->data();
    }
    check((to_ntt ? b2_ntt_forward_ptrs : b2_ntt_inverse_ptrs)(ctx->raw(), in.data(), out.data(), es.size(), es[0]->polynomial_count(), es[0]->coeff_modulus_size(),
                                                                B2_TABLES_COMPONENTWISE, 0, 0, ctx->stream), "[Evaluator::transform_batched]");
    for (Ciphertext *e : es) e->is_ntt_form() = to_ntt;
}
void Evaluator::transform_to_ntt_inplace_batched(const std::vector<Ciphertext *> &es) const { transform_batched(context_, es, true); }
void Evaluator::transform_from_ntt_inplace_batched(const std::vector<Ciphertext *> &es) const { transform_batched(context_, es, false); }

// Evaluator::pack_rlwe_ciphertexts_batched / _new (src/evaluator_lwes.cu:305-700): argument checks of the reference, then ONE C-ABI call
void Evaluator::pack_rlwe_ciphertexts_batched(const std::vector<std::vector<const Ciphertext *>> &groups, const GaloisKeys &akeys, size_t shift, size_t input_interval,
                                              size_t output_interval, const std::vector<Ciphertext *> &outputs, bool apply_field_trace) const {
    if (groups.size() != outputs.size()) throw std::invalid_argument("[Evaluator::pack_rlwe_ciphertexts_batched] Input groups and outputs should have same size.");
    if (groups.empty()) return;
    if (groups[0].empty() || !groups[0][0]) throw std::invalid_argument("[Evaluator::pack_rlwe_ciphertexts_batched] Input group 0 is empty.");
    const Ciphertext &first = *groups[0][0];
    check_ct("[Evaluator::pack_rlwe_ciphertexts_batched]", first);
    const size_t n = context_->poly_modulus_degree(), nl = first.coeff_modulus_size();
    auto pow2 = [](size_t v) { return v != 0 && (v & (v - 1)) == 0; };
    if (input_interval > n) throw std::invalid_argument("[Evaluator::pack_rlwe_ciphertexts_batched] input_interval must be less than poly_modulus_degree.");
    if (output_interval > input_interval) throw std::invalid_argument("[Evaluator::pack_rlwe_ciphertexts_batched] output_interval must be less than input_interval.");
    if (!pow2(input_interval)) throw std::invalid_argument("[Evaluator::pack_rlwe_ciphertexts_batched] input_interval must be power of two.");
    if (!pow2(output_interval)) throw std::invalid_argument("[Evaluator::pack_rlwe_ciphertexts_batched] output_interval must be power of two.");
    const size_t M = input_interval / output_interval;
    std::vector<const uint64_t *> ptrs;
    std::vector<size_t> counts;
    for (size_t g = 0; g < groups.size(); g++) {
        if (groups[g].empty()) throw std::invalid_argument("[Evaluator::pack_rlwe_ciphertexts_batched] Input group " + std::to_string(g) + " is empty.");
        if (groups[g].size() > M) throw std::invalid_argument("[Evaluator::pack_rlwe_ciphertexts_batched] Input group " + std::to_string(g) + " has more than input_interval / output_interval.");
        for (const Ciphertext *c : groups[g]) {
            if (!c || c->empty() || c->coeff_modulus_size() != nl) throw std::invalid_argument("[Evaluator::pack_rlwe_ciphertexts_batched] Input has different parms_id.");
            if (c->is_ntt_form() != first.is_ntt_form()) throw std::invalid_argument("[Evaluator::pack_rlwe_ciphertexts_batched] Input has different ntt_form.");
            if (c->polynomial_count() != 2) throw std::invalid_argument("[Evaluator::pack_rlwe_ciphertexts_batched] Input has different polynomial count.");
            if (c->contains_seed()) throw std::invalid_argument("[Evaluator::pack_rlwe_ciphertexts_batched] Ciphertext contains a seed.");
            ptrs.push_back(c->data());
        }
        counts.push_back(groups[g].size());
    }
    // the automorphism keys of the tree and of the trace
    size_t layers = 0;
    while (((size_t)1 << layers) < M) layers++;
    std::vector<uint64_t> elts;
    for (size_t l = 0; l < layers; l++) elts.push_back((n / input_interval) * ((uint64_t)1 << (l + 1)) + 1);
    if (output_interval != 1 && apply_field_trace) for (size_t deg = n; deg > n / output_interval; deg >>= 1) elts.push_back(deg + 1);
    std::vector<const b2_kswitch_key *> keys;
    for (uint64_t e : elts) {
        if (!akeys.has_key((size_t)e)) throw std::invalid_argument("[GaloisKeys::key] Galois key not present");
        keys.push_back(akeys.key(GaloisKeys::get_index((size_t)e)));
    }
    const bool out_ntt = context_->scheme() != SchemeType::BFV;
    std::vector<Ciphertext> res(groups.size());
    std::vector<uint64_t *> op(groups.size());
    for (size_t g = 0; g < groups.size(); g++) {
        res[g] = Ciphertext::like(*groups[g][0], 2);
        res[g].is_ntt_form() = out_ntt;
        op[g] = res[g].data();
    }
    check(b2_pack_rlwes(context_->raw(), ptrs.data(), counts.data(), groups.size(), first.is_ntt_form() ? 1 : 0, nl, shift, input_interval, output_interval, elts.data(), keys.data(),
                        elts.size(), apply_field_trace ? 1 : 0, out_ntt ? 1 : 0, op.data(), context_->stream), "[Evaluator::pack_rlwe_ciphertexts_batched]");
    for (size_t g = 0; g < groups.size(); g++) *outputs[g] = std::move(res[g]);
}
Ciphertext Evaluator::pack_rlwe_ciphertexts_new(const std::vector<const Ciphertext *> &ciphers, const GaloisKeys &akeys, size_t shift, size_t input_interval, size_t output_interval,
                                                bool apply_field_trace) const {
    Ciphertext out;
    pack_rlwe_ciphertexts_batched({ciphers}, akeys, shift, input_interval, output_interval, {&out}, apply_field_trace);
    return out;
}

// ---- LWE (src/lwe_ciphertext.cu, src/evaluator_lwes.cu:51-94, 201-303) ------------------------------------------------------------------
LWECiphertext::LWECiphertext(const HeContextPointer &ctx, size_t coeff_modulus_size)
    : ctx_(ctx), limbs_(coeff_modulus_size), c0_(ctx, (coeff_modulus_size + 1) & ~(size_t)1), c1_(ctx, coeff_modulus_size * ctx->poly_modulus_degree()) {}
std::vector<uint64_t> LWECiphertext::c0_to_host() const { std::vector<uint64_t> v(limbs_); if (limbs_) c0_.download(v.data(), limbs_); return v; }
std::vector<uint64_t> LWECiphertext::c1_to_host() const { std::vector<uint64_t> v(c1_.size()); if (!v.empty()) c1_.download(v.data(), v.size()); return v; }
Ciphertext LWECiphertext::assemble_lwe() const {
    if (empty()) throw std::invalid_argument("[LWECiphertext::assemble_lwe] LWE ciphertext is empty.");
    Ciphertext r(ctx_, 2, limbs_, false);
    r.scale() = scale_; r.correction_factor() = cf_;
    const uint64_t *p0 = c0_.data(), *p1 = c1_.data();
    uint64_t *o = r.data();
    check(b2_assemble_lwes(ctx_->raw(), &p0, &p1, &o, 1, limbs_, ctx_->stream), "[LWECiphertext::assemble_lwe]");
    return r;
}
LWECiphertext Evaluator::extract_lwe_new(const Ciphertext &e, size_t term) const {
    check_ct("[Evaluator::extract_lwe_new]", e);
    if (e.contains_seed()) throw std::invalid_argument("[Evaluator::extract_lwe_new] Ciphertext contains a seed.");
    if (e.polynomial_count() != 2) throw std::invalid_argument("[Evaluator::extract_lwe_new] Ciphertext size must be 2.");
    if (e.is_ntt_form()) {          // evaluator_lwes.cu:56-60
        Ciphertext t;
        transform_from_ntt(e, t);
        return extract_lwe_new(t, term);
    }
    if (term >= context_->poly_modulus_degree()) throw std::invalid_argument("[Evaluator::extract_lwe_new] term out of range.");
    LWECiphertext r(context_, e.coeff_modulus_size());
    r.scale() = e.scale(); r.correction_factor() = e.correction_factor();
    check(b2_extract_lwe(context_->raw(), e.data(), term, e.coeff_modulus_size(), r.c0(), r.c1(), context_->stream), "[Evaluator::extract_lwe_new]");
    return r;
}
void Evaluator::pack_lwe_ciphertexts_batched(const std::vector<std::vector<const LWECiphertext *>> &groups, const GaloisKeys &akeys, const std::vector<Ciphertext *> &outputs,
                                             bool apply_field_trace) const {
    if (groups.size() != outputs.size()) throw std::invalid_argument("[Evaluator::pack_lwe_ciphertexts_batched] Input groups and outputs should have same size.");
    if (groups.empty()) return;
    const size_t n = context_->poly_modulus_degree();
    std::vector<const uint64_t *> c0s, c1s;
    std::vector<size_t> counts;
    size_t nl = 0, maxc = 0;
    for (const auto &g : groups) {
        if (g.empty()) throw std::invalid_argument("[Evaluator::pack_lwe_ciphertexts_new] LWE ciphertexts must not be empty.");
        if (g.size() > n) throw std::invalid_argument("[Evaluator::pack_lwe_ciphertexts_new] LWE ciphertexts count must be less than poly_modulus_degree.");
        for (const LWECiphertext *l : g) {
            if (!l || l->empty()) throw std::invalid_argument("[Evaluator::pack_lwe_ciphertexts_new] LWE ciphertexts must not be empty.");
            if (nl == 0) nl = l->coeff_modulus_size();
            if (l->coeff_modulus_size() != nl || l->context().get() != context_.get()) throw std::invalid_argument("[Evaluator::pack_lwe_ciphertexts_new] LWE ciphertexts must have same parms id.");
            c0s.push_back(l->c0()); c1s.push_back(l->c1());
        }
        counts.push_back(g.size()); maxc = std::max(maxc, g.size());
    }
    size_t l2 = 0;
    while (((size_t)1 << l2) < maxc) l2++;
    const size_t output_interval = n >> l2;
    std::vector<uint64_t> elts;
    for (size_t l = 0; l < l2; l++) elts.push_back(((uint64_t)1 << (l + 1)) + 1);
    if (output_interval != 1 && apply_field_trace) for (size_t deg = n; deg > n / output_interval; deg >>= 1) elts.push_back(deg + 1);
    std::vector<const b2_kswitch_key *> keys;
    for (uint64_t e : elts) {
        if (!akeys.has_key((size_t)e)) throw std::invalid_argument("[GaloisKeys::key] Galois key not present");
        keys.push_back(akeys.key(GaloisKeys::get_index((size_t)e)));
    }
    const bool out_ntt = context_->scheme() != SchemeType::BFV;
    std::vector<Ciphertext> res(groups.size());
    std::vector<uint64_t *> op(groups.size());
    for (size_t g = 0; g < groups.size(); g++) {
        res[g] = Ciphertext(context_, 2, nl, out_ntt);
        res[g].scale() = groups[g][0]->scale(); res[g].correction_factor() = groups[g][0]->correction_factor();
        op[g] = res[g].data();
    }
    check(b2_pack_lwes(context_->raw(), c0s.data(), c1s.data(), counts.data(), groups.size(), nl, elts.data(), keys.data(), elts.size(), apply_field_trace ? 1 : 0, out_ntt ? 1 : 0,
                       op.data(), context_->stream), "[Evaluator::pack_lwe_ciphertexts_batched]");
    for (size_t g = 0; g < groups.size(); g++) *outputs[g] = std::move(res[g]);
}
Ciphertext Evaluator::pack_lwe_ciphertexts_new(const std::vector<const LWECiphertext *> &lwes, const GaloisKeys &akeys, bool apply_field_trace) const {
    Ciphertext out;
    pack_lwe_ciphertexts_batched({lwes}, akeys, {&out}, apply_field_trace);
    return out;
}

// bfv_centralize / transform_plain_to_ntt (src/evaluator_transform_ntt.cu:35-129)
Plaintext Evaluator::centralize(const Plaintext &p, size_t nl, uint64_t scalar) const {
    if (p.empty() || !p.is_mod_t()) throw std::invalid_argument("[Evaluator::bfv_centralize] Plaintext must be modulo t (no parms_id).");
    if (context_->plain_modulus() < 2) throw std::invalid_argument("[Evaluator::bfv_centralize] The context has no plain modulus.");
    Plaintext r = Plaintext::rns(context_, nl, false);
    check(b2_centralize_scaled(context_->raw(), p.data(), p.coeff_count(), nl, scalar, r.data(), context_->stream), "[Evaluator::bfv_centralize]");
    return r;
}
Plaintext Evaluator::transform_plain_to_ntt(const Plaintext &p, size_t nl) const {
    Plaintext c = centralize(p, nl);
    Plaintext r = Plaintext::rns(context_, nl, true);
    check(b2_ntt_forward(context_->raw(), c.data(), r.data(), 1, 1, nl, B2_TABLES_COMPONENTWISE, 0, 0, context_->stream), "[Evaluator::transform_plain_to_ntt]");
    check(b2_stream_sync(context_->raw(), context_->stream), "[Evaluator::transform_plain_to_ntt]");      // the temporary is freed on return
    return r;
}

// Evaluator::translate_plain (src/evaluator_translate_plain.cu:95-170)
void Evaluator::translate_plain(const Ciphertext &e, const Plaintext &p, Ciphertext &d, bool subtract) const {
    check_ct("[Evaluator::translate_plain]", e);
    if (e.contains_seed()) throw std::invalid_argument("[Evaluator::translate_plain] Ciphertext contains a seed; expand it first.");
    const SchemeType scheme = context_->scheme();
    if (scheme == SchemeType::BGV) {          // plain * correction_factor mod t -> centralize -> NTT -> c_0 +/- it     (evaluator_translate_plain.cu:79-88)
        if (!e.is_ntt_form()) throw std::invalid_argument("[Evaluator::translate_plain] Ciphertext is not in NTT form.");
        if (p.is_ntt_form() || !p.is_mod_t()) throw std::invalid_argument("[Evaluator::translate_plain] Plaintext is in NTT form.");
        const size_t L = e.coeff_modulus_size();
        Plaintext lifted = centralize(p, L, e.correction_factor());
        b2_ctx *c = context_->raw();
        check(b2_ntt_forward(c, lifted.data(), lifted.data(), 1, 1, L, B2_TABLES_COMPONENTWISE, 0, 0, context_->stream), "[Evaluator::translate_plain]");
        Ciphertext r = e.clone();
        check((subtract ? b2_sub : b2_add)(c, r.poly(0), lifted.data(), r.poly(0), 1, L, 0, context_->stream), "[Evaluator::translate_plain]");
        check(b2_stream_sync(c, context_->stream), "[Evaluator::translate_plain]");
        d = std::move(r);
        return;
    }
    if (e.is_ntt_form() != p.is_ntt_form()) throw std::invalid_argument("[Evaluator::translate_plain] Plaintext and ciphertext are not in the same NTT form.");
    if (scheme == SchemeType::BFV && p.is_mod_t() && e.is_ntt_form()) throw std::invalid_argument("[Evaluator::translate_plain] When plain is mod t, encrypted must not be in NTT form.");
    if (scheme == SchemeType::CKKS) {
        if (!e.is_ntt_form()) throw std::invalid_argument("[Evaluator::translate_plain_inplace] Ciphertext is not in NTT form.");
        if (std::fabs(p.scale() - e.scale()) > 1e-9 * std::fabs(e.scale())) throw std::invalid_argument("[Evaluator::translate_plain] Plaintext scale is not equal to the scale of the ciphertext.");
    }
    if (!p.is_mod_t() && p.coeff_modulus_size() != e.coeff_modulus_size()) throw std::invalid_argument("[Evaluator::translate_plain] Plaintext and ciphertext parameters do not match.");
    const size_t L = e.coeff_modulus_size();
    b2_ctx *c = context_->raw();
    Ciphertext r = e.clone();
    if (p.is_mod_t()) check(b2_bfv_scale_up(c, p.data(), p.coeff_count(), L, r.poly(0), subtract ? 1 : 0, r.poly(0), context_->stream), "[Evaluator::translate_plain]");
    else if (subtract) check(b2_sub(c, r.poly(0), p.data(), r.poly(0), 1, L, 0, context_->stream), "[Evaluator::translate_plain]");
    else check(b2_add(c, r.poly(0), p.data(), r.poly(0), 1, L, 0, context_->stream), "[Evaluator::translate_plain]");
    d = std::move(r);
}

// Evaluator::multiply_plain (src/evaluator_multiply_plain.cu:309-328): the four form combinations of the reference.  A Plaintext of this
// mirror is RNS-limbed (what transform_plain_to_ntt / the encoders hand to the evaluator), so the coefficient-form product is
// NTT(ct), NTT(pt), dyadic product, INTT (multiply_plain_normal, :13-68, after its centralize step).
void Evaluator::multiply_plain(const Ciphertext &e, const Plaintext &p, Ciphertext &d) const {
    check_ct("[Evaluator::multiply_plain]", e);
    if (p.is_mod_t()) {      // a message modulo t: centralize it to the ciphertext's level first (multiply_plain_normal, evaluator_multiply_plain.cu:13-68)
        Plaintext lifted = centralize(p, e.coeff_modulus_size());
        multiply_plain(e, lifted, d);
        check(b2_stream_sync(context_->raw(), context_->stream), "[Evaluator::multiply_plain]");
        return;
    }
    if (p.coeff_modulus_size() != e.coeff_modulus_size()) throw std::invalid_argument("[Evaluator::multiply_plain_ntt] Plaintext parms_id mismatch.");
    const size_t pc = e.polynomial_count(), L = e.coeff_modulus_size();
    b2_ctx *c = context_->raw();
    Ciphertext r = Ciphertext::like(e, pc);
    const uint64_t *ct = e.data(), *pt = p.data();
    DeviceArray pt_ntt;
    if (!e.is_ntt_form()) {          // transform a copy of the ciphertext (into the result buffer)
        check(b2_ntt_forward(c, e.data(), r.data(), 1, pc, L, B2_TABLES_COMPONENTWISE, 0, 0, context_->stream), "[Evaluator::multiply_plain]");
        ct = r.data();
    }
    if (!p.is_ntt_form()) {
        pt_ntt = DeviceArray(context_, L * context_->poly_modulus_degree());
        check(b2_ntt_forward(c, p.data(), pt_ntt.data(), 1, 1, L, B2_TABLES_COMPONENTWISE, 0, 0, context_->stream), "[Evaluator::multiply_plain]");
        pt = pt_ntt.data();
    }
    check(b2_dyadic_broadcast(c, ct, pt, r.data(), pc, L, 0, context_->stream), "[Evaluator::multiply_plain]");
    if (!e.is_ntt_form()) check(b2_ntt_inverse(c, r.data(), r.data(), 1, pc, L, B2_TABLES_COMPONENTWISE, 0, 0, context_->stream), "[Evaluator::multiply_plain]");
    if (pt_ntt.data()) check(b2_stream_sync(c, context_->stream), "[Evaluator::multiply_plain]");      // the temporary is freed on scope exit
    d = std::move(r);
}
static void check_plain_batch(const std::vector<const Ciphertext *> &es, const std::vector<const Plaintext *> &ps) {
    for (size_t i = 0; i < es.size(); i++) {
        if (!es[i] || !ps[i] || es[i]->empty()) throw std::invalid_argument("[Evaluator::multiply_plain_batched] Empty operand.");
        if (!es[i]->is_ntt_form() || !ps[i]->is_ntt_form())
            throw std::invalid_argument("[Evaluator::multiply_plain_batched] The batched forms cover NTT-form operands (multiply_plain_ntt_batched); transform the batch first.");
        if (es[i]->polynomial_count() != es[0]->polynomial_count() || es[i]->coeff_modulus_size() != es[0]->coeff_modulus_size() ||
            ps[i]->coeff_modulus_size() != es[0]->coeff_modulus_size())
            throw std::invalid_argument("[Evaluator::multiply_plain_batched] Operand shape mismatch.");
    }
}
// multiply_plain_ntt_batched (:216-256): ONE launch over the scattered operands
void Evaluator::multiply_plain_batched(const std::vector<const Ciphertext *> &es, const std::vector<const Plaintext *> &ps, const std::vector<Ciphertext *> &ds) const {
    if (es.size() != ps.size() || es.size() != ds.size()) throw std::invalid_argument("[Evaluator::multiply_plain_batched] Size mismatch.");
    if (es.empty()) return;
    check_plain_batch(es, ps);
    std::vector<Ciphertext> out(es.size());
    std::vector<const uint64_t *> c(es.size()), p(es.size());
    std::vector<uint64_t *> o(es.size());
    for (size_t i = 0; i < es.size(); i++) {
        out[i] = Ciphertext::like(*es[i], es[i]->polynomial_count());
        c[i] = es[i]->data(); p[i] = ps[i]->data(); o[i] = out[i].data();
    }
    check(b2_multiply_plain_accumulate(context_->raw(), c.data(), p.data(), o.data(), es.size(), es[0]->polynomial_count(), es[0]->coeff_modulus_size(), 0, 0, context_->stream),
          "[Evaluator::multiply_plain_batched]");
    for (size_t i = 0; i < es.size(); i++) *ds[i] = std::move(out[i]);
}
// multiply_plain_ntt_accumulate (:258-307): destination[i] (+)= encrypted[i] * plain[i]; destinations may REPEAT (all terms with the
// same destination are summed into it); ONE launch.  set_zero: the destinations are (re)created and start from zero.
void Evaluator::multiply_plain_accumulate(const std::vector<const Ciphertext *> &es, const std::vector<const Plaintext *> &ps, const std::vector<Ciphertext *> &ds,
                                          bool set_zero) const {
    if (es.size() != ps.size() || es.size() != ds.size()) throw std::invalid_argument("[Evaluator::multiply_plain_ntt_batched] Input vectors have different sizes.");
    if (es.empty()) return;
    check_plain_batch(es, ps);
    std::vector<const uint64_t *> c(es.size()), p(es.size());
    std::vector<uint64_t *> o(es.size());
    for (size_t i = 0; i < es.size(); i++) {
        if (!ds[i]) throw std::invalid_argument("[Evaluator::multiply_plain_accumulate] Null destination.");
        if (set_zero) {
            bool seen = false;
            for (size_t k = 0; k < i && !seen; k++) seen = ds[k] == ds[i];
            if (!seen) *ds[i] = Ciphertext::like(*es[i], es[i]->polynomial_count());
        } else if (ds[i]->empty() || ds[i]->coeff_modulus_size() != es[i]->coeff_modulus_size() || ds[i]->polynomial_count() != es[i]->polynomial_count() ||
                   ds[i]->is_ntt_form() != es[i]->is_ntt_form()) {
            throw std::invalid_argument("[Evaluator::multiply_plain_normal_accumulate] Destination parameters do not match.");
        }
    }
    for (size_t i = 0; i < es.size(); i++) { c[i] = es[i]->data(); p[i] = ps[i]->data(); o[i] = ds[i]->data(); }
    check(b2_multiply_plain_accumulate(context_->raw(), c.data(), p.data(), o.data(), es.size(), es[0]->polynomial_count(), es[0]->coeff_modulus_size(), 0,
                                       set_zero ? 0 : 1, context_->stream), "[Evaluator::multiply_plain_accumulate]");
}
void Evaluator::multiply_plain_accumulate(const std::vector<const Ciphertext *> &es, const std::vector<const Plaintext *> &ps, Ciphertext &d, bool set_zero) const {
    if (es.size() != ps.size()) throw std::invalid_argument("[Evaluator::multiply_plain_accumulate] Size mismatch.");
    if (es.empty()) return;
    std::vector<Ciphertext *> ds(es.size(), &d);
    multiply_plain_accumulate(es, ps, ds, set_zero || d.empty());
}

void Evaluator::negacyclic_shift(const Ciphertext &e, size_t shift, Ciphertext &d) const {
    check_ct("[Evaluator::negacyclic_shift]", e);
    if (e.is_ntt_form()) throw std::invalid_argument("[Evaluator::negacyclic_shift] Ciphertext must not be in NTT form.");
    Ciphertext r = Ciphertext::like(e, e.polynomial_count());
    check(b2_negacyclic_shift(context_->raw(), e.data(), shift, r.data(), e.polynomial_count(), e.coeff_modulus_size(), 0, context_->stream), "[Evaluator::negacyclic_shift]");
    d = std::move(r);
}

}  // namespace b2troy
