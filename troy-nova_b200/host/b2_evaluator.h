// b2_evaluator.h — C++ host mirror of Troy-Nova's SEAL-style surface for the RNS hot path (SURVEY.md §8a rows K6, X1 and
// the public Evaluator methods above them).  Pure C++17: it includes no CUDA header and reaches the GPU only through the
// C ABI of libb2rns.so (include/b2rns.h) — "host code stays C++ and calls CUDA through a thin C-ABI layer".
//
// It mirrors, with the reference's names, argument meaning and error behaviour (std::invalid_argument on shape / form /
// key mismatches, cf. src/evaluator.cu, src/evaluator_keyswitching.cu:11-361, src/evaluator_modswitch.cu,
// src/evaluator_transform_ntt.cu:469-653, src/evaluator_multiply_plain.cu):
//   HeContext, Ciphertext, Plaintext, KSwitchKeys / RelinKeys / GaloisKeys, and
//   Evaluator::{negate, add, sub, multiply, square, relinearize, apply_keyswitching, apply_galois, rotate_rows, rotate_vector,
//               rotate_columns, complex_conjugate, mod_switch_to_next, rescale_to_next, transform_to_ntt, transform_from_ntt,
//               multiply_plain, negacyclic_shift} with their _inplace / _new / _batched forms where the reference has them.
// Encryptor / Decryptor / RandomGenerator and the key holders are in b2_crypto.h.  Not here (SURVEY §8f "next" rows): encoders,
// KeyGenerator, serialisation.  Keys and ciphertexts are imported as raw RNS words in the reference's own
// layouts (ciphertext [poly][limb][N], src/ciphertext.h:211-252; key digit j = [2][K][N], src/kswitch_keys.h:36-57).
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/b2rns.h"

namespace b2troy {

enum class SchemeType { Nil = 0, BFV = 1, CKKS = 2, BGV = 3 };     // src/encryption_parameters.h:7-12

// throws std::runtime_error carrying b2_last_error() when a C-ABI call fails (the reference throws from its CUDA wrappers)
void check(b2_status st, const char *where);

class HeContext;
typedef std::shared_ptr<HeContext> HeContextPointer;

// Device-resident tables for one key-level modulus chain q_0 .. q_{K-2}, P.  (HeContext::create + to_device_inplace,
// src/he_context.cu:96-145; parameter validation beyond "prime, = 1 mod 2N, < 2^61" is the reference's business.)
class HeContext {
public:
    static HeContextPointer create(SchemeType scheme, size_t poly_modulus_degree, const std::vector<uint64_t> &coeff_modulus,
                                   uint64_t plain_modulus = 0, int device = 0);
    ~HeContext();
    HeContext(const HeContext &) = delete;
    HeContext &operator=(const HeContext &) = delete;

    b2_ctx *raw() const { return ctx_; }
    SchemeType scheme() const { return scheme_; }
    size_t poly_modulus_degree() const { return n_; }
    size_t key_modulus_size() const { return moduli_.size(); }            // K
    size_t first_modulus_size() const { return moduli_.size() > 1 ? moduli_.size() - 1 : 1; }   // limbs of a fresh ciphertext
    const std::vector<uint64_t> &coeff_modulus() const { return moduli_; }
    uint64_t plain_modulus() const { return t_; }
    bool on_device() const { return true; }                               // there is no host branch
    void *stream = nullptr;                                               // cudaStream_t every call is enqueued on
    void synchronize() const;

private:
    HeContext() = default;
    b2_ctx *ctx_ = nullptr;
    SchemeType scheme_ = SchemeType::Nil;
    size_t n_ = 0;
    std::vector<uint64_t> moduli_;
    uint64_t t_ = 0;
};

// RAII device allocation (the reference's utils::DynamicArray on a MemoryPool, src/utils/dynamic_array.h)
class DeviceArray {
public:
    DeviceArray() = default;
    DeviceArray(const HeContextPointer &ctx, size_t words);
    ~DeviceArray();
    DeviceArray(DeviceArray &&o) noexcept { *this = std::move(o); }
    DeviceArray &operator=(DeviceArray &&o) noexcept;
    DeviceArray(const DeviceArray &) = delete;
    DeviceArray &operator=(const DeviceArray &) = delete;
    uint64_t *data() { return p_; }
    const uint64_t *data() const { return p_; }
    size_t size() const { return words_; }
    DeviceArray clone() const;
    void upload(const uint64_t *host, size_t words);
    void download(uint64_t *host, size_t words) const;

private:
    HeContextPointer ctx_;
    uint64_t *p_ = nullptr;
    size_t words_ = 0;
};

// src/ciphertext.h: polynomial_count x coeff_modulus_size x N words, NTT-form flag, CKKS scale, BGV correction factor.
class Ciphertext {
public:
    Ciphertext() = default;
    Ciphertext(const HeContextPointer &ctx, size_t polynomial_count, size_t coeff_modulus_size, bool is_ntt_form);
    static Ciphertext like(const Ciphertext &other, size_t polynomial_count);
    static Ciphertext from_host(const HeContextPointer &ctx, const uint64_t *words, size_t polynomial_count, size_t coeff_modulus_size,
                                bool is_ntt_form, double scale = 1.0, uint64_t correction_factor = 1);
    Ciphertext clone() const;
    std::vector<uint64_t> to_host() const;

    const HeContextPointer &context() const { return ctx_; }
    size_t polynomial_count() const { return polys_; }
    size_t coeff_modulus_size() const { return limbs_; }
    size_t poly_modulus_degree() const { return ctx_ ? ctx_->poly_modulus_degree() : 0; }
    bool is_ntt_form() const { return ntt_; }
    bool &is_ntt_form() { return ntt_; }
    double scale() const { return scale_; }
    double &scale() { return scale_; }
    uint64_t correction_factor() const { return cf_; }
    uint64_t &correction_factor() { return cf_; }
    uint64_t seed() const { return seed_; }              // != 0: c_1 is to be re-expanded from this seed (src/ciphertext.h contains_seed)
    uint64_t &seed() { return seed_; }
    bool contains_seed() const { return seed_ != 0; }
    uint64_t *data() { return buf_.data(); }
    const uint64_t *data() const { return buf_.data(); }
    uint64_t *poly(size_t k) { return buf_.data() + k * limbs_ * poly_modulus_degree(); }
    const uint64_t *poly(size_t k) const { return buf_.data() + k * limbs_ * poly_modulus_degree(); }
    bool empty() const { return buf_.size() == 0; }

private:
    HeContextPointer ctx_;
    size_t polys_ = 0, limbs_ = 0;
    bool ntt_ = false;
    double scale_ = 1.0;
    uint64_t cf_ = 1, seed_ = 0;
    DeviceArray buf_;
};

// src/plaintext.h, the two forms the reference has:
//   RNS form  (parms_id set)      [coeff_modulus_size][N] words, coefficient or NTT form — what multiply_plain / CKKS encryption consume
//   mod-t form (parms_id zero)    coeff_count <= N words modulo the plain modulus — BFV / BGV messages, what Decryptor returns
class Plaintext {
public:
    Plaintext() = default;
    static Plaintext from_host(const HeContextPointer &ctx, const uint64_t *words, size_t coeff_modulus_size, bool is_ntt_form);
    static Plaintext from_host_mod_t(const HeContextPointer &ctx, const uint64_t *words, size_t coeff_count);
    static Plaintext rns(const HeContextPointer &ctx, size_t coeff_modulus_size, bool is_ntt_form);      // uninitialised
    static Plaintext mod_t(const HeContextPointer &ctx, size_t coeff_count);                             // uninitialised
    std::vector<uint64_t> to_host() const;
    const HeContextPointer &context() const { return ctx_; }
    bool is_mod_t() const { return limbs_ == 0; }                        // the reference's parms_id == parms_id_zero
    size_t coeff_modulus_size() const { return limbs_; }
    size_t coeff_count() const { return coeff_count_; }
    bool is_ntt_form() const { return ntt_; }
    double scale() const { return scale_; }
    double &scale() { return scale_; }
    const uint64_t *data() const { return buf_.data(); }
    uint64_t *data() { return buf_.data(); }
    bool empty() const { return buf_.size() == 0; }

private:
    HeContextPointer ctx_;
    size_t limbs_ = 0, coeff_count_ = 0;
    bool ntt_ = false;
    double scale_ = 1.0;
    DeviceArray buf_;
};

// src/lwe_ciphertext.h: c0 = coeff_modulus_size scalars, c1 = [coeff_modulus_size][N]; produced by Evaluator::extract_lwe_new, consumed by
// assemble_lwe / Evaluator::pack_lwe_ciphertexts_new
class LWECiphertext {
public:
    LWECiphertext() = default;
    LWECiphertext(const HeContextPointer &ctx, size_t coeff_modulus_size);
    Ciphertext assemble_lwe() const;                                     // RLWE with c_0 = the constant polynomial c0, c_1 = c1, coefficient form
    const HeContextPointer &context() const { return ctx_; }
    size_t coeff_modulus_size() const { return limbs_; }
    double scale() const { return scale_; }
    double &scale() { return scale_; }
    uint64_t correction_factor() const { return cf_; }
    uint64_t &correction_factor() { return cf_; }
    const uint64_t *c0() const { return c0_.data(); }
    uint64_t *c0() { return c0_.data(); }
    const uint64_t *c1() const { return c1_.data(); }
    uint64_t *c1() { return c1_.data(); }
    std::vector<uint64_t> c0_to_host() const;
    std::vector<uint64_t> c1_to_host() const;
    bool empty() const { return c1_.size() == 0; }

private:
    HeContextPointer ctx_;
    size_t limbs_ = 0;
    double scale_ = 1.0;
    uint64_t cf_ = 1;
    DeviceArray c0_, c1_;
};

// src/kswitch_keys.h: keys()[index] = vector of `digits` public keys, each a 2-polynomial key-level NTT-form ciphertext.
class KSwitchKeys {
public:
    KSwitchKeys() = default;
    explicit KSwitchKeys(const HeContextPointer &ctx) : ctx_(ctx) {}
    ~KSwitchKeys();
    KSwitchKeys(KSwitchKeys &&) noexcept;
    KSwitchKeys &operator=(KSwitchKeys &&) noexcept;
    KSwitchKeys(const KSwitchKeys &) = delete;
    KSwitchKeys &operator=(const KSwitchKeys &) = delete;
    // words: [digits][2][K][N] (host memory), uploaded and kept resident
    void set_key(size_t index, const uint64_t *words, size_t digits);
    // the same from a device array that already holds [digits][2][K][N] (KeyGenerator)
    void adopt_key(size_t index, DeviceArray &&storage, size_t digits);
    std::vector<uint64_t> key_to_host(size_t index) const;
    bool has_index(size_t index) const { return index < keys_.size() && keys_[index] != nullptr; }
    const b2_kswitch_key *key(size_t index) const { return has_index(index) ? keys_[index] : nullptr; }
    size_t digits(size_t index) const { return index < digits_.size() ? digits_[index] : 0; }
    size_t size() const { return keys_.size(); }
    const HeContextPointer &context() const { return ctx_; }

protected:
    HeContextPointer ctx_;
    std::vector<b2_kswitch_key *> keys_;
    std::vector<size_t> digits_;
    std::vector<DeviceArray> storage_;
};
// RelinKeys: index 0 holds the key for s^2 (src/kswitch_keys.h:92-116, get_index(key_power) = key_power - 2)
class RelinKeys : public KSwitchKeys {
public:
    using KSwitchKeys::KSwitchKeys;
    static size_t get_index(size_t key_power) {
        if (key_power < 2) throw std::invalid_argument("[RelinKeys::get_index] key_power must be at least 2.");
        return key_power - 2;
    }
};
// GaloisKeys: index (galois_element - 1) / 2 (src/kswitch_keys.h:130-160)
class GaloisKeys : public KSwitchKeys {
public:
    using KSwitchKeys::KSwitchKeys;
    static size_t get_index(size_t galois_element) {
        if ((galois_element & 1) == 0) throw std::invalid_argument("[GaloisKeys::get_index] Galois element must be odd.");
        return (galois_element - 1) >> 1;
    }
    bool has_key(size_t galois_element) const { return has_index(get_index(galois_element)); }
    void set_galois_key(size_t galois_element, const uint64_t *words, size_t digits) { set_key(get_index(galois_element), words, digits); }
};

// utils::naf (src/utils/number_theory.cu:6-20) and GaloisTool::get_element_from_step (src/utils/galois.cu:44-63)
std::vector<int> naf(int value);
size_t galois_element_from_step(size_t poly_modulus_degree, int step);

class Evaluator {
public:
    enum SwitchKeyDestinationAssignMethod { Overwrite = 0, AddInplace = 1, OverwriteExceptFirst = 2 };    // src/evaluator.h:18-22

    explicit Evaluator(HeContextPointer context) : context_(std::move(context)) {}
    const HeContextPointer &context() const { return context_; }
    bool on_device() const { return true; }

    // ---- element-wise (src/evaluator.cu:425-640)
    void negate(const Ciphertext &encrypted, Ciphertext &destination) const;
    void negate_inplace(Ciphertext &encrypted) const;
    Ciphertext negate_new(const Ciphertext &encrypted) const { Ciphertext d; negate(encrypted, d); return d; }
    void add(const Ciphertext &encrypted1, const Ciphertext &encrypted2, Ciphertext &destination) const { translate(encrypted1, encrypted2, destination, false); }
    void sub(const Ciphertext &encrypted1, const Ciphertext &encrypted2, Ciphertext &destination) const { translate(encrypted1, encrypted2, destination, true); }
    void add_inplace(Ciphertext &encrypted1, const Ciphertext &encrypted2) const { translate_inplace(encrypted1, encrypted2, false); }
    void sub_inplace(Ciphertext &encrypted1, const Ciphertext &encrypted2) const { translate_inplace(encrypted1, encrypted2, true); }
    Ciphertext add_new(const Ciphertext &a, const Ciphertext &b) const { Ciphertext d; add(a, b, d); return d; }
    Ciphertext sub_new(const Ciphertext &a, const Ciphertext &b) const { Ciphertext d; sub(a, b, d); return d; }

    // ---- ct x ct (src/evaluator.cu:29-424)
    void multiply(const Ciphertext &encrypted1, const Ciphertext &encrypted2, Ciphertext &destination) const;
    void multiply_inplace(Ciphertext &encrypted1, const Ciphertext &encrypted2) const { Ciphertext d; multiply(encrypted1, encrypted2, d); encrypted1 = std::move(d); }
    Ciphertext multiply_new(const Ciphertext &a, const Ciphertext &b) const { Ciphertext d; multiply(a, b, d); return d; }
    void square(const Ciphertext &encrypted, Ciphertext &destination) const;
    void square_inplace(Ciphertext &encrypted) const { Ciphertext d; square(encrypted, d); encrypted = std::move(d); }
    Ciphertext square_new(const Ciphertext &a) const { Ciphertext d; square(a, d); return d; }

    // ---- keyswitching family (src/evaluator_keyswitching.cu:11-361)
    void apply_keyswitching(const Ciphertext &encrypted, const KSwitchKeys &kswitch_keys, Ciphertext &destination) const;
    void apply_keyswitching_inplace(Ciphertext &encrypted, const KSwitchKeys &kswitch_keys) const;
    void relinearize(const Ciphertext &encrypted, const RelinKeys &relin_keys, Ciphertext &destination) const;
    void relinearize_inplace(Ciphertext &encrypted, const RelinKeys &relin_keys) const { Ciphertext d; relinearize(encrypted, relin_keys, d); encrypted = std::move(d); }
    Ciphertext relinearize_new(const Ciphertext &encrypted, const RelinKeys &relin_keys) const { Ciphertext d; relinearize(encrypted, relin_keys, d); return d; }
    void relinearize_batched(const std::vector<const Ciphertext *> &encrypted, const RelinKeys &relin_keys, const std::vector<Ciphertext *> &destination) const;
    void apply_galois(const Ciphertext &encrypted, size_t galois_element, const GaloisKeys &galois_keys, Ciphertext &destination) const;
    void apply_galois_inplace(Ciphertext &encrypted, size_t galois_element, const GaloisKeys &galois_keys) const {
        Ciphertext d; apply_galois(encrypted, galois_element, galois_keys, d); encrypted = std::move(d);
    }
    void apply_galois_batched(const std::vector<const Ciphertext *> &encrypted, size_t galois_element, const GaloisKeys &galois_keys,
                              const std::vector<Ciphertext *> &destination) const;
    // BFV / BGV: rotate_rows, rotate_columns; CKKS: rotate_vector, complex_conjugate (scheme checked as in src/evaluator.h:758-940)
    void rotate_rows(const Ciphertext &encrypted, int steps, const GaloisKeys &galois_keys, Ciphertext &destination) const;
    void rotate_vector(const Ciphertext &encrypted, int steps, const GaloisKeys &galois_keys, Ciphertext &destination) const;
    void rotate_columns(const Ciphertext &encrypted, const GaloisKeys &galois_keys, Ciphertext &destination) const;
    void complex_conjugate(const Ciphertext &encrypted, const GaloisKeys &galois_keys, Ciphertext &destination) const;
    Ciphertext rotate_rows_new(const Ciphertext &e, int steps, const GaloisKeys &k) const { Ciphertext d; rotate_rows(e, steps, k, d); return d; }
    Ciphertext rotate_vector_new(const Ciphertext &e, int steps, const GaloisKeys &k) const { Ciphertext d; rotate_vector(e, steps, k, d); return d; }
    void rotate_rows_batched(const std::vector<const Ciphertext *> &encrypted, int steps, const GaloisKeys &galois_keys, const std::vector<Ciphertext *> &destination) const;
    void rotate_vector_batched(const std::vector<const Ciphertext *> &encrypted, int steps, const GaloisKeys &galois_keys, const std::vector<Ciphertext *> &destination) const;

    // ---- level changes (src/evaluator_modswitch.cu:14-461)
    void mod_switch_to_next(const Ciphertext &encrypted, Ciphertext &destination) const;
    void mod_switch_to_next_inplace(Ciphertext &encrypted) const { Ciphertext d; mod_switch_to_next(encrypted, d); encrypted = std::move(d); }
    Ciphertext mod_switch_to_next_new(const Ciphertext &e) const { Ciphertext d; mod_switch_to_next(e, d); return d; }
    void rescale_to_next(const Ciphertext &encrypted, Ciphertext &destination) const;
    void rescale_to_next_inplace(Ciphertext &encrypted) const { Ciphertext d; rescale_to_next(encrypted, d); encrypted = std::move(d); }
    Ciphertext rescale_to_next_new(const Ciphertext &e) const { Ciphertext d; rescale_to_next(e, d); return d; }

    // ---- form changes (src/evaluator_transform_ntt.cu:469-653)
    void transform_to_ntt_inplace(Ciphertext &encrypted) const;
    void transform_from_ntt_inplace(Ciphertext &encrypted) const;
    void transform_to_ntt(const Ciphertext &encrypted, Ciphertext &destination) const { destination = encrypted.clone(); transform_to_ntt_inplace(destination); }
    void transform_from_ntt(const Ciphertext &encrypted, Ciphertext &destination) const { destination = encrypted.clone(); transform_from_ntt_inplace(destination); }
    void transform_to_ntt_inplace_batched(const std::vector<Ciphertext *> &encrypted) const;
    void transform_from_ntt_inplace_batched(const std::vector<Ciphertext *> &encrypted) const;

    // ---- RLWE packing (src/evaluator_lwes.cu:305-700): pack up to input_interval / output_interval ciphertexts into one (what the BumbleBee /
    //      Cheetah output packing and pack_lwe_ciphertexts call), with the field trace; groups are packed side by side in one call
    Ciphertext pack_rlwe_ciphertexts_new(const std::vector<const Ciphertext *> &ciphers, const GaloisKeys &automorphism_keys, size_t shift, size_t input_interval,
                                         size_t output_interval, bool apply_field_trace = true) const;
    void pack_rlwe_ciphertexts_batched(const std::vector<std::vector<const Ciphertext *>> &cipher_groups, const GaloisKeys &automorphism_keys, size_t shift,
                                       size_t input_interval, size_t output_interval, const std::vector<Ciphertext *> &outputs, bool apply_field_trace = true) const;

    // ---- LWE extraction and packing (src/evaluator_lwes.cu:51-94, 201-303)
    LWECiphertext extract_lwe_new(const Ciphertext &encrypted, size_t term) const;
    Ciphertext pack_lwe_ciphertexts_new(const std::vector<const LWECiphertext *> &lwes, const GaloisKeys &automorphism_keys, bool apply_field_trace = true) const;
    void pack_lwe_ciphertexts_batched(const std::vector<std::vector<const LWECiphertext *>> &lwes_groups, const GaloisKeys &automorphism_keys,
                                      const std::vector<Ciphertext *> &outputs, bool apply_field_trace = true) const;

    // ---- ct +/- pt (src/evaluator_translate_plain.cu:13-170): BFV with a mod-t plaintext (scale-up fused with the addition, one launch)
    //      or an RNS plaintext of the same form; CKKS with an NTT-form RNS plaintext of equal scale
    void add_plain(const Ciphertext &encrypted, const Plaintext &plain, Ciphertext &destination) const { translate_plain(encrypted, plain, destination, false); }
    void sub_plain(const Ciphertext &encrypted, const Plaintext &plain, Ciphertext &destination) const { translate_plain(encrypted, plain, destination, true); }
    void add_plain_inplace(Ciphertext &encrypted, const Plaintext &plain) const { Ciphertext d; translate_plain(encrypted, plain, d, false); encrypted = std::move(d); }
    void sub_plain_inplace(Ciphertext &encrypted, const Plaintext &plain) const { Ciphertext d; translate_plain(encrypted, plain, d, true); encrypted = std::move(d); }

    // ---- plaintext lifts (src/evaluator_transform_ntt.cu:35-129): a mod-t plaintext -> the RNS plaintext of a level (centralize), in NTT form
    Plaintext transform_plain_to_ntt(const Plaintext &plain, size_t coeff_modulus_size) const;
    Plaintext centralize(const Plaintext &plain, size_t coeff_modulus_size, uint64_t scalar = 1) const;

    // ---- ct x pt (src/evaluator_multiply_plain.cu:13-383): multiply_plain in all four form combinations, NTT-form batched / accumulate forms
    void multiply_plain(const Ciphertext &encrypted, const Plaintext &plain, Ciphertext &destination) const;
    void multiply_plain_inplace(Ciphertext &encrypted, const Plaintext &plain) const { Ciphertext d; multiply_plain(encrypted, plain, d); encrypted = std::move(d); }
    void multiply_plain_batched(const std::vector<const Ciphertext *> &encrypted, const std::vector<const Plaintext *> &plain, const std::vector<Ciphertext *> &destination) const;
    // destination += sum_i encrypted[i] * plain[i]   (multiply_plain_accumulate with set_zero)
    void multiply_plain_accumulate(const std::vector<const Ciphertext *> &encrypted, const std::vector<const Plaintext *> &plain, Ciphertext &destination, bool set_zero) const;
    // the reference's signature (src/evaluator_multiply_plain.cu:258-307): one destination pointer per term, destinations may repeat; one launch
    void multiply_plain_accumulate(const std::vector<const Ciphertext *> &encrypted, const std::vector<const Plaintext *> &plain,
                                   const std::vector<Ciphertext *> &destination, bool set_zero) const;

    // ---- LWE-packing helper (src/evaluator_lwes.cu; utils::negacyclic_shift_ps)
    void negacyclic_shift(const Ciphertext &encrypted, size_t shift, Ciphertext &destination) const;

private:
    HeContextPointer context_;
    void check_ct(const char *who, const Ciphertext &c) const;
    void translate_plain(const Ciphertext &encrypted, const Plaintext &plain, Ciphertext &destination, bool subtract) const;
    void translate(const Ciphertext &a, const Ciphertext &b, Ciphertext &destination, bool subtract) const;
    void translate_inplace(Ciphertext &a, const Ciphertext &b, bool subtract) const;
    void switch_key_internal(const Ciphertext &encrypted, const uint64_t *target, const KSwitchKeys &keys, size_t index,
                             SwitchKeyDestinationAssignMethod assign, Ciphertext &destination) const;
    void rotate_internal(const Ciphertext &encrypted, int steps, const GaloisKeys &galois_keys, Ciphertext &destination) const;
    void rotate_internal_batched(const std::vector<const Ciphertext *> &encrypted, int steps, const GaloisKeys &galois_keys,
                                 const std::vector<Ciphertext *> &destination) const;
};

}  // namespace b2troy
