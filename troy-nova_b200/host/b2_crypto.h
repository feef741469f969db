// b2_crypto.h — C++ host mirror of the reference's key holders, RandomGenerator, Encryptor and Decryptor on the C ABI
// of libb2rns.so.  Same names, argument meaning and error behaviour as src/encryptor.h / src/encryptor.cu, src/decryptor.h /
// src/decryptor.cu, src/utils/random_generator.h, src/key.h.  With the same generator state the ciphertext words equal the reference's.
// Keys are imported as raw words in the reference's layouts: secret key [K][N] and public key [2][K][N], NTT form at the key level
// (src/key.h:9-160).  Key GENERATION stays with the reference (KeyGenerator is host-side set-up code); a public key can be
// derived here from a secret key because it is one symmetric encryption of zero at the key level (KeyGenerator::create_public_key,
// src/key_generator.cu:57-66).
#pragma once
#include <mutex>

#include "b2_evaluator.h"

namespace b2troy {

// utils::RandomGenerator (src/utils/random_generator.h:45-93): AES-128 in counter mode keyed by a 128-bit seed
class RandomGenerator {
public:
    RandomGenerator() : RandomGenerator(0, 0) {}
    explicit RandomGenerator(uint64_t seed_lo, uint64_t seed_hi = 0) { st_.seed_lo = seed_lo; st_.seed_hi = seed_hi; st_.counter_lo = 0; st_.counter_hi = 0; }
    void reset_seed(uint64_t seed_lo, uint64_t seed_hi = 0) { st_.seed_lo = seed_lo; st_.seed_hi = seed_hi; st_.counter_lo = 0; st_.counter_hi = 0; }
    void set_counter(uint64_t lo, uint64_t hi = 0) { st_.counter_lo = lo; st_.counter_hi = hi; }
    uint64_t counter_lo() const { return st_.counter_lo; }
    uint64_t counter_hi() const { return st_.counter_hi; }
    uint64_t sample_uint64() { return b2_prng_sample_uint64(&st_); }
    b2_prng *raw() { return &st_; }

private:
    b2_prng st_;
};

// src/key.h: SecretKey = plaintext [K][N] in NTT form; PublicKey = ciphertext [2][K][N] in NTT form, both at the key level
class SecretKey {
public:
    SecretKey() = default;
    static SecretKey from_host(const HeContextPointer &ctx, const uint64_t *words);
    static SecretKey uninitialised(const HeContextPointer &ctx);
    SecretKey clone() const;
    uint64_t *data() { return buf_.data(); }
    std::vector<uint64_t> to_host() const;
    const uint64_t *data() const { return buf_.data(); }
    const HeContextPointer &context() const { return ctx_; }
    bool empty() const { return buf_.size() == 0; }
private:
    HeContextPointer ctx_;
    DeviceArray buf_;
};
class PublicKey {
public:
    PublicKey() = default;
    static PublicKey from_host(const HeContextPointer &ctx, const uint64_t *words);
    // KeyGenerator::create_public_key: one symmetric zero encryption at the key level, NTT form (src/key_generator.cu:57-66)
    static PublicKey create(const HeContextPointer &ctx, const SecretKey &sk, RandomGenerator &prng);
    const uint64_t *data() const { return buf_.data(); }
    std::vector<uint64_t> to_host() const;
    bool empty() const { return buf_.size() == 0; }
private:
    HeContextPointer ctx_;
    DeviceArray buf_;
};

// src/batch_encoder.h (BFV / BGV with a batching plain modulus): SIMD slots <-> the plaintext polynomial modulo t
class BatchEncoder {
public:
    explicit BatchEncoder(HeContextPointer context);
    size_t slot_count() const { return context_->poly_modulus_degree(); }
    void encode(const std::vector<uint64_t> &values, Plaintext &destination) const;
    Plaintext encode_new(const std::vector<uint64_t> &values) const { Plaintext p; encode(values, p); return p; }
    std::vector<uint64_t> decode_new(const Plaintext &plain) const;
    // encode_polynomial (src/batch_encoder.cu:296-343): the values ARE the coefficients (reduced modulo t)
    Plaintext encode_polynomial_new(const std::vector<uint64_t> &values) const;
private:
    HeContextPointer context_;
};

// src/key_generator.h: every key is drawn from the generator passed in, in call order, exactly as the reference's KeyGenerator draws from the context's
// generator — same seed and call order => the reference's keys word for word (tests/test_keygen_gpu.py)
class KeyGenerator {
public:
    // samples a fresh secret key (NTT(ternary)), or adopts an existing one
    KeyGenerator(HeContextPointer context, RandomGenerator &prng);
    KeyGenerator(HeContextPointer context, RandomGenerator &prng, const SecretKey &secret_key);
    const SecretKey &secret_key() const { return sk_; }
    PublicKey create_public_key() { return PublicKey::create(context_, sk_, *prng_); }
    RelinKeys create_relin_keys();                                               // the key for s^2 at RelinKeys::get_index(2)
    GaloisKeys create_galois_keys(const std::vector<size_t> &galois_elements);
    GaloisKeys create_automorphism_keys();                                       // 2^k + 1, k = 1 .. log N: what pack_lwe_ciphertexts needs

private:
    void one_kswitch_key(const uint64_t *new_key, KSwitchKeys &dest, size_t index);
    HeContextPointer context_;
    RandomGenerator *prng_;
    SecretKey sk_;
};

// src/encryptor.h.  `prng` is the context's generator of the reference (HeContext::random_generator()): it supplies the noise and, when no
// u_prng is passed, everything else.  Levels are named by their number of limbs (the reference's parms_id); 0 = the first level.
class Encryptor {
public:
    Encryptor(HeContextPointer context, uint64_t seed_lo = 0, uint64_t seed_hi = 0) : context_(std::move(context)), prng_(seed_lo, seed_hi) {}
    void set_public_key(PublicKey pk) { public_key_ = std::move(pk); }
    void set_secret_key(SecretKey sk) { secret_key_ = std::move(sk); }
    RandomGenerator &random_generator() { return prng_; }

    void encrypt_zero_asymmetric(Ciphertext &destination, size_t coeff_modulus_size = 0, RandomGenerator *u_prng = nullptr);
    void encrypt_zero_symmetric(bool save_seed, Ciphertext &destination, size_t coeff_modulus_size = 0, RandomGenerator *u_prng = nullptr);
    void encrypt_asymmetric(const Plaintext &plain, Ciphertext &destination, RandomGenerator *u_prng = nullptr);
    void encrypt_symmetric(const Plaintext &plain, bool save_seed, Ciphertext &destination, RandomGenerator *u_prng = nullptr);
    Ciphertext encrypt_zero_asymmetric_new(size_t coeff_modulus_size = 0, RandomGenerator *u_prng = nullptr) { Ciphertext d; encrypt_zero_asymmetric(d, coeff_modulus_size, u_prng); return d; }
    Ciphertext encrypt_zero_symmetric_new(bool save_seed, size_t coeff_modulus_size = 0, RandomGenerator *u_prng = nullptr) { Ciphertext d; encrypt_zero_symmetric(save_seed, d, coeff_modulus_size, u_prng); return d; }
    Ciphertext encrypt_asymmetric_new(const Plaintext &plain, RandomGenerator *u_prng = nullptr) { Ciphertext d; encrypt_asymmetric(plain, d, u_prng); return d; }
    Ciphertext encrypt_symmetric_new(const Plaintext &plain, bool save_seed, RandomGenerator *u_prng = nullptr) { Ciphertext d; encrypt_symmetric(plain, save_seed, d, u_prng); return d; }

private:
    void encrypt_zero_internal(size_t nl, bool is_ntt_form, bool is_asymmetric, bool save_seed, RandomGenerator *u_prng, Ciphertext &destination);
    void encrypt_internal(const Plaintext &plain, bool is_asymmetric, bool save_seed, RandomGenerator *u_prng, Ciphertext &destination);
    HeContextPointer context_;
    RandomGenerator prng_;
    PublicKey public_key_;
    SecretKey secret_key_;
};

// src/decryptor.h: holds the powers of the secret key (NTT form), extended on demand under a lock as the reference does
class Decryptor {
public:
    Decryptor(HeContextPointer context, const SecretKey &secret_key);
    void decrypt(const Ciphertext &encrypted, Plaintext &destination) const;
    Plaintext decrypt_new(const Ciphertext &encrypted) const { Plaintext p; decrypt(encrypted, p); return p; }
    // one call for the whole batch: 3 (BFV) / 2 (BGV) / 1 (CKKS) launches; all ciphertexts of one level and size
    void decrypt_batched(const std::vector<const Ciphertext *> &encrypted, const std::vector<Plaintext *> &destination) const;
    // c_0 + sum c_k s^k in the form of the ciphertext (bfv_decrypt_without_scaling_down / the CKKS result), as an RNS plaintext
    void phase(const Ciphertext &encrypted, Plaintext &destination) const;

private:
    const uint64_t *powers(size_t max_power) const;
    HeContextPointer context_;
    mutable std::mutex mu_;
    mutable DeviceArray powers_;
    mutable size_t n_powers_ = 0;
};

}  // namespace b2troy
