/*
 * b2rns.h — C ABI of libb2rns.so: the B200-native (sm_100a) RNS polynomial hot path that slots in
 * behind Troy-Nova's L2 free-function seam.
 *
 * Conventions (mirroring the reference; citations are relative to the reference tree):
 *   - polynomials are row-major uint64 arrays [poly][limb][N]            (src/ciphertext.h:211-252)
 *   - a context holds the KEY-LEVEL modulus chain q_0..q_{K-2}, P (special prime last); a ciphertext
 *     at a level with `nl` limbs uses moduli q_0..q_{nl-1}              (src/he_context.cu:5-45)
 *   - every stored output is the canonical residue in [0, q_i)
 *   - all data pointers are DEVICE pointers unless the function name ends in _host
 *   - calls are asynchronous on `stream` (a cudaStream_t cast to void*; NULL = legacy default
 *     stream); the reference launches on the per-thread default stream and never synchronises
 *     (src/utils/memory_pool.cu:10-19), the adapter passes cudaStreamPerThread
 *   - errors: the reference throws std::invalid_argument / std::runtime_error on the host
 *     (e.g. src/fgk/ntt_grouped.cu:261-263); here every entry point returns a b2_status and never
 *     throws; b2_last_error() gives the message for the calling thread
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns
 *     B2_ERR_NO_DEVICE.
 */
#ifndef B2RNS_H
#define B2RNS_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    B2_OK = 0,
    B2_ERR_INVALID_ARG = 1,   /* reference: std::invalid_argument */
    B2_ERR_CUDA = 2,          /* reference: std::runtime_error from cudaMalloc/cudaMemcpy wrappers */
    B2_ERR_NO_DEVICE = 3,
    B2_ERR_UNSUPPORTED = 4,
    B2_ERR_OUT_OF_MEMORY = 5
} b2_status;

/* SchemeType, src/encryption_parameters.h:7-12 */
typedef enum { B2_SCHEME_BFV = 1, B2_SCHEME_CKKS = 2, B2_SCHEME_BGV = 3 } b2_scheme;
/* Evaluator::SwitchKeyDestinationAssignMethod, src/evaluator.h:18-22 */
typedef enum { B2_ASSIGN_OVERWRITE = 0, B2_ASSIGN_ADD_INPLACE = 1, B2_ASSIGN_OVERWRITE_EXCEPT_FIRST = 2 } b2_assign;
/* NTTTableIndexer::NTTTableIndexerType, src/utils/ntt.h:93-97 */
typedef enum { B2_TABLES_COMPONENTWISE = 0, B2_TABLES_KS_SET_PRODUCTS = 1, B2_TABLES_KS_SKIP_FINALS = 2 } b2_table_mode;

/* which rounding division b2_mod_switch_to_next performs */
typedef enum {
    B2_MODSWITCH_BFV = 0,   /* RNSTool::divide_and_round_q_last, coefficient form   (src/utils/rns_tool.cu:374-497) */
    B2_MODSWITCH_CKKS = 1,  /* RNSTool::divide_and_round_q_last_ntt = rescale        (src/utils/rns_tool.cu:499-760) */
    B2_MODSWITCH_BGV = 2,   /* RNSTool::mod_t_and_divide_q_last_ntt                  (src/utils/rns_tool.cu:1540-1833) */
    B2_MODSWITCH_DROP = 3   /* CKKS mod_switch_to_next: drop the last limb           (src/evaluator_modswitch.cu:152-171) */
} b2_modswitch_kind;

typedef struct b2_ctx b2_ctx;
typedef struct b2_kswitch_key b2_kswitch_key;

typedef struct {
    uint32_t log_n;            /* poly_modulus_degree = 1 << log_n, 1 <= log_n <= 17 (src/utils/constants.h:13) */
    uint32_t n_moduli;         /* key-level modulus count K (special prime last), 1..64 */
    const uint64_t *moduli;    /* host pointer, K primes = 1 mod 2N, < 2^61 (src/modulus.cu:15-16) */
    uint64_t plain_modulus;    /* t (BFV/BGV), 0 for CKKS */
    uint32_t scheme;           /* b2_scheme */
} b2_params;

const char *b2_last_error(void);
const char *b2_version(void);
int         b2_device_count(void);
/* kernels launched by this library in this process so far (instrumentation for bench.py) */
unsigned long long b2_launch_count(void);

/* ---- context: replaces the device-resident tables of HeContext / ContextData
 *      (src/he_context.cu:46-145: NTTTables, RNSTool constants, GaloisTool tables) ---- */
b2_status b2_ctx_create(int device, const b2_params *params, b2_ctx **out);
void      b2_ctx_destroy(b2_ctx *ctx);
/* host-side helpers (pure host setup code, mirrors CoeffModulus::create / PlainModulus::batching,
 * src/coeff_modulus.cu:64-107) */
b2_status b2_coeff_modulus_create(size_t n, const uint32_t *bit_sizes, size_t count, uint64_t *out);
/* psi_i: minimal primitive 2N-th root of modulus i (src/utils/number_theory.cu:68-87) */
b2_status b2_ctx_root(const b2_ctx *ctx, size_t i, uint64_t *psi);

/* ---- device memory helpers for hosts that have no CUDA runtime of their own ---- */
b2_status b2_malloc(b2_ctx *ctx, size_t bytes, void **dptr);
b2_status b2_free(b2_ctx *ctx, void *dptr);
b2_status b2_malloc_host(size_t bytes, void **hptr);       /* pinned */
b2_status b2_free_host(void *hptr);
b2_status b2_memcpy_h2d(b2_ctx *ctx, void *dst, const void *src, size_t bytes, void *stream);
b2_status b2_memcpy_d2h(b2_ctx *ctx, void *dst, const void *src, size_t bytes, void *stream);
b2_status b2_memcpy_d2d(b2_ctx *ctx, void *dst, const void *src, size_t bytes, void *stream);
b2_status b2_stream_sync(b2_ctx *ctx, void *stream);

/* ---- N1/N2: utils::ntt_transfer_to_rev / ntt_transfer_from_rev (+_inplace when in == out)
 *      (src/utils/ntt.h:132-144).  data [pcount][ncomp][N]; the table of element (k, j) is chosen as
 *      NTTTableIndexer::get does (src/utils/ntt.h:105-123): COMPONENTWISE -> table_base + j;
 *      KS_SET_PRODUCTS -> (k == decomp ? K-1 : k); KS_SKIP_FINALS -> (j == decomp ? K-1 : j).
 *      `batch` independent operands are laid out contiguously ([batch][pcount][ncomp][N]); the
 *      reference's *_batched variants (src/fgk/ntt_grouped.cu:297-344) take a vector of slices,
 *      b2_ntt_*_ptrs takes the same thing as a HOST array of device pointers.
 *      forward: input < 4p, output canonical, bit-reversed order; inverse: input < 2p, output
 *      canonical and scaled by N^-1. ---- */
b2_status b2_ntt_forward(b2_ctx *ctx, const uint64_t *in, uint64_t *out, size_t batch, size_t pcount, size_t ncomp,
                         int table_mode, size_t table_base, size_t decomp, void *stream);
b2_status b2_ntt_inverse(b2_ctx *ctx, const uint64_t *in, uint64_t *out, size_t batch, size_t pcount, size_t ncomp,
                         int table_mode, size_t table_base, size_t decomp, void *stream);
b2_status b2_ntt_forward_ptrs(b2_ctx *ctx, const uint64_t *const *in, uint64_t *const *out, size_t batch, size_t pcount,
                              size_t ncomp, int table_mode, size_t table_base, size_t decomp, void *stream);
b2_status b2_ntt_inverse_ptrs(b2_ctx *ctx, const uint64_t *const *in, uint64_t *const *out, size_t batch, size_t pcount,
                              size_t ncomp, int table_mode, size_t table_base, size_t decomp, void *stream);
/* same with HOST buffers: H2D, transform, D2H on `stream`, synchronised on return (what a host-memory
 * caller of the reference does with Ciphertext::to_device / to_host around the op) */
b2_status b2_ntt_forward_host(b2_ctx *ctx, const uint64_t *in, uint64_t *out, size_t batch, size_t pcount, size_t ncomp,
                              int table_mode, size_t table_base, size_t decomp, void *stream);
b2_status b2_ntt_inverse_host(b2_ctx *ctx, const uint64_t *in, uint64_t *out, size_t batch, size_t pcount, size_t ncomp,
                              int table_mode, size_t table_base, size_t decomp, void *stream);

/* ---- K1-K5: Evaluator::switch_key_internal (src/evaluator_keyswitching_core.cu:757-1052).
 *      A key is `digits` ciphertexts key_j = [2][K][N] in NTT form at the key level; the reference keeps
 *      a device array of their base pointers (src/kswitch_keys.h:36-57) and so does b2_kswitch_key.
 *      digit_ptrs is a HOST array of `digits` DEVICE pointers (not copied: the key data stays owned by
 *      the caller and must outlive the handle).
 *      b2_keyswitch: target [decomp][N] (NTT form iff is_ntt), dest [2][decomp][N] written per `assign`
 *      (b2_assign).  BFV ciphertexts are coefficient form, CKKS / BGV are NTT form. ---- */
b2_status b2_kswitch_key_create(b2_ctx *ctx, const uint64_t *const *digit_ptrs, size_t digits, b2_kswitch_key **out);
b2_status b2_kswitch_key_create_contiguous(b2_ctx *ctx, const uint64_t *key /* [digits][2][K][N] */, size_t digits, b2_kswitch_key **out);
/* zero-copy variant for callers that already keep the pointer table on the device (KSwitchKeys::get_data_ptrs,
 * src/kswitch_keys.h:50-57): device_ptr_table = DEVICE array of `digits` device pointers, digit_ptrs = the same pointers
 * in a HOST array.  Nothing is allocated or copied; both arrays' contents must outlive the handle. */
b2_status b2_kswitch_key_wrap(b2_ctx *ctx, const uint64_t *const *device_ptr_table, const uint64_t *const *digit_ptrs, size_t digits, b2_kswitch_key **out);
void      b2_kswitch_key_destroy(b2_kswitch_key *key);
b2_status b2_keyswitch(b2_ctx *ctx, const uint64_t *target, size_t decomp, int is_ntt, const b2_kswitch_key *key,
                       uint64_t *dest, int assign, void *stream);
/* same with HOST buffers for target and dest (key resident on the device): H2D, keyswitch, D2H on `stream`, synchronised on return */
b2_status b2_keyswitch_host(b2_ctx *ctx, const uint64_t *target, size_t decomp, int is_ntt, const b2_kswitch_key *key,
                            uint64_t *dest, int assign, void *stream);
/* K5: Evaluator::switch_key_internal_batched (src/evaluator_keyswitching_core.cu:1055-1267): `batch` ciphertexts sharing one
 * key.  targets / dests are HOST arrays of `batch` DEVICE pointers (the reference's ConstSliceVec / vector<Ciphertext*>);
 * the same five launches serve up to 16 ciphertexts at a time, pointer tables travel as kernel arguments. */
b2_status b2_keyswitch_batched(b2_ctx *ctx, const uint64_t *const *targets, uint64_t *const *dests, size_t batch, size_t decomp, int is_ntt,
                               const b2_kswitch_key *key, int assign, void *stream);

/* ---- limb-sharded keyswitch across the GPUs of one box (no counterpart in the reference, which only
 * replicates whole contexts per device).  The decomp + 1 output moduli of Evaluator::switch_key_internal
 * (src/evaluator_keyswitching_core.cu:757-1052) are partitioned over `world` ranks; each rank keeps only its moduli's
 * ciphertext limbs ([poly][owned limbs][N]) and key rows.  Per keyswitch: prepare -> all-gather (caller, NCCL) ->
 * products -> broadcast of u from the special prime's owner (caller) -> finish.  Bit-identical to b2_keyswitch. */
b2_status b2_limb_partition(size_t n_out_moduli, int world, int rank, size_t *first, size_t *count);
/* partition weighted by modulus cost (integer-pipe moduli count double, the special prime's owner once more): contiguous ranges */
b2_status b2_limb_partition_weighted(const b2_ctx *ctx, size_t decomp, int world, int rank, size_t *first, size_t *count);
/* P2P exchange for the limb-sharded keyswitch (peer buffers mapped by the caller: symmetric memory / CUDA IPC; HOST arrays of `world`
 * DEVICE pointers): push copies `words` (even) u64 from src to peer_dst[r] for every r, then stores `epoch` into peer_flags[r][flag_slot]
 * with system-scope release; wait spins (on the device, stream-ordered) until flags[slots[i]] >= epoch for every i.
 * b2_keyswitch_shard_products with world == 0 takes `gathered` as the complete coefficient-form target [decomp][N]. */
b2_status b2_shard_push(b2_ctx *ctx, const uint64_t *src, uint64_t *const *peer_dst, int world, size_t words, uint64_t *const *peer_flags, size_t flag_slot,
                        uint64_t epoch, void *stream);
b2_status b2_shard_wait(b2_ctx *ctx, const uint64_t *flags, const int *slots, int n_slots, uint64_t epoch, void *stream);
b2_status b2_kswitch_key_shard(b2_ctx *ctx, const b2_kswitch_key *full, size_t decomp, size_t first, size_t count, b2_kswitch_key **out);
b2_status b2_keyswitch_shard_prepare(b2_ctx *ctx, const uint64_t *target_local /* [owned ct limbs][N] */, size_t decomp, int is_ntt,
                                     size_t first, size_t count, uint64_t *send /* [pad][N] */, void *stream);
b2_status b2_keyswitch_shard_products(b2_ctx *ctx, const uint64_t *gathered /* [world][pad][N] */, int world, size_t pad, size_t decomp,
                                      size_t first, size_t count, const b2_kswitch_key *shard, uint64_t *u /* [2][N] */, void *stream);
b2_status b2_keyswitch_shard_finish(b2_ctx *ctx, const uint64_t *u /* [2][N] */, size_t decomp, int is_ntt, size_t first, size_t count,
                                    uint64_t *dest_local /* [2][owned ct limbs][N] */, int assign, void *stream);


/* ---- R1-R4: mod_switch_to_next / rescale_to_next (src/evaluator_modswitch.cu:14-149,445-461).
 *      in [count][nl][N] -> out [count][nl-1][N]; kind is a b2_modswitch_kind. ---- */
b2_status b2_mod_switch_to_next(b2_ctx *ctx, const uint64_t *in, uint64_t *out, size_t count, size_t nl, int kind, void *stream);

/* ---- M1, M3, E1, G1: element-wise kernels.  Operands [count][ncomp][N], limb j uses modulus table_base + j. ---- */
/* utils::add_ps / sub_ps / negate_ps (src/utils/poly_small_mod.cu:263-420) */
b2_status b2_add(b2_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t count, size_t ncomp, size_t table_base, void *stream);
b2_status b2_sub(b2_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t count, size_t ncomp, size_t table_base, void *stream);
b2_status b2_negate(b2_ctx *ctx, const uint64_t *a, uint64_t *out, size_t count, size_t ncomp, size_t table_base, void *stream);
/* fgk::dyadic_convolute::dyadic_convolute / dyadic_square (src/fgk/dyadic_convolute.cu:6-148):
 * a [batch][na][ncomp][N], b [batch][nb][ncomp][N] -> out [batch][na+nb-1][ncomp][N]; na, nb <= 3 */
b2_status b2_dyadic_convolute(b2_ctx *ctx, const uint64_t *a, size_t na, const uint64_t *b, size_t nb, uint64_t *out, size_t batch, size_t ncomp, void *stream);
b2_status b2_dyadic_square(b2_ctx *ctx, const uint64_t *a, uint64_t *out, size_t batch, size_t ncomp, void *stream);
/* fgk::dyadic_convolute::dyadic_broadcast_product(_accumulate)_ps (src/fgk/dyadic_convolute.cu:151-292):
 * out[k] (+)= a[k] o plain, k < pcount; plain [ncomp][N] */
b2_status b2_dyadic_broadcast(b2_ctx *ctx, const uint64_t *a, const uint64_t *plain, uint64_t *out, size_t pcount, size_t ncomp, int accumulate, void *stream);
/* ... _bps (src/fgk/dyadic_convolute.cu:223-292), Evaluator::multiply_plain_batched on contiguous batches:
 * a, out [batch][pcount][ncomp][N], plain [batch][ncomp][N]; ciphertext b is multiplied by plaintext b, one launch */
b2_status b2_dyadic_broadcast_batched(b2_ctx *ctx, const uint64_t *a, const uint64_t *plain, uint64_t *out, size_t batch, size_t pcount, size_t ncomp,
                                      int accumulate, void *stream);
/* Evaluator::multiply_plain_ntt_batched / multiply_plain_ntt_accumulate over SCATTERED operands
 * (src/evaluator_multiply_plain.cu:216-307 -> kernel_dyadic_broadcast_product_bps, src/fgk/dyadic_convolute.cu:167-171).
 * ct / pt / dest are HOST arrays of n_terms DEVICE pointers: ct[i] [pcount][ncomp][N], pt[i] [ncomp][N], dest[i]
 * [pcount][ncomp][N]; limb j uses modulus table_base + j.  Destinations may REPEAT: all terms with the same destination are
 * summed into it (the reference loops the terms in-thread in order; modular sums are order-independent, so the stored words
 * are identical).  add_to_dest != 0 keeps the destinations' contents as starting values (the reference's set_zero == false).
 * One launch per 65535 distinct destinations; every operand word is read once. */
b2_status b2_multiply_plain_accumulate(b2_ctx *ctx, const uint64_t *const *ct, const uint64_t *const *pt, uint64_t *const *dest, size_t n_terms,
                                       size_t pcount, size_t ncomp, size_t table_base, int add_to_dest, void *stream);
/* ---- Decryptor ----
 * sk_powers: DEVICE array [max_power][K][N], s, s^2, ... in NTT form at the key level — the reference's Decryptor::secret_key_array_
 * (src/decryptor.h, filled by KeyGenerator::compute_secret_key_powers, src/key_generator.cu:84-110).  b2_secret_key_powers builds it
 * from the secret key sk [K][N] (NTT form; out == sk + 0 is allowed: power 1 is copied unless out == sk). */
b2_status b2_secret_key_powers(b2_ctx *ctx, const uint64_t *sk, size_t max_power, uint64_t *out, void *stream);
/* Decryptor::dot_product_ct_sk_array[_batched] (src/decryptor.cu:27-231): phase[b] = c_0 + sum_{k=1}^{pcount-1} c_k s^k, [nl][N], in the
 * form of the ciphertexts.  ct / phase: HOST arrays of `batch` DEVICE pointers (ct[b] [pcount][nl][N]).  NTT form: one launch;
 * coefficient form: two launches (forward transform with the key product fused, inverse transform with the sum and c_0 fused). */
b2_status b2_decrypt_phase(b2_ctx *ctx, const uint64_t *const *ct, uint64_t *const *phase, size_t batch, size_t pcount, size_t nl, int is_ntt_form,
                           const uint64_t *sk_powers, void *stream);
/* RNSTool::decrypt_scale_and_round[_batched] (src/utils/rns_tool.cu:1334-1430): BFV, phase[b] [nl][N] coefficient form ->
 * plain[b] [N] = round(t/q * phase) mod t.  gamma is derived as the reference does (rns_tool.cu:68-76). */
b2_status b2_decrypt_scale_and_round(b2_ctx *ctx, const uint64_t *const *phase, uint64_t *const *plain, size_t batch, size_t nl, void *stream);
/* RNSTool::decrypt_mod_t (BaseConverter::exact_convey_array, src/utils/rns_base.cu:445-583) and the correction factor of
 * scaling_variant::decentralize (src/utils/scaling_variant.cu:415-431): BGV, phase[b] [nl][N] coefficient form -> plain[b] [N].
 * correction_factors: HOST array [batch] or NULL (all 1); B2_ERR_INVALID_ARG if one is not invertible modulo t. */
b2_status b2_decrypt_mod_t(b2_ctx *ctx, const uint64_t *const *phase, uint64_t *const *plain, size_t batch, size_t nl, const uint64_t *correction_factors,
                           void *stream);
/* Decryptor::decrypt / decrypt_batched (src/decryptor.cu:233-535) for the scheme of the context: BFV (coefficient-form
 * ciphertexts) and BGV (NTT form) -> plain[b] [N] modulo t; CKKS (NTT form) -> plain[b] [nl][N] in NTT form.
 * 3 / 2 / 1 launches per call for any batch size. */
b2_status b2_decrypt(b2_ctx *ctx, const uint64_t *const *ct, uint64_t *const *plain, size_t batch, size_t pcount, size_t nl, const uint64_t *sk_powers,
                     const uint64_t *correction_factors, void *stream);
/* ---- Encryptor / RLWE sampling and the plaintext lifts (rank 3) ----
 * b2_prng mirrors utils::RandomGenerator (src/utils/random_generator.h:45-93): AES-128 in counter mode, key = the 128-bit seed, one
 * 16-byte block per counter value.  Every call that draws from a generator advances its counter exactly as the reference does, so
 * a sequence of calls with the same starting state reproduces the reference's ciphertext words. */
typedef struct { uint64_t seed_lo, seed_hi, counter_lo, counter_hi; } b2_prng;
typedef enum { B2_SAMPLE_TERNARY = 1, B2_SAMPLE_CENTERED_BINOMIAL = 2, B2_SAMPLE_UNIFORM = 3 } b2_sample_kind;
/* RandomGenerator::fill_bytes (random_generator.cu:128-141): `bytes` bytes at the DEVICE address out */
b2_status b2_prng_fill(b2_ctx *ctx, b2_prng *g, void *out, size_t bytes, void *stream);
/* RandomGenerator::sample_uint64 (host) */
uint64_t b2_prng_sample_uint64(b2_prng *g);
/* sample_poly_ternary / _centered_binomial / _uniform (random_generator.cu:282-505), and their _batched forms: out [count][nl][N] */
b2_status b2_sample_poly(b2_ctx *ctx, int kind, b2_prng *g, uint64_t *out, size_t count, size_t nl, void *stream);
/* rlwe::asymmetric_with_u_prng (src/utils/rlwe.cu:11-92) with the level handling of Encryptor::encrypt_zero_internal
 * (src/encryptor.cu:45-98: below the key level the encryption happens one level up and is switched down).
 * pk: DEVICE [2][K][N], NTT form, key level.  ct: DEVICE [2][nl][N].  u_prng draws u, e_prng the two noise polynomials; both may be
 * the same object (the reference's u_prng == nullptr: everything from the context's generator). */
b2_status b2_encrypt_zero_asymmetric(b2_ctx *ctx, const uint64_t *pk, size_t nl, int is_ntt_form, b2_prng *u_prng, b2_prng *e_prng, uint64_t *ct, void *stream);
/* rlwe::symmetric_with_c1_prng (src/utils/rlwe.cu:218-322): sk DEVICE [K][N] NTT form; ct = (-(a s + e), a).  *seed_out (may be NULL)
 * receives the seed of a's generator, which the reference stores in the ciphertext when save_seed != 0. */
b2_status b2_encrypt_zero_symmetric(b2_ctx *ctx, const uint64_t *sk, size_t nl, int is_ntt_form, int save_seed, b2_prng *c1_prng, b2_prng *e_prng, uint64_t *ct,
                                    uint64_t *seed_out, void *stream);
/* BatchEncoder (src/batch_encoder.cu:14-60,169-227,407-460): SIMD slots <-> the plaintext polynomial modulo t.  The context's plain modulus must be a
 * batching prime (= 1 mod 2N); the tables of the negacyclic transform modulo t and the slot map are built by the first call.
 * encode: values DEVICE [count <= N] -> plain DEVICE [N] (coefficient form, modulo t); decode: plain [plain_count <= N] -> values [N]. */
b2_status b2_batch_encode(b2_ctx *ctx, const uint64_t *values, size_t count, uint64_t *plain, void *stream);
b2_status b2_batch_decode(b2_ctx *ctx, const uint64_t *plain, size_t plain_count, uint64_t *values, void *stream);
/* KeyGenerator (src/key_generator.cu): with the generator in the state of the reference's context generator the keys equal the reference's.
 *   secret key (:31-58)                sk [K][N] = NTT(ternary sample)
 *   public key (:60-78)                = b2_encrypt_zero_symmetric(sk, K, ntt = 1, save_seed = 0) at the key level
 *   key-switching key (:136-157)       out [K-1][2][K][N]: digit i = symmetric encryption of zero at the key level + (P mod q_i) new_key in limb i of c_0;
 *                                      new_key [K][N] NTT form = s^2 (relinearisation; b2_secret_key_powers) or the Galois image of s (rotations; b2_galois) */
b2_status b2_generate_secret_key(b2_ctx *ctx, b2_prng *g, uint64_t *sk, void *stream);
b2_status b2_generate_kswitch_key(b2_ctx *ctx, const uint64_t *sk, const uint64_t *new_key, b2_prng *g, uint64_t *out, void *stream);
/* scaling_variant::scale_up = multiply_translate_plain (src/fgk/translate_plain.cu:6-70): BFV, plain = plain_count words modulo t;
 * dest[j][x] = from[j][x] +/- round(q/t m_x) mod q_j for x < plain_count (from == NULL: zero; from == dest allowed), [nl][N].
 * This is the c_0 update of Encryptor::encrypt and of Evaluator::add_plain / sub_plain. */
b2_status b2_bfv_scale_up(b2_ctx *ctx, const uint64_t *plain, size_t plain_count, size_t nl, const uint64_t *from, int subtract, uint64_t *dest, void *stream);
/* scaling_variant::centralize (src/utils/scaling_variant.cu:326-357): plain modulo t -> centred residues modulo q_j, out [nl][N] */
b2_status b2_centralize(b2_ctx *ctx, const uint64_t *plain, size_t plain_count, size_t nl, uint64_t *out, void *stream);
/* the same on (scalar * m mod t): BGV Evaluator::add_plain / sub_plain multiply by the ciphertext's correction factor first (src/evaluator_translate_plain.cu:79-82) */
b2_status b2_centralize_scaled(b2_ctx *ctx, const uint64_t *plain, size_t plain_count, size_t nl, uint64_t scalar, uint64_t *out, void *stream);
/* Encryptor::encrypt_internal (src/encryptor.cu:247-338): BFV / BGV plain = plain_count words modulo t (nl = first level);
 * CKKS plain = [nl][N] NTT form.  key = pk (asymmetric != 0) or sk. */
b2_status b2_encrypt(b2_ctx *ctx, const uint64_t *plain, size_t plain_count, const uint64_t *key, int asymmetric, size_t nl, b2_prng *u_prng, b2_prng *e_prng,
                     uint64_t *ct, void *stream);
/* ---- RLWE packing ----
 * One layer of Evaluator::pack_rlwe_ciphertexts_new (src/evaluator_lwes.cu:437-479) over `pairs` (even, odd) pairs, coefficient form:
 *   out[b] = (even + x^shift odd) + apply_galois(even - x^shift odd, galois_element)
 * as ONE element-wise launch (shift, add, sub and the automorphism fused; the reference issues six) and ONE batched keyswitch.
 * in: DEVICE [2 pairs][2][nl][N] (pair b = slots 2b, 2b+1); out: DEVICE [pairs][2][nl][N]; target: DEVICE scratch [pairs][nl][N]. */
b2_status b2_pack_layer(b2_ctx *ctx, const uint64_t *in, uint64_t *out, uint64_t *target, size_t pairs, size_t nl, size_t shift, uint64_t galois_element,
                        const b2_kswitch_key *key, void *stream);
/* Evaluator::pack_rlwe_ciphertexts_new / pack_rlwe_ciphertexts_batched (src/evaluator_lwes.cu:305-700; what pack_lwe_ciphertexts_new :201-229
 * and the BumbleBee / Cheetah output packing src/app/matmul.cu:572-619 call) including field_trace_inplace[_batched] (:96-130), for `groups`
 * groups packed side by side: every layer is one fused launch + one batched keyswitch over all groups.
 * ciphers: HOST array of sum(counts) DEVICE pointers, group after group, each [2][nl][N]; counts: HOST [groups];
 * galois_elements / keys: the automorphism keys as parallel HOST arrays (B2_ERR_INVALID_ARG "Galois key not present" if one is missing);
 * outs: HOST array of `groups` DEVICE pointers, each [2][nl][N], NTT form iff output_ntt_form (the reference: CKKS / BGV).
 * The tree runs in coefficient form (CKKS too: the automorphism and the keyswitch compute the same ring element in either form, which drops
 * the two transforms per layer of the reference's CKKS branch; BGV transforms the keyswitch operand per layer as the reference does); the
 * stored words are the reference's. */
b2_status b2_pack_rlwes(b2_ctx *ctx, const uint64_t *const *ciphers, const size_t *counts, size_t groups, int input_ntt_form, size_t nl, size_t shift,
                        size_t input_interval, size_t output_interval, const uint64_t *galois_elements, const b2_kswitch_key *const *keys, size_t n_keys,
                        int apply_field_trace, int output_ntt_form, uint64_t *const *outs, void *stream);
/* LWE ciphertexts (src/lwe_ciphertext.h): c0 = nl scalars, c1 = [nl][N].
 * Evaluator::extract_lwe_new (src/evaluator_lwes.cu:51-94) on a coefficient-form ct [2][nl][N]: c1 = x^(2N - term) c_1, c0[j] = c_0[j][term]; one launch */
b2_status b2_extract_lwe(b2_ctx *ctx, const uint64_t *ct, size_t term, size_t nl, uint64_t *c0, uint64_t *c1, void *stream);
/* LWECiphertext::assemble_lwe[_batched] (src/lwe_ciphertext.cu:12-140): c0s / c1s / outs HOST arrays of `count` DEVICE pointers; outs[b] [2][nl][N], one launch */
b2_status b2_assemble_lwes(b2_ctx *ctx, const uint64_t *const *c0s, const uint64_t *const *c1s, uint64_t *const *outs, size_t count, size_t nl, void *stream);
/* Evaluator::pack_lwe_ciphertexts_new / _batched (src/evaluator_lwes.cu:201-303): counts[g] LWEs per group (c0s / c1s: all groups, one after the other)
 * -> outs[g] [2][nl][N]; input_interval = N, output_interval = N / 2^ceil(log2 max count).  One assemble launch, then the tree of b2_pack_rlwes. */
b2_status b2_pack_lwes(b2_ctx *ctx, const uint64_t *const *c0s, const uint64_t *const *c1s, const size_t *counts, size_t groups, size_t nl,
                       const uint64_t *galois_elements, const b2_kswitch_key *const *keys, size_t n_keys, int apply_field_trace, int output_ntt_form,
                       uint64_t *const *outs, void *stream);
/* Evaluator::bfv_multiply / bfv_square device branch (src/evaluator.cu:29-116,199-282) = BEHZ RNS multiplication:
 * RNSTool::fast_b_conv_m_tilde_sm_mrq + NTT + dyadic convolution + INTT + RNSTool::fast_floor_fast_b_conv_sk
 * (src/fgk/rns_tool.cu:7-343).  a, b [batch][2][nl][N] coefficient form -> out [batch][3][nl][N] coefficient form.
 * The context must be BFV with a plain modulus; the auxiliary bases B, m_sk (61-bit primes = 1 mod 2N descending from
 * 2^61, src/utils/rns_tool.cu:59-90) are derived inside the library and can be read back with b2_behz_bases. */
b2_status b2_bfv_multiply(b2_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t batch, size_t nl, void *stream);
b2_status b2_behz_bases(b2_ctx *ctx, size_t nl, uint64_t *bsk /* >= K+2 */, size_t *bsk_count, uint64_t *gamma);
/* utils::ntt_multiply_inv_degree (src/utils/ntt.cu:78-149): x <- x * (N^-1 * scalar) mod q_j, in place;
 * data [count][ncomp][N], limb j uses modulus table_base + j */
b2_status b2_scale_inv_degree(b2_ctx *ctx, uint64_t *data, size_t count, size_t ncomp, size_t table_base, uint64_t scalar, void *stream);
/* utils::negacyclic_shift_ps (src/utils/poly_small_mod.cu:902-949); out of place */
b2_status b2_negacyclic_shift(b2_ctx *ctx, const uint64_t *a, size_t shift, uint64_t *out, size_t count, size_t ncomp, size_t table_base, void *stream);
/* GaloisTool::apply_ps / apply_ntt_ps (src/utils/galois.cu:141-300); out of place */
b2_status b2_galois(b2_ctx *ctx, const uint64_t *in, uint64_t galois_elt, uint64_t *out, size_t count, size_t ncomp, size_t table_base, int is_ntt, void *stream);

#ifdef __cplusplus
}
#endif
#endif
