// b2_host_test.cpp — scripted driver of the C++ host mirror (b2_evaluator.h) for tests/test_host_mirror_gpu.py.
//   b2_host_test <input.bin> <output.bin>
// input  (uint64 words): scheme, log_n, K, t, G, galois_element[G], moduli[K], relin key [L][2][K][N],
//                        galois keys G x [L][2][K][N], ct a [2][L][N], ct b [2][L][N], plain [L][N] (NTT form),
//                        secret key [K][N], public key [2][K][N], message (BFV / BGV: [N] words modulo t; CKKS: [L][N] NTT form)
// output (uint64 words): the result ciphertexts of the script below, concatenated, then one word per error-behaviour probe.
// The python test recomputes every result with the CPU oracle and compares bit for bit.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "b2_crypto.h"

using namespace b2troy;

static std::vector<uint64_t> read_all(const char *path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", path); std::exit(2); }
    const size_t bytes = (size_t)f.tellg();
    std::vector<uint64_t> v(bytes / 8);
    f.seekg(0);
    f.read((char *)v.data(), (std::streamsize)(v.size() * 8));
    return v;
}

int main(int argc, char **argv) {
    if (argc != 3) { std::fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    const std::vector<uint64_t> in = read_all(argv[1]);
    size_t pos = 0;
    auto take = [&](size_t n) { const uint64_t *p = in.data() + pos; pos += n; if (pos > in.size()) { std::fprintf(stderr, "input truncated\n"); std::exit(2); } return p; };
    const uint64_t *hdr = take(5);
    const SchemeType scheme = (SchemeType)hdr[0];
    const size_t log_n = hdr[1], K = hdr[2], G = hdr[4], n = (size_t)1 << log_n, L = K - 1;
    const uint64_t t = hdr[3];
    const uint64_t *gelts = take(G);
    const uint64_t *moduli = take(K);
    const size_t key_words = L * 2 * K * n;
    std::vector<uint64_t> out;
    auto emit = [&](const Ciphertext &c) { const std::vector<uint64_t> h = c.to_host(); out.insert(out.end(), h.begin(), h.end()); };
    try {
        HeContextPointer ctx = HeContext::create(scheme, n, std::vector<uint64_t>(moduli, moduli + K), t, 0);
        RelinKeys rk(ctx);
        rk.set_key(RelinKeys::get_index(2), take(key_words), L);
        GaloisKeys gk(ctx);
        for (size_t i = 0; i < G; i++) gk.set_galois_key((size_t)gelts[i], take(key_words), L);
        const bool ntt = scheme != SchemeType::BFV;
        const double scale = scheme == SchemeType::CKKS ? 1048576.0 : 1.0;
        Ciphertext a = Ciphertext::from_host(ctx, take(2 * L * n), 2, L, ntt, scale);
        Ciphertext b = Ciphertext::from_host(ctx, take(2 * L * n), 2, L, ntt, scale);
        const uint64_t *plain_words = take(L * n);
        Plaintext plain = Plaintext::from_host(ctx, plain_words, L, true);
        Evaluator ev(ctx);

        Ciphertext m = ev.multiply_new(a, b);                     emit(m);        // 1
        Ciphertext r = ev.relinearize_new(m, rk);                 emit(r);        // 2
        Ciphertext rot1, rot3, conj;
        if (scheme == SchemeType::CKKS) { ev.rotate_vector(r, 1, gk, rot1); ev.rotate_vector(r, 3, gk, rot3); ev.complex_conjugate(r, gk, conj); }
        else { ev.rotate_rows(r, 1, gk, rot1); ev.rotate_rows(r, 3, gk, rot3); ev.rotate_columns(r, gk, conj); }
        emit(rot1); emit(rot3); emit(conj);                                       // 3, 4 (NAF: -1 then 4), 5
        Ciphertext down = scheme == SchemeType::CKKS ? ev.rescale_to_next_new(r) : ev.mod_switch_to_next_new(r);
        emit(down);                                                               // 6
        Ciphertext sq = ev.square_new(a);                         emit(sq);       // 7
        emit(ev.add_new(a, b)); emit(ev.sub_new(a, b)); emit(ev.negate_new(a));   // 8, 9, 10
        {                                                                         // 11: multiply_plain on NTT-form operands
            Ciphertext an = a.clone();
            if (!ntt) ev.transform_to_ntt_inplace(an);
            Ciphertext mp; ev.multiply_plain(an, plain, mp);
            if (!ntt) ev.transform_from_ntt_inplace(mp);
            emit(mp);
        }
        {                                                                         // 12, 13: relinearize_batched
            Ciphertext o0, o1;
            ev.relinearize_batched({&m, &sq}, rk, {&o0, &o1});
            emit(o0); emit(o1);
        }
        {                                                                         // 14, 15: rotate batched (direct key), 16, 17: NAF path
            Ciphertext o0, o1, p0, p1;
            if (scheme == SchemeType::CKKS) { ev.rotate_vector_batched({&r, &rot1}, 1, gk, {&o0, &o1}); ev.rotate_vector_batched({&r, &rot1}, 3, gk, {&p0, &p1}); }
            else { ev.rotate_rows_batched({&r, &rot1}, 1, gk, {&o0, &o1}); ev.rotate_rows_batched({&r, &rot1}, 3, gk, {&p0, &p1}); }
            emit(o0); emit(o1); emit(p0); emit(p1);
        }
        {                                                                         // 18: apply_keyswitching with the relin key as a plain kswitch key
            Ciphertext ks; ev.apply_keyswitching(r, rk, ks); emit(ks);
        }
        {                                                                         // 19, 20: multiply_plain with coefficient-form operands
            Ciphertext an = a.clone(), ac = a.clone();
            if (!ntt) ev.transform_to_ntt_inplace(an); else ev.transform_from_ntt_inplace(ac);
            Ciphertext pw = Ciphertext::from_host(ctx, plain_words, 1, L, true, 1.0);
            ev.transform_from_ntt_inplace(pw);
            Plaintext plain_c = Plaintext::from_host(ctx, pw.to_host().data(), L, false);
            Ciphertext cc, nc;
            ev.multiply_plain(ac, plain_c, cc); emit(cc);                         //   both coefficient form -> coefficient form
            ev.multiply_plain(an, plain_c, nc); emit(nc);                         //   NTT-form ciphertext, coefficient-form plaintext -> NTT form
            Ciphertext bn = b.clone();
            if (!ntt) ev.transform_to_ntt_inplace(bn);
            Ciphertext o0, o1;                                                    // 21, 22: multiply_plain_batched, one launch
            ev.multiply_plain_batched({&an, &bn}, {&plain, &plain}, {&o0, &o1});
            emit(o0); emit(o1);
            Ciphertext d0, d1;                                                    // 23, 24: multiply_plain_accumulate with a repeated destination
            ev.multiply_plain_accumulate({&an, &bn, &bn}, {&plain, &plain, &plain}, std::vector<Ciphertext *>{&d0, &d1, &d0}, true);
            ev.multiply_plain_accumulate({&an}, {&plain}, d1, false);
            emit(d0); emit(d1);
        }
        std::vector<uint64_t> crypto_probes;
        {                                                                         // 25..: Encryptor / Decryptor / add_plain
            const uint64_t *sk_words = take(K * n), *pk_words = take(2 * K * n);
            const bool ckks = scheme == SchemeType::CKKS;
            const uint64_t *msg_words = take(ckks ? L * n : n);
            SecretKey sk = SecretKey::from_host(ctx, sk_words);
            Encryptor enc(ctx, 4242);
            auto cprobe = [&](auto &&fn) { try { fn(); crypto_probes.push_back(0); } catch (const std::invalid_argument &) { crypto_probes.push_back(1); } };
            Plaintext msg = ckks ? Plaintext::from_host(ctx, msg_words, L, true) : Plaintext::from_host_mod_t(ctx, msg_words, n);
            cprobe([&] { Ciphertext d; enc.encrypt_asymmetric(msg, d); });               // no public key set
            cprobe([&] { Ciphertext d; enc.encrypt_zero_symmetric(false, d); });         // no secret key set
            enc.set_public_key(PublicKey::from_host(ctx, pk_words));
            enc.set_secret_key(SecretKey::from_host(ctx, sk_words));
            Ciphertext c1 = enc.encrypt_asymmetric_new(msg);            emit(c1);        // 25
            Ciphertext c2 = enc.encrypt_symmetric_new(msg, false);      emit(c2);        // 26: the same generator continues
            Ciphertext z = enc.encrypt_zero_symmetric_new(true, 1);     emit(z);         // 27: last level, seed saved
            out.push_back(z.seed());
            Decryptor dec(ctx, sk);
            auto emit_plain = [&](const Plaintext &p) { const std::vector<uint64_t> h = p.to_host(); out.insert(out.end(), h.begin(), h.end()); };
            emit_plain(dec.decrypt_new(c1));                                            // 28
            Ciphertext prod = ev.multiply_new(c1, c2);
            emit_plain(dec.decrypt_new(prod));                                          // 29: three polynomials, s^2 computed on demand
            {
                Plaintext p0, p1;                                                       // 30, 31: decrypt_batched
                dec.decrypt_batched({&c2, &c1}, {&p0, &p1});
                emit_plain(p0); emit_plain(p1);
            }
            {                                                                           // 32-34: add_plain / sub_plain / multiply_plain with the message
                Ciphertext ap, sp, mp2;                                                 //        (BFV / BGV: modulo t -> scale-up resp. centralize + NTT; CKKS: RNS)
                ev.add_plain(c1, msg, ap); ev.sub_plain(c1, msg, sp); ev.multiply_plain(c1, msg, mp2);
                emit(ap); emit(sp); emit(mp2);
            }
            {                                                                           // 34: RLWE packing of three ciphertexts (two layers, no trace): keys 3 and 5
                Ciphertext packed = ev.pack_rlwe_ciphertexts_new({&a, &b, &r}, gk, 0, n, n / 4, false);
                emit(packed);
                LWECiphertext l0 = ev.extract_lwe_new(a, 0), l7 = ev.extract_lwe_new(b, 7), l9 = ev.extract_lwe_new(r, 9);       // 35-37: NTT-form inputs are transformed first
                { const std::vector<uint64_t> v0 = l7.c0_to_host(), v1 = l7.c1_to_host(); out.insert(out.end(), v0.begin(), v0.end()); out.insert(out.end(), v1.begin(), v1.end()); }
                emit(l9.assemble_lwe());
                emit(ev.pack_lwe_ciphertexts_new({&l0, &l7, &l9}, gk, false));                                                       //        three LWEs -> N / 4 slots, keys 3 and 5
                GaloisKeys none(ctx);
                cprobe([&] { ev.pack_rlwe_ciphertexts_new({&a, &b}, none, 0, n, n / 2, false); });          // key missing
                cprobe([&] { ev.pack_rlwe_ciphertexts_new({&a, &b, &r}, gk, 0, n, n / 2, false); });        // more ciphertexts than slots
            }
            {                                                                           // 38-41: KeyGenerator from a seed: secret, public, relinearisation, Galois(3) keys
                RandomGenerator kprng(0x5eed);
                KeyGenerator kg(ctx, kprng);
                { const std::vector<uint64_t> v = kg.secret_key().to_host(); out.insert(out.end(), v.begin(), v.end()); }
                { const std::vector<uint64_t> v = kg.create_public_key().to_host(); out.insert(out.end(), v.begin(), v.end()); }
                RelinKeys krk = kg.create_relin_keys();
                { const std::vector<uint64_t> v = krk.key_to_host(RelinKeys::get_index(2)); out.insert(out.end(), v.begin(), v.end()); }
                GaloisKeys kgk = kg.create_galois_keys({3, 3});
                { const std::vector<uint64_t> v = kgk.key_to_host(GaloisKeys::get_index(3)); out.insert(out.end(), v.begin(), v.end()); }
                // and they work: a fresh encryption under the generated public key decrypts to the message under the generated secret key
                Encryptor kenc(ctx, 99);
                kenc.set_public_key(kg.create_public_key());
                Decryptor kdec(ctx, kg.secret_key());
                emit_plain(kdec.decrypt_new(kenc.encrypt_asymmetric_new(msg)));
            }
            if (!ckks) {                                                                // BatchEncoder: encode the message words as slots, decode them back
                BatchEncoder be(ctx);
                const std::vector<uint64_t> slots(msg_words, msg_words + n / 2 + 1);
                Plaintext ep = be.encode_new(slots);
                emit_plain(ep);
                const std::vector<uint64_t> back = be.decode_new(ep);
                out.insert(out.end(), back.begin(), back.end());
            }
            cprobe([&] { Plaintext d; dec.decrypt(z, d); });                             // contains a seed
            cprobe([&] { Ciphertext t2 = c1.clone(); t2.is_ntt_form() = !t2.is_ntt_form(); Plaintext d; dec.decrypt(t2, d); });   // wrong form for the scheme
        }
        // error behaviour probes: 1 = std::invalid_argument thrown as the reference does
        auto probe = [&](auto &&fn) { try { fn(); out.push_back(0); } catch (const std::invalid_argument &) { out.push_back(1); } };
        probe([&] { Ciphertext d; ev.relinearize(m, RelinKeys(ctx), d); });                       // key not present
        probe([&] { Ciphertext d; if (scheme == SchemeType::CKKS) ev.rotate_vector(r, 5, gk, d); else ev.rotate_rows(r, 5, gk, d); });   // NAF needs a missing key (4 + 1 ok? 5 = 1 + 4: present) -> expect 0
        probe([&] { Ciphertext d; if (scheme == SchemeType::CKKS) ev.rotate_vector(r, 2, gk, d); else ev.rotate_rows(r, 2, gk, d); });   // single-term NAF, key missing -> 1
        probe([&] { Ciphertext d; ev.add(a, down, d); });                                          // level mismatch
        probe([&] { Ciphertext d; if (scheme == SchemeType::CKKS) ev.rotate_rows(r, 1, gk, d); else ev.rotate_vector(r, 1, gk, d); });   // wrong scheme
        probe([&] { Ciphertext c3 = m.clone(); Ciphertext d; ev.apply_galois(c3, (size_t)gelts[0], gk, d); });                           // size-3 ciphertext
        out.insert(out.end(), crypto_probes.begin(), crypto_probes.end());
        ctx->synchronize();
    } catch (const std::exception &e) {
        std::fprintf(stderr, "b2_host_test: %s\n", e.what());
        return 1;
    }
    std::ofstream f(argv[2], std::ios::binary);
    f.write((const char *)out.data(), (std::streamsize)(out.size() * 8));
    return 0;
}
