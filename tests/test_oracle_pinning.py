"""CPU tests that pin the oracle (oracle/rns_oracle.c):
 (i)  the reference's own known-answer vectors (test/utils/ntt.cu:33-75, test/utils/uint_small_mod.cu:395-404,
      :468-475, :680-690, test/utils/galois.cu:20-76),
 (ii) golden vectors produced by the unmodified reference (tests/golden/, generator oracle/make_golden.py),
 (iii) live comparison with oracle/_ref when it is built (this container), incl. the SURVEY §8c fingerprints.
"""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- (i) reference KATs -----------------------------------------------------------------------
def test_kat_primitive_roots_and_small_ntt(po):
    p = 0xFFFFFFFFFFC0001
    o1 = po.Oracle(1, [p], 0, po.CKKS)
    assert o1.root_powers(0)[0, 0] == 1 and o1.root_powers(0)[1, 0] == 288794978602139552
    assert o1.root_powers(0, True)[1, 0] == po.lib().ro_invmod(288794978602139552, p)
    o2 = po.Oracle(2, [p], 0, po.CKKS)
    assert [int(v) for v in o2.root_powers(0)[:, 0]] == [1, 288794978602139552, 178930308976060547, 748001537669050592]
    z = np.array([0, 0], dtype=np.uint64); assert list(o1.ntt_fwd(z, 1, 1)) == [0, 0]
    assert list(o1.ntt_fwd(np.array([1, 0], dtype=np.uint64), 1, 1)) == [1, 1]
    assert list(o1.ntt_fwd(np.array([1, 1], dtype=np.uint64), 1, 1)) == [288794978602139553, 864126526004445282]
    o5 = po.Oracle(5, [p], 0, po.CKKS)
    x = np.arange(32, dtype=np.uint64)
    assert np.array_equal(o5.ntt_inv(o5.ntt_fwd(x, 1, 1), 1, 1), x)
    assert not o5.ntt_inv(np.zeros(32, dtype=np.uint64), 1, 1).any()


def test_kat_uint_small_mod(po):
    L = po.lib()
    m = 13131313131313
    assert L.ro_barrett_reduce128(0, 0, m) == 0 and L.ro_barrett_reduce128(1, 0, m) == 1
    assert L.ro_barrett_reduce128(123, 456, m) == 8722750765283
    assert L.ro_barrett_reduce128(24242424242424, 79797979797979, m) == 1010101010101
    q = 2305843009211596801
    assert L.ro_mulmod(1152921504605798400, 1152921504605798401, q) == 576460752302899200
    assert L.ro_mulmod(1152921504605798401, 1152921504605798401, q) == 1729382256908697601
    assert L.ro_mulmod(2305843009211596800, 2305843009211596800, q) == 1
    assert L.ro_mulmod(6, 7, 10) == 2
    assert L.ro_powmod(2424242424, 16, 131313131313) == 39418477653
    assert L.ro_powmod(2, 0xFFFFFFFFFFFFFFFF, 5) == 3
    # Shoup multiply agrees with plain modmul on random operands (lazy range [0, 2p))
    rng = np.random.default_rng(0)
    for p in (q, 1099511480321, 1032193):
        for _ in range(200):
            x = int(rng.integers(0, 1 << 63)) * 2 + 1; w = int(rng.integers(0, p))
            r = L.ro_shoup_mul_lazy(x, w, L.ro_shoup_quotient(w, p), p)
            assert r < 2 * p and r % p == (x * w) % p
            assert L.ro_barrett_reduce64(x, p) == x % p


def test_kat_galois(po):
    assert [po.lib().ro_galois_element_from_step(8, s) for s in (0, 1, -3, 2, -2, 3, -1)] == [15, 3, 3, 9, 9, 11, 11]
    o = po.Oracle(3, [17], 0, po.CKKS)
    x = np.arange(8, dtype=np.uint64)
    assert list(o.galois(x, 3, 1, 1, False)) == [0, 14, 6, 1, 13, 7, 2, 12]
    assert list(o.galois(x, 3, 1, 1, True)) == [4, 5, 7, 6, 1, 0, 2, 3]


def test_prime_generation_matches_reference_defaults(po):
    # CoeffModulus::bfv_default(32768) first / last primes (src/coeff_modulus.cu:53-56) are what get_primes finds
    assert int(po.coeff_modulus_create(4096, [50, 40, 40, 50, 20])[4]) == 0xFC001
    m = po.coeff_modulus_create(4096, [50, 40, 40, 50])
    assert [hex(int(v)) for v in m] == ["0x3fffffffcc001", "0xfffffc6001", "0xfffffdc001", "0x3ffffffffc001"]


# ---- (ii) golden vectors from the unmodified reference -----------------------------------------
@pytest.mark.parametrize("name", ["bfv", "ckks", "bgv"])
def test_oracle_reproduces_reference_golden(po, name):
    g = np.load(os.path.join(GOLD, "ref_%s_n256.npz" % name))
    n, scheme = int(g["n"]), int(g["scheme"])
    moduli = po.coeff_modulus_create(n, [int(b) for b in g["bits"]])
    assert np.array_equal(moduli, g["moduli"])
    t = po.plain_modulus_batching(n, 20) if scheme != po.CKKS else 0
    assert t == int(g["t"])
    o = po.Oracle(8, moduli, t, scheme)
    K = o.nm; L = K - 1
    for i in range(K):
        assert o.root(i) == int(g["roots"][i])
        assert np.array_equal(o.root_powers(i), g["root_powers_%d" % i])
        assert np.array_equal(o.root_powers(i, True), g["inv_root_powers_%d" % i])
    gal = int(g["galois_elt"]); assert gal == o.galois_element_from_step(1)
    ntt = scheme != po.BFV
    for lvl in (L, L - 1):
        s = "_l%d" % lvl
        a, b, c3 = g["a" + s], g["b" + s], g["c3" + s]
        sz = lvl * n
        if ntt:
            assert np.array_equal(o.dyadic_convolute(a, 2, b, 2, lvl), g["mul" + s])
            assert np.array_equal(o.dyadic_square(a, lvl), g["square" + s])
        else:       # BFV: BEHZ multiply (src/evaluator.cu:29-116), pinned word for word
            assert np.array_equal(o.bfv_multiply(a, b, lvl), g["mul" + s])
            assert np.array_equal(o.bfv_multiply(a, a, lvl), g["square" + s])
        assert np.array_equal(o.keyswitch(c3[2 * sz:], lvl, ntt, g["relin_key"], dest=c3[:2 * sz], assign=po.ASSIGN_ADD_INPLACE), g["relin" + s])
        ga = o.galois(a, gal, 2, lvl, ntt)
        assert np.array_equal(o.keyswitch(ga[sz:], lvl, ntt, g["galois_key"], dest=ga, assign=po.ASSIGN_OVERWRITE_EXCEPT_FIRST), g["rotate1" + s])
        fn = {po.BFV: o.divide_round_q_last, po.CKKS: o.divide_round_q_last_ntt, po.BGV: o.mod_t_divide_q_last_ntt}[scheme]
        assert np.array_equal(fn(a, 2, lvl), g["modswitch" + s])
        if scheme == po.CKKS:
            assert np.array_equal(o.drop_last_limb(a, 2, lvl), g["drop" + s])
        assert np.array_equal(o.galois(a, gal, 2, lvl, False), g["galois_coeff" + s])
        assert np.array_equal(o.galois(a, gal, 2, lvl, True), g["galois_ntt" + s])
        an = a if ntt else o.ntt_fwd(a, 2, lvl)
        assert np.array_equal(o.dyadic_broadcast(an, g["plain" + s], 2, lvl), g["mulplain_ntt" + s])
        assert np.array_equal(o.ntt_fwd(a, 2, lvl), g["to_ntt" + s]) and np.array_equal(o.ntt_inv(a, 2, lvl), g["from_ntt" + s])
    assert np.array_equal(o.ntt_fwd(g["ks_set_in"], L + 1, L, po.TABLES_KS_SET_PRODUCTS, 0, L), g["ks_set_out"])
    assert np.array_equal(o.ntt_inv(g["ks_skip_in"], 2, L + 1, po.TABLES_KS_SKIP_FINALS, 0, L), g["ks_skip_out"])
    # the reference's own ciphertext decrypts to ~0 under its own secret key with the oracle's arithmetic
    ph = o.phase(g["secret_key"], g["ct"], 2, L, ntt)
    q0 = int(moduli[0]); c = ph[:n].astype(object) % q0
    if scheme == po.BGV:      # BGV noise is t*e
        pass
    else:
        assert max(min(int(v), q0 - int(v)) for v in c) < 2 ** 16


def test_oracle_decryptor_reproduces_reference_golden(po):
    """Decryptor (src/decryptor.cu:27-535, rns_tool.cu:1118-1430, rns_base.cu:445-583): golden vectors from the unmodified reference"""
    g = np.load(os.path.join(GOLD, "ref_decrypt_n256.npz"))
    n = int(g["n"])
    for scheme, name in ((po.BFV, "bfv"), (po.CKKS, "ckks"), (po.BGV, "bgv")):
        moduli = g[name + "_moduli"]
        o = po.Oracle(8, moduli, int(g[name + "_t"]) if scheme != po.CKKS else 0, scheme)
        L = o.nm - 1
        skp = o.secret_key_powers(g[name + "_sk"], 2)
        assert np.array_equal(o.decrypt(g[name + "_zero"], 2, L, skp), g[name + "_zero_dec"])
        if scheme != po.CKKS:
            assert not g[name + "_zero_dec"].any()
        for nl in (L, 1):
            for pc in (2, 3):
                key = "%s_l%d_p%d" % (name, nl, pc)
                ct = g[key + "_ct"]
                assert np.array_equal(o.decrypt(ct, pc, nl, skp), g[key + "_dec"]), key
                if scheme == po.BGV:
                    assert np.array_equal(o.decrypt(ct, pc, nl, skp, 7), g[key + "_dec_cf7"]), key
                if scheme == po.BFV:
                    ph = o.decrypt_phase(ct, pc, nl, False, skp)
                    assert np.array_equal(ph, g[key + "_phase"]), key
                    assert np.array_equal(o.decrypt_scale_and_round(ph, nl), g[key + "_dec"]), key


def test_oracle_encryptor_reproduces_reference_golden(po):
    """generator, samplers, Encryptor (src/utils/random_generator.cu:12-505, src/utils/rlwe.cu:11-322, src/encryptor.cu:12-338): golden
    vectors from the unmodified reference with explicit seeds"""
    g = np.load(os.path.join(GOLD, "ref_encrypt_n256.npz"))
    for scheme, name in ((po.BFV, "bfv"), (po.CKKS, "ckks"), (po.BGV, "bgv")):
        o = po.Oracle(8, g[name + "_moduli"], int(g[name + "_t"]) if scheme != po.CKKS else 0, scheme)
        K = o.nm; L = K - 1
        st = o.prng_state(77, 3, 9)
        assert np.array_equal(o.prng_fill_bytes(st, 16 * 7 + 3), g[name + "_bytes"])
        for kn in ("ternary", "cbd", "uniform"):
            assert np.array_equal(o.sample_poly(kn, st, K), g["%s_%s" % (name, kn)]), kn
        assert np.array_equal(st, g[name + "_state_after"])
        pk, sk = g[name + "_pk"], g[name + "_sk"]
        ntt = scheme != po.BFV
        for nl in (K, L, 1):
            assert np.array_equal(o.encrypt_zero_asymmetric(pk, nl, ntt, o.prng_state(1234, 5)), g["%s_l%d_asym" % (name, nl)]), (name, nl)
            for save in (0, 1):
                ct, seed = o.encrypt_zero_symmetric(sk, nl, ntt, save, o.prng_state(99))
                assert np.array_equal(ct, g["%s_l%d_sym%d" % (name, nl, save)]), (name, nl, save)
                assert (seed if save else 0) == int(g["%s_l%d_sym%d_seed" % (name, nl, save)])
        for asym in (0, 1):
            assert np.array_equal(o.encrypt(g[name + "_plain"], pk if asym else sk, asym, L, o.prng_state(4242)), g["%s_enc%d" % (name, asym)]), (name, asym)
        # and the loop closes inside the oracle: decrypt(encrypt(m)) == m
        if scheme != po.CKKS:
            m = g[name + "_plain"]
            dec = o.decrypt(g[name + "_enc1"], 2, L, o.secret_key_powers(sk, 1))
            assert np.array_equal(dec[:m.size], m) and not dec[m.size:].any()


def test_oracle_packing_reproduces_reference_golden(po):
    """Evaluator::pack_rlwe_ciphertexts_new + field trace (src/evaluator_lwes.cu:96-105,305-491): golden vectors from the unmodified reference"""
    g = np.load(os.path.join(GOLD, "ref_pack_n64.npz"))
    n = int(g["n"])
    for scheme, name in ((po.BFV, "bfv"), (po.CKKS, "ckks"), (po.BGV, "bgv")):
        o = po.Oracle(6, g[name + "_moduli"], int(g[name + "_t"]) if scheme != po.CKKS else 0, scheme)
        L = o.nm - 1
        keys = {int(e): g["%s_key_%d" % (name, int(e))] for e in g[name + "_elts"]}
        for ci, (count, shift, ii, oi, ft) in enumerate(g["cases"]):
            cts = np.split(g["%s_c%d_in" % (name, ci)], int(count))
            for inp_ntt in ((0, 1) if scheme != po.BFV else (0,)):
                got = o.pack_rlwes(cts, bool(inp_ntt), L, int(shift), int(ii), int(oi), keys, bool(ft))
                assert np.array_equal(got, g["%s_c%d_out%d" % (name, ci, inp_ntt)]), (name, ci, inp_ntt)


def _replay_keygen(po, o, seed, n):
    K = o.nm
    st = o.prng_state(seed)
    sk = o.keygen_secret_prng(st)
    pk, _ = o.encrypt_zero_symmetric(sk, K, True, False, st)
    relin = o.keygen_kswitch_prng(sk, o.secret_key_powers(sk, 2)[K * n:], st)
    g3 = o.keygen_kswitch_prng(sk, o.galois(sk, 3, 1, K, True), st)
    gc = o.keygen_kswitch_prng(sk, o.galois(sk, 2 * n - 1, 1, K, True), st)
    return sk, pk, relin, g3, gc


def test_oracle_keygen_reproduces_reference_golden(po):
    """KeyGenerator from a context seed (src/key_generator.cu:31-78,136-256): every key the unmodified reference generated, in its drawing order"""
    g = np.load(os.path.join(GOLD, "ref_keygen_n256.npz"))
    n = int(g["n"])
    for scheme, name in ((po.BFV, "bfv"), (po.CKKS, "ckks"), (po.BGV, "bgv")):
        o = po.Oracle(8, g[name + "_moduli"], int(g[name + "_t"]) if scheme != po.CKKS else 0, scheme)
        sk, pk, relin, g3, gc = _replay_keygen(po, o, int(g["seed"]), n)
        for got, key in ((sk, "_sk"), (pk, "_pk"), (relin, "_relin"), (g3, "_galois_3"), (gc, "_galois_conj")):
            assert np.array_equal(got, g[name + key]), (name, key)


# ---- (iii) live against the compiled reference --------------------------------------------------
SURVEY_FPR = {2: ("3866d51961505bd9", "c59ae60bda39e030", "a0e0a04249a08cd3", "4a186e8776597c7e", "eb819b8149c577a2"),
              3: ("82ff3df13af52835", "2046b27467209ab9", "a967cf538de6dc20", "def3d7ec042efe31", "e11839bb23d6a573"),
              1: ("6f35b18b9c970541", "18a28bad21f66e14", "e677dc9a95a7f003", "437f3e313696d5a6", "e670adc1a8c7459c")}


@pytest.mark.parametrize("scheme", [2, 3, 1])
def test_live_reference_fingerprints_and_oracle_parity(po, scheme):
    if not po.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference; run __graft_entry__.build())")
    n = 4096
    r = po.Reference(scheme, n, [50, 40, 40, 50], t_bits=20, seed=0x123)
    r.gen_relin(); gal = r.galois_element_from_step(1); r.gen_galois([gal])
    L = r.nm - 1
    c = r.encrypt_zero(); m = r.multiply(c, c, L); rl = r.relinearize(m, L); ro = r.rotate(rl, 1, L); ms = r.mod_switch_scale_to_next(rl, 2, L)
    assert tuple("%016x" % po.fnv1a64(x) for x in (c, m, rl, ro, ms)) == SURVEY_FPR[scheme]
    o = po.Oracle(12, r.moduli, r.t, scheme)
    ntt = scheme != po.BFV
    sz = L * n
    assert np.array_equal(o.keyswitch(m[2 * sz:], L, ntt, r.relin_key(), dest=m[:2 * sz], assign=po.ASSIGN_ADD_INPLACE), rl)
    ga = o.galois(rl, gal, 2, L, ntt)
    assert np.array_equal(o.keyswitch(ga[sz:], L, ntt, r.galois_key(gal), dest=ga, assign=po.ASSIGN_OVERWRITE_EXCEPT_FIRST), ro)
    fn = {po.BFV: o.divide_round_q_last, po.CKKS: o.divide_round_q_last_ntt, po.BGV: o.mod_t_divide_q_last_ntt}[scheme]
    assert np.array_equal(fn(rl, 2, L), ms)
    if ntt:
        assert np.array_equal(o.dyadic_convolute(c, 2, c, 2, L), m)
    else:
        assert np.array_equal(o.bfv_multiply(c, c, L), m)


@pytest.mark.parametrize("scheme,log_n,bits,t_bits", [(1, 12, [40, 40, 50, 55], 20), (1, 10, [60, 60, 60], 33), (2, 11, [50, 40, 40, 55], 20),
                                                        (3, 12, [40, 45, 50, 55], 20), (3, 10, [60, 60], 17), (1, 9, [60] * 6, 59)])
def test_live_reference_decryptor(po, scheme, log_n, bits, t_bits):
    """oracle Decryptor == unmodified reference, word for word, on random ciphertexts (2 and 3 polynomials, first and last level,
    BGV correction factors) under the reference's secret key"""
    if not po.ref_available():
        pytest.skip("oracle/_ref not built")
    n = 1 << log_n
    r = po.Reference(scheme, n, bits, t_bits=t_bits, seed=11)
    o = po.Oracle(log_n, r.moduli, r.t if scheme != po.CKKS else 0, scheme)
    rng = np.random.default_rng(scheme * 100 + log_n)
    skp = o.secret_key_powers(r.secret_key(), 2)
    for nl in (len(bits) - 1, 1):
        for pc in (2, 3):
            ct = o.random_poly(rng, pc, nl)
            for cf in ((1, 5) if scheme == po.BGV else (1,)):
                assert np.array_equal(o.decrypt(ct, pc, nl, skp, cf), r.decrypt_raw(ct, pc, nl, cf)), (nl, pc, cf)
            if scheme != po.BGV:
                assert np.array_equal(o.decrypt_phase(ct, pc, nl, scheme == po.CKKS, skp), r.decrypt_phase_raw(ct, pc, nl))


@pytest.mark.parametrize("scheme,bits,t_bits", [(1, [40, 40, 50, 55], 20), (2, [50, 40, 40, 55], 20), (3, [40, 45, 50, 55], 20), (1, [60, 60, 60], 33), (3, [60] * 4, 59)])
def test_live_reference_encryptor(po, scheme, bits, t_bits):
    """oracle generator / samplers / Encryptor == unmodified reference with the same seeds, word for word"""
    if not po.ref_available():
        pytest.skip("oracle/_ref not built")
    n, logn = 1024, 10
    r = po.Reference(scheme, n, bits, t_bits=t_bits)
    o = po.Oracle(logn, r.moduli, r.t if scheme != po.CKKS else 0, scheme)
    K, L = r.nm, r.nm - 1
    st = o.prng_state(77, 3, 2 ** 64 - 3); st2 = st.copy()          # the counter crosses a 64-bit boundary
    assert np.array_equal(r.prng(0, st, nbytes=16 * 7 + 3), o.prng_fill_bytes(st2, 16 * 7 + 3)) and np.array_equal(st, st2)
    for kind, name in ((1, "ternary"), (2, "cbd"), (3, "uniform")):
        assert np.array_equal(r.prng(kind, st, nl=K), o.sample_poly(name, st2, K)), name
        assert np.array_equal(st, st2), name
    pk, sk = r.public_key(), r.secret_key()
    ntt = scheme != po.BFV
    for nl in (K, L, 1):
        assert np.array_equal(r.encrypt_seeded(None, True, nl, 1234, 5)[0], o.encrypt_zero_asymmetric(pk, nl, ntt, o.prng_state(1234, 5))), ("asym", nl)
        for save in (False, True):
            want, seed = r.encrypt_seeded(None, False, nl, 99, 0, save_seed=save)
            got, gseed = o.encrypt_zero_symmetric(sk, nl, ntt, save, o.prng_state(99))
            assert np.array_equal(want, got) and seed == (gseed if save else 0), ("sym", nl, save)
    rng = np.random.default_rng(scheme)
    plains = [o.random_poly(rng, 1, L)] if scheme == po.CKKS else [rng.integers(0, r.t, size=c, dtype=np.uint64) for c in (n, n // 2 + 3)]
    for plain in plains:
        for asym in (True, False):
            assert np.array_equal(r.encrypt_seeded(plain, asym, L, 4242)[0], o.encrypt(plain, pk if asym else sk, asym, L, o.prng_state(4242))), ("enc", asym)


@pytest.mark.parametrize("scheme,bits", [(1, [40, 40, 50, 55]), (2, [50, 40, 40, 55]), (3, [60, 60, 60])])
def test_live_reference_keygen(po, scheme, bits):
    if not po.ref_available():
        pytest.skip("oracle/_ref not built")
    n, S = 1024, 0x77aa55
    r = po.Reference(scheme, n, bits, t_bits=20, seed=S)
    r.gen_relin(); r.gen_galois([3, 2 * n - 1])
    o = po.Oracle(10, r.moduli, r.t if scheme != po.CKKS else 0, scheme)
    sk, pk, relin, g3, gc = _replay_keygen(po, o, S, n)
    assert np.array_equal(sk, r.secret_key()) and np.array_equal(pk, r.public_key())
    assert np.array_equal(relin, r.relin_key()) and np.array_equal(g3, r.galois_key(3)) and np.array_equal(gc, r.galois_key(2 * n - 1))


@pytest.mark.parametrize("scheme,bits", [(1, [40, 40, 50, 55]), (3, [40, 45, 50, 55]), (1, [60, 60, 60])])
def test_live_reference_plaintext_ops(po, scheme, bits):
    """multiply_plain / add_plain / sub_plain with a plaintext modulo t (centralize, BFV scale-up, the BGV correction factor):
    oracle compositions == the unmodified reference (src/evaluator_multiply_plain.cu:13-68, src/evaluator_translate_plain.cu:95-170)"""
    if not po.ref_available():
        pytest.skip("oracle/_ref not built")
    n = 512
    r = po.Reference(scheme, n, bits, t_bits=20)
    o = po.Oracle(9, r.moduli, r.t, scheme)
    rng = np.random.default_rng(scheme)
    for nl in (len(bits) - 1, 1):
        ct = o.random_poly(rng, 2, nl)
        for cnt in (n, n // 2 + 1):
            plain = rng.integers(0, r.t, size=cnt, dtype=np.uint64)
            for op in (0, 1, 2):
                for cf in ((1, 7) if scheme == po.BGV and op else (1,)):
                    assert np.array_equal(r.plain_op(op, ct, nl, plain, cf), o.plain_op(op, ct, nl, plain, cf)), (nl, cnt, op, cf)


@pytest.mark.parametrize("scheme,log_n", [(1, 10), (3, 8), (1, 13)])
def test_live_reference_batch_encoder(po, scheme, log_n):
    """BatchEncoder::encode / decode (src/batch_encoder.cu:14-60,169-227,407-460): oracle slot map + transform modulo t == the unmodified reference"""
    if not po.ref_available():
        pytest.skip("oracle/_ref not built")
    n = 1 << log_n
    r = po.Reference(scheme, n, [40, 40, 50], t_bits=20)
    o = po.Oracle(log_n, r.moduli, r.t, scheme)
    rng = np.random.default_rng(n)
    for cnt in (n, n // 2 + 1, 1):
        v = rng.integers(0, r.t, size=cnt, dtype=np.uint64)
        e = r.batch_encode(v)
        assert np.array_equal(e, o.batch_encode(v)) and np.array_equal(r.batch_decode(e), o.batch_decode(e)) and np.array_equal(o.batch_decode(e)[:cnt], v)
    p = rng.integers(0, r.t, size=n // 3, dtype=np.uint64)
    assert np.array_equal(r.batch_decode(p), o.batch_decode(p))


@pytest.mark.parametrize("scheme,bits", [(1, [40, 40, 50]), (2, [40, 40, 50]), (3, [40, 45, 55])])
def test_live_reference_packing_and_lwe(po, scheme, bits):
    """oracle pack_rlwes / extract_lwe / pack_lwes == the unmodified reference's Evaluator::pack_rlwe_ciphertexts_new, extract_lwe_new and
    pack_lwe_ciphertexts_new with the reference's automorphism keys (src/evaluator_lwes.cu:51-94,201-229,305-491)"""
    if not po.ref_available():
        pytest.skip("oracle/_ref not built")
    n, logn = 64, 6
    r = po.Reference(scheme, n, bits, t_bits=20)
    o = po.Oracle(logn, r.moduli, r.t if scheme != po.CKKS else 0, scheme)
    L = r.nm - 1
    elts = [e for e in sorted({(n >> k) + 1 for k in range(0, logn)} | {2 ** k + 1 for k in range(1, logn + 1)}) if e < 2 * n]
    r.gen_galois(elts)
    keys = {e: r.galois_key(e) for e in elts}
    rng = np.random.default_rng(scheme)
    for (count, shift, ii, oi, ft) in ((4, 0, n, n // 4, True), (8, 5, n, n // 8, False), (5, 2 * n - 3, n, n // 8, True), (1, 0, n, n, True)):
        cts = [o.random_poly(rng, 2, L) for _ in range(count)]
        for inp_ntt in ((False, True) if scheme != po.BFV else (False,)):
            assert np.array_equal(r.pack_rlwes(cts, L, inp_ntt, shift, ii, oi, ft), o.pack_rlwes(cts, inp_ntt, L, shift, ii, oi, keys, ft)), (count, shift, ii, oi, ft, inp_ntt)
    ct = o.random_poly(rng, 2, L)
    for term in (0, 1, 17, n - 1):
        w0, w1 = r.extract_lwe(ct, term, L, False)
        g0, g1 = o.extract_lwe(ct, term, L)
        assert np.array_equal(w0, g0) and np.array_equal(w1, g1), term
    for count in (1, 3, 7):
        lwes = [(o.random_poly(rng, 1, L).reshape(L, n)[:, 0].copy(), o.random_poly(rng, 1, L)) for _ in range(count)]
        assert np.array_equal(r.pack_lwes(lwes, L, True), o.pack_lwes(lwes, L, keys, True)), count


@pytest.mark.parametrize("log_n,bits,t_bits", [(6, [40, 40, 40], 20), (10, [60, 40, 40, 60], 20), (9, [50, 36, 45, 60, 55], 30), (7, [60] * 6, 59)])
def test_live_reference_behz_multiply(po, log_n, bits, t_bits):
    """BEHZ multiply at every level of several modulus chains, incl. a 60-bit plain modulus (|B| = |q| + 1 branch,
    src/utils/rns_tool.cu:59-62): oracle == unmodified reference, word for word."""
    if not po.ref_available():
        pytest.skip("oracle/_ref not built")
    n = 1 << log_n
    r = po.Reference(po.BFV, n, bits, t_bits=t_bits, seed=7)
    o = po.Oracle(log_n, r.moduli, r.t, po.BFV)
    rng = np.random.default_rng(log_n)
    for nl in range(len(bits) - 1, 0, -1):
        a = o.random_poly(rng, 2, nl); b = o.random_poly(rng, 2, nl)
        assert np.array_equal(o.bfv_multiply(a, b, nl), r.multiply(a, b, nl)), nl
        assert np.array_equal(o.bfv_multiply(a, a, nl), r.square(a, nl)), nl
