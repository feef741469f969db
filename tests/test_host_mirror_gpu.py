"""C++ host mirror (troy-nova_b200/host/b2_evaluator.{h,cpp}: HeContext / Ciphertext / KSwitchKeys / Evaluator with the
reference's method names) driven by its scripted test binary and checked bit for bit against the CPU oracle:
multiply, relinearize, rotate (direct key and NAF decomposition), conjugate / rotate_columns, rescale / mod_switch,
square, add / sub / negate, multiply_plain, the batched forms, and the reference's error behaviour."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, make_oracle

pytestmark = pytest.mark.gpu
DRIVER = os.path.join(ROOT, "troy-nova_b200", "host", "b2_host_test")


def _rotate(o, po, ct, L, is_ntt, g, gkey):
    rot = o.galois(ct, g, 2, L, is_ntt)
    return o.keyswitch(rot[L * o.n:], L, is_ntt, gkey, dest=rot, assign=po.ASSIGN_OVERWRITE_EXCEPT_FIRST)


@pytest.mark.parametrize("scheme", [1, 2, 3])
def test_host_mirror_script(po, scheme, tmp_path):
    if not os.path.exists(DRIVER):
        pytest.fail("troy-nova_b200/host/b2_host_test not built (make -C troy-nova_b200)")
    o = make_oracle(po, scheme=scheme, log_n=12, bits=[50, 40, 40, 55], t_bits=20 if scheme != 2 else 0)
    n, K = o.n, o.nm
    L = K - 1
    is_ntt = scheme != 1
    rng = np.random.default_rng(300 + scheme)
    steps = [1, -1, 4, 0]
    gelts = [o.galois_element_from_step(s) for s in steps] + [5]          # 3 (= step 1) and 5: the automorphisms of a two-layer packing
    relin = o.random_key(rng)
    gkeys = {g: o.random_key(rng) for g in gelts}
    a = o.random_poly(rng, 2, L); b = o.random_poly(rng, 2, L)
    plain = o.random_poly(rng, 1, L)
    hdr = np.array([scheme, o.log_n, K, o.t, len(gelts)] + gelts + [int(m) for m in o.moduli], dtype=np.uint64)
    sk, pk = o.random_poly(rng, 1, K), o.random_poly(rng, 2, K)
    msg = o.random_poly(rng, 1, L) if scheme == 2 else rng.integers(0, o.t, size=n, dtype=np.uint64)
    blob = np.concatenate([hdr, relin] + [gkeys[g] for g in gelts] + [a, b, plain, sk, pk, msg])
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    blob.tofile(fin)
    res = subprocess.run([DRIVER, str(fin), str(fout)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    out = np.fromfile(fout, dtype=np.uint64)

    # ---- the same script on the oracle
    mul = (lambda x, y: o.bfv_multiply(x, y, L)) if scheme == 1 else (lambda x, y: o.dyadic_convolute(x, 2, y, 2, L))
    w = L * n
    m = mul(a, b)
    relinearize = lambda c3: o.keyswitch(c3[2 * w:], L, is_ntt, relin, dest=c3[:2 * w], assign=po.ASSIGN_ADD_INPLACE)
    r = relinearize(m)
    rot = lambda c, s: _rotate(o, po, c, L, is_ntt, o.galois_element_from_step(s), gkeys[o.galois_element_from_step(s)])
    rot1 = rot(r, 1)
    rot3 = rot(rot(r, -1), 4)
    conj = rot(r, 0)
    down = {1: o.divide_round_q_last, 2: o.divide_round_q_last_ntt, 3: o.mod_t_divide_q_last_ntt}[scheme](r, 2, L)
    sq = o.bfv_multiply(a, a, L) if scheme == 1 else o.dyadic_square(a, L)
    an = a if is_ntt else o.ntt_fwd(a.copy(), 2, L)
    mp = o.dyadic_broadcast(an, plain, 2, L)
    if not is_ntt:
        mp = o.ntt_inv(mp, 2, L)
    ks = o.keyswitch(r[w:], L, is_ntt, relin, dest=r, assign=po.ASSIGN_OVERWRITE_EXCEPT_FIRST)
    want = [("multiply", m), ("relinearize", r), ("rotate 1", rot1), ("rotate 3 (NAF -1, 4)", rot3), ("conjugate / rotate_columns", conj),
            ("rescale / mod_switch_to_next", down), ("square", sq), ("add", o.add(a, b, 2, L)), ("sub", o.sub(a, b, 2, L)), ("negate", o.negate(a, 2, L)),
            ("multiply_plain", mp), ("relinearize_batched[0]", r), ("relinearize_batched[1]", relinearize(sq)),
            ("rotate_batched[0]", rot1), ("rotate_batched[1]", rot(rot1, 1)), ("rotate_batched NAF[0]", rot3), ("rotate_batched NAF[1]", rot(rot(rot1, -1), 4)),
            ("apply_keyswitching", ks)]
    bn = b if is_ntt else o.ntt_fwd(b.copy(), 2, L)
    mp_ntt = o.dyadic_broadcast(an, plain, 2, L)
    mpb_ntt = o.dyadic_broadcast(bn, plain, 2, L)
    both = o.add(mp_ntt, mpb_ntt, 2, L)
    want += [("multiply_plain coefficient x coefficient", o.ntt_inv(mp_ntt.copy(), 2, L)), ("multiply_plain ntt x coefficient", mp_ntt),
             ("multiply_plain_batched[0]", mp_ntt), ("multiply_plain_batched[1]", mpb_ntt),
             ("multiply_plain_accumulate d0 (repeated destination)", both), ("multiply_plain_accumulate d1 (+= second call)", both)]
    # ---- Encryptor / Decryptor / add_plain: one generator (seed 4242) drives the three encryptions in order
    st = o.prng_state(4242)
    c1 = o.encrypt(msg, pk, True, L, st)
    c2 = o.encrypt(msg, sk, False, L, st)
    z, zseed = o.encrypt_zero_symmetric(sk, 1, is_ntt, True, st)
    skp = o.secret_key_powers(sk, 2)
    prod = mul(c1, c2)
    want += [("encrypt_asymmetric", c1), ("encrypt_symmetric", c2), ("encrypt_zero_symmetric (saved seed, last level)", z), ("saved seed", np.array([zseed], dtype=np.uint64)),
             ("decrypt", o.decrypt(c1, 2, L, skp)), ("decrypt of a product", o.decrypt(prod, 3, L, skp)),
             ("decrypt_batched[0]", o.decrypt(c2, 2, L, skp)), ("decrypt_batched[1]", o.decrypt(c1, 2, L, skp))]
    if scheme == 2:
        want += [("add_plain", np.concatenate([o.add(c1[:w], msg, 1, L), c1[w:]])), ("sub_plain", np.concatenate([o.sub(c1[:w], msg, 1, L), c1[w:]])),
                 ("multiply_plain (message)", o.dyadic_broadcast(c1, msg, 2, L))]
    else:
        want += [("add_plain", o.plain_op(1, c1, L, msg)), ("sub_plain", o.plain_op(2, c1, L, msg)), ("multiply_plain (message modulo t)", o.plain_op(0, c1, L, msg))]
    want += [("pack_rlwe_ciphertexts_new", o.pack_rlwes([a, b, r], is_ntt, L, 0, n, n // 4, {3: gkeys[3], 5: gkeys[5]}, False))]
    coeff = (lambda c: c) if not is_ntt else (lambda c: o.ntt_inv(c.copy(), 2, L))
    l0, l7, l9 = o.extract_lwe(coeff(a), 0, L), o.extract_lwe(coeff(b), 7, L), o.extract_lwe(coeff(r), 9, L)
    want += [("extract_lwe c0", l7[0]), ("extract_lwe c1", l7[1]), ("assemble_lwe", o.assemble_lwe(l9[0], l9[1], L)),
             ("pack_lwe_ciphertexts_new", o.pack_lwes([l0, l7, l9], L, {3: gkeys[3], 5: gkeys[5]}, False))]
    kst = o.prng_state(0x5eed)
    ksk = o.keygen_secret_prng(kst)
    kpk, _ = o.encrypt_zero_symmetric(ksk, K, True, False, kst)
    krk = o.keygen_kswitch_prng(ksk, o.secret_key_powers(ksk, 2)[K * n:], kst)
    kgk = o.keygen_kswitch_prng(ksk, o.galois(ksk, 3, 1, K, True), kst)
    kpk2, _ = o.encrypt_zero_symmetric(ksk, K, True, False, kst)                 # the second create_public_key call of the script
    kct = o.encrypt(msg, kpk2, True, L, o.prng_state(99))
    want += [("KeyGenerator secret key", ksk), ("public key", kpk), ("relinearisation key", krk), ("Galois key", kgk),
             ("decrypt(encrypt(m)) under generated keys", o.decrypt(kct, 2, L, o.secret_key_powers(ksk, 1)))]
    if scheme != 2:
        slots = msg[:n // 2 + 1]
        want += [("BatchEncoder::encode", o.batch_encode(slots)), ("BatchEncoder::decode", np.concatenate([slots, np.zeros(n - slots.size, dtype=np.uint64)]))]
    pos = 0
    for name, exp in want:
        got = out[pos:pos + exp.size]; pos += exp.size
        assert np.array_equal(got, exp), f"{name} differs from the oracle (scheme {scheme})"
    probes = out[pos:]
    assert probes.tolist() == [1, 0, 1, 1, 1, 1] + [1, 1, 1, 1, 1, 1], f"error behaviour probes: {probes.tolist()}"
