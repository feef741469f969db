"""KeyGenerator on the GPU (b2_generate_secret_key, b2_generate_kswitch_key, public key = symmetric encryption of zero at the key level) and the
whole pipeline FROM A SEED: with the generator in the state of the reference's context generator, the natively generated secret / public /
relinearisation / Galois keys, a native encryption, multiply + relinearize + rotate and the native decryption reproduce the unmodified
reference word for word.  Reference: src/key_generator.cu:31-78,136-256, src/utils/rlwe.cu, src/encryptor.cu, src/decryptor.cu."""
import numpy as np
import pytest

from conftest import make_oracle, to_dev, to_np

pytestmark = pytest.mark.gpu


def _zeros(n):
    import torch
    return torch.zeros(n, dtype=torch.int64, device="cuda")


@pytest.mark.parametrize("name,kw", [("C1", {}), ("C4", {}), (None, dict(scheme=2, log_n=12, bits=[50, 40, 40, 55], t_bits=0)), (None, dict(scheme=2, log_n=16, bits=[60, 40, 60], t_bits=0))])
def test_keygen_matches_oracle(b2, po, name, kw):
    o = make_oracle(po, name, **kw)
    ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
    n, K = o.n, o.nm
    g = b2.Prng(0x5eed, 0)
    st = o.prng_state(0x5eed)
    sk = _zeros(K * n)
    ctx.generate_secret_key(g, sk)
    hsk = o.keygen_secret_prng(st)
    assert np.array_equal(to_np(sk), hsk) and g.state() == tuple(int(v) for v in st), "secret key"
    pk = _zeros(2 * K * n)
    ctx.encrypt_zero_symmetric(sk, K, True, pk, g)
    hpk, _ = o.encrypt_zero_symmetric(hsk, K, True, False, st)
    assert np.array_equal(to_np(pk), hpk), "public key"
    skp = _zeros(2 * K * n)
    ctx.secret_key_powers(sk, 2, skp)
    rk = _zeros((K - 1) * 2 * K * n)
    ctx.generate_kswitch_key(sk, skp[K * n:], g, rk)
    assert np.array_equal(to_np(rk), o.keygen_kswitch_prng(hsk, o.secret_key_powers(hsk, 2)[K * n:], st)), "relinearisation key"
    rot = _zeros(K * n)
    ctx.galois(sk, 3, rot, 1, K, True)
    gk = _zeros((K - 1) * 2 * K * n)
    ctx.generate_kswitch_key(sk, rot, g, gk)
    assert np.array_equal(to_np(gk), o.keygen_kswitch_prng(hsk, o.galois(hsk, 3, 1, K, True), st)), "Galois key"
    assert g.state() == tuple(int(v) for v in st), "generator state after key generation"


@pytest.mark.parametrize("scheme,bits", [(1, [40, 40, 50, 55]), (2, [50, 40, 40, 55]), (3, [40, 45, 50, 55])])
def test_pipeline_from_a_seed_equals_the_reference(b2, po, scheme, bits):
    if not po.ref_available():
        pytest.skip("oracle/_ref not built")
    n, logn, S = 2048, 11, 0xabcdef12345
    R = po.Reference(scheme, n, bits, t_bits=20, seed=S)
    R.gen_relin(); R.gen_galois([3])
    t = R.t if scheme != 2 else 0
    ctx = b2.Context(logn, R.moduli, t, scheme)
    o = po.Oracle(logn, R.moduli, t, scheme)
    K, L = R.nm, R.nm - 1
    # ---- keys, in the order the reference drew them: secret key, public key, relinearisation key, Galois key
    g = b2.Prng(S, 0)
    sk, pk = _zeros(K * n), _zeros(2 * K * n)
    ctx.generate_secret_key(g, sk)
    ctx.encrypt_zero_symmetric(sk, K, True, pk, g)
    skp = _zeros(2 * K * n)
    ctx.secret_key_powers(sk, 2, skp)
    rk, gk, rot = _zeros(L * 2 * K * n), _zeros(L * 2 * K * n), _zeros(K * n)
    ctx.generate_kswitch_key(sk, skp[K * n:], g, rk)
    ctx.galois(sk, 3, rot, 1, K, True)
    ctx.generate_kswitch_key(sk, rot, g, gk)
    assert np.array_equal(to_np(sk), R.secret_key()) and np.array_equal(to_np(pk), R.public_key()), "secret / public key differ from the reference's"
    assert np.array_equal(to_np(rk), R.relin_key()) and np.array_equal(to_np(gk), R.galois_key(3)), "relinearisation / Galois key differ from the reference's"
    # ---- encrypt (a fresh explicit seed on both sides), multiply, relinearize, rotate, decrypt
    rng = np.random.default_rng(9)
    if scheme == 2:
        plain = o.ntt_fwd(np.concatenate([rng.integers(0, 1 << 20, size=n, dtype=np.uint64) % np.uint64(R.moduli[j]) for j in range(L)]), 1, L)
    else:
        plain = rng.integers(0, R.t, size=n, dtype=np.uint64)
    want_ct, _ = R.encrypt_seeded(plain, True, L, 4242)
    ct = _zeros(2 * L * n)
    ctx.encrypt(to_dev(plain), n, pk, True, L, ct, b2.Prng(4242))
    assert np.array_equal(to_np(ct), want_ct), "ciphertext differs from the reference's"
    is_ntt = scheme != 1
    prod = _zeros(3 * L * n)
    if scheme == 1:
        ctx.bfv_multiply(ct, ct, prod, L)
    else:
        ctx.dyadic_convolute(ct, 2, ct, 2, prod, L)
    want_prod = R.multiply(want_ct, want_ct, L)
    assert np.array_equal(to_np(prod), want_prod), "product"
    kh = ctx.kswitch_key(rk.view(L, 2, K, n))
    rel = prod[:2 * L * n].clone()
    ctx.keyswitch(prod[2 * L * n:], L, is_ntt, kh, rel, b2.ASSIGN_ADD_INPLACE)
    want_rel = R.relinearize(want_prod, L)
    assert np.array_equal(to_np(rel), want_rel), "relinearize"
    gh = ctx.kswitch_key(gk.view(L, 2, K, n))
    perm, rotd = _zeros(2 * L * n), None
    ctx.galois(rel, 3, perm, 2, L, is_ntt)
    ctx.keyswitch(perm[L * n:], L, is_ntt, gh, perm, b2.ASSIGN_OVERWRITE_EXCEPT_FIRST)
    assert np.array_equal(to_np(perm), R.rotate(want_rel, 1, L)), "rotate"
    out = _zeros(L * n if scheme == 2 else n)
    ctx.decrypt([perm], [out], 2, L, skp)
    assert np.array_equal(to_np(out), R.decrypt_raw(to_np(perm), 2, L)), "decrypt"
