"""BASELINE config 1 — "BFV N=8192 coeff {40,40,40} plain 20-bit: encode / encrypt / multiply / relinearize / decrypt (quickstart)" — run natively end to end
on the GPU from a seed: BatchEncoder, KeyGenerator, Encryptor, BEHZ multiply, relinearize, Decryptor, decode.  Every intermediate equals the oracle's
(which is pinned to the reference), the slots of the result are the slot-wise products of the inputs, and — when oracle/_ref is present — the ciphertexts
equal the reference's own for the same seeds.  Also the BGV variant (NTT-form ciphertexts, mod-t decryption).
Reference: src/batch_encoder.cu, src/key_generator.cu, src/encryptor.cu, src/evaluator.cu:29-116, src/evaluator_keyswitching.cu, src/decryptor.cu."""
import numpy as np
import pytest

from conftest import make_oracle, to_dev, to_np

pytestmark = pytest.mark.gpu


def _zeros(n):
    import torch
    return torch.zeros(n, dtype=torch.int64, device="cuda")


def test_batch_encoder_matches_oracle(b2, po):
    for name, kw in (("C1", {}), (None, dict(scheme=3, log_n=10, bits=[40, 40, 50], t_bits=20)), ("C2", {})):
        o = make_oracle(po, name, **kw)
        ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
        n = o.n
        rng = np.random.default_rng(n)
        for cnt in (n, n // 2 + 1, 1):
            v = rng.integers(0, o.t, size=cnt, dtype=np.uint64)
            plain, back = _zeros(n), _zeros(n)
            before = b2.launch_count()
            ctx.batch_encode(to_dev(np.concatenate([v, np.zeros(cnt & 1, dtype=np.uint64)])), cnt, plain)
            assert b2.launch_count() - before == 2
            assert np.array_equal(to_np(plain), o.batch_encode(v)), (name, cnt)
            ctx.batch_decode(plain, n, back)
            got = to_np(back)
            assert np.array_equal(got[:cnt], v) and not got[cnt:].any()
        short = rng.integers(0, o.t, size=n // 3 + (n // 3 & 1), dtype=np.uint64)
        back = _zeros(n)
        ctx.batch_decode(to_dev(short), short.size, back)
        assert np.array_equal(to_np(back), o.batch_decode(short))
    ck = make_oracle(po, scheme=2, log_n=12, bits=[40, 40, 50], t_bits=0)
    cctx = b2.Context(ck.log_n, ck.moduli, 0, ck.scheme)
    with pytest.raises(Exception):                       # batch_encoder.cu:191-193: no batching without a plain modulus = 1 mod 2N
        cctx.batch_encode(_zeros(8), 8, _zeros(ck.n))


@pytest.mark.parametrize("scheme", [1, 3])
def test_quickstart_from_a_seed(b2, po, scheme):
    o = make_oracle(po, "C1", scheme=scheme)
    ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
    n, K, L, t = o.n, o.nm, o.nm - 1, o.t
    is_ntt = scheme != 1
    rng = np.random.default_rng(2026)
    x = rng.integers(0, t, size=n, dtype=np.uint64); y = rng.integers(0, t, size=n, dtype=np.uint64)
    # keys from a seed
    g = b2.Prng(0x123, 0); st = o.prng_state(0x123)
    sk, pk, skp = _zeros(K * n), _zeros(2 * K * n), _zeros(2 * K * n)
    ctx.generate_secret_key(g, sk); ctx.encrypt_zero_symmetric(sk, K, True, pk, g); ctx.secret_key_powers(sk, 2, skp)
    rk = _zeros(L * 2 * K * n)
    ctx.generate_kswitch_key(sk, skp[K * n:], g, rk)
    hsk = o.keygen_secret_prng(st); hpk, _ = o.encrypt_zero_symmetric(hsk, K, True, False, st)
    hskp = o.secret_key_powers(hsk, 2); hrk = o.keygen_kswitch_prng(hsk, hskp[K * n:], st)
    assert np.array_equal(to_np(rk), hrk)
    # encode, encrypt (the generator continues)
    px, py, cx, cy = _zeros(n), _zeros(n), _zeros(2 * L * n), _zeros(2 * L * n)
    ctx.batch_encode(to_dev(x), n, px); ctx.batch_encode(to_dev(y), n, py)
    ctx.encrypt(px, n, pk, True, L, cx, g); ctx.encrypt(py, n, pk, True, L, cy, g)
    hcx = o.encrypt(o.batch_encode(x), hpk, True, L, st); hcy = o.encrypt(o.batch_encode(y), hpk, True, L, st)
    assert np.array_equal(to_np(cx), hcx) and np.array_equal(to_np(cy), hcy), "ciphertexts"
    # multiply, relinearize
    prod = _zeros(3 * L * n)
    if scheme == 1:
        ctx.bfv_multiply(cx, cy, prod, L); hprod = o.bfv_multiply(hcx, hcy, L)
    else:
        ctx.dyadic_convolute(cx, 2, cy, 2, prod, L); hprod = o.dyadic_convolute(hcx, 2, hcy, 2, L)
    assert np.array_equal(to_np(prod), hprod), "product"
    kh = ctx.kswitch_key(rk.view(L, 2, K, n))
    rel = prod[:2 * L * n].clone()
    ctx.keyswitch(prod[2 * L * n:], L, is_ntt, kh, rel, b2.ASSIGN_ADD_INPLACE)
    hrel = o.keyswitch(hprod[2 * L * n:], L, is_ntt, hrk, dest=hprod[:2 * L * n], assign=po.ASSIGN_ADD_INPLACE)
    assert np.array_equal(to_np(rel), hrel), "relinearize"
    # decrypt, decode: the slots of the result are the slot-wise products
    dec, out = _zeros(n), _zeros(n)
    ctx.decrypt([rel], [dec], 2, L, skp)
    assert np.array_equal(to_np(dec), o.decrypt(hrel, 2, L, hskp)), "decrypt"
    ctx.batch_decode(dec, n, out)
    want = np.array([(int(a) * int(b)) % t for a, b in zip(x, y)], dtype=np.uint64)
    assert np.array_equal(to_np(out), want), "decode(decrypt(relinearize(enc(x) * enc(y)))) != x * y slot-wise"
