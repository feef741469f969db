"""§8f rank 2, Decryptor: b2_decrypt / b2_decrypt_phase / b2_decrypt_scale_and_round / b2_decrypt_mod_t / b2_secret_key_powers against the
oracle restatement (oracle/rns_oracle_decrypt.inc, pinned to the reference's Decryptor in tests/test_oracle_pinning.py) and, when
oracle/_ref is present, live against the reference's own Decryptor with the reference's keys and a real encryption.
Reference: src/decryptor.cu:27-535, src/utils/rns_tool.cu:1118-1430, src/utils/rns_base.cu:445-583, src/key_generator.cu:84-110."""
import numpy as np
import pytest

from conftest import make_oracle, to_dev, to_np

pytestmark = pytest.mark.gpu

CASES = [
    ("C1", {}),                                                             # BFV 8192, quickstart
    ("C2", {}),                                                             # BFV 16384, 60/40-bit mix
    ("C5", {}),                                                             # BFV 8192 (matmul)
    ("C4", {}),                                                             # BGV 32768, 8 limbs
    (None, dict(scheme=3, log_n=12, bits=[60, 60, 60], t_bits=33)),         # BGV, all 60-bit, wide t
    (None, dict(scheme=1, log_n=12, bits=[60, 60, 60], t_bits=33)),         # BFV, all 60-bit, wide t
    (None, dict(scheme=2, log_n=13, bits=[50, 40, 40, 55], t_bits=0)),      # CKKS
    (None, dict(scheme=1, log_n=16, bits=[60, 40, 60], t_bits=20)),         # BFV at N = 2^16 (cluster / queue kernels with the fused functors)
]


def _zeros(n):
    import torch
    return torch.zeros(n, dtype=torch.int64, device="cuda")


@pytest.mark.parametrize("name,kw", CASES)
def test_decrypt_matches_oracle(b2, po, name, kw):
    o = make_oracle(po, name, **kw)
    ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
    n, K, L = o.n, o.nm, o.nm - 1
    rng = np.random.default_rng(17)
    sk = o.random_poly(rng, 1, K)                        # any NTT-form "key": decryption is a deterministic function of it
    d_skp = _zeros(2 * K * n)
    ctx.secret_key_powers(to_dev(sk), 2, d_skp)
    skp = o.secret_key_powers(sk, 2)
    assert np.array_equal(to_np(d_skp), skp), "secret key powers"
    is_ntt = o.scheme != 1
    out_words = (lambda nl: nl * n) if o.scheme == 2 else (lambda nl: n)
    for nl in sorted({L, max(1, L - 1), 1}, reverse=True):
        for pc, batch in ((2, 1), (2, 5), (3, 3)):
            cts = [o.random_poly(rng, pc, nl) for _ in range(batch)]
            cfs = [1 if (o.scheme != 3 or b % 2 == 0) else 3 + 2 * b for b in range(batch)]
            d_cts = [to_dev(c) for c in cts]
            # the phase alone
            d_ph = [_zeros(nl * n) for _ in range(batch)]
            before = b2.launch_count()
            ctx.decrypt_phase(d_cts, d_ph, pc, nl, is_ntt, d_skp)
            assert b2.launch_count() - before == (1 if is_ntt else 2), "phase launches"
            for b in range(batch):
                assert np.array_equal(to_np(d_ph[b]), o.decrypt_phase(cts[b], pc, nl, is_ntt, skp)), f"phase nl={nl} pc={pc} item {b}"
                assert np.array_equal(to_np(d_cts[b]), cts[b]), "ciphertext modified"
            # the whole decryption, one call for the batch
            d_out = [_zeros(out_words(nl)) for _ in range(batch)]
            before = b2.launch_count()
            ctx.decrypt(d_cts, d_out, pc, nl, d_skp, cfs if o.scheme == 3 else None)
            assert b2.launch_count() - before == {1: 3, 2: 1, 3: 2}[o.scheme], "decrypt launches (any batch size)"
            for b in range(batch):
                assert np.array_equal(to_np(d_out[b]), o.decrypt(cts[b], pc, nl, skp, cfs[b])), f"decrypt nl={nl} pc={pc} item {b}"
            # the scaling steps on their own (coefficient-form phases)
            if o.scheme == 1:
                d_pl = [_zeros(n) for _ in range(batch)]
                ctx.decrypt_scale_and_round(d_ph, d_pl, nl)
                for b in range(batch):
                    assert np.array_equal(to_np(d_pl[b]), o.decrypt_scale_and_round(to_np(d_ph[b]), nl))
            if o.scheme == 3:
                phs = [o.random_poly(rng, 1, nl) for _ in range(batch)]
                d_pl = [_zeros(n) for _ in range(batch)]
                ctx.decrypt_mod_t([to_dev(p) for p in phs], d_pl, nl, cfs)
                for b in range(batch):
                    assert np.array_equal(to_np(d_pl[b]), o.decrypt_mod_t(phs[b], nl, cfs[b]))


def test_decrypt_error_behaviour(b2, po):
    o = make_oracle(po, "C4")
    ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
    n, K = o.n, o.nm
    d_skp = _zeros(K * n)
    ct, out = _zeros(2 * 2 * n), _zeros(n)
    with pytest.raises(Exception):            # scaling_variant.cu:426-428: correction factor not invertible (t is prime: a multiple of t)
        ctx.decrypt([ct], [out], 2, 2, d_skp, [o.t])
    with pytest.raises(Exception):            # decryptor.cu:237: fewer than two polynomials
        ctx.decrypt([ct], [out], 1, 2, d_skp)
    ck = make_oracle(po, scheme=2, log_n=12, bits=[40, 40, 50], t_bits=0)
    cctx = b2.Context(ck.log_n, ck.moduli, 0, ck.scheme)
    with pytest.raises(Exception):            # no plain modulus: the BFV scaling is undefined
        cctx.decrypt_scale_and_round([_zeros(2 * ck.n)], [_zeros(ck.n)], 2)


@pytest.mark.parametrize("scheme,bits,t_bits", [(1, [40, 40, 50, 55], 20), (3, [40, 45, 50, 55], 20), (2, [50, 40, 40, 55], 0)])
def test_decrypt_against_reference_with_its_keys(b2, po, scheme, bits, t_bits):
    """the reference's own secret key and a real encryption of zero from the reference's Encryptor: the GPU decryption equals the
    reference's Decryptor bit for bit, and (BFV / BGV) the plaintext is zero"""
    if not po.ref_available():
        pytest.skip("oracle/_ref not built")
    n = 4096
    R = po.Reference(scheme, n, bits, t_bits=t_bits or 20)
    t = R.t if scheme != 2 else 0
    ctx = b2.Context(12, R.moduli, t, scheme)
    K, L = R.nm, R.nm - 1
    d_skp = _zeros(2 * K * n)
    ctx.secret_key_powers(to_dev(R.secret_key()), 2, d_skp)
    ct = R.encrypt_zero()
    words = L * n if scheme == 2 else n
    d_out = _zeros(words)
    ctx.decrypt([to_dev(ct)], [d_out], 2, L, d_skp)
    got = to_np(d_out)
    assert np.array_equal(got, R.decrypt_raw(ct, 2, L)), "GPU decryption differs from the reference's Decryptor"
    if scheme != 2:
        assert not got.any(), "an encryption of zero must decrypt to zero"
    # a product ciphertext (three polynomials), decrypted without relinearisation
    c3 = R.multiply(ct, ct, L)
    d_out3 = _zeros(words)
    ctx.decrypt([to_dev(c3)], [d_out3], 3, L, d_skp)
    if scheme != 2:      # the reference's CKKS decrypt checks the scale of a product against the modulus; the raw shim keeps the input scale
        assert np.array_equal(to_np(d_out3), R.decrypt_raw(c3, 3, L))
