"""§8f rank 2 / 3, Encryptor: the AES-128-CTR generator, the three samplers, asymmetric / symmetric zero encryption at every level, the BFV
scale-up, centralize and Encryptor::encrypt against the oracle restatement (oracle/rns_oracle_encrypt.inc, pinned to the reference in
tests/test_oracle_pinning.py) with the same (seed, counter); and, when oracle/_ref is present, live against the reference's own Encryptor
with the reference's keys, followed by a decryption on the GPU that returns the message.
Reference: src/utils/random_generator.cu:12-505, src/utils/rlwe.cu:11-322, src/encryptor.cu:12-338, src/fgk/translate_plain.cu:6-70."""
import numpy as np
import pytest

from conftest import make_oracle, to_dev, to_np

pytestmark = pytest.mark.gpu

CASES = [
    ("C1", {}),
    ("C2", {}),
    ("C4", {}),
    (None, dict(scheme=2, log_n=13, bits=[50, 40, 40, 55], t_bits=0)),
    (None, dict(scheme=1, log_n=12, bits=[60, 60, 60], t_bits=33)),
    (None, dict(scheme=3, log_n=11, bits=[60, 60, 60], t_bits=33)),
    (None, dict(scheme=2, log_n=16, bits=[60, 40, 40, 60], t_bits=0)),      # N = 2^16: the fused epilogues on the cluster / queue kernels
]


def _zeros(n):
    import torch
    return torch.zeros(n, dtype=torch.int64, device="cuda")


def test_generator_and_samplers(b2, po):
    import torch
    o = make_oracle(po, "C1")
    ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
    n, K = o.n, o.nm
    g = b2.Prng(0x0123456789abcdef, 0xfedcba9876543210, 5)
    st = o.prng_state(0x0123456789abcdef, 0xfedcba9876543210, 5)
    nbytes = 16 * 1000 + 7
    buf = torch.zeros(16 * 1001, dtype=torch.uint8, device="cuda")
    ctx.prng_fill(g, buf, nbytes)
    assert np.array_equal(buf.cpu().numpy()[:nbytes], o.prng_fill_bytes(st, nbytes)) and g.state() == tuple(int(v) for v in st)
    assert ctx.prng_sample_uint64(g) == int(po.lib().ro_prng_sample_uint64(st.ctypes.data_as(po.u64p))) and g.state() == tuple(int(v) for v in st)
    for kind, name in ((b2.SAMPLE_TERNARY, "ternary"), (b2.SAMPLE_CENTERED_BINOMIAL, "cbd"), (b2.SAMPLE_UNIFORM, "uniform")):
        for nl in (K, 1):
            count = 3
            d = _zeros(count * nl * n)
            ctx.sample_poly(kind, g, d, nl, count=count)
            want = np.concatenate([o.sample_poly(name, st, nl) for _ in range(count)])      # the batched form = consecutive polynomials of one stream
            assert np.array_equal(to_np(d), want), (name, nl)
            assert g.state() == tuple(int(v) for v in st), "counter after " + name


@pytest.mark.parametrize("name,kw", CASES)
def test_encrypt_matches_oracle(b2, po, name, kw):
    o = make_oracle(po, name, **kw)
    ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
    n, K, L = o.n, o.nm, o.nm - 1
    rng = np.random.default_rng(23)
    pk, sk = o.random_poly(rng, 2, K), o.random_poly(rng, 1, K)      # any NTT-form "keys": encryption is a deterministic function of them
    d_pk, d_sk = to_dev(pk), to_dev(sk)
    is_ntt = o.scheme != 1
    for nl in sorted({K, L, 1}, reverse=True):
        for form in ({is_ntt, True} if o.scheme == 1 else {is_ntt}):
            ct = _zeros(2 * nl * n)
            g = b2.Prng(1234, 5)
            before = b2.launch_count()
            ctx.encrypt_zero_asymmetric(d_pk, nl, form, ct, g)
            launches = b2.launch_count() - before
            st = o.prng_state(1234, 5)
            assert np.array_equal(to_np(ct), o.encrypt_zero_asymmetric(pk, nl, form, st)), ("asymmetric", nl, form)
            assert g.state() == tuple(int(v) for v in st)
            assert launches <= (4 if nl == K else 7), f"asymmetric zero encryption took {launches} launches"
            for save in (False, True):
                g = b2.Prng(99)
                seed = ctx.encrypt_zero_symmetric(d_sk, nl, form, ct, g, save_seed=save)
                st = o.prng_state(99)
                want, wseed = o.encrypt_zero_symmetric(sk, nl, form, save, st)
                assert np.array_equal(to_np(ct), want), ("symmetric", nl, form, save)
                assert seed == wseed and g.state() == tuple(int(v) for v in st)
    # plaintext lifts and Encryptor::encrypt
    if o.scheme == 2:
        for nl in (L, 1):
            plain = o.random_poly(rng, 1, nl)
            for asym in (True, False):
                ct = _zeros(2 * nl * n)
                ctx.encrypt(to_dev(plain), n, d_pk if asym else d_sk, asym, nl, ct, b2.Prng(4242))
                assert np.array_equal(to_np(ct), o.encrypt(plain, pk if asym else sk, asym, nl, o.prng_state(4242))), ("encrypt", nl, asym)
    else:
        for cnt in (n, n // 2 + 3, 1):
            plain = rng.integers(0, o.t, size=cnt, dtype=np.uint64)
            plain[0] = o.t - 1
            d_plain = to_dev(np.concatenate([plain, np.zeros(2, dtype=np.uint64)]))      # 16-byte multiple
            if o.scheme == 1:
                base = o.random_poly(rng, 1, L)
                for sub in (False, True):
                    d = _zeros(L * n)
                    ctx.bfv_scale_up(d_plain, cnt, L, d, frm=to_dev(base), subtract=sub)
                    assert np.array_equal(to_np(d), o.bfv_scale_up(plain, L, base, sub)), ("scale_up", cnt, sub)
                d = _zeros(L * n)
                ctx.bfv_scale_up(d_plain, cnt, L, d)
                assert np.array_equal(to_np(d), o.bfv_scale_up(plain, L))
            d = _zeros(L * n)
            ctx.centralize(d_plain, cnt, L, d)
            assert np.array_equal(to_np(d), o.centralize(plain, L)), ("centralize", cnt)
            for asym in (True, False):
                ct = _zeros(2 * L * n)
                ctx.encrypt(d_plain, cnt, d_pk if asym else d_sk, asym, L, ct, b2.Prng(4242))
                assert np.array_equal(to_np(ct), o.encrypt(plain, pk if asym else sk, asym, L, o.prng_state(4242))), ("encrypt", cnt, asym)


@pytest.mark.parametrize("scheme,bits,t_bits", [(1, [40, 40, 50, 55], 20), (3, [40, 45, 50, 55], 20), (2, [50, 40, 40, 55], 20)])
def test_encrypt_against_reference_and_round_trip(b2, po, scheme, bits, t_bits):
    """same seed, the reference's keys: the GPU Encryptor's ciphertext equals the reference Encryptor's word for word; decrypting it on
    the GPU returns the message (BFV / BGV), or message + small noise (CKKS)"""
    if not po.ref_available():
        pytest.skip("oracle/_ref not built")
    n = 2048
    R = po.Reference(scheme, n, bits, t_bits=t_bits)
    t = R.t if scheme != 2 else 0
    ctx = b2.Context(11, R.moduli, t, scheme)
    K, L = R.nm, R.nm - 1
    pk, sk = R.public_key(), R.secret_key()
    d_pk, d_sk = to_dev(pk), to_dev(sk)
    d_skp = _zeros(K * n)
    ctx.secret_key_powers(d_sk, 1, d_skp)
    rng = np.random.default_rng(5)
    if scheme == 2:
        o = po.Oracle(11, R.moduli, 0, scheme)
        msg = rng.integers(0, 1 << 30, size=n, dtype=np.uint64)
        plain = o.ntt_fwd(np.concatenate([msg % np.uint64(R.moduli[j]) for j in range(L)]), 1, L)
    else:
        plain = rng.integers(0, R.t, size=n, dtype=np.uint64)
    d_plain = to_dev(plain)
    for asym in (True, False):
        want, _ = R.encrypt_seeded(plain, asym, L, 777, 1)
        ct = _zeros(2 * L * n)
        ctx.encrypt(d_plain, n, d_pk if asym else d_sk, asym, L, ct, b2.Prng(777, 1))
        assert np.array_equal(to_np(ct), want), ("differs from the reference Encryptor", scheme, asym)
        out = _zeros(L * n if scheme == 2 else n)
        ctx.decrypt([ct], [out], 2, L, d_skp)
        if scheme != 2:
            assert np.array_equal(to_np(out), plain), "decrypt(encrypt(m)) != m"
        else:
            dec = o.ntt_inv(to_np(out), 1, L)[:n].astype(np.int64)
            q0 = int(R.moduli[0])
            err = (dec - msg.astype(np.int64) + q0 // 2) % q0 - q0 // 2
            assert np.abs(err).max() < 1 << 12, "CKKS round trip noise too large"
