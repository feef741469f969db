"""K1-K5, R1-R4, M1/M3/E1/G1 parity: libb2rns vs the CPU oracle (bit-exact), through the C ABI."""
import numpy as np
import pytest

from conftest import make_oracle, to_dev, to_np

pytestmark = pytest.mark.gpu

SCHEMES = {1: "bfv", 2: "ckks", 3: "bgv"}


def _ctx(b2, o):
    return b2.Context(o.log_n, o.moduli, o.t, o.scheme)


@pytest.mark.parametrize("scheme", [1, 2, 3])
@pytest.mark.parametrize("log_n,bits", [(5, [40, 40, 40]), (12, [60, 40, 40, 60]), (13, [50, 30, 45, 61, 55])])
def test_keyswitch_all_assign_modes(b2, po, scheme, log_n, bits):
    o = make_oracle(po, scheme=scheme, log_n=log_n, bits=bits, t_bits=20 if scheme != 2 else 0)
    ctx = _ctx(b2, o)
    rng = np.random.default_rng(100 * scheme + log_n)
    key = o.random_key(rng)
    dkey = to_dev(key)
    kh = ctx.kswitch_key(dkey)
    is_ntt = scheme != 1
    for decomp in sorted({o.nm - 1, max(1, o.nm - 2), 1}):
        target = o.random_poly(rng, 1, decomp)
        for assign in (b2.ASSIGN_OVERWRITE, b2.ASSIGN_ADD_INPLACE, b2.ASSIGN_OVERWRITE_EXCEPT_FIRST):
            dest0 = o.random_poly(rng, 2, decomp)
            want = o.keyswitch(target, decomp, is_ntt, key, dest=dest0, assign=assign)
            d = to_dev(dest0)
            ctx.keyswitch(to_dev(target), decomp, is_ntt, kh, d, assign)
            assert np.array_equal(to_np(d), want), f"{SCHEMES[scheme]} decomp={decomp} assign={assign}"
            # the host-buffer entry point (target and destination in host memory; the destination's old contents travel under the keyswitch)
            hd = dest0.copy()
            ctx.keyswitch_host(np.ascontiguousarray(target), decomp, is_ntt, kh, hd, assign)
            assert np.array_equal(hd, want), f"{SCHEMES[scheme]} host buffers decomp={decomp} assign={assign}"


def test_keyswitch_bfv_both_forms(b2, po):
    """the reference also accepts NTT-form BFV ciphertexts (is_ntt_form branch, keyswitching_core.cu:959-969)"""
    o = make_oracle(po, "C1")
    ctx = _ctx(b2, o)
    rng = np.random.default_rng(9)
    key = o.random_key(rng); kh = ctx.kswitch_key(to_dev(key))
    decomp = o.nm - 1
    target = o.random_poly(rng, 1, decomp)
    for is_ntt in (False, True):
        want = o.keyswitch(target, decomp, is_ntt, key)
        d = to_dev(np.zeros(2 * decomp * o.n, dtype=np.uint64))
        ctx.keyswitch(to_dev(target), decomp, is_ntt, kh, d, b2.ASSIGN_OVERWRITE)
        assert np.array_equal(to_np(d), want)


@pytest.mark.parametrize("name", ["C2", "C4"])
def test_keyswitch_baseline_configs(b2, po, name):
    o = make_oracle(po, name)
    ctx = _ctx(b2, o)
    rng = np.random.default_rng(11)
    key = o.random_key(rng); kh = ctx.kswitch_key(to_dev(key))
    decomp = o.nm - 1
    is_ntt = o.scheme != 1
    target = o.random_poly(rng, 1, decomp); dest0 = o.random_poly(rng, 2, decomp)
    want = o.keyswitch(target, decomp, is_ntt, key, dest=dest0, assign=po.ASSIGN_ADD_INPLACE)
    d = to_dev(dest0)
    ctx.keyswitch(to_dev(target), decomp, is_ntt, kh, d, b2.ASSIGN_ADD_INPLACE)
    assert np.array_equal(to_np(d), want)


def test_keyswitch_c3_full_size(b2, po):
    """BASELINE config C3: CKKS N=2^16, 20 limbs + special prime.  Bit-exact against the oracle."""
    o = make_oracle(po, "C3")
    ctx = _ctx(b2, o)
    rng = np.random.default_rng(12)
    key = o.random_key(rng); kh = ctx.kswitch_key(to_dev(key))
    decomp = o.nm - 1
    target = o.random_poly(rng, 1, decomp); dest0 = o.random_poly(rng, 2, decomp)
    want = o.keyswitch(target, decomp, True, key, dest=dest0, assign=po.ASSIGN_OVERWRITE_EXCEPT_FIRST)
    d = to_dev(dest0)
    ctx.keyswitch(to_dev(target), decomp, True, kh, d, b2.ASSIGN_OVERWRITE_EXCEPT_FIRST)
    assert np.array_equal(to_np(d), want)


def test_keyswitch_semantics_real_key(b2, po):
    """decrypt-level check with a structurally real key (oracle toy RLWE): switching c1 from s' to s
    must give phase(out) = c1*s' + small noise."""
    o = make_oracle(po, scheme=2, log_n=12, bits=[50, 40, 40, 50])
    ctx = _ctx(b2, o)
    sk = o.keygen_secret(1); sk2 = o.keygen_secret(2)
    key = o.keygen_kswitch(sk, sk2, 3)
    kh = ctx.kswitch_key(to_dev(key))
    decomp = o.nm - 1
    rng = np.random.default_rng(5)
    c1 = o.random_poly(rng, 1, decomp)               # NTT form
    d = to_dev(np.zeros(2 * decomp * o.n, dtype=np.uint64))
    ctx.keyswitch(to_dev(c1), decomp, True, kh, d, b2.ASSIGN_OVERWRITE)
    out = to_np(d)
    ph = o.phase(sk, out, 2, decomp, True)            # out0 + out1*s  (coefficient form)
    # expected: c1 * s'  (NTT-domain product, back to coefficients)
    exp = o.ntt_inv(o.dyadic_broadcast(c1, sk2[: decomp * o.n], 1, decomp), 1, decomp)
    q0 = int(o.moduli[0])
    diff = (ph[: o.n].astype(object) - exp[: o.n].astype(object)) % q0
    diff = np.array([int(v) if v < q0 // 2 else int(v) - q0 for v in diff])
    assert np.abs(diff).max() < 2 ** 20, "keyswitch noise too large: %d" % np.abs(diff).max()


@pytest.mark.parametrize("log_n,bits", [(4, [40, 40, 40]), (12, [60, 40, 40, 60]), (14, [60, 40, 45, 55, 60])])
def test_mod_switch_to_next_kinds(b2, po, log_n, bits):
    rng = np.random.default_rng(log_n)
    for scheme, kind, fn in ((1, 0, "divide_round_q_last"), (2, 1, "divide_round_q_last_ntt"), (3, 2, "mod_t_divide_q_last_ntt"), (2, 3, "drop_last_limb")):
        o = make_oracle(po, scheme=scheme, log_n=log_n, bits=bits, t_bits=20 if scheme != 2 else 0)
        ctx = _ctx(b2, o)
        for nl in range(o.nm, 1, -1):
            for count in (1, 3):
                x = o.random_poly(rng, count, nl)
                want = getattr(o, fn)(x, count, nl)
                out = to_dev(np.zeros(count * (nl - 1) * o.n, dtype=np.uint64))
                ctx.mod_switch_to_next(to_dev(x), out, count, nl, kind)
                assert np.array_equal(to_np(out), want), (fn, nl, count)


def test_bgv_modswitch_chain_c4(b2, po):
    """BASELINE config C4: BGV N=32768, 8 limbs: mod_switch chain to the last level."""
    o = make_oracle(po, "C4")
    ctx = _ctx(b2, o)
    rng = np.random.default_rng(44)
    nl = o.nm - 1
    x = o.random_poly(rng, 2, nl)
    d = to_dev(x)
    while nl > 1:
        want = o.mod_t_divide_q_last_ntt(x, 2, nl)
        out = to_dev(np.zeros(2 * (nl - 1) * o.n, dtype=np.uint64))
        ctx.mod_switch_to_next(d, out, 2, nl, b2.MODSWITCH_BGV)
        assert np.array_equal(to_np(out), want), nl
        x, d, nl = want, out, nl - 1


@pytest.mark.parametrize("log_n", [3, 12, 15])
def test_elementwise(b2, po, log_n):
    o = make_oracle(po, scheme=2, log_n=log_n, bits=[60, 40, 50, 61])
    ctx = _ctx(b2, o)
    rng = np.random.default_rng(log_n)
    nc = o.nm
    a = o.random_poly(rng, 2, nc); b = o.random_poly(rng, 2, nc); c3 = o.random_poly(rng, 3, nc)
    da, db = to_dev(a), to_dev(b)
    out = to_dev(np.zeros_like(a))
    assert np.array_equal(to_np(ctx.add(da, db, out, 2, nc)), o.add(a, b, 2, nc))
    assert np.array_equal(to_np(ctx.sub(da, db, out, 2, nc)), o.sub(a, b, 2, nc))
    az = a.copy(); az[::7] = 0
    assert np.array_equal(to_np(ctx.negate(to_dev(az), out, 2, nc)), o.negate(az, 2, nc))
    o3 = to_dev(np.zeros(3 * nc * o.n, dtype=np.uint64))
    assert np.array_equal(to_np(ctx.dyadic_convolute(da, 2, db, 2, o3, nc)), o.dyadic_convolute(a, 2, b, 2, nc))
    assert np.array_equal(to_np(ctx.dyadic_square(da, o3, nc)), o.dyadic_square(a, nc))
    o4 = to_dev(np.zeros(4 * nc * o.n, dtype=np.uint64))
    assert np.array_equal(to_np(ctx.dyadic_convolute(to_dev(c3), 3, db, 2, o4, nc)), o.dyadic_convolute(c3, 3, b, 2, nc))
    pl = o.random_poly(rng, 1, nc)
    assert np.array_equal(to_np(ctx.dyadic_broadcast(da, to_dev(pl), out, 2, nc)), o.dyadic_broadcast(a, pl, 2, nc))
    acc = o.random_poly(rng, 2, nc); dacc = to_dev(acc)
    assert np.array_equal(to_np(ctx.dyadic_broadcast(da, to_dev(pl), dacc, 2, nc, accumulate=True)), o.dyadic_broadcast(a, pl, 2, nc, acc_into=acc))
    # batched multiply_plain (one plaintext per ciphertext) == per-item, with and without accumulation
    Bp = 3
    ap = np.concatenate([o.random_poly(rng, 2, nc) for _ in range(Bp)]); pp = np.concatenate([o.random_poly(rng, 1, nc) for _ in range(Bp)])
    accp = np.concatenate([o.random_poly(rng, 2, nc) for _ in range(Bp)])
    per_ct, per_pt = 2 * nc * o.n, nc * o.n
    outp = to_dev(np.zeros_like(ap)); daccp = to_dev(accp)
    ctx.dyadic_broadcast_batched(to_dev(ap), to_dev(pp), outp, Bp, 2, nc)
    ctx.dyadic_broadcast_batched(to_dev(ap), to_dev(pp), daccp, Bp, 2, nc, accumulate=True)
    for i in range(Bp):
        ai, pi = ap[i * per_ct:(i + 1) * per_ct], pp[i * per_pt:(i + 1) * per_pt]
        assert np.array_equal(to_np(outp)[i * per_ct:(i + 1) * per_ct], o.dyadic_broadcast(ai, pi, 2, nc))
        assert np.array_equal(to_np(daccp)[i * per_ct:(i + 1) * per_ct], o.dyadic_broadcast(ai, pi, 2, nc, acc_into=accp[i * per_ct:(i + 1) * per_ct]))
    # batched convolution == per-item
    B = 3
    ab = np.concatenate([o.random_poly(rng, 2, nc) for _ in range(B)]); bb = np.concatenate([o.random_poly(rng, 2, nc) for _ in range(B)])
    ob = to_dev(np.zeros(B * 3 * nc * o.n, dtype=np.uint64))
    ctx.dyadic_convolute(to_dev(ab), 2, to_dev(bb), 2, ob, nc, batch=B)
    per = 2 * nc * o.n
    want = np.concatenate([o.dyadic_convolute(ab[i * per:(i + 1) * per], 2, bb[i * per:(i + 1) * per], 2, nc) for i in range(B)])
    assert np.array_equal(to_np(ob), want)
    for shift in (0, 1, o.n - 1, o.n, o.n + 5 if o.n > 8 else o.n + 1):
        assert np.array_equal(to_np(ctx.negacyclic_shift(to_dev(az), shift, out, 2, nc)), o.negacyclic_shift(az, shift, 2, nc)), shift
    for g in (3, 2 * o.n - 1, o.galois_element_from_step(1), (5 * 3) % (2 * o.n) | 1):
        assert np.array_equal(to_np(ctx.galois(to_dev(az), g, out, 2, nc, False)), o.galois(az, g, 2, nc, False)), ("coeff", g)
        assert np.array_equal(to_np(ctx.galois(da, g, out, 2, nc, True)), o.galois(a, g, 2, nc, True)), ("ntt", g)


def test_galois_commutes_with_ntt(b2, po):
    """size-independent property at C3 size: apply_ntt(NTT(a)) == NTT(apply(a))"""
    o = make_oracle(po, "C3")
    ctx = _ctx(b2, o)
    rng = np.random.default_rng(6)
    nc = 3
    a = o.random_poly(rng, 1, nc)
    g = o.galois_element_from_step(7)
    da = to_dev(a); t1 = to_dev(np.zeros_like(a)); t2 = to_dev(np.zeros_like(a))
    ctx.galois(da, g, t1, 1, nc, False); ctx.ntt_forward(t1, pcount=1, ncomp=nc)
    ctx.ntt_forward(da, pcount=1, ncomp=nc); ctx.galois(da, g, t2, 1, nc, True)
    assert np.array_equal(to_np(t1), to_np(t2))
