"""§8f rank 1, RLWE packing: b2_pack_rlwes / b2_pack_layer against the oracle's step-by-step restatement of the reference's device branch
(oracle/pyoracle.py pack_rlwes, pinned to Evaluator::pack_rlwe_ciphertexts_new by tests/golden/ref_pack_n64.npz), including the claim that the
tree can run in coefficient form for the NTT-form schemes and still store the reference's words.
Reference: src/evaluator_lwes.cu:96-105 (field trace), :305-491 (pack_rlwe_ciphertexts_new)."""
import numpy as np
import pytest

from conftest import make_oracle, to_dev, to_np

pytestmark = pytest.mark.gpu


def _zeros(n):
    import torch
    return torch.zeros(n, dtype=torch.int64, device="cuda")


def _keys(o, ctx, rng):
    n, logn = o.n, o.log_n
    elts = sorted({(n >> k) + 1 for k in range(0, logn)} | {2 ** k + 1 for k in range(1, logn + 1)})
    elts = [e for e in elts if e < 2 * n]
    host = {e: o.random_key(rng) for e in elts}
    return host, {e: ctx.kswitch_key(to_dev(host[e])) for e in elts}


@pytest.mark.parametrize("name,kw,cases", [
    (None, dict(scheme=1, log_n=10, bits=[60, 40, 40, 60], t_bits=20), [(8, 0, 1024, 128, True), (5, 3, 1024, 128, True), (4, 0, 512, 128, False), (1, 0, 1024, 1024, True)]),
    (None, dict(scheme=2, log_n=10, bits=[50, 40, 40, 55], t_bits=0), [(8, 0, 1024, 128, True), (3, 7, 1024, 256, True)]),
    (None, dict(scheme=3, log_n=10, bits=[50, 40, 40, 55], t_bits=20), [(4, 0, 1024, 256, True), (7, 0, 1024, 128, False)]),
    ("C5", {}, [(64, 0, 8192, 128, True)]),                         # BumbleBee output packing shape: 64 ciphertexts -> one, then the trace
])
def test_pack_rlwes_matches_oracle(b2, po, name, kw, cases):
    o = make_oracle(po, name, **kw)
    ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
    n, L = o.n, o.nm - 1
    rng = np.random.default_rng(31)
    hkeys, dkeys = _keys(o, ctx, rng)
    for (count, shift, ii, oi, ft) in cases:
        cts = [o.random_poly(rng, 2, L) for _ in range(count)]
        for inp_ntt in ((False, True) if o.scheme != 1 else (False,)):
            want = o.pack_rlwes(cts, inp_ntt, L, shift, ii, oi, hkeys, ft)
            d_cts = [to_dev(c) for c in cts]
            out = _zeros(2 * L * n)
            before = b2.launch_count()
            ctx.pack_rlwes(d_cts, inp_ntt, L, shift, ii, oi, dkeys, out, apply_field_trace=ft)
            launches = b2.launch_count() - before
            assert np.array_equal(to_np(out), want), (count, shift, ii, oi, ft, inp_ntt)
            for i in range(count):
                assert np.array_equal(to_np(d_cts[i]), cts[i]), "input modified"
            layers = (ii // oi).bit_length() - 1
            print("pack %d -> 1 (%d layers): %d launches" % (count, layers, launches))


@pytest.mark.parametrize("scheme", [1, 2, 3])
def test_pack_rlwes_groups_side_by_side(b2, po, scheme):
    """pack_rlwe_ciphertexts_batched: several groups (different fill levels) in one call == each group packed on its own"""
    o = make_oracle(po, scheme=scheme, log_n=10, bits=[50, 40, 40, 55], t_bits=20 if scheme != 2 else 0)
    ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
    n, L = o.n, o.nm - 1
    rng = np.random.default_rng(77)
    hkeys, dkeys = _keys(o, ctx, rng)
    counts, shift, ii, oi = [8, 3, 1, 6], 2 * n - 7, n, n // 8           # the matmul's inherent shift lies in [N, 2N)
    groups = [[o.random_poly(rng, 2, L) for _ in range(c)] for c in counts]
    inp_ntt = scheme != 1
    outs = [_zeros(2 * L * n) for _ in counts]
    before = b2.launch_count()
    ctx.pack_rlwes([[to_dev(c) for c in g] for g in groups], inp_ntt, L, shift, ii, oi, dkeys, outs)
    launches = b2.launch_count() - before
    for g, grp in enumerate(groups):
        assert np.array_equal(to_np(outs[g]), o.pack_rlwes(grp, inp_ntt, L, shift, ii, oi, hkeys, True)), "group %d" % g
    one = _zeros(2 * L * n)
    before = b2.launch_count()
    ctx.pack_rlwes([to_dev(c) for c in groups[0]], inp_ntt, L, shift, ii, oi, dkeys, one)
    assert launches <= (b2.launch_count() - before) + 8, "groups must share the launches of one group (plus the per-group result copies)"


@pytest.mark.parametrize("scheme", [1, 2, 3])
def test_lwe_extract_assemble_pack(b2, po, scheme):
    """Evaluator::extract_lwe_new, LWECiphertext::assemble_lwe_batched and pack_lwe_ciphertexts_new / _batched (src/evaluator_lwes.cu:51-94,201-303,
    src/lwe_ciphertext.cu:12-140) against the oracle compositions pinned to the reference (tests/test_oracle_pinning.py)"""
    o = make_oracle(po, scheme=scheme, log_n=10, bits=[50, 40, 40, 55], t_bits=20 if scheme != 2 else 0)
    ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
    n, L = o.n, o.nm - 1
    rng = np.random.default_rng(5)
    hkeys, dkeys = _keys(o, ctx, rng)
    ct = o.random_poly(rng, 2, L)
    d_ct = to_dev(ct)
    lwes = []
    for term in (0, 1, 333, n - 1):
        c0, c1 = _zeros(L + (L & 1)), _zeros(L * n)
        before = b2.launch_count()
        ctx.extract_lwe(d_ct, term, L, c0, c1)
        assert b2.launch_count() - before == 1
        w0, w1 = o.extract_lwe(ct, term, L)
        assert np.array_equal(to_np(c0)[:L], w0) and np.array_equal(to_np(c1), w1), term
        lwes.append((c0, c1, w0, w1))
    outs = [_zeros(2 * L * n) for _ in lwes]
    before = b2.launch_count()
    ctx.assemble_lwes([l[0] for l in lwes], [l[1] for l in lwes], outs, L)
    assert b2.launch_count() - before == 1
    for (c0, c1, w0, w1), out in zip(lwes, outs):
        assert np.array_equal(to_np(out), o.assemble_lwe(w0, w1, L))
    # two groups of LWEs packed side by side (sizes 4 and 3 -> both into N / 4 slots), with the trace
    more = [(o.random_poly(rng, 1, L).reshape(L, n)[:, 0].copy(), o.random_poly(rng, 1, L)) for _ in range(3)]
    g0 = [(l[2], l[3]) for l in lwes]
    d_groups = [[(l[0], l[1]) for l in lwes], [(to_dev(np.concatenate([c0, np.zeros(L & 1, dtype=np.uint64)])), to_dev(c1)) for c0, c1 in more]]
    pouts = [_zeros(2 * L * n), _zeros(2 * L * n)]
    ctx.pack_lwes(d_groups, L, dkeys, pouts)
    assert np.array_equal(to_np(pouts[0]), o.pack_lwes(g0, L, hkeys, True)), "group 0"
    assert np.array_equal(to_np(pouts[1]), o.pack_lwes(more + [(np.zeros(L, dtype=np.uint64), np.zeros(L * n, dtype=np.uint64))][:0], L, hkeys, True)), "group 1"


def test_pack_layer_and_errors(b2, po):
    o = make_oracle(po, scheme=1, log_n=11, bits=[60, 40, 60], t_bits=20)
    ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
    n, L = o.n, o.nm - 1
    rng = np.random.default_rng(2)
    hkeys, dkeys = _keys(o, ctx, rng)
    pairs, shift, g = 3, n // 4, 5
    slots = [o.random_poly(rng, 2, L) for _ in range(2 * pairs)]
    out, tgt = _zeros(pairs * 2 * L * n), _zeros(pairs * L * n)
    before = b2.launch_count()
    ctx.pack_layer(to_dev(np.concatenate(slots)), out, tgt, pairs, L, shift, g, dkeys[g])
    assert b2.launch_count() - before <= 1 + 6, "a layer is one fused element-wise launch + one batched keyswitch"
    got = to_np(out)
    for b in range(pairs):
        even, odd = slots[2 * b], slots[2 * b + 1]
        temp = o.negacyclic_shift(odd, shift, 2, L)
        want = o.add(o.add(even, temp, 2, L), o.apply_galois(o.sub(even, temp, 2, L), g, L, False, hkeys[g]), 2, L)
        assert np.array_equal(got[b * 2 * L * n:(b + 1) * 2 * L * n], want), b
    # a missing automorphism key is reported before anything runs (the reference throws from GaloisKeys::key)
    d_cts = [to_dev(s) for s in slots[:4]]
    with pytest.raises(Exception):
        ctx.pack_rlwes(d_cts, False, L, 0, n, n // 4, {5: dkeys[5]}, _zeros(2 * L * n))
    with pytest.raises(Exception):      # evaluator_lwes.cu:337-339
        ctx.pack_rlwes(d_cts, False, L, 0, n, n // 2, dkeys, _zeros(2 * L * n))
