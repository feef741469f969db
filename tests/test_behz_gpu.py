"""M2 (BFV BEHZ multiply) and N3 parity: libb2rns vs the CPU oracle (bit-exact), through the C ABI.
The oracle's BEHZ restatement is pinned against the unmodified reference in tests/test_oracle_pinning.py."""
import os

import numpy as np
import pytest

from conftest import make_oracle, to_dev, to_np

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ctx(b2, o):
    return b2.Context(o.log_n, o.moduli, o.t, o.scheme)


@pytest.mark.parametrize("log_n,bits,t_bits", [(5, [40, 40, 40], 20), (12, [60, 40, 40, 60], 20), (13, [50, 36, 45, 60, 55], 30), (7, [60] * 6, 59),
                                               (11, [45] * 10, 20)])
def test_bfv_multiply_all_levels(b2, po, log_n, bits, t_bits):
    o = make_oracle(po, scheme=1, log_n=log_n, bits=bits, t_bits=t_bits)
    ctx = _ctx(b2, o)
    rng = np.random.default_rng(log_n)
    for nl in range(o.nm, 0, -1):
        assert ctx.behz_bases(nl) == o.behz_bases(nl), "auxiliary BEHZ base differs from the reference's"
        a = o.random_poly(rng, 2, nl); b = o.random_poly(rng, 2, nl)
        out = to_dev(np.zeros(3 * nl * o.n, dtype=np.uint64))
        ctx.bfv_multiply(to_dev(a), to_dev(b), out, nl)
        assert np.array_equal(to_np(out), o.bfv_multiply(a, b, nl)), nl
        da = to_dev(a)
        ctx.bfv_multiply(da, da, out, nl)           # square == multiply(a, a)  (bfv_square, evaluator.cu:199-282)
        assert np.array_equal(to_np(out), o.bfv_multiply(a, a, nl)), ("square", nl)


def test_bfv_multiply_reference_golden(b2, po):
    """golden vectors produced by the unmodified reference (tests/golden/ref_bfv_n256.npz)"""
    g = np.load(os.path.join(GOLD, "ref_bfv_n256.npz"))
    n = int(g["n"])
    o = po.Oracle(8, g["moduli"], int(g["t"]), po.BFV)
    ctx = _ctx(b2, o)
    L = o.nm - 1
    for lvl in (L, L - 1):
        s = "_l%d" % lvl
        out = to_dev(np.zeros(3 * lvl * n, dtype=np.uint64))
        ctx.bfv_multiply(to_dev(g["a" + s]), to_dev(g["b" + s]), out, lvl)
        assert np.array_equal(to_np(out), g["mul" + s])
        ctx.bfv_multiply(to_dev(g["a" + s]), to_dev(g["a" + s]), out, lvl)
        assert np.array_equal(to_np(out), g["square" + s])


@pytest.mark.parametrize("name", ["C1", "C2", "C5"])
def test_bfv_multiply_baseline_configs_batched(b2, po, name):
    o = make_oracle(po, name)
    ctx = _ctx(b2, o)
    rng = np.random.default_rng(21)
    nl, B = o.nm - 1, 3
    a = np.concatenate([o.random_poly(rng, 2, nl) for _ in range(B)]); b = np.concatenate([o.random_poly(rng, 2, nl) for _ in range(B)])
    out = to_dev(np.zeros(B * 3 * nl * o.n, dtype=np.uint64))
    ctx.bfv_multiply(to_dev(a), to_dev(b), out, nl, batch=B)
    per = 2 * nl * o.n
    want = np.concatenate([o.bfv_multiply(a[i * per:(i + 1) * per], b[i * per:(i + 1) * per], nl) for i in range(B)])
    assert np.array_equal(to_np(out), want)


def test_bfv_multiply_errors(b2, po):
    o = make_oracle(po, scheme=2, log_n=12, bits=[40, 40, 40])
    ctx = _ctx(b2, o)
    d = to_dev(np.zeros(3 * 2 * o.n, dtype=np.uint64))
    with pytest.raises(b2.B2Error):
        ctx.bfv_multiply(d, d, d, 2)                # CKKS context


@pytest.mark.parametrize("log_n", [4, 13])
def test_scale_inv_degree(b2, po, log_n):
    o = make_oracle(po, scheme=2, log_n=log_n, bits=[60, 40, 50, 61])
    ctx = _ctx(b2, o)
    rng = np.random.default_rng(log_n)
    x = o.random_poly(rng, 2, o.nm)
    for scalar in (1, 3, (1 << 64) - 5):
        d = to_dev(x)
        ctx.scale_inv_degree(d, 2, o.nm, scalar)
        assert np.array_equal(to_np(d), o.ntt_multiply_inv_degree(x, 2, o.nm, scalar))
    y = o.random_poly(rng, 1, 1, base=2)
    d = to_dev(y); ctx.scale_inv_degree(d, 1, 1, 7, table_base=2)
    assert np.array_equal(to_np(d), o.ntt_multiply_inv_degree(y, 1, 1, 7, base=2))
