"""N1/N2 parity: libb2rns NTT / inverse NTT vs the CPU oracle (bit-exact), through the C ABI."""
import numpy as np
import pytest

from conftest import make_oracle, to_dev, to_np

pytestmark = pytest.mark.gpu


def _ctx(b2, o):
    return b2.Context(o.log_n, o.moduli, o.t, o.scheme)


@pytest.mark.parametrize("log_n", [1, 2, 5, 8, 10, 11, 12, 13, 14, 15, 16, 17])
def test_ntt_roundtrip_and_parity_componentwise(b2, po, log_n):
    n = 1 << log_n
    bits = [60, 40, 50, 61 if log_n <= 16 else 60]
    o = make_oracle(po, scheme=2, log_n=log_n, bits=bits)
    ctx = _ctx(b2, o)
    for i in range(o.nm):
        assert ctx.root(i) == o.root(i)
    rng = np.random.default_rng(log_n)
    pc, nc = 2, o.nm
    x = o.random_poly(rng, pc, nc)
    d = to_dev(x)
    ctx.ntt_forward(d, pcount=pc, ncomp=nc)
    got = to_np(d)
    assert np.array_equal(got, o.ntt_fwd(x, pc, nc)), "forward NTT differs from oracle"
    ctx.ntt_inverse(d, pcount=pc, ncomp=nc)
    assert np.array_equal(to_np(d), x), "inverse(forward(x)) != x"
    # inverse on arbitrary canonical input, out of place
    y = o.random_poly(rng, pc, nc)
    dy, dz = to_dev(y), to_dev(np.zeros_like(y))
    ctx.ntt_inverse(dy, dz, pcount=pc, ncomp=nc)
    assert np.array_equal(to_np(dz), o.ntt_inv(y, pc, nc)), "inverse NTT differs from oracle"
    assert np.array_equal(to_np(dy), y), "out-of-place transform modified its input"


@pytest.mark.parametrize("log_n", [12, 13, 16])
def test_ntt_lazy_inputs_below_4p(b2, po, log_n):
    """the reference's forward transform accepts inputs in [0, 4p) (src/fgk/ntt_grouped.cu:31-36)"""
    o = make_oracle(po, scheme=2, log_n=log_n, bits=[60, 40, 61])
    ctx = _ctx(b2, o)
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.integers(0, 4 * int(p), size=o.n, dtype=np.uint64) for p in o.moduli])
    d = to_dev(x)
    ctx.ntt_forward(d, pcount=1, ncomp=o.nm)
    assert np.array_equal(to_np(d), o.ntt_fwd(x, 1, o.nm))


@pytest.mark.parametrize("log_n", [5, 12, 14])
def test_ntt_keyswitch_table_modes(b2, po, log_n):
    o = make_oracle(po, scheme=2, log_n=log_n, bits=[50, 40, 40, 40, 50])
    ctx = _ctx(b2, o)
    rng = np.random.default_rng(2)
    K = o.nm
    for decomp in (K - 1, 2):
        # KeySwitchingSetProducts: [decomp+1][decomp][N], table by poly index, last poly -> special prime
        x = rng.integers(0, 1 << 39, size=(decomp + 1) * decomp * o.n, dtype=np.uint64)
        d = to_dev(x)
        ctx.ntt_forward(d, pcount=decomp + 1, ncomp=decomp, mode=b2.TABLES_KS_SET_PRODUCTS, decomp=decomp)
        assert np.array_equal(to_np(d), o.ntt_fwd(x, decomp + 1, decomp, po.TABLES_KS_SET_PRODUCTS, 0, decomp))
        # KeySwitchingSkipFinals: [2][decomp+1][N], table by component, last component -> special prime
        x = rng.integers(0, 1 << 39, size=2 * (decomp + 1) * o.n, dtype=np.uint64)
        d = to_dev(x)
        ctx.ntt_inverse(d, pcount=2, ncomp=decomp + 1, mode=b2.TABLES_KS_SKIP_FINALS, decomp=decomp)
        assert np.array_equal(to_np(d), o.ntt_inv(x, 2, decomp + 1, po.TABLES_KS_SKIP_FINALS, 0, decomp))
    # componentwise with a table_base (single special-prime limb)
    x = rng.integers(0, int(o.moduli[K - 1]), size=o.n, dtype=np.uint64)
    d = to_dev(x)
    ctx.ntt_forward(d, pcount=1, ncomp=1, table_base=K - 1)
    assert np.array_equal(to_np(d), o.ntt_fwd(x, 1, 1, 0, K - 1, 0))


def test_ntt_batched_matches_single(b2, po):
    o = make_oracle(po, "C2")
    ctx = _ctx(b2, o)
    rng = np.random.default_rng(3)
    B, pc, nc = 8, 2, o.nm - 1
    x = np.concatenate([o.random_poly(rng, pc, nc) for _ in range(B)])
    d = to_dev(x)
    ctx.ntt_forward(d, batch=B, pcount=pc, ncomp=nc)
    assert np.array_equal(to_np(d), o.ntt_fwd(x, B * pc, nc))
    ctx.ntt_inverse(d, batch=B, pcount=pc, ncomp=nc)
    assert np.array_equal(to_np(d), x)


def test_ntt_full_size_properties_c3(b2, po):
    """BASELINE config C3 (N=2^16, 21 moduli), all limbs of a key-level polynomial pair: parity on the
    whole operand plus linearity and the convolution property (size-independent checks)."""
    o = make_oracle(po, "C3")
    ctx = _ctx(b2, o)
    rng = np.random.default_rng(4)
    K = o.nm
    a = o.random_poly(rng, 1, K); b = o.random_poly(rng, 1, K)
    da, db = to_dev(a), to_dev(b)
    ctx.ntt_forward(da, pcount=1, ncomp=K); ctx.ntt_forward(db, pcount=1, ncomp=K)
    A, Bv = to_np(da), to_np(db)
    assert np.array_equal(A, o.ntt_fwd(a, 1, K))
    s = o.add(a, b, 1, K)
    ds = to_dev(s); ctx.ntt_forward(ds, pcount=1, ncomp=K)
    assert np.array_equal(to_np(ds), o.add(A, Bv, 1, K)), "NTT(a+b) != NTT(a)+NTT(b)"
    # x * a  (negacyclic shift by one) <-> pointwise product with NTT(x)
    sh = o.negacyclic_shift(a, 1, 1, K)
    xpoly = np.zeros(K * o.n, dtype=np.uint64)
    for j in range(K):
        xpoly[j * o.n + 1] = 1
    dx = to_dev(xpoly); ctx.ntt_forward(dx, pcount=1, ncomp=K)
    prod = o.dyadic_broadcast(A, to_np(dx), 1, K)
    dsh = to_dev(sh); ctx.ntt_forward(dsh, pcount=1, ncomp=K)
    assert np.array_equal(to_np(dsh), prod), "NTT(x*a) != NTT(x) o NTT(a)"


def test_ntt_host_entry_point(b2, po):
    o = make_oracle(po, "C1")
    ctx = _ctx(b2, o)
    rng = np.random.default_rng(5)
    x = o.random_poly(rng, 2, o.nm)
    out = np.zeros_like(x)
    ctx.ntt_host(True, x, out, pcount=2, ncomp=o.nm)
    assert np.array_equal(out, o.ntt_fwd(x, 2, o.nm))
    back = np.zeros_like(x)
    ctx.ntt_host(False, out, back, pcount=2, ncomp=o.nm)
    assert np.array_equal(back, x)


def test_ntt_argument_errors(b2, po):
    o = make_oracle(po, scheme=2, log_n=12, bits=[40, 40])
    ctx = _ctx(b2, o)
    d = to_dev(np.zeros(3 * o.n, dtype=np.uint64))
    with pytest.raises(b2.B2Error):
        ctx.ntt_forward(d, pcount=1, ncomp=3)           # more components than moduli
    with pytest.raises(b2.B2Error):
        ctx.ntt_forward(d, pcount=1, ncomp=1, mode=7)   # unknown indexer
    with pytest.raises(b2.B2Error):
        b2.Context(12, [15, 17], 0, 2)                  # not NTT primes
