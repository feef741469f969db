"""Parity at BASELINE size (VERDICT r01 "what's weak" 1 and 2).

(1) the headline instantiation — ntt_persistent_kernel<8, ., 3, PlainLoad, PlainStore> over 420 limbs at N = 2^16 — against
    the C oracle, forward and inverse (every other plain N = 2^16 test has <= 32 limbs and takes the cluster kernel);
(2) the GPU against the UNMODIFIED REFERENCE itself (oracle/_ref host branch through ref_shim) at the full C3 / C4 / C2 / C5
    parameter sets, with the reference's own context (seed 0x123), KeyGenerator keys and Encryptor ciphertexts: multiply,
    relinearize, rotate(1), rotate(7) (NAF: -1 then 8), rescale / mod-switch chain, word for word.
    Reference op list: test/bench/he_operations.cu:695-1023; keyswitch src/evaluator_keyswitching_core.cu:757-1052."""
import numpy as np
import pytest

from conftest import CONFIGS, make_oracle, to_dev, to_np

pytestmark = pytest.mark.gpu


def test_ntt_420_limbs_persistent_kernel_vs_oracle(b2, po):
    o = make_oracle(po, "C3")
    ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
    rng = np.random.default_rng(420)
    pc, nc = 20, o.nm                       # the [20][21][N] key-level operand of bench.py
    x = o.random_poly(rng, pc, nc)
    d = to_dev(x)
    ctx.ntt_forward(d, pcount=pc, ncomp=nc)
    want = o.ntt_fwd(x, pc, nc)
    assert np.array_equal(to_np(d), want), "420-limb forward NTT differs from the oracle"
    ctx.ntt_inverse(d, pcount=pc, ncomp=nc)
    assert np.array_equal(to_np(d), x), "inverse(forward(x)) != x at 420 limbs"
    y = o.random_poly(rng, pc, nc)
    dy, dz = to_dev(y), to_dev(np.zeros_like(y))
    ctx.ntt_inverse(dy, dz, pcount=pc, ncomp=nc)
    assert np.array_equal(to_np(dz), o.ntt_inv(y, pc, nc)), "420-limb inverse NTT differs from the oracle"
    # the keyswitch table mode at the same size: [21][20][N], table by polynomial, last polynomial -> special prime
    K = o.nm
    z = rng.integers(0, 1 << 39, size=K * (K - 1) * o.n, dtype=np.uint64)
    dzz = to_dev(z)
    ctx.ntt_forward(dzz, pcount=K, ncomp=K - 1, mode=b2.TABLES_KS_SET_PRODUCTS, decomp=K - 1)
    assert np.array_equal(to_np(dzz), o.ntt_fwd(z, K, K - 1, mode=po.TABLES_KS_SET_PRODUCTS, decomp=K - 1))


def _naf(v):        # utils::naf, src/utils/number_theory.cu:6-21
    res, sign, v, i = [], v < 0, abs(v), 0
    while v > 0:
        zi = (2 - (v & 3)) if (v & 1) else 0
        v = (v - zi) >> 1
        if zi:
            res.append((-zi if sign else zi) << i)
        i += 1
    return res


class _Gpu:
    """the reference's Evaluator flows expressed over the C ABI (what host/b2_evaluator.cpp does in C++)"""

    def __init__(self, b2, ref, cfg):
        self.b2, self.ref = b2, ref
        self.n, self.K = ref.n, ref.nm
        self.scheme = cfg["scheme"]
        self.ctx = b2.Context(cfg["log_n"], [int(m) for m in ref.moduli], ref.t, self.scheme)
        self.ntt_form = self.scheme != b2.BFV
        self.keys = {}

    def key(self, name, arr):
        if name not in self.keys:
            self.keys[name] = self.ctx.kswitch_key(to_dev(arr))
        return self.keys[name]

    def multiply(self, a, b, L):
        out = to_dev(np.zeros(3 * L * self.n, dtype=np.uint64))
        if self.scheme == self.b2.BFV:
            self.ctx.bfv_multiply(to_dev(a), to_dev(b), out, L)
        else:
            self.ctx.dyadic_convolute(to_dev(a), 2, to_dev(b), 2, out, L)
        return to_np(out)

    def relinearize(self, c3, L, key):
        n = self.n
        d = to_dev(c3)
        dest = d[: 2 * L * n].clone()
        self.ctx.keyswitch(d[2 * L * n:], L, self.ntt_form, key, dest, self.b2.ASSIGN_ADD_INPLACE)
        return to_np(dest)

    def apply_galois(self, c2, L, elt, key):
        d = to_dev(c2)
        dest = to_dev(np.zeros_like(c2))
        self.ctx.galois(d, elt, dest, 2, L, self.ntt_form)
        tmp = dest[L * self.n:].clone()
        self.ctx.keyswitch(tmp, L, self.ntt_form, key, dest, self.b2.ASSIGN_OVERWRITE_EXCEPT_FIRST)
        return to_np(dest)


@pytest.mark.parametrize("name", ["C3", "C4", "C2", "C5"])
def test_gpu_vs_reference_host_branch_full_size(b2, po, name):
    if not po.ref_available():
        pytest.skip("oracle/_ref (compiled reference) not present")
    cfg = CONFIGS[name]
    n = 1 << cfg["log_n"]
    ref = po.Reference(cfg["scheme"], n, cfg["bits"], cfg["t_bits"] or 20, seed=0x123, device=False)
    L = ref.nm - 1
    g = _Gpu(b2, ref, cfg)
    ref.gen_relin()
    rk = g.key("relin", ref.relin_key())
    steps = [1] + _naf(7)
    elts = [ref.galois_element_from_step(s) for s in steps]
    ref.gen_galois(sorted(set(elts)))
    ct = ref.encrypt_zero()                                   # [2][L][N] from the reference Encryptor
    ct2 = ref.encrypt_zero()
    # multiply
    prod = ref.multiply(ct, ct2, L)
    assert np.array_equal(g.multiply(ct, ct2, L), prod), name + ": multiply differs from the reference"
    # relinearize
    relin = ref.relinearize(prod, L)
    assert np.array_equal(g.relinearize(prod, L, rk), relin), name + ": relinearize differs from the reference"
    # rotate(1): one Galois key;  rotate(7): NAF composition (-1, 8)
    rot1 = ref.rotate(relin, 1, L)
    got = g.apply_galois(relin, L, elts[0], g.key(elts[0], ref.galois_key(elts[0])))
    assert np.array_equal(got, rot1), name + ": rotate(1) differs from the reference"
    rot7 = ref.rotate(relin, 7, L)
    cur = relin
    for e in elts[1:]:
        cur = g.apply_galois(cur, L, e, g.key(e, ref.galois_key(e)))
    assert np.array_equal(cur, rot7), name + ": rotate(7) (NAF) differs from the reference"
    # rescale (CKKS) / mod_switch_to_next (BFV, BGV), down the whole chain
    kind = {b2.CKKS: b2.MODSWITCH_CKKS, b2.BFV: b2.MODSWITCH_BFV, b2.BGV: b2.MODSWITCH_BGV}[cfg["scheme"]]
    cur_ref, nl = relin, L
    while nl >= 2:
        nxt = ref.mod_switch_scale_to_next(cur_ref, 2, nl)
        out = to_dev(np.zeros(2 * (nl - 1) * n, dtype=np.uint64))
        g.ctx.mod_switch_to_next(to_dev(cur_ref), out, 2, nl, kind)
        assert np.array_equal(to_np(out), nxt), "%s: mod switch %d -> %d differs from the reference" % (name, nl, nl - 1)
        cur_ref, nl = nxt, nl - 1
        if name == "C3" and nl < L - 2:
            break                                            # three levels are enough at 20 limbs (CPU time)
