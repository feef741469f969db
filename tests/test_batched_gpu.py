"""K5 / B1: Evaluator::switch_key_internal_batched (src/evaluator_keyswitching_core.cu:1055-1267) through
b2_keyswitch_batched — batches of ciphertexts sharing one key, given as host arrays of device pointers (the reference's
ConstSliceVec), including batches larger than one 16-ciphertext launch set, and empty inputs."""
import numpy as np
import pytest
import torch

from conftest import make_oracle, to_dev, to_np

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scheme", [1, 2, 3])
@pytest.mark.parametrize("batch", [1, 3, 17])
def test_keyswitch_batched_matches_oracle(b2, po, scheme, batch):
    o = make_oracle(po, scheme=scheme, log_n=12, bits=[55, 40, 40, 60], t_bits=20 if scheme != 2 else 0)
    ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
    rng = np.random.default_rng(500 + 10 * scheme + batch)
    key = o.random_key(rng); kh = ctx.kswitch_key(to_dev(key))
    is_ntt = scheme != 1
    for decomp in (o.nm - 1, 2):
        for assign in (b2.ASSIGN_OVERWRITE, b2.ASSIGN_ADD_INPLACE, b2.ASSIGN_OVERWRITE_EXCEPT_FIRST):
            targets = [o.random_poly(rng, 1, decomp) for _ in range(batch)]
            dests0 = [o.random_poly(rng, 2, decomp) for _ in range(batch)]
            dts = [to_dev(t) for t in targets]; dds = [to_dev(d) for d in dests0]
            ctx.keyswitch_batched(dts, decomp, is_ntt, kh, dds, assign)
            for i in range(batch):
                want = o.keyswitch(targets[i], decomp, is_ntt, key, dest=dests0[i], assign=assign)
                assert np.array_equal(to_np(dds[i]), want), f"scheme {scheme} batch {batch} item {i} decomp {decomp} assign {assign}"


def test_c2_batched_rotate_rows_slice(b2, po):
    """BASELINE config C2 (BFV N=16384, {60,40,40}+60): a 20-ciphertext slice of the 256-ciphertext rotate_rows batch."""
    o = make_oracle(po, "C2")
    ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
    rng = np.random.default_rng(77)
    key = o.random_key(rng); kh = ctx.kswitch_key(to_dev(key))
    L, n, B = o.nm - 1, o.n, 20
    g = o.galois_element_from_step(1)
    cts = [o.random_poly(rng, 2, L) for _ in range(B)]
    dev = torch.stack([to_dev(c) for c in cts]).reshape(B, 2, L * n)
    rot = torch.empty_like(dev)
    ctx.galois(dev, g, rot, 2 * B, L, False)
    ctx.keyswitch_batched([rot[b, 1] for b in range(B)], L, False, kh, [rot[b] for b in range(B)], b2.ASSIGN_OVERWRITE_EXCEPT_FIRST)
    for b in range(B):
        r = o.galois(cts[b], g, 2, L, False)
        want = o.keyswitch(r[L * n:], L, False, key, dest=r, assign=po.ASSIGN_OVERWRITE_EXCEPT_FIRST)
        assert np.array_equal(to_np(rot[b].reshape(-1)), want), f"ciphertext {b}"


def test_empty_inputs_are_no_ops(b2, po):
    o = make_oracle(po, scheme=2, log_n=12, bits=[50, 40, 55])
    ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
    rng = np.random.default_rng(3)
    kh = ctx.kswitch_key(to_dev(o.random_key(rng)))
    before = b2.launch_count()
    ctx.keyswitch_batched([], 2, True, kh, [])                                   # empty batch (reference: early return, :1063)
    d = to_dev(np.zeros(2 * o.n, dtype=np.uint64))
    ctx.dyadic_broadcast_batched(d, d, d, 0, 2, 1)
    ctx.add(d, d, d, 0, 1)
    ctx.dyadic_convolute(d, 2, d, 2, d, 1, batch=0)
    assert b2.launch_count() == before, "empty inputs must not launch kernels"
