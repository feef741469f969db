"""B1 / M3 / X1: the scattered-pointer batched entry points run as ONE launch and match the oracle bit for bit
(reference: ntt_transfer_*_batched, src/fgk/ntt_grouped.cu:297-344,642-685; multiply_plain_ntt_accumulate with repeated
destinations, src/evaluator_multiply_plain.cu:258-307); plus the re-entrancy of one context across streams / threads."""
import threading

import numpy as np
import pytest

from conftest import make_oracle, to_dev, to_np

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("log_n,batch", [(10, 5), (13, 37), (14, 256), (16, 3), (16, 24)])
def test_ntt_ptrs_single_launch(b2, po, log_n, batch):
    import torch
    o = make_oracle(po, scheme=1, log_n=log_n, bits=[60, 40, 40, 60], t_bits=20)
    ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
    rng = np.random.default_rng(batch)
    pc, nc = 2, o.nm - 1
    per = pc * nc * o.n
    xs = [o.random_poly(rng, pc, nc) for _ in range(batch)]
    # scattered: separate allocations in shuffled order, with gaps
    order = rng.permutation(batch)
    bufs = [None] * batch
    for i in order:
        pad = torch.empty(int(rng.integers(1, 64)) * 16, dtype=torch.int64, device="cuda")     # noqa: F841  (keeps addresses irregular)
        bufs[i] = to_dev(xs[i])
    outs = [torch.zeros(per, dtype=torch.int64, device="cuda") for _ in range(batch)]
    before = b2.launch_count()
    ctx.ntt_ptrs(True, bufs, outs, pcount=pc, ncomp=nc)
    assert b2.launch_count() - before == 1, "scattered batched NTT must be one launch"
    for i in range(batch):
        assert np.array_equal(to_np(outs[i]), o.ntt_fwd(xs[i], pc, nc)), "item %d forward" % i
        assert np.array_equal(to_np(bufs[i]), xs[i]), "input modified"
    before = b2.launch_count()
    ctx.ntt_ptrs(False, outs, None, pcount=pc, ncomp=nc)                  # in place
    assert b2.launch_count() - before == 1
    for i in range(batch):
        assert np.array_equal(to_np(outs[i]), xs[i]), "item %d round trip" % i


@pytest.mark.parametrize("name,n_dest,terms_per_dest", [("C5", 6, 5), ("C2", 3, 70), ("C3", 2, 3)])
def test_multiply_plain_accumulate_repeated_destinations(b2, po, name, n_dest, terms_per_dest):
    o = make_oracle(po, name)
    ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
    rng = np.random.default_rng(7)
    pc, nc = 2, o.nm - 1
    nt = n_dest * terms_per_dest
    cts = [o.random_poly(rng, pc, nc) for _ in range(nt)]
    pts = [o.random_poly(rng, 1, nc) for _ in range(nt)]
    init = [o.random_poly(rng, pc, nc) for _ in range(n_dest)]
    which = rng.permutation(np.repeat(np.arange(n_dest), terms_per_dest))    # term i -> destination which[i], interleaved
    for add in (False, True):
        want = [x.copy() if add else None for x in init]
        for i in range(nt):                       # the reference's order: dest[i] (+)= ct[i] o pt[i], i ascending
            d = int(which[i])
            want[d] = o.dyadic_broadcast(cts[i], pts[i], pc, nc) if want[d] is None else o.dyadic_broadcast(cts[i], pts[i], pc, nc, acc_into=want[d])
        dd = [to_dev(x) for x in init]
        dc, dp = [to_dev(x) for x in cts], [to_dev(x) for x in pts]
        before = b2.launch_count()
        ctx.multiply_plain_accumulate(dc, dp, [dd[int(w)] for w in which], pc, nc, add_to_dest=add)
        assert b2.launch_count() - before == 1, "multiply_plain_accumulate must be one launch"
        for d in range(n_dest):
            assert np.array_equal(to_np(dd[d]), want[d]), "destination %d (add=%s)" % (d, add)


def test_context_is_reentrant_across_streams_and_threads(b2, po):
    """two host threads, each on its own stream, hammer ONE context with keyswitches and large transforms (the persistent
    NTT's work queue and the workspace arena used to be per context: ADVICE r01 high)"""
    import torch
    o = make_oracle(po, scheme=2, log_n=14, bits=[60, 40, 40, 40, 60])
    ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
    rng = np.random.default_rng(3)
    L = o.nm - 1
    key = o.random_key(rng)
    kh = ctx.kswitch_key(to_dev(key))
    jobs = []
    for w in range(2):
        tgt = o.random_poly(rng, 1, L)
        big = o.random_poly(rng, 16, o.nm)
        jobs.append((tgt, o.keyswitch(tgt, L, True, key), big, o.ntt_fwd(big, 16, o.nm)))
    errs = []

    def worker(w):
        try:
            s = torch.cuda.Stream()
            tgt, want_ks, big, want_ntt = jobs[w]
            with torch.cuda.stream(s):
                for it in range(25):
                    dt, dest = to_dev(tgt), to_dev(np.zeros(2 * L * o.n, dtype=np.uint64))
                    db = to_dev(big)
                    ctx.keyswitch(dt, L, True, kh, dest, stream=s.cuda_stream)
                    ctx.ntt_forward(db, pcount=16, ncomp=o.nm, stream=s.cuda_stream)
                    s.synchronize()
                    if not np.array_equal(to_np(dest), want_ks):
                        errs.append("thread %d iteration %d: keyswitch corrupted" % (w, it)); return
                    if not np.array_equal(to_np(db), want_ntt):
                        errs.append("thread %d iteration %d: NTT corrupted" % (w, it)); return
        except Exception as e:      # noqa: BLE001
            errs.append("thread %d: %r" % (w, e))

    ts = [threading.Thread(target=worker, args=(w,)) for w in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs


def test_small_n_keyswitch_mixed_37_40_bit_primes(b2, po):
    """at N <= 2048 the integer-only kernel runs, so the mod-up loader must not skip the reduction of a
    digit modulus that is within 4 bits of — but more than 4x — the output modulus"""
    for log_n in (10, 11, 12):
        o = make_oracle(po, scheme=2, log_n=log_n, bits=[40, 37, 40])
        ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
        rng = np.random.default_rng(log_n)
        L = o.nm - 1
        key = o.random_key(rng)
        kh = ctx.kswitch_key(to_dev(key))
        tgt = o.random_poly(rng, 1, L)
        dest = to_dev(np.zeros(2 * L * o.n, dtype=np.uint64))
        ctx.keyswitch(to_dev(tgt), L, True, kh, dest)
        assert np.array_equal(to_np(dest), o.keyswitch(tgt, L, True, key)), "log_n=%d" % log_n
