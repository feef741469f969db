"""Limb-sharded keyswitch on the GPU through the C ABI: every simulated world size is run rank by
rank on ONE device (the all-gather and the broadcast are emulated by copies), bit-exact against the oracle's unsharded
keyswitch.  The NCCL choreography itself is covered by tests/test_sharded_cpu.py (gloo) and tools/sharded_check.py
(torchrun, N GPUs)."""
import importlib

import numpy as np
import pytest
import torch

from conftest import make_oracle, to_dev, to_np

pytestmark = pytest.mark.gpu


def _run_world(b2, sh, o, ctx, kfull, key, world, is_ntt, assign, rng):
    n, K = o.n, o.nm
    decomp = K - 1
    target = o.random_poly(rng, 1, decomp).reshape(decomp, n)
    dest0 = o.random_poly(rng, 2, decomp).reshape(2, decomp, n)
    want = o.keyswitch(target.reshape(-1).copy(), decomp, is_ntt, key, dest=dest0.reshape(-1).copy(), assign=assign).reshape(2, decomp, n)
    parts = [sh.partition(decomp + 1, world, r) for r in range(world)]
    ncts = [sh.owned_ct_limbs(decomp, f, c) for f, c in parts]
    pad = max(ncts)
    gathered = torch.zeros(world * pad * n, dtype=torch.int64, device="cuda")
    # stage 1 on every rank, "all-gather" = each rank's send buffer lands in its slot
    for r, (f, c) in enumerate(parts):
        if ncts[r]:
            ctx.keyswitch_shard_prepare(to_dev(target[f:f + ncts[r]]), decomp, is_ntt, f, c, gathered[r * pad * n:])
    # stage 2 + 3 rank by rank (prod lives in the context workspace between the two calls)
    u = torch.zeros(2 * n, dtype=torch.int64, device="cuda")
    owner = next(r for r, (f, c) in enumerate(parts) if c and f + c == decomp + 1)
    order = [owner] + [r for r in range(world) if r != owner]        # the owner's products produce u ("broadcast")
    shards = {r: ctx.kswitch_key_shard(kfull, decomp, *parts[r]) for r in range(world) if parts[r][1]}
    got = np.zeros_like(want)
    ctx.keyswitch_shard_products(gathered, world, pad, decomp, *parts[owner], shards[owner], u)
    u_owner = u.clone()
    for r in order:
        f, c = parts[r]
        if c == 0:
            continue
        scratch = torch.zeros(2 * n, dtype=torch.int64, device="cuda")
        ctx.keyswitch_shard_products(gathered, world, pad, decomp, f, c, shards[r], scratch)
        if ncts[r]:
            dl = to_dev(np.ascontiguousarray(dest0[:, f:f + ncts[r]]))
            ctx.keyswitch_shard_finish(u_owner, decomp, is_ntt, f, c, dl, assign)
            got[:, f:f + ncts[r]] = to_np(dl).reshape(2, ncts[r], n)
    assert np.array_equal(got, want), f"world={world} is_ntt={is_ntt} assign={assign}"


@pytest.mark.parametrize("scheme", [1, 2, 3])
def test_sharded_keyswitch_small(b2, po, scheme):
    sh = importlib.import_module("troy-nova_b200.sharded")
    o = make_oracle(po, scheme=scheme, log_n=12, bits=[50, 30, 45, 40, 55], t_bits=20 if scheme != 2 else 0)
    ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
    rng = np.random.default_rng(40 + scheme)
    key = o.random_key(rng)
    kfull = ctx.kswitch_key(to_dev(key))
    for world in (1, 2, 3, 5, 8):
        for assign in (b2.ASSIGN_OVERWRITE, b2.ASSIGN_ADD_INPLACE, b2.ASSIGN_OVERWRITE_EXCEPT_FIRST):
            _run_world(b2, sh, o, ctx, kfull, key, world, scheme != 1, assign, rng)


def test_sharded_keyswitch_c3_full_size(b2, po):
    """BASELINE config C3 (CKKS N=2^16, 20 + 1 moduli) cut over 8 ranks: 3,3,3,3,3,2,2,2 output moduli."""
    sh = importlib.import_module("troy-nova_b200.sharded")
    o = make_oracle(po, "C3")
    ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
    rng = np.random.default_rng(41)
    key = o.random_key(rng)
    kfull = ctx.kswitch_key(to_dev(key))
    _run_world(b2, sh, o, ctx, kfull, key, 8, True, b2.ASSIGN_ADD_INPLACE, rng)


def _run_world_p2p(b2, sh, o, ctx, kfull, key, world, is_ntt, assign, rng, calls=3):
    """the P2P data plane (b2_shard_push / b2_shard_wait, weighted partition) with every simulated rank's buffers in this process:
    each rank pushes into all ranks' buffers; stages run rank by rank in an order in which every flag is raised before it is awaited"""
    n, K = o.n, o.nm
    decomp = K - 1
    sizes = sh.P2PExchange.sizes(decomp, n, world)
    bufs = {name: [torch.zeros(words, dtype=torch.int64, device="cuda") for _ in range(world)] for name, words in sizes.items()}
    peers = {name: [t.data_ptr() for t in ts] for name, ts in bufs.items()}
    ranks = [sh.P2PLimbShardedKeyswitch(ctx, decomp, sh.P2PExchange(r, world, {k: v[r] for k, v in bufs.items()}, peers)) for r in range(world)]
    parts = ranks[0].parts
    assert sum(c for _, c in parts) == decomp + 1 and all(parts[r][0] + parts[r][1] == parts[r + 1][0] for r in range(world - 1))
    shards = {r: ctx.kswitch_key_shard(kfull, decomp, *parts[r]) for r in range(world) if parts[r][1]}
    owner = ranks[0].special_owner
    for _ in range(calls):                                  # several calls: buffer parity and epochs
        target = o.random_poly(rng, 1, decomp).reshape(decomp, n)
        dest0 = o.random_poly(rng, 2, decomp).reshape(2, decomp, n)
        want = o.keyswitch(target.reshape(-1).copy(), decomp, is_ntt, key, dest=dest0.reshape(-1).copy(), assign=assign).reshape(2, decomp, n)
        for r, ks in enumerate(ranks):
            f = ks.first
            ks.stage_prepare_push(to_dev(target[f:f + ks.nct]) if ks.nct else None, is_ntt)
        # one context = one workspace here: a rank's products must be followed by its own finish before the next rank's products;
        # the owner goes first so that u has been pushed to everybody (its second run below pushes the same u again)
        ranks[owner].stage_products(shards.get(owner))
        got = np.zeros_like(want)
        for r, ks in enumerate(ranks):
            ks.stage_products(shards.get(r))
            if ks.nct:
                dl = to_dev(np.ascontiguousarray(dest0[:, ks.first:ks.first + ks.nct]))
                ks.stage_finish(is_ntt, dl, assign)
                got[:, ks.first:ks.first + ks.nct] = to_np(dl).reshape(2, ks.nct, n)
        torch.cuda.synchronize()
        assert np.array_equal(got, want), f"p2p world={world} is_ntt={is_ntt} assign={assign}"


@pytest.mark.parametrize("scheme", [1, 2, 3])
def test_p2p_sharded_keyswitch_small(b2, po, scheme):
    sh = importlib.import_module("troy-nova_b200.sharded")
    o = make_oracle(po, scheme=scheme, log_n=12, bits=[50, 30, 45, 40, 55], t_bits=20 if scheme != 2 else 0)
    ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
    rng = np.random.default_rng(70 + scheme)
    key = o.random_key(rng)
    kfull = ctx.kswitch_key(to_dev(key))
    for world in (1, 2, 3, 5, 8):
        for assign in (b2.ASSIGN_OVERWRITE, b2.ASSIGN_ADD_INPLACE):
            _run_world_p2p(b2, sh, o, ctx, kfull, key, world, scheme != 1, assign, rng, calls=2)


def test_p2p_sharded_keyswitch_c3_full_size(b2, po):
    """C3 over 8 ranks with the weighted partition: q_0 (60-bit) shares a rank with one 40-bit modulus, the special prime has a rank of its own"""
    sh = importlib.import_module("troy-nova_b200.sharded")
    o = make_oracle(po, "C3")
    ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
    parts = [ctx.limb_partition_weighted(o.nm - 1, 8, r) for r in range(8)]
    assert parts[0] == (0, 2) and parts[7] == (20, 1), parts
    rng = np.random.default_rng(43)
    key = o.random_key(rng)
    kfull = ctx.kswitch_key(to_dev(key))
    _run_world_p2p(b2, sh, o, ctx, kfull, key, 8, True, b2.ASSIGN_ADD_INPLACE, rng, calls=2)


def test_sharded_argument_errors(b2, po):
    o = make_oracle(po, scheme=2, log_n=12, bits=[50, 40, 55])
    ctx = b2.Context(o.log_n, o.moduli, o.t, o.scheme)
    rng = np.random.default_rng(1)
    kfull = ctx.kswitch_key(to_dev(o.random_key(rng)))
    with pytest.raises(b2.B2Error):
        ctx.kswitch_key_shard(kfull, 2, 2, 2)                 # range past the special prime
    shard = ctx.kswitch_key_shard(kfull, 2, 0, 2)
    d = torch.zeros(2 * 2 * o.n, dtype=torch.int64, device="cuda")
    with pytest.raises(b2.B2Error):                           # a shard is not a full key
        ctx.keyswitch(d, 2, True, shard, d)
    with pytest.raises(b2.B2Error):                           # shard range mismatch
        ctx.keyswitch_shard_products(d, 1, 2, 2, 1, 2, shard, d)
