"""Limb-sharded keyswitch: host logic on CPU (no GPU).  world_size-2 `gloo` run of the collective choreography
(troy-nova_b200/sharded.py) with the three local stages implemented from the oracle (test infrastructure), checked
against the oracle's unsharded keyswitch; plus the partition function of the C ABI."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _mod_mul(a, b, m):
    return ((a.astype(object) * b.astype(object)) % m).astype(np.uint64)


class OracleStages:
    """prepare / products / finish on the CPU, restated from SURVEY.md §9 F with oracle transforms (CKKS / BFV)."""

    def __init__(self, o, po):
        self.o, self.po, self.n = o, po, o.n
        self.prod = None

    def empty(self, words):
        return torch.zeros(int(words), dtype=torch.int64)

    @staticmethod
    def _np(t):
        return t.numpy().view(np.uint64)

    def prepare(self, target_local, decomp, is_ntt, first, count, send):
        n, s, t = self.n, self._np(send), self._np(target_local)
        nct = 0 if first >= decomp else min(first + count, decomp) - first
        for jl in range(nct):
            limb = t[jl * n:(jl + 1) * n].copy()
            s[jl * n:(jl + 1) * n] = self.o.ntt_inv(limb, 1, 1, base=first + jl) if is_ntt else limb

    def products(self, gathered, world, pad, decomp, first, count, shard, u):
        from importlib import import_module
        sh = import_module("troy-nova_b200.sharded")
        n, K, q = self.n, self.o.nm, [int(x) for x in self.o.moduli]
        g = self._np(gathered)
        c = np.zeros((decomp, n), dtype=np.uint64)
        for r in range(world):
            f, cnt = sh.partition(decomp + 1, world, r)
            for jl in range(sh.owned_ct_limbs(decomp, f, cnt)):
                c[f + jl] = g[(r * pad + jl) * n:(r * pad + jl + 1) * n]
        self.prod = np.zeros((2, count, n), dtype=np.uint64)
        for il in range(count):
            i = first + il
            ki = K - 1 if i == decomp else i
            m = q[ki]
            for j in range(decomp):
                d = self.o.ntt_fwd((c[j] % np.uint64(m)).astype(np.uint64), 1, 1, base=ki)
                for k in range(2):
                    self.prod[k, il] = (self.prod[k, il].astype(object) + _mod_mul(d, shard[j, k, il], m).astype(object)) % m
        if first + count == decomp + 1:
            uu = self._np(u)
            for k in range(2):
                uu[k * n:(k + 1) * n] = self.o.ntt_inv(self.prod[k, count - 1].copy(), 1, 1, base=K - 1)

    def finish(self, u, decomp, is_ntt, first, count, dest_local, assign):
        n, K, q = self.n, self.o.nm, [int(x) for x in self.o.moduli]
        P = q[K - 1]; half = P >> 1
        uu, d = self._np(u), self._np(dest_local)
        nct = 0 if first >= decomp else min(first + count, decomp) - first
        for jl in range(nct):
            j = first + jl; qj = q[j]; pinv = pow(P % qj, -1, qj)
            for k in range(2):
                r = (uu[k * n:(k + 1) * n].astype(object) + half) % P
                t = ((r % qj) + qj - (half % qj)) % qj
                if is_ntt:
                    s = self.prod[k, jl].astype(object)
                    t = self.o.ntt_fwd(np.array(t, dtype=np.uint64), 1, 1, base=j).astype(object)
                else:
                    s = self.o.ntt_inv(self.prod[k, jl].copy(), 1, 1, base=j).astype(object)
                out = ((s - t) * pinv) % qj
                sl = slice((k * nct + jl) * n, (k * nct + jl + 1) * n)
                add = assign == 1 or (assign == 2 and k == 0)
                d[sl] = np.array(((d[sl].astype(object) + out) % qj) if add else out, dtype=np.uint64)


def _worker(rank, world, port, scheme_name, is_ntt, assign, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pyoracle as po
        sh = importlib.import_module("troy-nova_b200.sharded")
        log_n, n = 8, 256
        bits = [40, 30, 30, 40]
        moduli = po.coeff_modulus_create(n, bits)
        scheme = getattr(po, scheme_name)
        t = po.plain_modulus_batching(n, 20) if scheme != po.CKKS else 0
        o = po.Oracle(log_n, moduli, t, scheme)
        K = o.nm; decomp = K - 1
        rng = np.random.default_rng(7)                      # same stream on every rank
        key = o.random_key(rng).reshape(decomp, 2, K, n)
        target = o.random_poly(rng, 1, decomp).reshape(decomp, n)
        dest0 = o.random_poly(rng, 2, decomp).reshape(2, decomp, n)
        want = o.keyswitch(target.reshape(-1).copy(), decomp, is_ntt, key.reshape(-1).copy(), dest=dest0.reshape(-1).copy(), assign=assign).reshape(2, decomp, n)

        ks = sh.LimbShardedKeyswitch(OracleStages(o, po), decomp)
        f, c, nct = ks.first, ks.count, ks.nct
        rows = [K - 1 if f + il == decomp else f + il for il in range(c)]
        shard = np.ascontiguousarray(key[:, :, rows, :])
        tl = torch.from_numpy(np.ascontiguousarray(target[f:f + nct]).reshape(-1).view(np.int64).copy()) if nct else torch.zeros(2, dtype=torch.int64)
        dl = torch.from_numpy(np.ascontiguousarray(dest0[:, f:f + nct]).reshape(-1).view(np.int64).copy()) if nct else torch.zeros(2, dtype=torch.int64)
        ks(tl, is_ntt, shard, dl, assign)
        ok = True
        if nct:
            got = dl.numpy().view(np.uint64).reshape(2, nct, n)
            ok = bool(np.array_equal(got, want[:, f:f + nct]))
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(int(flag.item()))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.parametrize("scheme_name,is_ntt,assign", [("CKKS", True, 0), ("CKKS", True, 1), ("BFV", False, 2)])
@pytest.mark.parametrize("world", [2, 3])
def test_limb_sharded_keyswitch_gloo(world, scheme_name, is_ntt, assign):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, scheme_name, is_ntt, assign, ret)) for port in [_free_port()] for r in range(world)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert ret.get(timeout=5) == 1, "sharded keyswitch differs from the unsharded oracle"


def test_partition_matches_c_abi():
    b2 = importlib.import_module("troy-nova_b200")
    sh = importlib.import_module("troy-nova_b200.sharded")
    for n_out in (1, 2, 4, 9, 21, 36):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for r in range(world):
                f, c = sh.partition(n_out, world, r)
                assert (f, c) == b2.limb_partition(n_out, world, r)
                covered += list(range(f, f + c))
            assert covered == list(range(n_out))
    assert [sh.partition(21, 8, r)[1] for r in range(8)] == [3, 3, 3, 3, 3, 2, 2, 2]
