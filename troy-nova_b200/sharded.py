"""Limb-sharded keyswitch across the GPUs of one box — the collective choreography.

One process per GPU (`torchrun`); `torch.distributed` carries the two exchanges a keyswitch needs when the RNS limbs are
partitioned over ranks: the ALL-GATHER of the coefficient-form target (decomp * N words in total) before mod-up, and the
BROADCAST of the special-prime limb u (2 N words) before mod-down.  Everything else — transforms, key inner product,
rounding division — is local to the rank that owns the modulus and runs in libb2rns (b2_keyswitch_shard_*).

Layout on rank r: owned output moduli [first, first + count) of 0..decomp (decomp = special prime);
ciphertext polynomials as [poly][owned ciphertext limbs][N]; key rows of the owned moduli only (b2_kswitch_key_shard).

The class is written against a small `stages` interface (prepare / products / finish + buffer allocation) so that the
choreography itself is testable without a GPU: tests/test_sharded_cpu.py drives it over `gloo` with world size 2 and a
CPU stage implementation built from the oracle (test infrastructure), bench.py / the GPU tests use `DeviceStages`.
No arithmetic lives in this file.
"""
import torch
import torch.distributed as dist


def partition(n_out_moduli, world, rank):
    """Contiguous balanced ranges: the first (n mod world) ranks own one modulus more (mirrors b2_limb_partition)."""
    base, rem = divmod(n_out_moduli, world)
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def owned_ct_limbs(decomp, first, count):
    return 0 if first >= decomp else min(first + count, decomp) - first


class DeviceStages:
    """The three local stages on this rank's GPU through the C ABI (b2_keyswitch_shard_*)."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.n = ctx.n

    def empty(self, words):
        return torch.empty(max(int(words), 2), dtype=torch.int64, device="cuda")

    # a rank that owns no modulus (world > decomp + 1) only takes part in the collectives
    def prepare(self, target_local, decomp, is_ntt, first, count, send):
        if count:
            self.ctx.keyswitch_shard_prepare(target_local, decomp, is_ntt, first, count, send)

    def products(self, gathered, world, pad, decomp, first, count, shard, u):
        if count:
            self.ctx.keyswitch_shard_products(gathered, world, pad, decomp, first, count, shard, u)

    def finish(self, u, decomp, is_ntt, first, count, dest_local, assign):
        if count:
            self.ctx.keyswitch_shard_finish(u, decomp, is_ntt, first, count, dest_local, assign)


class LimbShardedKeyswitch:
    """keyswitch(target_local, key_shard, dest_local): Evaluator::switch_key_internal
    (src/evaluator_keyswitching_core.cu:757-1052) with the output moduli partitioned over the process group."""

    def __init__(self, stages, decomp, group=None):
        self.stages, self.decomp, self.group = stages, decomp, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.first, self.count = partition(decomp + 1, self.world, self.rank)
        self.nct = owned_ct_limbs(decomp, self.first, self.count)
        # every rank contributes `pad` limbs to the all-gather (NCCL all-gather needs equal shares)
        self.pad = max(owned_ct_limbs(decomp, *partition(decomp + 1, self.world, r)) for r in range(self.world))
        self.special_owner = next(r for r in range(self.world) if partition(decomp + 1, self.world, r)[1] > 0 and sum(partition(decomp + 1, self.world, r)) == decomp + 1)
        n = stages.n
        self.send = stages.empty(self.pad * n)
        self.gathered = stages.empty(self.world * self.pad * n)
        self.u = stages.empty(2 * n)

    def __call__(self, target_local, is_ntt, key_shard, dest_local, assign=0):
        s = self.stages
        s.prepare(target_local, self.decomp, is_ntt, self.first, self.count, self.send)
        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered, self.send, group=self.group)
        else:
            self.gathered.copy_(self.send)
        s.products(self.gathered, self.world, self.pad, self.decomp, self.first, self.count, key_shard, self.u)
        if self.world > 1:
            dist.broadcast(self.u, src=dist.get_global_rank(self.group, self.special_owner) if self.group is not None else self.special_owner,
                           group=self.group)
        s.finish(self.u, self.decomp, is_ntt, self.first, self.count, dest_local, assign)
        return dest_local


class P2PExchange:
    """Buffers of the P2P data plane on one rank, with every peer's copy mapped into this process.

    `peers[name]` is the list (by rank) of raw device pointers of buffer `name`; buffer `name` of this rank is `local[name]`.
    from_symmetric_memory() allocates them with torch.distributed._symmetric_memory (one process per GPU, NVLink / NVSwitch);
    tests build the same object over plain tensors of ONE process to drive the kernels rank by rank."""

    def __init__(self, rank, world, local, peers):
        self.rank, self.world, self.local, self.peers = rank, world, local, peers

    @staticmethod
    def sizes(decomp, n, world):
        # tc: two parities of the coefficient-form target [decomp][N]; u: two parities of [2][N]; flags: [2 exchanges][world]
        return {"tc": 2 * decomp * n, "u": 2 * 2 * n, "flags": 2 * max(world, 2)}

    @classmethod
    def from_symmetric_memory(cls, decomp, n, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = group or dist.group.WORLD
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        local, peers, keep = {}, {}, []
        for name, words in cls.sizes(decomp, n, world).items():
            t = symm_mem.empty(words, dtype=torch.int64, device=torch.device("cuda", torch.cuda.current_device()))
            t.zero_()
            h = symm_mem.rendezvous(t, group)
            local[name] = t
            peers[name] = [int(p) for p in h.buffer_ptrs]
            keep.append(h)
        torch.cuda.synchronize()
        dist.barrier(group)
        ex = cls(rank, world, local, peers)
        ex._keep = keep
        return ex


class P2PLimbShardedKeyswitch:
    """The limb-sharded keyswitch with the exchanges done by the kernels themselves (b2_shard_push / b2_shard_wait): every rank
    pushes its coefficient-form limbs straight into all peers' target buffers (final layout: no all-gather call, no compaction),
    the special prime's owner pushes u; flags with system-scope release / acquire order the stages.  Partition weighted by
    modulus cost (b2_limb_partition_weighted).  Buffers alternate by the parity of the call so that a fast rank's next push can
    never land in a buffer a slow rank is still reading (the u exchange of call k orders every rank behind call k - 1)."""

    def __init__(self, ctx, decomp, exchange, epoch0=0):
        self.ctx, self.decomp, self.ex = ctx, decomp, exchange
        self.world, self.rank = exchange.world, exchange.rank
        self.parts = [ctx.limb_partition_weighted(decomp, self.world, r) for r in range(self.world)]
        self.first, self.count = self.parts[self.rank]
        self.nct = owned_ct_limbs(decomp, self.first, self.count)
        self.ct_owners = [r for r, (f, c) in enumerate(self.parts) if owned_ct_limbs(decomp, f, c) > 0]
        self.special_owner = next(r for r, (f, c) in enumerate(self.parts) if c and f + c == decomp + 1)
        self.n = ctx.n
        self.send = torch.empty(max(self.nct, 1) * self.n, dtype=torch.int64, device="cuda")
        self.u_local = torch.empty(2 * self.n, dtype=torch.int64, device="cuda")
        self.epoch = epoch0

    def stage_prepare_push(self, target_local, is_ntt):
        """INTT of the owned ciphertext limbs, pushed into every peer's tc[parity] at this rank's limb offset; raises flag A"""
        self.epoch += 1
        par, n, W = self.epoch & 1, self.n, self.world
        if self.nct:
            self.ctx.keyswitch_shard_prepare(target_local, self.decomp, is_ntt, self.first, self.count, self.send)
        off = (par * self.decomp + self.first) * n * 8
        self.ctx.shard_push(self.send, [p + off for p in self.ex.peers["tc"]], self.nct * n, self.ex.peers["flags"], self.rank, self.epoch)

    def stage_products(self, key_shard):
        par, n, W = self.epoch & 1, self.n, self.world
        if self.count == 0:
            return
        self.ctx.shard_wait(self.ex.local["flags"], self.ct_owners, self.epoch)
        tc = self.ex.local["tc"][par * self.decomp * n:]
        self.ctx.keyswitch_shard_products(tc, 0, self.decomp, self.decomp, self.first, self.count, key_shard, self.u_local)
        if self.rank == self.special_owner:
            self.ctx.shard_push(self.u_local, [p + par * 2 * n * 8 for p in self.ex.peers["u"]], 2 * n, [p + W * 8 for p in self.ex.peers["flags"]], 0, self.epoch)

    def stage_finish(self, is_ntt, dest_local, assign):
        par, n, W = self.epoch & 1, self.n, self.world
        if self.nct == 0:
            return dest_local
        self.ctx.shard_wait(self.ex.local["flags"][W:], [0], self.epoch)
        self.ctx.keyswitch_shard_finish(self.ex.local["u"][par * 2 * n:], self.decomp, is_ntt, self.first, self.count, dest_local, assign)
        return dest_local

    def __call__(self, target_local, is_ntt, key_shard, dest_local, assign=0):
        self.stage_prepare_push(target_local, is_ntt)
        self.stage_products(key_shard)
        return self.stage_finish(is_ntt, dest_local, assign)
