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
