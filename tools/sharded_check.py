"""torchrun check + timing of the limb-sharded keyswitch over NCCL (one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py [--small]
Every rank builds the same synthetic key / ciphertext (same seed), keeps only its limbs, runs relinearize-style
keyswitches through LimbShardedKeyswitch and rank 0 compares the gathered result with the single-GPU b2_keyswitch."""
import importlib, os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b2 = importlib.import_module("troy-nova_b200")
sh = importlib.import_module("troy-nova_b200.sharded")

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
small = "--small" in sys.argv
log_n = 13 if small else 16
bits = [60, 40, 40, 40, 60] if small else [60] + [40] * 19 + [60]
n = 1 << log_n
moduli = b2.coeff_modulus_create(n, bits)
K = len(moduli); L = K - 1
ctx = b2.Context(log_n, moduli, 0, b2.CKKS, device=local)
g = torch.Generator(device="cuda"); g.manual_seed(5)
key = torch.empty((L, 2, K, n), dtype=torch.int64, device="cuda")
for i in range(K):
    key[:, :, i, :] = torch.randint(0, moduli[i], (L, 2, n), dtype=torch.int64, device="cuda", generator=g)
ct = torch.empty((3, L, n), dtype=torch.int64, device="cuda")
for j in range(L):
    ct[:, j, :] = torch.randint(0, moduli[j], (3, n), dtype=torch.int64, device="cuda", generator=g)
kfull = ctx.kswitch_key(key)
# single-GPU result (every rank computes it; it is the checker here)
want = ct[:2].clone()
ctx.keyswitch(ct[2], L, True, kfull, want, b2.ASSIGN_ADD_INPLACE)
torch.cuda.synchronize()

ks = sh.LimbShardedKeyswitch(sh.DeviceStages(ctx), L)
f, c, nct = ks.first, ks.count, ks.nct
shard = ctx.kswitch_key_shard(kfull, L, f, c)
del kfull
tl = ct[2, f:f + nct].contiguous() if nct else torch.zeros(2, dtype=torch.int64, device="cuda")
dl0 = ct[:2, f:f + nct].contiguous() if nct else torch.zeros(2, dtype=torch.int64, device="cuda")
dl = dl0.clone()
ks(tl, True, shard, dl, b2.ASSIGN_ADD_INPLACE)
torch.cuda.synchronize()
ok = torch.tensor([1 if (nct == 0 or torch.equal(dl, want[:, f:f + nct])) else 0], device="cuda")
if world > 1:
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
# timing: max over ranks of CUDA-event time
def run():
    dl.copy_(dl0)
    ks(tl, True, shard, dl, b2.ASSIGN_ADD_INPLACE)
for _ in range(5): run()
if world > 1: dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 50
e0.record()
for _ in range(iters): run()
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device="cuda")
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("sharded relinearize: world=%d N=2^%d L=%d  bit-exact vs single-GPU: %s   %.1f us per op (max over ranks)  owned moduli %s"
          % (world, log_n, L, "YES" if int(ok.item()) == 1 else "NO", float(t.item()), [sh.partition(L + 1, world, r)[1] for r in range(world)]), flush=True)
# ---- the same keyswitch with the P2P data plane (b2_shard_push / b2_shard_wait over symmetric memory, weighted partition)
ok2 = torch.tensor([1], device="cuda")
if world > 1:
    kfull = ctx.kswitch_key(key)
    ex = sh.P2PExchange.from_symmetric_memory(L, n)
    kp = sh.P2PLimbShardedKeyswitch(ctx, L, ex)
    f, c, nct = kp.first, kp.count, kp.nct
    shard2 = ctx.kswitch_key_shard(kfull, L, f, c) if c else None
    del kfull
    tl2 = ct[2, f:f + nct].contiguous() if nct else torch.zeros(2, dtype=torch.int64, device="cuda")
    dl20 = ct[:2, f:f + nct].contiguous() if nct else torch.zeros(2, dtype=torch.int64, device="cuda")
    dl2 = dl20.clone()
    kp(tl2, True, shard2, dl2, b2.ASSIGN_ADD_INPLACE)
    torch.cuda.synchronize()
    ok2 = torch.tensor([1 if (nct == 0 or torch.equal(dl2, want[:, f:f + nct])) else 0], device="cuda")
    dist.all_reduce(ok2, op=dist.ReduceOp.MIN)
    def run2():
        dl2.copy_(dl20)
        kp(tl2, True, shard2, dl2, b2.ASSIGN_ADD_INPLACE)
    for _ in range(5): run2()
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(iters): run2()
    e1.record(); torch.cuda.synchronize()
    t2 = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device="cuda")
    dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("P2P sharded relinearize: world=%d  bit-exact: %s   %.1f us per op (max over ranks)  owned moduli %s (weighted)"
              % (world, "YES" if int(ok2.item()) == 1 else "NO", float(t2.item()), [p[1] for p in kp.parts]), flush=True)
    dist.barrier()
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if int(ok.item()) == 1 and int(ok2.item()) == 1 else 1)
