"""Developer timing of the persistent cluster / DSMEM transform against the queue kernels: FP64-path and mixed operands, CUDA events.
   B2_NTT_CLUSTER_MAX=1000000 routes every N = 2^16 launch through the cluster kernel; B2_NTT_CLUSTERS caps the resident clusters."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b2 = importlib.import_module("troy-nova_b200")
n = 1 << 16
moduli = b2.coeff_modulus_create(n, [60] + [40] * 19 + [60])
ctx = b2.Context(16, moduli, 0, b2.CKKS)
def timeit(fn, warm=3, it=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e-3
tag = f"cluster_max={os.environ.get('B2_NTT_CLUSTER_MAX','-')} clusters={os.environ.get('B2_NTT_CLUSTERS','-')}"
for pc, nc, tb in ((20, 19, 1), (20, 21, 0), (1, 20, 0), (1, 21, 0), (3, 21, 0), (200, 21, 0)):
    x = torch.randint(0, 1 << 39, (pc * nc * n,), dtype=torch.int64, device="cuda")
    ref = None
    tf = timeit(lambda: ctx.ntt_forward(x, pcount=pc, ncomp=nc, table_base=tb)); ti = timeit(lambda: ctx.ntt_inverse(x, pcount=pc, ncomp=nc, table_base=tb))
    L = pc * nc
    print(f"{tag} {L:5d} limbs ({'FP64' if tb else 'mixed'}) fwd {tf*1e6:8.1f} us ({L/tf/1e6:6.3f} Mlimb/s)  inv {ti*1e6:8.1f} us ({L/ti/1e6:6.3f} Mlimb/s)", flush=True)
