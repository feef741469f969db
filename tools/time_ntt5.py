"""Developer timing of the TMA NTT kernel alone: 380 FP64-path limbs (19 moduli x 20), CUDA events."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b2 = importlib.import_module("troy-nova_b200")
n = 1 << 16
moduli = b2.coeff_modulus_create(n, [40] * 19 + [60])
ctx = b2.Context(16, moduli, 0, b2.CKKS)
def timeit(fn, warm=3, it=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e-3
for pc in (20, 200):
    x = torch.randint(0, 1 << 39, (pc * 19 * n,), dtype=torch.int64, device="cuda")
    tf = timeit(lambda: ctx.ntt_forward(x, pcount=pc, ncomp=19)); ti = timeit(lambda: ctx.ntt_inverse(x, pcount=pc, ncomp=19))
    L = pc * 19
    print(f"dbg={os.environ.get('B2_NTT5_DBG','0')} {L:5d} FP64 limbs fwd {tf*1e6:8.1f} us ({L/tf/1e6:6.3f} Mlimb/s)  inv {ti*1e6:8.1f} us ({L/ti/1e6:6.3f} Mlimb/s)")
