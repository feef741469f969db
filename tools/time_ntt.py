"""Quick on-box timing of libb2rns NTT vs the reference's device branch (developer tool, not the bench)."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
b2 = importlib.import_module("troy-nova_b200")
import pyoracle as po

def timeit(fn, warm=3, it=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e-3

log_n = 16; n = 1 << log_n
bits = [60] + [40] * 19 + [60]
moduli = b2.coeff_modulus_create(n, bits)
ctx = b2.Context(log_n, moduli, 0, b2.CKKS)
for (pc, nc, label) in [(1, 1, "1 limb"), (1, 21, "21 limbs"), (20, 21, "420 limbs"), (200, 21, "4200 limbs")]:
    x = torch.randint(0, 1 << 39, (pc * nc * n,), dtype=torch.int64, device="cuda")
    tf = timeit(lambda: ctx.ntt_forward(x, pcount=pc, ncomp=nc))
    ti = timeit(lambda: ctx.ntt_inverse(x, pcount=pc, ncomp=nc))
    L = pc * nc
    print(f"b2 NTT {label:12s} fwd {tf*1e6:9.1f} us ({L/tf/1e6:6.3f} Mlimb/s, {L*n*16/tf/1e9:7.1f} GB/s)   inv {ti*1e6:9.1f} us ({L/ti/1e6:6.3f} Mlimb/s)")
# keyswitch-mode call: [21][20][N] set-products
x = torch.randint(0, 1 << 39, (21 * 20 * n,), dtype=torch.int64, device="cuda")
tf = timeit(lambda: ctx.ntt_forward(x, pcount=21, ncomp=20, mode=b2.TABLES_KS_SET_PRODUCTS, decomp=20))
print(f"b2 NTT set-products 420 limbs fwd {tf*1e6:9.1f} us ({420/tf/1e6:6.3f} Mlimb/s)")
if po.ref_available() and "--ref" in sys.argv:
    r = po.Reference(po.CKKS, n, bits, seed=0x123, device=True)
    for pc in (1, 20):
        t0 = r.time_op(0, 21, pc, 2, 10); t1 = r.time_op(1, 21, pc, 2, 10)
        print(f"reference device NTT {pc*21} limbs: fwd {t0*1e6:.1f} us ({pc*21/t0/1e6:.3f} Mlimb/s) inv {t1*1e6:.1f} us")
    r.gen_relin(); g = r.galois_element_from_step(1); r.gen_galois([g])
    for op, nm in ((2, "multiply"), (3, "relinearize"), (4, "rotate"), (5, "rescale")):
        t = r.time_op(op, 20, 1, 2, 10)
        print(f"reference device {nm}: {t*1e6:.1f} us")
