"""Developer timing of the forward / inverse transform across ring degrees (the work-queue lag is a number of ITEMS, tools note in
b2_ntt_launch.cuh): ~1 GiB of limbs per degree, FP64-path (40-bit) and integer-path (60-bit) moduli, CUDA events."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b2 = importlib.import_module("troy-nova_b200")
def timeit(fn, warm=2, it=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e-3
for logn in (12, 13, 14, 15, 16):
    n = 1 << logn
    for bits, name in (([40, 40, 40, 40], "40-bit"), ([60, 60, 60, 60], "60-bit")):
        moduli = b2.coeff_modulus_create(n, bits)
        ctx = b2.Context(logn, moduli, 0, b2.CKKS)
        pc = (1 << 27) // (4 * n)                    # 1 GiB of limbs
        x = torch.randint(0, 1 << 39, (pc * 4 * n,), dtype=torch.int64, device="cuda")
        tf = timeit(lambda: ctx.ntt_forward(x, pcount=pc, ncomp=4)); ti = timeit(lambda: ctx.ntt_inverse(x, pcount=pc, ncomp=4))
        L = pc * 4
        print(f"N=2^{logn} {name}: {L:7d} limbs  fwd {L/tf/1e6:7.2f} Mlimb/s ({L*n*16/tf/1e9:6.0f} GB/s)  inv {L/ti/1e6:7.2f} Mlimb/s ({L*n*16/ti/1e9:6.0f} GB/s)", flush=True)
        del x, ctx
