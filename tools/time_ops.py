"""On-box timing of the non-NTT hot-path kernels against their HBM floors (developer tool, not the bench).
C3 = CKKS N=2^16 L=20(+1); C2 = BFV N=2^14 L=3(+1)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b2 = importlib.import_module("troy-nova_b200")
PEAK = 6.55e12

def timeit(fn, warm=3, it=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e-3

def rnd(moduli, prefix, n):
    x = torch.empty(tuple(prefix) + (len(moduli), n), dtype=torch.int64, device="cuda")
    for j, q in enumerate(moduli):
        x[..., j, :] = torch.randint(0, q, tuple(prefix) + (n,), dtype=torch.int64, device="cuda")
    return x

def report(name, t, nbytes):
    print(f"{name:42s} {t*1e6:9.1f} us   {nbytes/1e6:8.1f} MB algorithmic   {nbytes/t/1e9:7.0f} GB/s = {nbytes/t/PEAK*100:5.1f} % of HBM peak")

# ---- C3
log_n, n = 16, 1 << 16
moduli = b2.coeff_modulus_create(n, [60] + [40] * 19 + [60]); K = len(moduli); L = K - 1
ctx = b2.Context(log_n, moduli, 0, b2.CKKS)
a = rnd(moduli[:L], (2,), n); b = rnd(moduli[:L], (2,), n); m = torch.empty((3, L, n), dtype=torch.int64, device="cuda")
W = L * n * 8
report("C3 dyadic_convolute (ct x ct)", timeit(lambda: ctx.dyadic_convolute(a, 2, b, 2, m, L)), 7 * W)
report("C3 dyadic_square", timeit(lambda: ctx.dyadic_square(a, m, L)), 5 * W)
o = torch.empty_like(a)
report("C3 add", timeit(lambda: ctx.add(a, b, o, 2, L)), 6 * W)
p = rnd(moduli[:L], (), n)
report("C3 multiply_plain (ntt)", timeit(lambda: ctx.dyadic_broadcast(a, p, o, 2, L)), 5 * W)
report("C3 galois (ntt form, 2 polys)", timeit(lambda: ctx.galois(a, 3, o, 2, L, True)), 4 * W)
r = torch.empty((2, L - 1, n), dtype=torch.int64, device="cuda")
report("C3 rescale_to_next (2 + 38 transforms)", timeit(lambda: ctx.mod_switch_to_next(a, r, 2, L, b2.MODSWITCH_CKKS)), 8 * n * (2 * L + 2 * (L - 1)))
ac = a.clone()
report("C3 transform_to_ntt (40 limbs)", timeit(lambda: ctx.ntt_forward(ac, pcount=2, ncomp=L)), 2 * 2 * W)
del ctx
# ---- C2 BFV multiply
log_n, n = 14, 1 << 14
moduli = b2.coeff_modulus_create(n, [60, 40, 40, 60]); K = len(moduli); L = K - 1
ctx = b2.Context(log_n, moduli, 786433, b2.BFV)
for B in (1, 16, 256):
    a = rnd(moduli[:L], (B, 2), n); b = rnd(moduli[:L], (B, 2), n); m = torch.empty((B, 3, L, n), dtype=torch.int64, device="cuda")
    t = timeit(lambda: ctx.bfv_multiply(a, b, m, L, batch=B), it=5)
    print(f"C2 bfv_multiply (BEHZ) batch {B:4d}             {t*1e6:9.1f} us   {t/B*1e6:8.1f} us per ciphertext")
