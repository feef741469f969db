"""Developer timing of the Encryptor / Decryptor pieces at C3 (CKKS N = 2^16, 20 limbs + special prime), CUDA events."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b2 = importlib.import_module("troy-nova_b200")
n = 1 << 16
moduli = b2.coeff_modulus_create(n, [60] + [40] * 19 + [60])
ctx = b2.Context(16, moduli, 0, b2.CKKS)
K, L = 21, 20
def timeit(fn, warm=3, it=30):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
g = b2.Prng(1, 2)
buf = torch.zeros(2 * K * n, dtype=torch.int64, device="cuda")
pk = torch.randint(0, 1 << 39, (2 * K * n,), dtype=torch.int64, device="cuda")
ct = torch.zeros(2 * K * n, dtype=torch.int64, device="cuda")
out = torch.zeros(L * n, dtype=torch.int64, device="cuda")
print("sample ternary  [21][N]      %7.1f us" % timeit(lambda: ctx.sample_poly(b2.SAMPLE_TERNARY, g, buf, K)))
print("sample cbd   [2][21][N]      %7.1f us" % timeit(lambda: ctx.sample_poly(b2.SAMPLE_CENTERED_BINOMIAL, g, buf, K, count=2)))
print("sample uniform  [21][N]      %7.1f us" % timeit(lambda: ctx.sample_poly(b2.SAMPLE_UNIFORM, g, buf, K)))
print("encrypt_zero_asymmetric key level (21 limbs, no switch) %7.1f us" % timeit(lambda: ctx.encrypt_zero_asymmetric(pk, K, True, ct, g)))
print("encrypt_zero_asymmetric first level (20 limbs)          %7.1f us" % timeit(lambda: ctx.encrypt_zero_asymmetric(pk, L, True, ct, g)))
print("encrypt_zero_symmetric  first level (20 limbs)          %7.1f us" % timeit(lambda: ctx.encrypt_zero_symmetric(pk, L, True, ct, g)))
print("decrypt (2 polynomials, 20 limbs)                       %7.1f us" % timeit(lambda: ctx.decrypt([ct], [out], 2, L, pk)))
print("mod_switch_to_next 21 -> 20 limbs, 2 polys              %7.1f us" % timeit(lambda: ctx.mod_switch_to_next(pk, ct, 2, K, b2.MODSWITCH_CKKS)))
