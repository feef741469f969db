"""Developer timing of one C3 keyswitch (relinearize shape), CUDA events, data resident."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b2 = importlib.import_module("troy-nova_b200")
n = 1 << 16
moduli = b2.coeff_modulus_create(n, [60] + [40] * 19 + [60]); K = len(moduli); L = K - 1
ctx = b2.Context(16, moduli, 0, b2.CKKS)
g = torch.Generator(device="cuda"); g.manual_seed(1)
key = torch.empty((L, 2, K, n), dtype=torch.int64, device="cuda")
for i in range(K):
    key[:, :, i, :] = torch.randint(0, moduli[i], (L, 2, n), dtype=torch.int64, device="cuda", generator=g)
kh = ctx.kswitch_key(key)
ct = torch.empty((3, L, n), dtype=torch.int64, device="cuda")
for j in range(L):
    ct[:, j, :] = torch.randint(0, moduli[j], (3, n), dtype=torch.int64, device="cuda", generator=g)
dest = ct[:2].clone()
def f(): ctx.keyswitch(ct[2], L, True, kh, dest, b2.ASSIGN_ADD_INPLACE)
for _ in range(5): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): f()
e1.record(); torch.cuda.synchronize()
print("keyswitch C3: %.1f us" % (e0.elapsed_time(e1) / 50 * 1e3), {k: v for k, v in os.environ.items() if k.startswith("B2_")})
