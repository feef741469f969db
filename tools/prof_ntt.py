import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b2 = importlib.import_module("troy-nova_b200")
log_n = 16; n = 1 << log_n
moduli = b2.coeff_modulus_create(n, [60] + [40] * 19 + [60])
ctx = b2.Context(log_n, moduli, 0, b2.CKKS)
pc, nc = 20, 21
x = torch.randint(0, 1 << 39, (pc * nc * n,), dtype=torch.int64, device="cuda")
for _ in range(3):
    ctx.ntt_forward(x, pcount=pc, ncomp=nc)
    ctx.ntt_inverse(x, pcount=pc, ncomp=nc)
torch.cuda.synchronize()
