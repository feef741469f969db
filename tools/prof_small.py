"""Small-launch transforms (1, 2, 20, 40 limbs at N = 2^16) for ncu duration lists (developer tool)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b2 = importlib.import_module("troy-nova_b200")
log_n = 16; n = 1 << log_n
moduli = b2.coeff_modulus_create(n, [60] + [40] * 19 + [60])
ctx = b2.Context(log_n, moduli, 0, b2.CKKS)
for pc, nc in ((1, 1), (2, 1), (1, 20), (2, 20)):
    x = torch.randint(0, 1 << 39, (pc * nc * n,), dtype=torch.int64, device="cuda")
    for _ in range(3):
        ctx.ntt_forward(x, pcount=pc, ncomp=nc, table_base=1 if nc == 1 else 0)
        ctx.ntt_inverse(x, pcount=pc, ncomp=nc, table_base=1 if nc == 1 else 0)
torch.cuda.synchronize()
