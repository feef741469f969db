"""Two C3 keyswitches (CKKS N=2^16, L=20) for `ncu --set full` captures: the first warms, the second is profiled."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b2 = importlib.import_module("troy-nova_b200")
log_n = 16; n = 1 << log_n
moduli = b2.coeff_modulus_create(n, [60] + [40] * 19 + [60])
ctx = b2.Context(log_n, moduli, 0, b2.CKKS)
K = len(moduli); L = K - 1
g = torch.Generator(device="cuda"); g.manual_seed(1)
key = torch.empty((L, 2, K, n), dtype=torch.int64, device="cuda")
for i in range(K):
    key[:, :, i, :] = torch.randint(0, moduli[i], (L, 2, n), dtype=torch.int64, device="cuda", generator=g)
kh = ctx.kswitch_key(key)
ct = torch.empty((3, L, n), dtype=torch.int64, device="cuda")
for j in range(L):
    ct[:, j, :] = torch.randint(0, moduli[j], (3, n), dtype=torch.int64, device="cuda", generator=g)
dest = ct[:2].clone()
torch.cuda.synchronize()
l0 = b2.launch_count()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    ctx.keyswitch(ct[2], L, True, kh, dest, b2.ASSIGN_ADD_INPLACE)
torch.cuda.synchronize()
print("launches per keyswitch:", (b2.launch_count() - l0) // (int(sys.argv[1]) if len(sys.argv) > 1 else 2))
