"""Developer timing of b2_multiply_plain_accumulate at the BumbleBee 1024^3 shape (C5: N = 8192, 3 ciphertext limbs): I x J x B terms,
term (i, j, b) = ct a[b][i] x pt w[i][j] -> dest r[b][j]."""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b2 = importlib.import_module("troy-nova_b200")
n = 8192
moduli = b2.coeff_modulus_create(n, [60, 40, 40, 60])
ctx = b2.Context(13, moduli, 0, b2.CKKS)
L = 3
I, J, B = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (64, 256, 8)))
a = torch.randint(0, 1 << 39, (B * I, 2 * L * n), dtype=torch.int64, device="cuda")
w = torch.randint(0, 1 << 39, (I * J, L * n), dtype=torch.int64, device="cuda")
r = torch.zeros((B * J, 2 * L * n), dtype=torch.int64, device="cuda")
cts, pts, dests = [], [], []
for i in range(I):
    for j in range(J):
        for b in range(B):
            cts.append(a[b * I + i].data_ptr()); pts.append(w[i * J + j].data_ptr()); dests.append(r[b * J + j].data_ptr())
print("terms", len(cts), "destinations", B * J, "pt GB", w.numel() * 8 / 1e9, "ct MB", a.numel() * 8 / 1e6, flush=True)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.multiply_plain_accumulate(cts, pts, dests, 2, L, add_to_dest=False)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("call returned after %.1f ms, device done after %.1f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3), flush=True)
