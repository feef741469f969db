"""e2e (host-buffer) NTT throughput vs the host link's duplex copy rate (developer tool)."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b2 = importlib.import_module("troy-nova_b200")
n = 1 << 16
moduli = b2.coeff_modulus_create(n, [60] + [40] * 19 + [60])
ctx = b2.Context(16, moduli, 0, b2.CKKS)
pc, nc = 20, 21
hb = torch.randint(0, 1 << 39, (pc, nc, n), dtype=torch.int64).pin_memory()
h = hb.numpy().view(np.uint64).reshape(-1)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    ctx.ntt_host(True, h, h, pcount=pc, ncomp=nc, stream=st); ctx.ntt_host(False, h, h, pcount=pc, ncomp=nc, stream=st)
torch.cuda.synchronize(); t0 = time.perf_counter(); k = 6
for _ in range(k):
    ctx.ntt_host(True, h, h, pcount=pc, ncomp=nc, stream=st); ctx.ntt_host(False, h, h, pcount=pc, ncomp=nc, stream=st)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("chunk %s MiB: %.1f k limbs/s, %.1f GB/s per direction" % (os.environ.get("B2_HOST_CHUNK_MB", "4"), 2 * k * pc * nc / dt / 1e3, 2 * k * pc * nc * n * 8 / dt / 1e9))
