"""Developer timing of b2_pack_rlwes at the C5 shape (BFV N = 8192, 3 limbs + special prime): 64 -> 1 with the trace, 1 / 16 / 128 groups."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b2 = importlib.import_module("troy-nova_b200")
log_n, n = 13, 1 << 13
moduli = b2.coeff_modulus_create(n, [60, 40, 40, 60]); K = len(moduli); L = K - 1
ctx = b2.Context(log_n, moduli, 1032193, b2.BFV)
elts = sorted({(n >> k) + 1 for k in range(0, log_n)} | {2 ** k + 1 for k in range(1, log_n + 1)})
keys = {e: ctx.kswitch_key(torch.randint(0, 1 << 39, (L, 2, K, n), dtype=torch.int64, device="cuda")) for e in elts if e < 2 * n}
def timeit(fn, warm=2, it=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
M = 64
for G in (1, 16, 128):
    cts = torch.randint(0, 1 << 39, (G * M, 2 * L * n), dtype=torch.int64, device="cuda")
    groups = [[cts[g * M + i] for i in range(M)] for g in range(G)]
    outs = [torch.empty(2 * L * n, dtype=torch.int64, device="cuda") for _ in range(G)]
    before = b2.launch_count()
    ctx.pack_rlwes(groups, False, L, 0, n, n // M, keys, outs)
    launches = b2.launch_count() - before
    t = timeit(lambda: ctx.pack_rlwes(groups, False, L, 0, n, n // M, keys, outs))
    print(f"{G:4d} groups of {M} -> 1 with trace: {t:9.1f} us total, {t / G:8.1f} us per group, {launches} launches", flush=True)
