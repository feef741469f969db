#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native RNS hot path (contract: see the task statement).

Metric (BASELINE.json): NTT limbs/s at N = 2^16, on the CKKS N=65536 / 21-moduli parameter set (config C3);
relinearize / rotate ops/s on the same parameter set are reported in the same JSON line under "extra".

A "step" is one pass of the hot path over one batch of synthetic input: a forward negacyclic NTT followed
by an inverse NTT over a key-level operand of 20 x 21 = 420 limbs (210 MiB, larger than the 126 MB L2, so
every step streams from HBM).  840 limb-transforms per step per GPU.

  value  : limb-transforms/s with the operand already resident in HBM (CUDA events, max over ranks)
  e2e    : the same metric through the host-buffer C-ABI entry points b2_ntt_forward_host /
           b2_ntt_inverse_host (pinned host memory -> H2D -> kernels -> D2H inside the timed region)
  roofline: algorithmic bytes (16 N per limb-transform) / CUDA-event duration of the
           transform launches, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline: the reference's own host (CPU) NTT, timed on this box's cores through oracle/_ref

Multi-GPU (torchrun, one rank per GPU): the batch is sharded by ciphertext (independent limbs, no data-path
collective) -> weak scaling; NCCL is only used for the barrier and the max-over-ranks of the timings.

The second half of BASELINE's metric — relinearize / rotate ops/s on the same parameter set — is reported as TOP-LEVEL
blocks "relinearize" and "rotate", each with its own value (data resident), roofline (492.8 MB compulsory bytes per keyswitch,
SURVEY.md §8d), cpu_baseline (the reference's host Evaluator::relinearize / rotate_vector through oracle/_ref, threads stated)
and e2e (b2_keyswitch_host: ciphertext in pinned host memory, key resident).  "gpu_baseline" times the reference's own DEVICE
branch (its generic kernels recompiled for sm_100, oracle/_ref) in the same run: the bar SURVEY.md §8d sets.

--impl reference : times the reference's CPU implementation of the same path (oracle/_ref when built, else
the C oracle port) on all host cores, bounded sample per step.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

LOG_N = 16
BITS = [60] + [40] * 19 + [60]          # CKKS C3: 20 ciphertext limbs + special prime
PCOUNT, NCOMP = 20, 21                   # one key-level operand [20][21][N] = 420 limbs = 210 MiB
METRIC = "ntt_limbs_per_s_n65536"
UNIT = "limbs/s"


def measured_traffic():
    """dram__bytes_read + dram__bytes_write of one forward launch of the workload (ncu --set full capture of the same
    command, committed under profiles/): {"bytes_per_launch": ..., "source": ...} or None."""
    p = os.path.join(ROOT, "profiles", "ntt_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def ks_traffic(name):
    """dram bytes of one C3 keyswitch from the committed ncu launch list (profiles/ks_traffic.json), or None"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ks_traffic.json")))["bytes_per_keyswitch"]
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
def cpu_reference_ntt(seconds_budget=12.0, threads=None):
    """Reference host NTT (oracle/_ref) on `threads` cores: limb-transforms/s.  Falls back to the C oracle port."""
    import numpy as np
    import pyoracle as po
    n = 1 << LOG_N
    threads = threads or min(os.cpu_count() or 1, 64)
    limbs_per_call = 6                      # fwd over [2][3][N] per call, then inv
    if po.ref_available():
        kind = "reference"
        results = [None] * threads

        def work(i):
            r = po.Reference(po.CKKS, n, [60, 40, 40, 60], seed=0x123 + i)
            r.time_op(0, 3, 2, 0, 1)       # warm
            t0 = time.perf_counter(); calls = 0; op_s = 0.0
            while time.perf_counter() - t0 < seconds_budget:
                op_s += r.time_op(0, 3, 2, 0, 1) + r.time_op(1, 3, 2, 0, 1); calls += 2
            results[i] = (calls * limbs_per_call, op_s)
        ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
        [t.start() for t in ths]; [t.join() for t in ths]
        total = sum(r[0] for r in results); wall = max(r[1] for r in results)
        rate = sum(r[0] / r[1] for r in results)
    else:
        kind = "port"
        os.environ["RO_THREADS"] = str(threads)
        moduli = po.coeff_modulus_create(n, [60, 40, 40, 60])
        o = po.Oracle(LOG_N, moduli, 0, po.CKKS)
        rng = np.random.default_rng(1)
        x = o.random_poly(rng, threads, 3)
        t0 = time.perf_counter(); total = 0
        while time.perf_counter() - t0 < seconds_budget:
            o.ntt_fwd(x, threads, 3); o.ntt_inv(x, threads, 3); total += 2 * threads * 3
        wall = time.perf_counter() - t0
        rate = total / wall
    return {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": "%d limb-transforms (N=2^16, fwd+inv, 60/40-bit primes) over %.1f s on %d threads" % (total, wall, threads)}


def cpu_reference_keyswitch(seconds_budget=6.0, threads=None):
    """Reference host Evaluator::relinearize and rotate_vector(1) at C3 (CKKS N=65536, 20+1 moduli) through oracle/_ref:
    one shared context (keys are read-only), `threads` host threads each timing its own ciphertexts.  ops/s per op kind."""
    import pyoracle as po
    if not po.ref_available():
        return None
    n = 1 << LOG_N
    threads = threads or min(os.cpu_count() or 1, 16)
    r = po.Reference(po.CKKS, n, BITS, seed=0x123)
    r.gen_relin()
    r.gen_galois([r.galois_element_from_step(1)])
    out = {}
    for op, name in ((3, "relinearize"), (4, "rotate")):
        res = [None] * threads

        def work(i):
            t0 = time.perf_counter(); k = 0; busy = 0.0
            while time.perf_counter() - t0 < seconds_budget or k == 0:
                busy += r.time_op(op, NCOMP - 1, 1, 0, 1); k += 1
            res[i] = (k, busy)
        ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
        [t.start() for t in ths]; [t.join() for t in ths]
        out[name] = {"value": sum(k / b for k, b in res), "unit": "ops/s", "cores": threads, "kind": "reference",
                     "sample": "%d x Evaluator::%s on the host branch (CKKS N=65536, 20 limbs + special prime, reference KeyGenerator keys, seed 0x123) "
                               "on %d threads, %.2f s per op per thread" % (sum(k for k, _ in res), "relinearize_new" if op == 3 else "rotate_vector_new(1)", threads,
                                                                             sum(b for _, b in res) / max(1, sum(k for k, _ in res)))}
    return out


def gpu_reference_baseline():
    """The reference's DEVICE branch (oracle/_ref = its generic kernels recompiled for sm_100) on this GPU, data resident:
    the GPU bar of SURVEY.md §8d.  Timed by the shim with cudaDeviceSynchronize around each op."""
    import pyoracle as po
    if not po.ref_available():
        return None
    n = 1 << LOG_N
    r = po.Reference(po.CKKS, n, BITS, seed=0x123, device=True)
    r.gen_relin()
    r.gen_galois([r.galois_element_from_step(1)])
    tf = r.time_op(0, NCOMP, PCOUNT, 2, 10); ti = r.time_op(1, NCOMP, PCOUNT, 2, 10)
    g = {"kind": "reference device branch (sm_100 recompile of the reference kernels, oracle/_ref)",
         "ntt_limbs_per_s": 2 * PCOUNT * NCOMP / (tf + ti), "ntt_forward_us": tf * 1e6, "ntt_inverse_us": ti * 1e6}
    for op, name in ((3, "relinearize"), (4, "rotate"), (2, "multiply"), (5, "rescale"), (7, "encrypt_zero_asymmetric"), (8, "decrypt")):
        t = r.time_op(op, NCOMP - 1, 1, 2, 10)
        g[name + "_us"] = t * 1e6; g[name + "_ops_per_s"] = 1.0 / t
    del r
    try:      # C5 packing on the reference's device branch: 64 BFV ciphertexts at N = 8192 into one, with the field trace
        n5 = 1 << 13
        r5 = po.Reference(po.BFV, n5, [60, 40, 40, 60], t_bits=20, seed=0x123, device=True)
        elts = sorted({(n5 >> k) + 1 for k in range(0, 13)} | {2 ** k + 1 for k in range(1, 14)})
        r5.gen_galois([e for e in elts if e < 2 * n5])
        t = r5.time_op(9, 3, 64, 1, 3)
        g["c5_pack_64_to_1_us"] = t * 1e6
    except Exception as ex:
        g["c5_pack_error"] = str(ex)[:120]
    return g


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    vals = []
    for s in range(args.warmup + args.steps):
        r = cpu_reference_ntt(seconds_budget=per_step)
        if s >= args.warmup:
            vals.append(r)
    v = sum(x["value"] for x in vals) / len(vals)
    cb = dict(vals[-1]); cb["value"] = v
    ks = None
    try:
        ks = cpu_reference_keyswitch(seconds_budget=min(8.0, per_step))
    except Exception as ex:                      # the NTT line stands on its own
        ks = {"error": str(ex)[:200]}
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "C3 CKKS N=65536, q={60,40x19,60}: forward+inverse negacyclic NTT of a [20][21][N] operand (420 limbs, 210 MiB) per GPU per step",
                       "sample": "reference host branch (utils::ntt_inplace_ps on host slices), %.1f s of [2][3][N] forward+inverse calls per step on every host thread, 60/40-bit primes" % per_step},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "relinearize": (ks or {}).get("relinearize"), "rotate": (ks or {}).get("rotate")}
    print(json.dumps(line), flush=True)


def _rand_limbs(torch, moduli, shape_prefix, n, gen):
    """uniform residues [*shape_prefix][len(moduli)][n]"""
    x = torch.empty(tuple(shape_prefix) + (len(moduli), n), dtype=torch.int64, device="cuda")
    for j, q in enumerate(moduli):
        x[..., j, :] = torch.randint(0, q, tuple(shape_prefix) + (n,), dtype=torch.int64, device="cuda", generator=gen)
    return x


def bench_c2(b2, torch, timed, local, world):
    """BASELINE config 2: BFV N=16384, q={60,40,40}+special 60, 256 ciphertexts: batched NTT / INTT, multiply_plain
    (NTT form), rotate_rows (Galois automorphism + batched keyswitch, 16 ciphertexts per launch set)."""
    log_n, n, B = 14, 1 << 14, 256
    moduli = b2.coeff_modulus_create(n, [60, 40, 40, 60]); K = len(moduli); L = K - 1
    ctx = b2.Context(log_n, moduli, 786433, b2.BFV, device=local)   # t = 3 * 2^18 + 1 = 1 mod 2N
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    cts = _rand_limbs(torch, moduli[:L], (B, 2), n, g)
    plain = _rand_limbs(torch, moduli[:L], (B,), n, g)          # one plaintext per ciphertext (multiply_plain_batched)
    key = torch.empty((L, 2, K, n), dtype=torch.int64, device="cuda")
    for i in range(K):
        key[:, :, i, :] = torch.randint(0, moduli[i], (L, 2, n), dtype=torch.int64, device="cuda", generator=g)
    kh = ctx.kswitch_key(key)
    out = torch.empty_like(cts); rot = torch.empty_like(cts)
    r = {}
    t = timed(lambda: ctx.ntt_forward(cts, batch=B, pcount=2, ncomp=L), 5); r["ntt_batched_limbs_per_s"] = world * B * 2 * L / t
    t = timed(lambda: ctx.ntt_inverse(cts, batch=B, pcount=2, ncomp=L), 5); r["intt_batched_limbs_per_s"] = world * B * 2 * L / t
    t = timed(lambda: ctx.dyadic_broadcast_batched(cts, plain, out, B, 2, L), 5); r["multiply_plain_ops_per_s"] = world * B / t
    gal = 3

    def rotate_rows():
        ctx.galois(cts, gal, rot, 2 * B, L, False)          # coefficient-form automorphism of all 2 B polynomials
        ctx.keyswitch_batched([rot[b, 1] for b in range(B)], L, False, kh, [rot[b] for b in range(B)], b2.ASSIGN_OVERWRITE_EXCEPT_FIRST)
    t = timed(rotate_rows, 3); r["rotate_rows_ops_per_s"] = world * B / t; r["rotate_rows_us_per_op"] = t / B * 1e6
    return r


def bench_c5_pack(b2, torch, timed, local, world):
    """BASELINE config 5 (BumbleBee output packing): BFV N=8192, q={60,40,40}+special 60: pack 64 ciphertexts into one with the
    field trace (Evaluator::pack_rlwe_ciphertexts_new, input_interval N, output_interval N/64), and 16 such groups side by side in one call."""
    log_n, n = 13, 1 << 13
    moduli = b2.coeff_modulus_create(n, [60, 40, 40, 60]); K = len(moduli); L = K - 1
    ctx = b2.Context(log_n, moduli, 1032193, b2.BFV, device=local)    # a 20-bit prime = 1 mod 2N
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    elts = sorted({(n >> k) + 1 for k in range(0, log_n)} | {2 ** k + 1 for k in range(1, log_n + 1)})
    elts = [e for e in elts if e < 2 * n]
    keys = {}
    for e in elts:
        key = torch.empty((L, 2, K, n), dtype=torch.int64, device="cuda")
        for i in range(K):
            key[:, :, i, :] = torch.randint(0, moduli[i], (L, 2, n), dtype=torch.int64, device="cuda", generator=g)
        keys[e] = ctx.kswitch_key(key)
    G, M = 16, 64
    cts = _rand_limbs(torch, moduli[:L], (G * M, 2), n, g)
    outs = [torch.empty((2, L, n), dtype=torch.int64, device="cuda") for _ in range(G)]
    r = {}
    t = timed(lambda: ctx.pack_rlwes([cts[i] for i in range(M)], False, L, 0, n, n // M, keys, outs[0]), 5)
    r["pack_64_to_1_us"] = t * 1e6; r["pack_64_to_1_ops_per_s"] = world / t
    t = timed(lambda: ctx.pack_rlwes([[cts[gi * M + i] for i in range(M)] for gi in range(G)], False, L, 0, n, n // M, keys, outs), 3)
    r["pack_16_groups_us_per_group"] = t / G * 1e6; r["pack_groups_per_s"] = world * G / t
    return r


def bench_c1_quickstart(b2, torch, timed, local, world):
    """BASELINE config 1 (the reference's quickstart, which it runs in host memory): BFV N=8192, q={40,40}+special 40, t 20 bits — BatchEncoder encode x2,
    asymmetric encrypt x2, BEHZ multiply, relinearize, decrypt, decode, natively on the GPU with keys generated from a seed."""
    log_n, n = 13, 1 << 13
    moduli = b2.coeff_modulus_create(n, [40, 40, 40]); K = len(moduli); L = K - 1
    ctx = b2.Context(log_n, moduli, 1032193, b2.BFV, device=local)
    z = lambda w: torch.zeros(w, dtype=torch.int64, device="cuda")
    g = b2.Prng(0x123, 0)
    sk, pk, skp, rk = z(K * n), z(2 * K * n), z(2 * K * n), z(L * 2 * K * n)
    ctx.generate_secret_key(g, sk); ctx.encrypt_zero_symmetric(sk, K, True, pk, g); ctx.secret_key_powers(sk, 2, skp)
    ctx.generate_kswitch_key(sk, skp[K * n:], g, rk)
    kh = ctx.kswitch_key(rk.view(L, 2, K, n))
    x = torch.randint(0, 1032193, (n,), dtype=torch.int64, device="cuda"); y = torch.randint(0, 1032193, (n,), dtype=torch.int64, device="cuda")
    px, py, cx, cy, prod, dec, out = z(n), z(n), z(2 * L * n), z(2 * L * n), z(3 * L * n), z(n), z(n)

    def pipeline():
        ctx.batch_encode(x, n, px); ctx.batch_encode(y, n, py)
        ctx.encrypt(px, n, pk, True, L, cx, g); ctx.encrypt(py, n, pk, True, L, cy, g)
        ctx.bfv_multiply(cx, cy, prod, L)
        ctx.keyswitch(prod[2 * L * n:], L, False, kh, prod[:2 * L * n], b2.ASSIGN_ADD_INPLACE)
        ctx.decrypt([prod[:2 * L * n]], [dec], 2, L, skp)
        ctx.batch_decode(dec, n, out)
    t = timed(pipeline, 10)
    ok = bool(torch.equal(out, (x * y) % 1032193))
    return {"pipeline_us": t * 1e6, "pipelines_per_s": world / t, "decoded_equals_slotwise_product": ok}


def bench_c4(b2, torch, timed, local, world):
    """BASELINE config 4: BGV N=32768, q={60,40x7}+special 60: multiply + relinearize, then the mod-switch chain to the last level."""
    log_n, n = 15, 1 << 15
    moduli = b2.coeff_modulus_create(n, [60] + [40] * 7 + [60]); K = len(moduli); L = K - 1
    ctx = b2.Context(log_n, moduli, 786433, b2.BGV, device=local)
    g = torch.Generator(device="cuda"); g.manual_seed(9)
    a = _rand_limbs(torch, moduli[:L], (2,), n, g); b = _rand_limbs(torch, moduli[:L], (2,), n, g)
    key = torch.empty((L, 2, K, n), dtype=torch.int64, device="cuda")
    for i in range(K):
        key[:, :, i, :] = torch.randint(0, moduli[i], (L, 2, n), dtype=torch.int64, device="cuda", generator=g)
    kh = ctx.kswitch_key(key)
    m = torch.empty((3, L, n), dtype=torch.int64, device="cuda"); d = torch.empty((2, L, n), dtype=torch.int64, device="cuda")
    r = {}
    t = timed(lambda: ctx.dyadic_convolute(a, 2, b, 2, m, L), 10); r["multiply_ops_per_s"] = world / t

    def relin():
        d.copy_(m[:2]); ctx.keyswitch(m[2], L, True, kh, d, b2.ASSIGN_ADD_INPLACE)
    t = timed(relin, 10); r["relinearize_ops_per_s"] = world / t; r["relinearize_us"] = t * 1e6
    bufs = [torch.empty((2, l, n), dtype=torch.int64, device="cuda") for l in range(L, 0, -1)]
    bufs[0].copy_(d)

    def chain():
        for i in range(L - 1):
            ctx.mod_switch_to_next(bufs[i], bufs[i + 1], 2, L - i, b2.MODSWITCH_BGV)
    t = timed(chain, 5); r["modswitch_chain_8_to_1_us"] = t * 1e6
    return r


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b2", choices=["b2", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b2" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    b2 = importlib.import_module("troy-nova_b200")

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — libb2rns has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = 1 << LOG_N
    moduli = b2.coeff_modulus_create(n, BITS)
    ctx = b2.Context(LOG_N, moduli, 0, b2.CKKS, device=local)
    limbs = PCOUNT * NCOMP
    g = torch.Generator(device="cuda"); g.manual_seed(1 + rank)
    # uniform residues below each limb's modulus
    x = torch.empty((PCOUNT, NCOMP, n), dtype=torch.int64, device="cuda")
    for j in range(NCOMP):
        x[:, j, :] = torch.randint(0, moduli[j], (PCOUNT, n), dtype=torch.int64, device="cuda", generator=g)
    x0 = x.clone()

    def step():
        ctx.ntt_forward(x, pcount=PCOUNT, ncomp=NCOMP)
        ctx.ntt_inverse(x, pcount=PCOUNT, ncomp=NCOMP)

    for _ in range(args.warmup):
        step()
    barrier()
    assert torch.equal(x, x0), "inverse(forward(x)) != x"       # round-trip sanity on the real workload
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = b2.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = b2.launch_count() - l0
    sec = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    if sampler:
        sampler.stop_flag = True; sampler.join(timeout=2)
    value = world * args.steps * 2 * limbs / sec

    # ---- roofline of the dominant kernels: forward transform launches alone, CUDA events on the launch stream
    ef0, ef1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(3, args.steps // 2)
    torch.cuda.synchronize(); ef0.record()
    for _ in range(reps):
        ctx.ntt_forward(x, pcount=PCOUNT, ncomp=NCOMP)
    ef1.record(); torch.cuda.synchronize()
    fwd_s = ef0.elapsed_time(ef1) * 1e-3 / reps
    peak, peak_src = measured_peaks()
    achieved = limbs * n * 16 / fwd_s / 1e9
    tr = measured_traffic()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": tr["bytes_per_launch"] if tr else None, "traffic_source": tr["source"] if tr else None,
                "kernel": "b2_ntt_forward over 420 limbs = ONE launch of ntt_persistent_kernel<8,true,3,PlainLoad,PlainStore> (the operand mixes 380 FP64-pipe and 40 integer-pipe limbs; "
                          "mixed launches take the cp.async queue kernel, all-FP64 launches the TMA kernel ntt5_kernel: csrc/b2_ntt5_launch.cuh run_ntt_split)",
                "algorithmic_bytes_per_call": limbs * n * 16, "call_ms": fwd_s * 1e3, "peak_source": peak_src}

    # ---- end to end through the host-buffer C ABI (pinned host memory, H2D + D2H inside the timed region)
    e2e = None
    e2e_pc = PCOUNT                                    # the same [20][21][N] operand: 210 MiB per direction per call
    hb = torch.empty((e2e_pc, NCOMP, n), dtype=torch.int64).pin_memory()
    hb.copy_(x0[:e2e_pc].cpu())
    hnp = hb.numpy().view(np.uint64).reshape(-1)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        ctx.ntt_host(True, hnp, hnp, pcount=e2e_pc, ncomp=NCOMP, stream=stream)
        ctx.ntt_host(False, hnp, hnp, pcount=e2e_pc, ncomp=NCOMP, stream=stream)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, args.steps // 2)
    for _ in range(e2e_steps):
        ctx.ntt_host(True, hnp, hnp, pcount=e2e_pc, ncomp=NCOMP, stream=stream)
        ctx.ntt_host(False, hnp, hnp, pcount=e2e_pc, ncomp=NCOMP, stream=stream)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    nbytes = e2e_pc * NCOMP * n * 8
    e2e = {"value": world * e2e_steps * 2 * e2e_pc * NCOMP / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 2 * nbytes, "d2h_bytes_per_step": 2 * nbytes,
           "note": "b2_ntt_forward_host + b2_ntt_inverse_host on the [20][21][N] operand in pinned host memory (chunked: H2D, transform and D2H overlap); result verified below"}
    assert torch.equal(hb, x0[:e2e_pc].cpu()), "e2e round trip mismatch"
    # what the host link can do at best: the same pinned buffer copied in both directions at once (plain cudaMemcpyAsync)
    try:
        hb2 = torch.empty_like(hb).pin_memory(); dA = torch.empty_like(x0[:e2e_pc]); dB = x0[:e2e_pc].clone()
        sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
        best = 0.0
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            with torch.cuda.stream(sA):
                dA.copy_(hb, non_blocking=True)
            with torch.cuda.stream(sB):
                hb2.copy_(dB, non_blocking=True)
            torch.cuda.synchronize()
            best = max(best, nbytes / (time.perf_counter() - t0) / 1e9)
        e2e["host_link_duplex_gbs_per_direction"] = best
        e2e["host_link_frac"] = (e2e["value"] / world) * n * 8 / 1e9 / best
        del hb2, dA, dB
    except Exception as ex:        # measurement aid only
        e2e["host_link_duplex_gbs_per_direction"] = None; e2e["host_link_note"] = str(ex)[:80]

    # ---- relinearize / rotate ops/s on the same parameter set (keyswitch core, data resident): top-level blocks of the line
    limb_sharded = None
    extra = {}
    ksblocks = {}
    if not args.no_extra:
        L = NCOMP - 1
        key = torch.empty((L, 2, NCOMP, n), dtype=torch.int64, device="cuda")
        for i in range(NCOMP):
            key[:, :, i, :] = torch.randint(0, moduli[i], (L, 2, n), dtype=torch.int64, device="cuda", generator=g)
        kh = ctx.kswitch_key(key)
        ct3 = x0[:3, :L, :].contiguous()                # [3][L][N] "NTT-form" ciphertext after multiply
        dest = ct3[:2].clone(); rot = torch.empty_like(dest)
        gal = pow(3, 1, 2 * n)

        def relin():
            dest.copy_(ct3[:2])
            ctx.keyswitch(ct3[2], L, True, kh, dest, b2.ASSIGN_ADD_INPLACE)

        def rotate():
            ctx.galois(ct3[:2], gal, rot, 2, L, True)
            ctx.keyswitch(rot[1], L, True, kh, rot, b2.ASSIGN_OVERWRITE_EXCEPT_FIRST)
        ks_bytes = 8 * n * (L + 2 * L * (L + 1) + 4 * L)          # target + key + dest write + dest read (accumulating forms)
        for name, fn in (("relinearize", relin), ("rotate", rotate)):
            for _ in range(3):
                fn()
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k = max(5, args.steps)
            a0.record()
            for _ in range(k):
                fn()
            a1.record(); barrier()
            s = max_over_ranks(a0.elapsed_time(a1) * 1e-3)
            extra[name + "_ops_per_s"] = world * k / s
            extra[name + "_us"] = s / k * 1e6
            ach = ks_bytes / (s / k) / 1e9
            ksblocks[name] = {"value": world * k / s, "unit": "ops/s", "us_per_op": s / k * 1e6, "steps": k,
                              "config": "C3 CKKS N=65536, 20 limbs + special prime, one ciphertext per GPU per op, operands and key resident in HBM",
                              "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": ks_traffic(name),
                                           "algorithmic_bytes_per_op": ks_bytes, "kernels": "INTT(target), mod-up NTT, key inner product, INTT(special), mod-down NTT (5 launches)",
                                           "peak_source": peak_src}}
        # end to end: the ciphertext lives in pinned host memory, the key stays resident (b2_keyswitch_host)
        try:
            h3 = torch.empty((3, L, n), dtype=torch.int64).pin_memory(); h3.copy_(ct3.cpu())
            hd = torch.empty((2, L, n), dtype=torch.int64).pin_memory()
            t_np = h3[2].numpy().view(np.uint64).reshape(-1); d_np = hd.numpy().view(np.uint64).reshape(-1)
            stream = torch.cuda.current_stream().cuda_stream

            def relin_host():
                hd.copy_(h3[:2])
                ctx.keyswitch_host(t_np, L, True, kh, d_np, b2.ASSIGN_ADD_INPLACE, stream=stream)
            for _ in range(2):
                relin_host()
            relin(); torch.cuda.synchronize()
            assert torch.equal(hd, dest.cpu()), "e2e relinearize differs from the resident result"
            # timed: the C-ABI call alone (the destination keeps accumulating: every word stays a canonical residue, and the call moves the same
            # bytes whatever they hold); the reset copy above is host-side test scaffolding, not part of the operation
            barrier(); t0 = time.perf_counter(); k = max(20, args.steps)
            for _ in range(k):
                ctx.keyswitch_host(t_np, L, True, kh, d_np, b2.ASSIGN_ADD_INPLACE, stream=stream)
            s = max_over_ranks(time.perf_counter() - t0)
            for nm in ("relinearize", "rotate"):
                ksblocks[nm]["e2e"] = {"value": world * k / s, "unit": "ops/s", "h2d_bytes_per_step": 3 * L * n * 8, "d2h_bytes_per_step": 2 * L * n * 8,
                                       "note": "b2_keyswitch_host: target + destination from pinned host memory, key resident, result back to the host (the destination's old contents "
                                               "are uploaded on a copy stream under the keyswitch and added by one launch); measured on relinearize (rotate uploads one polynomial less)"}
            del h3, hd
        except Exception as ex:
            for nm in ("relinearize", "rotate"):
                ksblocks[nm]["e2e"] = {"error": str(ex)[:160]}
        extra["keyswitch_compulsory_bytes"] = ks_bytes
        def timed(fn, k):
            for _ in range(2):
                fn()
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(k):
                fn()
            a1.record(); barrier()
            return max_over_ranks(a0.elapsed_time(a1) * 1e-3) / k

        # C3, batch 8: eight relinearizes sharing the key in one batched call (Evaluator::switch_key_internal_batched, K5)
        nb = 8
        cb = [x0[:3, :L, :].clone() for _ in range(nb)]
        db = [c[:2].clone() for c in cb]
        s8 = timed(lambda: ctx.keyswitch_batched([c[2] for c in cb], L, True, kh, db, b2.ASSIGN_ADD_INPLACE), max(3, args.steps // 4))
        extra["relinearize_batch8_ops_per_s"] = world * nb / s8
        extra["relinearize_batch8_us_per_op"] = s8 / nb * 1e6
        del cb, db
        # throughput regime of the headline metric: the same transforms over 10 x the operand (4200 limbs, 2.1 GiB)
        big = torch.empty((200, NCOMP, n), dtype=torch.int64, device="cuda")
        for i in range(10):
            big[20 * i:20 * (i + 1)].copy_(x0)
        tf = timed(lambda: ctx.ntt_forward(big, pcount=200, ncomp=NCOMP), 3)
        ti = timed(lambda: ctx.ntt_inverse(big, pcount=200, ncomp=NCOMP), 3)
        extra["ntt_4200_limbs"] = {"forward_limbs_per_s": world * 4200 / tf, "inverse_limbs_per_s": world * 4200 / ti,
                                   "forward_hbm_frac": 4200 * n * 16 / tf / 1e9 / peak, "inverse_hbm_frac": 4200 * n * 16 / ti / 1e9 / peak}
        del big
        # SURVEY §8f rank 2 on the same parameter set: Encryptor (asymmetric encryption of zero at the first level: sample u, e0, e1 with the
        # AES-CTR generator, NTT, products with the public key, noise, switch down from the key level) and Decryptor (c0 + c1 s), data resident
        pk = torch.empty((2, NCOMP, n), dtype=torch.int64, device="cuda")
        for i in range(NCOMP):
            pk[:, i, :] = torch.randint(0, moduli[i], (2, n), dtype=torch.int64, device="cuda", generator=g)
        skp = pk[0].contiguous()
        fresh = torch.empty((2, L, n), dtype=torch.int64, device="cuda")
        dec = torch.empty((L, n), dtype=torch.int64, device="cuda")
        prng = b2.Prng(2026, 7)
        te = timed(lambda: ctx.encrypt_zero_asymmetric(pk, L, True, fresh, prng), max(3, args.steps // 2))
        td = timed(lambda: ctx.decrypt([fresh], [dec], 2, L, skp), max(3, args.steps // 2))
        extra["encrypt_zero_asymmetric_us"] = te * 1e6; extra["encrypt_zero_asymmetric_ops_per_s"] = world / te
        extra["decrypt_us"] = td * 1e6; extra["decrypt_ops_per_s"] = world / td
        del pk, skp, fresh, dec
        # the other named parameter sets (BASELINE.json configs 2 and 4), data resident, per GPU x world
        extra["c2_bfv_n16384_batch256"] = bench_c2(b2, torch, timed, local, world)
        extra["c4_bgv_n32768"] = bench_c4(b2, torch, timed, local, world)
        extra["c5_bfv_n8192_packing"] = bench_c5_pack(b2, torch, timed, local, world)
        extra["c1_bfv_n8192_quickstart"] = bench_c1_quickstart(b2, torch, timed, local, world)
        if world > 1:
            # ONE relinearize cut over all GPUs by RNS limb: all-gather of the coefficient-form target and
            # broadcast of the special limb over NCCL, everything else local to the modulus owner
            sh = importlib.import_module("troy-nova_b200.sharded")
            gs = torch.Generator(device="cuda"); gs.manual_seed(4242)          # the SAME ciphertext and key on every rank
            for i in range(NCOMP):
                key[:, :, i, :] = torch.randint(0, moduli[i], (L, 2, n), dtype=torch.int64, device="cuda", generator=gs)
            cts = torch.empty((3, L, n), dtype=torch.int64, device="cuda")
            for j in range(L):
                cts[:, j, :] = torch.randint(0, moduli[j], (3, n), dtype=torch.int64, device="cuda", generator=gs)
            want = cts[:2].clone()
            ctx.keyswitch(cts[2], L, True, kh, want, b2.ASSIGN_ADD_INPLACE)     # single-GPU result: the checker
            ks = sh.LimbShardedKeyswitch(sh.DeviceStages(ctx), L)
            f, c, nct = ks.first, ks.count, ks.nct
            shard = ctx.kswitch_key_shard(kh, L, f, c) if c else None
            tl = cts[2, f:f + nct].contiguous() if nct else torch.zeros(2, dtype=torch.int64, device="cuda")
            dl0 = cts[:2, f:f + nct].contiguous() if nct else torch.zeros(2, dtype=torch.int64, device="cuda")
            dl = dl0.clone()

            def relin_sharded():
                dl.copy_(dl0)
                ks(tl, True, shard, dl, b2.ASSIGN_ADD_INPLACE)
            relin_sharded()
            ok = 1 if (nct == 0 or torch.equal(dl, want[:, f:f + nct])) else 0
            for _ in range(3):
                relin_sharded()
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k = max(5, args.steps)
            a0.record()
            for _ in range(k):
                relin_sharded()
            a1.record(); barrier()
            s = max_over_ranks(a0.elapsed_time(a1) * 1e-3)
            extra["relinearize_limb_sharded_us"] = s / k * 1e6
            extra["relinearize_limb_sharded_ops_per_s"] = k / s
            extra["relinearize_limb_sharded_speedup_vs_1gpu"] = extra["relinearize_us"] / (s / k * 1e6)
            extra["relinearize_limb_sharded_bit_exact"] = bool(max_over_ranks(1.0 - ok) == 0.0)
            # the same keyswitch with the P2P data plane: pushes over NVLink into the peers' symmetric buffers + flags, weighted partition
            try:
                ex = sh.P2PExchange.from_symmetric_memory(L, n)
                kp = sh.P2PLimbShardedKeyswitch(ctx, L, ex)
                f2, c2, nct2 = kp.first, kp.count, kp.nct
                shard2 = ctx.kswitch_key_shard(kh, L, f2, c2) if c2 else None
                tl2 = cts[2, f2:f2 + nct2].contiguous() if nct2 else torch.zeros(2, dtype=torch.int64, device="cuda")
                dl20 = cts[:2, f2:f2 + nct2].contiguous() if nct2 else torch.zeros(2, dtype=torch.int64, device="cuda")
                dl2 = dl20.clone()

                def relin_p2p():
                    dl2.copy_(dl20)
                    kp(tl2, True, shard2, dl2, b2.ASSIGN_ADD_INPLACE)
                relin_p2p()
                ok2 = 1 if (nct2 == 0 or torch.equal(dl2, want[:, f2:f2 + nct2])) else 0
                for _ in range(3):
                    relin_p2p()
                barrier()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(k):
                    relin_p2p()
                a1.record(); barrier()
                s2 = max_over_ranks(a0.elapsed_time(a1) * 1e-3)
                limb_sharded = {"relinearize_us": s2 / k * 1e6, "ops_per_s": k / s2, "speedup_vs_1gpu": extra["relinearize_us"] / (s2 / k * 1e6),
                                "bit_exact_vs_single_gpu": bool(max_over_ranks(1.0 - ok2) == 0.0), "scaling": "strong (one keyswitch cut over all GPUs by RNS limb)",
                                "data_plane": "P2P pushes into the peers' symmetric buffers + system-scope flags (b2_shard_push / b2_shard_wait), no collective call",
                                "partition": [p[1] for p in kp.parts], "nccl_path_us": s / k * 1e6}
            except Exception as ex_:        # symmetric memory unavailable: the NCCL path above stands
                limb_sharded = {"error": str(ex_)[:200], "nccl_path_us": s / k * 1e6}
            extra["limb_partition"] = [sh.partition(L + 1, world, r)[1] for r in range(world)]

    cpu = None
    gpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_ntt()
        if not args.no_extra:
            try:
                kcpu = cpu_reference_keyswitch()
                for nm in ("relinearize", "rotate"):
                    if kcpu and nm in ksblocks:
                        ksblocks[nm]["cpu_baseline"] = kcpu[nm]
            except Exception as ex:
                extra["cpu_keyswitch_baseline_error"] = str(ex)[:200]
    if rank == 0 and world == 1 and not args.no_gpu_baseline and not args.no_extra:
        try:
            gpu_base = gpu_reference_baseline()
            if gpu_base:
                gpu_base["speedup_ntt"] = value / gpu_base["ntt_limbs_per_s"]
                for nm in ("relinearize", "rotate"):
                    if nm in ksblocks:
                        gpu_base["speedup_" + nm] = ksblocks[nm]["value"] / gpu_base[nm + "_ops_per_s"]
        except Exception as ex:
            gpu_base = {"error": str(ex)[:200]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64 (40-bit limbs: exact FP64-pipe arithmetic; 60-bit limbs: integer pipe)", "data": "synthetic",
                "config": {"workload": "C3 CKKS N=65536, q={60,40x19,60}: forward+inverse negacyclic NTT of a [20][21][N] operand (420 limbs, 210 MiB) per GPU per step",
                           "l2": "inputs (210 MiB) larger than L2 (126 MB); no flush needed", "limbs_per_step_per_gpu": 2 * limbs,
                           "sharding": "by ciphertext batch, no data-path collective"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
                "clocks": sampler.summary() if sampler else None,
                "relinearize": ksblocks.get("relinearize") if not args.no_extra else None, "rotate": ksblocks.get("rotate") if not args.no_extra else None,
                "gpu_baseline": gpu_base, "limb_sharded_keyswitch": limb_sharded, "extra": extra}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
