"""Summarise an `ncu --page source --csv` export (SASS view): executed-instruction mix per opcode and the instructions
that collect the most stall samples.  Usage: python tools/ncu_source_summary.py prof_source.csv[.gz] [kernel-index] [top]"""
import csv, gzip, sys, collections
path = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
kernels = []; cur = None; hdr = None
for row in csv.reader(f):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}; kernels.append(cur); hdr = None; continue
    if row and row[0] == "Address":
        hdr = row; continue
    if cur is not None and hdr and len(row) >= len(hdr) - 2:
        cur["rows"].append(dict(zip(hdr, row)))
k = kernels[which]
print("kernel:", k["name"][:100], " instructions:", len(k["rows"]))
mix = collections.Counter(); samp = collections.Counter(); tot = 0; tots = 0
stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
stall_tot = collections.Counter()
for i, r in enumerate(k["rows"]):
    op = r["Source"].split()[0] if not r["Source"].strip().startswith("@") else r["Source"].split()[1]
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LD", "ST", "IMAD", "DFMA", "ATOM", "RED", "BAR")) and "." in op else "")
    n = int(r["Instructions Executed"] or 0); s = int(r["# Samples"] or 0)
    mix[op] += n; samp[op] += s; tot += n; tots += s
    for c in stall_cols: stall_tot[c] += int(r[c] or 0)
print("total warp-instructions executed: %d   samples: %d" % (tot, tots))
print("%-16s %12s %7s %8s %7s" % ("opcode", "executed", "share", "samples", "share"))
for op, n in mix.most_common(30):
    print("%-16s %12d %6.1f%% %8d %6.1f%%" % (op, n, 100.0 * n / tot, samp[op], 100.0 * samp[op] / max(1, tots)))
print("stall totals:", ", ".join("%s=%d" % (c[6:], v) for c, v in stall_tot.most_common(10)))
print("--- top instructions by stall samples")
rows = sorted(enumerate(k["rows"]), key=lambda t: -int(t[1]["# Samples"] or 0))[:top]
for i, r in rows:
    why = sorted(((int(r[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print("%5d %6s  %-70s %s" % (i, r["# Samples"], r["Source"].strip()[:70], " ".join("%s=%d" % (b, a) for a, b in why if a)))
