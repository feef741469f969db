"""Aggregate an `ncu --page source --csv` dump: executed warp instructions and stall samples per SASS opcode, per kernel."""
import csv, gzip, sys, collections
f = gzip.open(sys.argv[1], "rt") if sys.argv[1].endswith(".gz") else open(sys.argv[1])
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rd = csv.reader(f)
kern = None; hdr = None; agg = None; total = 0; tots = 0
def flush():
    if not agg: return
    print("==", kern[:110], " total warp-instr %d  samples %d" % (total, tots))
    for op, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
        print("   %-22s %12d %6.2f%%   stall samples %8d %6.2f%%" % (op, n, 100.0 * n / max(total, 1), s, 100.0 * s / max(tots, 1)))
for row in rd:
    if row and row[0] == "Kernel Name":
        flush(); kern = row[1]; agg = collections.defaultdict(lambda: [0, 0]); total = 0; tots = 0; hdr = None; continue
    if row and row[0] == "Address":
        hdr = {k: i for i, k in enumerate(row)}; continue
    if not hdr or len(row) < 6: continue
    sass = row[hdr["Source"]].strip()
    parts = sass.split()
    if not parts: continue
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    op = ".".join(op.split(".")[:2])
    n = int(row[hdr["Instructions Executed"]] or 0); s = int(row[hdr["Warp Stall Sampling (All Samples)"]] or 0)
    agg[op][0] += n; agg[op][1] += s; total += n; tots += s
flush()
