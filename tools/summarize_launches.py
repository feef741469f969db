"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv, re, sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
agg = OrderedDict()
total = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3 if unit in ("ms", "msecond") else v)
    name = r["Kernel Name"]
    short = re.sub(r"\(.*", "", name)
    m = re.match(r"(?:void )?(?:b2::)?(\w+)<(.*)", name)
    if m and m.group(1) == "ntt_pass_kernel":
        a = [x.strip() for x in m.group(2).split(",")]
        # <A, COL, FWD, FIRST, FINAL, ...>
        short = "ntt_pass_kernel[%s %s A=%s %s]" % ("fwd" if a[2].startswith(("true", "1")) else "inv", "col" if a[1].startswith(("true", "1")) else "row", a[0],
                                                    re.sub(r"b2::", "", ",".join(a[11:13]))[:60])
    elif m:
        short = m.group(1)
    k = (short, r["Grid Size"], r["Block Size"])
    e = agg.setdefault(k, [0, 0.0])
    e[0] += 1; e[1] += v_us; total += v_us
print("# %s : %d launches, %.1f us total device time (cold-cache, serialised: compare SHARES)" % (sys.argv[1], sum(e[0] for e in agg.values()), total))
print("%-100s %-16s %-12s %6s %12s %10s %7s" % ("kernel", "grid", "block", "count", "total_us", "avg_us", "share"))
for (short, g, b), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-100s %-16s %-12s %6d %12.1f %10.2f %6.1f%%" % (short[:100], g, b, c, t, t / c, 100 * t / total))
