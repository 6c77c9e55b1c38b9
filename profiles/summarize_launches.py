"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv [skip_first_n]"""
import csv, re, sys
from collections import defaultdict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[skip:]:
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void ", "", name)[:90]
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)   # -> us
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| `{k}` | {n} | {t:.0f} | {100 * t / tot:.1f}% |")
print(f"\ntotal {tot/1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches")
