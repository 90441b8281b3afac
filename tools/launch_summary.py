"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py file.csv"""
import csv
import re
import sys
from collections import defaultdict

tot, cnt = defaultdict(float), defaultdict(int)
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
ik, iv, iu = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
for r in rows[hdr + 1:]:
    if len(r) <= iv:
        continue
    name = re.sub(r"\(.*", "", r[ik]).replace("void qb::", "").replace("qb::", "")
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(r[iu], 1e-6)
    tot[name] += v
    cnt[name] += 1
all_ms = sum(tot.values())
print(f"{'kernel':60s} {'launches':>8s} {'total ms':>10s} {'avg ms':>9s} {'share':>7s}")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k[:60]:60s} {cnt[k]:8d} {tot[k]:10.2f} {tot[k]/cnt[k]:9.3f} {100*tot[k]/all_ms:6.1f}%")
print(f"{'TOTAL':60s} {sum(cnt.values()):8d} {all_ms:10.2f}")
