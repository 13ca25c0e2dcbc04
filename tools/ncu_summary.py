"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"_GLOBAL__N__[0-9a-f_]+cu_[0-9a-f]+::|\(anonymous namespace\)::", "", name)
    rows.append((int(r["ID"]), name, us))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
tot = sum(u for _, _, u in rows)
agg = collections.OrderedDict()
for _, n, u in rows:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += u
print(f"launches={len(rows)} total={tot:.1f} us")
for n, (c, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{u:10.1f} us {100*u/tot:5.1f}%  x{c:<4d} avg {u/c:8.2f} us  {n[:110]}")
