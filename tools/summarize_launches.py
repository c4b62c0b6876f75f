"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    if unit in ("us", "usecond"): v *= 1e3
    elif unit in ("ms", "msecond"): v *= 1e6
    tot[name] += v; cnt[name] += 1
allt = sum(tot.values())
print(f"{'kernel':40s} {'launches':>9s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k:40s} {cnt[k]:9d} {v / 1e6:10.3f} {v / allt:7.3f} {v / cnt[k] / 1e3:9.2f}")
print(f"{'TOTAL':40s} {sum(cnt.values()):9d} {allt / 1e6:10.3f}")
