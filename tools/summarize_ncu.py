"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, share."""
import csv, re, sys, collections
path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    tot[name][0] += 1
    tot[name][1] += v_us
total = sum(v[1] for v in tot.values())
print(f"total device time {total/1e3:.3f} ms over {sum(v[0] for v in tot.values())} launches")
print(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {n:8d} {t:12.1f} {t/n:10.2f} {100*t/total:6.2f}%")
