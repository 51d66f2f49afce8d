"""Summarise an ncu launch list (gpu__time_duration.sum per launch) by kernel name."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
tot = collections.defaultdict(float); cnt = collections.Counter()
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:70]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
    tot[name] += ms; cnt[name] += 1
total = sum(tot.values())
print("total %.2f ms over %d launches" % (total, sum(cnt.values())))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("%8.2f ms %5.1f%% %5d  %s" % (v, 100 * v / total, cnt[k], k))
