"""Summarise a per-launch ncu metrics CSV (--metrics gpu__time_duration.sum,dram__bytes_*.sum,sm__pipe_tensor_cycles_active...,
l1tex__m_xbar2l1tex_read_bytes.sum --csv) by kernel name: time, DRAM bytes and rate, L2->SM bytes and rate, time-weighted
tensor-pipe activity.   python tools/metrics_summary.py file.csv [top_n] [--launches NAME]"""
import collections
import csv
import re
import sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "nsecond": 1e-9, "us": 1e-6, "usecond": 1e-6,
         "ms": 1e-3, "msecond": 1e-3, "second": 1.0, "%": 1.0, "": 1.0}


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    launches = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if not row.get("ID"):
            continue
        d = launches.setdefault(int(row["ID"]), {"name": re.sub(r"\(.*", "", row["Kernel Name"])[:64], "grid": row.get("Grid Size", "")})
        try:
            d[row["Metric Name"]] = float(row["Metric Value"].replace(",", "")) * SCALE.get(row["Metric Unit"], 1.0)
        except ValueError:
            pass
    return list(launches.values())


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 40
    L = load(path)
    if "--launches" in sys.argv:
        pat = sys.argv[sys.argv.index("--launches") + 1]
        for d in L:
            if pat in d["name"]:
                t = d.get("gpu__time_duration.sum", 0.0)
                print("%9.1f us  grid %-16s dram %7.1f MB  l2->sm %8.1f MB  tensor %5.1f%%  %s" % (
                    t * 1e6, d["grid"], (d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)) / 1e6,
                    d.get("l1tex__m_xbar2l1tex_read_bytes.sum", 0) / 1e6,
                    d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0), d["name"]))
        return
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    for d in L:
        a = agg[d["name"]]
        t = d.get("gpu__time_duration.sum", 0.0)
        a["t"] += t
        a["n"] += 1
        a["dram"] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        a["l2sm"] += d.get("l1tex__m_xbar2l1tex_read_bytes.sum", 0)
        a["tens"] += t * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0)
    total = sum(a["t"] for a in agg.values())
    print("total %.2f ms over %d launches (ncu per-launch metrics, cold-cache, serialised: compare SHARES)" % (total * 1e3, len(L)))
    print("%9s %6s %6s  %9s %8s  %9s %8s  %7s  %s" % ("ms", "share", "n", "DRAM MB", "TB/s", "L2->SM MB", "TB/s", "tensor%", "kernel"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"])[:top]:
        t = a["t"]
        print("%9.2f %5.1f%% %6d  %9.0f %8.2f  %9.0f %8.2f  %7.1f  %s" % (
            t * 1e3, 100 * t / total, a["n"], a["dram"] / 1e6, a["dram"] / t / 1e12 if t else 0, a["l2sm"] / 1e6,
            a["l2sm"] / t / 1e12 if t else 0, a["tens"] / t if t else 0, k))


if __name__ == "__main__":
    main()
