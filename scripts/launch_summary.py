"""Aggregate an ncu --csv launch list (one row per launch per metric) by kernel name."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).split("::")[-1]
    m = row["Metric Name"]
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    if m.startswith("dram__bytes"):
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    a = agg.setdefault(name, collections.defaultdict(float))
    a[m] += v
    a["n_" + m] += 1
tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
print(f"{'kernel':44s} {'n':>4s} {'total ms':>9s} {'avg us':>9s} {'share':>6s} {'rd MB':>9s} {'wr MB':>9s} {'GB/s':>7s} {'occ%':>5s} {'sm%':>5s}")
for k, a in sorted(agg.items(), key=lambda x: -x[1]["gpu__time_duration.sum"]):
    n = int(a["n_gpu__time_duration.sum"])
    t = a["gpu__time_duration.sum"]
    rd, wr = a.get("dram__bytes_read.sum", 0) / n, a.get("dram__bytes_write.sum", 0) / n
    occ = a.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0) / n
    smp = a.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", 0) / n
    print(f"{k[:44]:44s} {n:4d} {t / 1e3:9.3f} {t / n:9.1f} {100 * t / tot:5.1f}% {rd:9.1f} {wr:9.1f} {(rd + wr) / (t / n) * 1e3 / 1e3:7.0f} {occ:5.1f} {smp:5.1f}")
