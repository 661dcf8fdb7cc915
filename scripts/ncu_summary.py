"""Summarise an `ncu --page raw --csv` export: one line per profiled launch."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [
    ("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"), ("lts__t_sectors_op_read.sum", "l2_rd_sect"),
    ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
]
idx = [(hdr.index(w), n) for w, n in want if w in hdr]
print(" | ".join(f"{n}[{units[i]}]" for i, n in idx))
for r in data:
    out = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = v.split("(")[0].split("::")[-1][:28]
        out.append(v)
    print(" | ".join(out))
