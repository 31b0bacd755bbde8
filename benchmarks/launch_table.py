"""Prints the kernels of the LAST step of an `ncu --metrics gpu__time_duration.sum --csv` launch list (our kernels only)."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
names = [(r[ki], float(r[vi].replace(",", "")), r[gi]) for r in rows[1:]]
# a step starts at init_bbox_kernel
starts = [i for i, (n, _, _) in enumerate(names) if "init_bbox_kernel" in n]
last = names[starts[-1]:] if starts else names
total = sum(v for _, v, _ in last)
for n, v, g in last:
    short = n.replace("<unnamed>::", "").split("(")[0][:70]
    print("%-72s %-18s %9.1f us %5.1f %%" % (short, g, v / 1000, 100 * v / total))
print("total %.1f us over %d launches" % (total / 1000, len(last)))
