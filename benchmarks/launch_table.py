"""Prints the kernels of the LAST step of an `ncu --metrics gpu__time_duration.sum --csv` launch list (our kernels only)."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
names = [(r[ki], float(r[vi].replace(",", "")), r[gi]) for r in rows[1:]]
# a step starts at init_bbox_kernel - or, in the two-stream step, at the object-branch kernels enqueued just before it
# (ncu serialises kernels in launch order: side-stream launches of a step come first)
OBJECT_BRANCH = ("seg_histogram", "view_table", "row_normalize", "build_score", "gemm_kernel", "view_weights", "refine_weights",
                 "segmented_wmean")
starts = [i for i, (n, _, _) in enumerate(names) if "init_bbox_kernel" in n]
firsts = []
for st in starts:
    first = st
    while first > 0 and any(k in names[first - 1][0] for k in OBJECT_BRANCH):
        first -= 1
    firsts.append(first)
steps = [names[a:b] for a, b in zip(firsts, firsts[1:] + [len(names)])] or [names]
# the last step over the reference's int64 instance maps (bench.py times a uint8-map variant afterwards)
wide = [st for st in steps if any("seg_histogram_ring" in n or "seg_histogram_kernel<long long>" in n for n, _, _ in st)]
headline = [st for st in wide if any("unpack_compact_wide" in n for n, _, _ in st)]  # uint8 masks (not the full-output variant)
ring = [st for st in headline if any("seg_histogram_ring" in n for n, _, _ in st)]  # the two-stream step, if the run has one
last = (ring or headline or wide or steps)[-1]
total = sum(v for _, v, _ in last)
for n, v, g in last:
    short = n.replace("<unnamed>::", "").split("(")[0][:70]
    print("%-72s %-18s %9.1f us %5.1f %%" % (short, g, v / 1000, 100 * v / total))
print("total %.1f us over %d launches" % (total / 1000, len(last)))
