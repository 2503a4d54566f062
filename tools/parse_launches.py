"""Print the per-kernel durations / DRAM bytes of ONE pipeline invocation from an ncu --csv launch list.
   python tools/parse_launches.py launches.csv first_kernel_substring [which_invocation]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ik, im, iv, iid = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("ID")
per = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > iv:
        per.setdefault(r[iid], {"k": r[ik]})[r[im]] = float(r[iv].replace(",", ""))
seq = list(per.values())
starts = [i for i, v in enumerate(seq) if sys.argv[2] in v["k"]]
w = int(sys.argv[3]) if len(sys.argv) > 3 else 1
a = starts[w]
b = starts[w + 1] if w + 1 < len(starts) else len(seq)
tot = 0.0
for v in seq[a:b]:
    t = v["gpu__time_duration.sum"] / 1e3
    tot += t
    print("%-44s %8.1f us  rd %7.1f MB  wr %7.1f MB" % (v["k"][:44], t, v.get("dram__bytes_read.sum", 0) / 1e6, v.get("dram__bytes_write.sum", 0) / 1e6))
print("total %.1f us over %d launches" % (tot, b - a))
