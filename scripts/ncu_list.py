"""Print one line per launch of an `ncu --metrics ... --csv --log-file x.csv` capture (run here, no GPU)."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
ki, mi, vi, ii, ui = (h.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
d = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Gbyte": 1e3}.get(r[ui], 1.0)
    d.setdefault((r[ii], r[ki].split("(")[0][-18:]), {})[r[mi]] = v
tot = collections.defaultdict(float)
for (i, k), m in d.items():
    tot[k] += m["gpu__time_duration.sum"]
    print(i, k, "us=%.1f" % m["gpu__time_duration.sum"], "inst=%.1fM" % (m.get("smsp__inst_executed.sum", 0) / 1e6),
          "rd=%.0fMB wr=%.0fMB" % (m.get("dram__bytes_read.sum", 0), m.get("dram__bytes_write.sum", 0)),
          "issue=%.0f%% warps=%.0f%%" % (m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0),
                                       m.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0)))
print({k: round(v, 1) for k, v in tot.items()})
