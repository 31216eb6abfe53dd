#!/bin/bash
# configs[3] leg: wall time per call and where the device time goes (ncu launch list, summed by kernel)
set -u
mkdir -p gpurun_out
timeout 900 python scripts/r2_pairs_prof.py 2> gpurun_out/pairs_prof.err | tee gpurun_out/pairs_prof.jsonl | cut -c1-200
NCU=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/pairs_launches.csv python scripts/r2_pairs_prof.py > gpurun_out/pairs_ncu.log 2>&1
python - <<'PY'
import collections, csv
rows = [r for r in csv.reader(open('gpurun_out/pairs_launches.csv')) if len(r) > 10]
h = rows[0]; ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    k = r[ki].split("(")[0][-28:]
    tot[k] += v; cnt[k] += 1
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k:30s} {cnt[k]:5d} launches {v/1e3:8.2f} ms")
print("total", sum(tot.values())/1e3, "ms in", sum(cnt.values()), "launches")
PY
