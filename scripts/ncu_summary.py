"""Summarise ncu captures into small text files for profiles/ (run here, no GPU needed):
    python scripts/ncu_summary.py rep  gpurun_out/x.ncu-rep   > profiles/x_summary.txt
    python scripts/ncu_summary.py list gpurun_out/launches.csv > profiles/launches_summary.txt
"""
import collections
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def rep(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# {path}: {len(rows) - 2} captured launch(es); ncu --set full --clock-control none")
    for r in rows[2:]:
        print(f"\n== {r[hdr.index('Kernel Name')]}  (launch id {r[0]})")
        for w in WANT:
            if w in hdr:
                print(f"  {w:70s} {r[hdr.index(w)]:>18s} {units[hdr.index(w)]}")
        st = [(float(r[i] or 0), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for i, h in enumerate(hdr)
              if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("not_issued")]
        tot = sum(v for v, _ in st) or 1.0
        st.sort(reverse=True)
        print("  warp stall samples: " + ", ".join(f"{h} {v / tot:.1%}" for v, h in st[:8]))


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        if r[ui] == "ns":
            v /= 1e3
        agg[r[ki]][0] += 1
        agg[r[ki]][1] += v
    tot = sum(v for _, v in agg.values())
    print(f"# {path}: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print(f"# total device time of listed launches: {tot:.1f} us")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v / tot:7.3%}  n={n:5d}  total_us={v:12.1f}  avg_us={v / n:9.2f}  {k[:110]}")


if __name__ == "__main__":
    {"rep": rep, "list": launches}[sys.argv[1]](sys.argv[2])
