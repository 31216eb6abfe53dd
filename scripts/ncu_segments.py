"""Split an ncu source-page capture of one launch into the code segments between barriers / calls / exits and
print each segment's share of executed instructions and stall samples (run here, no GPU):
    python scripts/ncu_segments.py gpurun_out/x.ncu-rep <launch-skip>"""
import csv, io, subprocess, sys
rep, skip = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[1]
ie, te, ss = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
seg, cur, seen = [], None, set()
def new(): return {"inst": 0.0, "thr": 0.0, "smp": 0.0, "first": None, "stalls": {}, "ld": 0}
cur = new()
for r in rows[2:]:
    try:
        a, v, t, s = r[0], float(r[ie]), float(r[te]), float(r[ss])
    except (ValueError, IndexError):
        continue
    if a in seen:
        continue
    seen.add(a)
    if cur["first"] is None:
        cur["first"] = a
    cur["inst"] += v; cur["thr"] += t; cur["smp"] += s
    for i, n in stall_cols:
        try:
            cur["stalls"][n] = cur["stalls"].get(n, 0.0) + float(r[i] or 0)
        except ValueError:
            pass
    op = r[1].split()
    if any(k in r[1] for k in ("BAR.SYNC", "EXIT")) or (len(op) and op[0] == "CALL.REL.NOINC"):
        seg.append((cur, r[1].strip()[:34]))
        cur = new()
seg.append((cur, "end"))
tot = sum(c["inst"] for c, _ in seg) or 1
tots = sum(c["smp"] for c, _ in seg) or 1
print(f"# {rep} launch {skip}: {tot/1e6:.1f}M warp instructions, {tots:.0f} samples")
for c, why in seg:
    if c["inst"] / tot < 0.002 and c["smp"] / tots < 0.005:
        continue
    top = sorted(c["stalls"].items(), key=lambda kv: -kv[1])[:3]
    st = ", ".join(f"{n[6:]} {v / max(c['smp'],1):.0%}" for n, v in top if v > 0)
    print(f"{c['first'][-5:]:>6} inst {c['inst']/tot:6.1%} ({c['inst']/1e6:6.2f}M) lanes {c['thr']/max(c['inst'],1):5.1f}  samples {c['smp']/tots:6.1%}  [{st}]  ends at {why}")
