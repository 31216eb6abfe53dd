"""Attribute an ncu capture's per-SASS-instruction counters to CUDA source lines (run here, no GPU):
    python scripts/ncu_lines.py gpurun_out/x.ncu-rep <launch-skip> [kernel-substring] [top-N]
Uses nvdisasm --print-line-info on the cubin inside icpslam_b200/lib/libb2icp.so (build with -lineinfo)."""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, skip = sys.argv[1], sys.argv[2]
ksub = sys.argv[3] if len(sys.argv) > 3 else "icp_sweep_p2pILi4"
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "icpslam_b200/lib/libb2icp.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
amap, cur, inside = {}, None, False
for ln in dis.splitlines():
    if ln.startswith("\t.section\t.text.") or ln.startswith(".text."):
        inside = ksub in ln
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    m2 = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
    if m2 and cur:
        amap.setdefault(int(m2.group(1), 16), cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[1]
ie, te, ss = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = tots = 0
base = None
seen = set()
for r in rows[2:]:
    try:
        a = int(r[0], 16) if r[0].startswith("0x") else int(r[0])
        v, t, s = float(r[ie]), float(r[te]), float(r[ss])
    except (ValueError, IndexError):
        continue
    if base is None:
        base = a
    if a in seen:
        continue
    seen.add(a)
    key = amap.get(a - base, ("?", 0))
    agg[key][0] += v; agg[key][1] += t; agg[key][2] += s
    tot += v; tots += s
print(f"# {rep} launch {skip} kernel~{ksub}: {tot:.0f} warp instructions, {tots:.0f} stall samples")
src = {}
def line(f, l):
    p = os.path.join(root, "icpslam_b200/csrc", f)
    if f not in src and os.path.exists(p):
        src[f] = open(p).read().splitlines()
    try:
        return src[f][l - 1].strip()[:100]
    except (KeyError, IndexError):
        return ""
key = (lambda kv: -kv[1][1]) if os.environ.get("BY") == "thr" else (lambda kv: -kv[1][2])
for k, (v, t, s) in sorted(agg.items(), key=key)[:topn]:
    print(f"{s / max(tots,1):6.2%} smp {v / max(tot,1):6.2%} inst {t/1e6:8.1f}M thr-inst lanes={t / max(v, 1):5.1f}  {k[0]}:{k[1]}  {line(*k)}")
