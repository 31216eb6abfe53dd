"""Per-kernel SASS digest of the built library (run here, no GPU):  python scripts/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections, os, re, subprocess
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "icpslam_b200/lib/libb2icp.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
KEEP = ("UBLKCP", "SYNCS", "CREDUX", "REDUX", "DADD", "DFMA", "DMUL", "LDG", "LDS", "STG", "STS", "SHFL", "MUFU", "VIMNMX", "ATOM", "RED", "LDL", "STL")
print("# cuobjdump -sass icpslam_b200/lib/libb2icp.so (sm_100a): per kernel the instruction count and the mnemonics that show what\n"
      "# it is built from (UBLKCP.S.G = cp.async.bulk global -> shared, SYNCS.* = mbarrier arrive / expect_tx / try_wait,\n"
      "# CREDUX / REDUX = warp reductions, LDL / STL = local memory).  Regenerate: python scripts/sass_summary.py\n")
blocks = re.split(r"\n\s*Function : ", sass)[1:]
for blk, name in zip(blocks, names):
    ins = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", blk)
    c = collections.Counter()
    for i in ins:
        base = i.split(".")[0]
        if base in KEEP:
            c["UBLKCP.S.G" if i.startswith("UBLKCP.S.G") else base] += 1
    print(f"{len(ins):6d} instr  {name.strip()}")
    print("              " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())))
