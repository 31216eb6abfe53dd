"""Run bench.py's resident leg for each kernel-tuning variant under icpslam_b200/lib/variants (dev tool)."""
import glob, json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for lib in sorted(glob.glob(os.path.join(root, "icpslam_b200/lib/variants/*.so"))):
    env = dict(os.environ, B2ICP_LIB=lib)
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "5", "--warmup", "3", "--cpu-sample", "0"] + sys.argv[1:],
                       env=env, capture_output=True, text=True)
    try:
        j = json.loads(r.stdout.strip().splitlines()[-1])
        print(os.path.basename(lib), "value=%.0f e2e=%.0f ms/step=%.2f launch_us=%.1f frac=%.4f" % (
            j["value"], j["e2e"]["value"], j["ms_per_step"], j["roofline"]["avg_launch_us"], j["roofline"]["frac"]), flush=True)
    except Exception as e:
        print(os.path.basename(lib), "FAILED", e, r.stderr[-500:], flush=True)
