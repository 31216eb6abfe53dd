# dev tool: GPU parity tests, then bench.py's resident leg for a few tuning variants (env overrides)
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
run() {
echo "== $*"
env "$@" python bench.py --steps 5 --warmup 3 --cpu-sample 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'iter_us',round(r['avg_launch_us'],1),'frac',round(r['frac'],4),'searched',round(r['searched_fraction'],4),'iters',d['config']['mean_iterations'])
"
}
for v in "$@"; do run $v; done
