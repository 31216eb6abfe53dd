python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for m in 0.1; do for q in 4 8; do
echo "== margin $m qpt $q"
B2ICP_MARGIN=$m B2ICP_QPT=$q python bench.py --steps 5 --warmup 3 --cpu-sample 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'launch_us',round(r['avg_launch_us'],1),'frac',round(r['frac'],4),'searched',round(r['searched_fraction'],4),'iters',d['config']['mean_iterations'])
"
done; done
for m in 0.0 0.05 0.25; do
echo "== margin $m qpt 4"
B2ICP_MARGIN=$m python bench.py --steps 5 --warmup 3 --cpu-sample 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'launch_us',round(r['avg_launch_us'],1),'frac',round(r['frac'],4),'searched',round(r['searched_fraction'],4),'iters',d['config']['mean_iterations'])
"
done
