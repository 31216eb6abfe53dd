# dev tool: bench.py's resident leg for several neighbour-grid cell sizes
for c in "$@"; do
echo "== cell $c"
python bench.py --steps 5 --warmup 3 --cpu-sample 0 --grid-cell $c 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'iter_us',round(r['avg_launch_us'],1),'frac',round(r['frac'],4),'searched',round(r['searched_fraction'],4),'occ',round(d['config']['grid_occupancy'],2))
"
done
