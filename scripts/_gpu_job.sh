python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 12 --warmup 3 --cpu-sample 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'sync',round(d['config']['synchronous_call_scans_per_s']),'iter_us',round(d['roofline']['avg_launch_us'],1),'frac',round(d['roofline']['frac'],4))"
