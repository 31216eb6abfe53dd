for v in "B2ICP_X=1" "B2ICP_QPT_SCHED=64"; do
echo "== $v"
env $v python bench.py --steps 16 --warmup 8 --cpu-sample 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'sync',round(d['config']['synchronous_call_scans_per_s']))"
done
