for f in 4 6 8; do
echo "== in flight $f"
python bench.py --steps 24 --warmup 8 --cpu-sample 0 --in-flight $f 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'sync',round(d['config']['synchronous_call_scans_per_s']))"
done
