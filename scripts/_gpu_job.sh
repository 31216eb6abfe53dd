for b in 8 16 32; do
echo "== batch $b"
python bench.py --steps 24 --warmup 4 --cpu-sample 0 --batch $b 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'sync',round(d['config']['synchronous_call_scans_per_s']))"
done
