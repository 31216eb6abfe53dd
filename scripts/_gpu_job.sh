python -m pytest tests -m gpu -x -q 2>&1 | tail -3
B2ICP_LIB=icpslam_b200/lib/variants/k3.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for v in "B2ICP_X=1" "B2ICP_LIB=icpslam_b200/lib/variants/k3.so"; do
echo "== $v"
env $v python bench.py --steps 16 --warmup 8 --cpu-sample 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'sync',round(d['config']['synchronous_call_scans_per_s']),'searched',round(d['roofline']['searched_fraction'],4))"
done
