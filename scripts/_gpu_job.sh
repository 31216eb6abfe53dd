python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 12 --warmup 3 2>gpurun_out/b.err | tail -1 > gpurun_out/bench_r01_stream4.json
python -c "
import json
d=json.loads(open('gpurun_out/bench_r01_stream4.json').read())
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'sync',round(d['config']['synchronous_call_scans_per_s']),'launches',d['gpu_launches'],'frac',d['roofline']['frac'],'cpu',d['cpu_baseline']['value'])"
