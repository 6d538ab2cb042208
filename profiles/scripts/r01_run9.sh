python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 > gpurun_out/b.json 2> gpurun_out/b.err || tail -5 gpurun_out/b.err
python - <<PY
import json
d=json.load(open('gpurun_out/b.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']), {k:round(v,3) for k,v in d['host_phase_thread_seconds']['e2e'].items()}, d['roofline']['kernel_us_per_step'], d['gpu_launches'])
PY
