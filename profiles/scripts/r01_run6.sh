python -m pytest tests/test_gpu_tracker.py -m gpu -x -q 2>&1 | tail -4
for cfg in "--groups 16 --threads 16" "--groups 32 --threads 16" "--groups 48 --threads 16" "--groups 32 --threads 8"; do
python bench.py --steps 20 --warmup 5 $cfg > gpurun_out/b.json 2> gpurun_out/b.err || tail -5 gpurun_out/b.err
python - <<PY
import json
d=json.load(open('gpurun_out/b.json'))
print("$cfg",'value',round(d['value']),'e2e',round(d['e2e']['value']), {k:round(v,3) for k,v in d['host_phase_thread_seconds']['e2e'].items()}, d['gpu_launches'])
PY
done
