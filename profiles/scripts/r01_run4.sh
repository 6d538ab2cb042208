python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for cfg in "--groups 16 --threads 16" "--groups 32 --threads 16"; do
python bench.py --steps 20 --warmup 5 $cfg > gpurun_out/b.json 2> gpurun_out/b.err || tail -5 gpurun_out/b.err
python - <<PY
import json
d=json.load(open('gpurun_out/b.json'))
print("$cfg", 'value',round(d['value']),'e2e',round(d['e2e']['value']), d['host_phase_thread_seconds'], d['gpu_launches'], d['roofline']['kernel_us_per_step'])
PY
done
