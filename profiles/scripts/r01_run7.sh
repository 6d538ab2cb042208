for mc in 8 32; do
for cfg in "--groups 16 --threads 16" "--groups 32 --threads 16" "--groups 64 --threads 16"; do
CUDA_DEVICE_MAX_CONNECTIONS=$mc python bench.py --steps 20 --warmup 5 $cfg > gpurun_out/b.json 2> gpurun_out/b.err || tail -5 gpurun_out/b.err
python - <<PY
import json
d=json.load(open('gpurun_out/b.json'))
print("MC=$mc $cfg",'value',round(d['value']),'e2e',round(d['e2e']['value']), {k:round(v,3) for k,v in d['host_phase_thread_seconds']['e2e'].items()}, d['gpu_launches'])
PY
done
done
