python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 40 --warmup 5 > gpurun_out/b20.json 2> gpurun_out/b20.err || tail -5 gpurun_out/b20.err
python - <<PY
import json
d=json.load(open('gpurun_out/b20.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']), {k:round(v,1) for k,v in d['roofline']['kernel_us_per_step'].items()}, 'us/gn', round(d['us_per_gn_iter'],2))
print({k:round(v,1) for k,v in d['align_and_feature_align_kernel_us_per_frame'].items()})
print(d['e2e'])
PY
