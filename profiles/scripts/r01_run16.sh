python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 40 --warmup 5 > gpurun_out/b16.json 2> gpurun_out/b16.err || tail -5 gpurun_out/b16.err
python - <<PY
import json
d=json.load(open('gpurun_out/b16.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']), d['roofline']['kernel_us_per_step'])
print({k:round(v,1) for k,v in d['align_and_feature_align_kernel_us_per_frame'].items()})
PY
python bench.py --steps 60 --warmup 5 --sweep 8x2,8x3,8x4,8x8 --sweep-device 2>&1 | grep sweep
