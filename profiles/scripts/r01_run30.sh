python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fast or filter" 2>&1 | tail -3
python bench.py --steps 60 --warmup 5 > gpurun_out/b30.json 2> gpurun_out/b30.err || tail -5 gpurun_out/b30.err
python - <<PY
import json
d=json.load(open('gpurun_out/b30.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']), {k:round(v,1) for k,v in d['roofline']['kernel_us_per_step'].items()})
PY
