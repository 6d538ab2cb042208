python bench.py --seqs 8 --groups 1 --threads 1 --steps 40 --warmup 5 > gpurun_out/b17.json 2> gpurun_out/b17.err || tail -5 gpurun_out/b17.err
python - <<PY
import json
d=json.load(open('gpurun_out/b17.json'))
print('8 seqs 1 group: value',round(d['value']),'e2e',round(d['e2e']['value']), {k:round(v,1) for k,v in d['roofline']['kernel_us_per_step'].items()})
print({k:round(v,1) for k,v in d['align_and_feature_align_kernel_us_per_frame'].items()})
PY
python bench.py --steps 60 --warmup 5 --sweep 8x4 --sweep-device --sweep-cycles 2>&1 | grep -A1 sweep
python bench.py --steps 60 --warmup 5 --sweep 8x4 --sweep-cycles 2>&1 | grep -A1 sweep
