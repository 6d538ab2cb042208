python profiles/scripts/undistort_timing.py 2>&1 | tail -4 | tee gpurun_out/undistort_timing.txt
python bench.py > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err || tail -3 gpurun_out/final2_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/final2_bench.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']), 'cpu', round(d['cpu_baseline']['value']), d['cpu_baseline']['sample'])
PY
