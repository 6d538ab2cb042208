# round-1 final check on one B200: GPU tests, smoke, both bench arms, ncu evidence
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/final_tests.log
python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/final_smoke.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err || tail -3 gpurun_out/final_ref.err
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err || tail -5 gpurun_out/final_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/final_bench.json')); r=json.load(open('gpurun_out/final_ref.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ref',round(r['value']), 'e2e/ref', round(d['e2e']['value']/r['value'],1))
print({k:round(v,1) for k,v in d['roofline']['kernel_us_per_step'].items()}, 'us/gn', round(d['us_per_gn_iter'],2), 'launches', d['gpu_launches'])
print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['pcie'])
PY
bash profiles/scripts/r01_final_profile.sh i > /dev/null 2>&1
ls gpurun_out | grep _i
