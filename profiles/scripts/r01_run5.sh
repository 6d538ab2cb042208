python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --groups 16 --threads 16 > gpurun_out/b.json 2> gpurun_out/b.err || tail -5 gpurun_out/b.err
python - <<PY
import json
d=json.load(open('gpurun_out/b.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']), d['host_phase_thread_seconds'], d['gpu_launches'], d['roofline']['kernel_us_per_step'])
PY
ncu --set full --clock-control none --import-source on -k regex:"fast_cells" -s 6 -c 1 -o gpurun_out/prof_fast python bench.py --steps 2 --warmup 3 --groups 1 > gpurun_out/ncu_fast.log 2>&1
