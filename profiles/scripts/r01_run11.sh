# state check after container re-creation: GPU tests, smoke, default bench, reference arm
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/g_tests.log
python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/g_smoke.log
python bench.py > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err || tail -5 gpurun_out/g_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/g_ref.json 2> gpurun_out/g_ref.err || tail -5 gpurun_out/g_ref.err
nproc; cat gpurun_out/g_bench.json | head -c 3000
