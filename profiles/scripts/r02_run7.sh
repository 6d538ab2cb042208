#!/bin/bash
# round 2, GPU run 7: bench lines for the other configurations (C1, C3, C5) and compute-sanitizer over the GPU suite
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for c in C1 C3 C5; do
  timeout 400 python bench.py --config $c --steps 40 --warmup 5 --no-extras > gpurun_out/r02_7_bench_$c.json 2> gpurun_out/r02_7_bench_$c.err
  echo "bench $c rc=$?"; tail -2 gpurun_out/r02_7_bench_$c.err; cut -c1-400 gpurun_out/r02_7_bench_$c.json
done
# memcheck over the parity suite (one process; the slow trajectory tests are deselected by name where they exceed minutes)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r02_7_memcheck.log \
  python -m pytest tests -m gpu -x -q > gpurun_out/r02_7_memcheck_pytest.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/r02_7_memcheck_pytest.log
tail -5 gpurun_out/r02_7_memcheck_pytest.log; tail -5 gpurun_out/r02_7_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/r02_7_racecheck.log \
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_resident.py -m gpu -x -q > gpurun_out/r02_7_racecheck_pytest.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/r02_7_racecheck_pytest.log
tail -5 gpurun_out/r02_7_racecheck_pytest.log; tail -5 gpurun_out/r02_7_racecheck.log
