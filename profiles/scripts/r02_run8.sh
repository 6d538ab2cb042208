#!/bin/bash
# round 2, GPU run 8: the whole GPU suite with the ORB mode wired end to end, then the C2 bench (no extras)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_8_tests.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r02_8_tests.log
grep -a "^ORB\|orb" gpurun_out/r02_8_tests.log | head -20
timeout 400 python bench.py --steps 40 --warmup 5 --no-extras > gpurun_out/r02_8_bench.json 2> gpurun_out/r02_8_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_8_bench.err; cut -c1-300 gpurun_out/r02_8_bench.json
