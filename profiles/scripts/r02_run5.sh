#!/bin/bash
# round 2, GPU run 5: parity suite on the restructured tracking chain, short bench, kernel timeline
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_5_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_5_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_5_tests.log
tail -15 gpurun_out/r02_5_tests.log
timeout 400 python bench.py --steps 40 --warmup 5 > gpurun_out/r02_5_bench.json 2> gpurun_out/r02_5_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_5_bench.err; cut -c1-1500 gpurun_out/r02_5_bench.json
timeout 200 python profiles/scripts/timeline.py 6 3 64 1 > gpurun_out/r02_5_timeline.txt 2>&1
timeout 60 python profiles/scripts/timeline_summary.py gpurun_out/timeline_6_1.json >> gpurun_out/r02_5_timeline.txt 2>&1
tail -25 gpurun_out/r02_5_timeline.txt
