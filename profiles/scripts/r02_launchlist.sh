#!/bin/bash
# ncu launch list (gpu__time_duration per launch) of a short single-context bench run
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-x}
B="python bench.py --steps 2 --warmup 3 --groups 1 --threads 1 --repeats 1 --no-extras"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02$TAG.csv $B > gpurun_out/ncu_launch_r02$TAG.log 2>&1
python profiles/scripts/summarize_ncu.py - gpurun_out/launches_r02$TAG.csv gpurun_out/r02_$TAG
cat gpurun_out/r02_${TAG}_launches.txt
