#!/bin/bash
# round-2 evidence: ncu launch list + ncu --set full of every kernel of the path (one context: no overlap)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-a}
B="python bench.py --steps 2 --warmup 3 --groups 1 --threads 1 --repeats 1 --no-extras"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02$TAG.csv $B > gpurun_out/ncu_launch_r02$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"upload|pyr_|fast_cells|fast_select|seq_align|search_seq|seq_post" -s 30 -c 27 -o gpurun_out/prof_r02$TAG $B > gpurun_out/ncu_full_r02$TAG.log 2>&1
ls -la gpurun_out | tail -4
tail -3 gpurun_out/ncu_full_r02$TAG.log
