#!/bin/bash
# round-2: ncu --set full of the ORB-mode kernels (bench.py --orb, one context)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-f}
B="python bench.py --steps 2 --warmup 3 --groups 1 --threads 1 --repeats 1 --no-extras --orb"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02${TAG}_orb.csv $B > gpurun_out/ncu_launch_r02${TAG}_orb.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"orb_frames|seq_orb_points|search_seq" -s 6 -c 6 -o gpurun_out/prof_r02${TAG}_orb $B > gpurun_out/ncu_full_r02${TAG}_orb.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
