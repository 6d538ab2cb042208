#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
: > gpurun_out/r02_sweep2.txt
for d in 1 2; do
  echo "depth $d" >> gpurun_out/r02_sweep2.txt
  timeout 600 python bench.py --steps 40 --warmup 5 --depth $d --sweep 6x3,6x6,8x8,12x6,12x12,16x8 --sweep-device >> gpurun_out/r02_sweep2.txt 2>&1
done
cat gpurun_out/r02_sweep2.txt
