#!/bin/bash
# host configuration sweep (groups x threads), frames resident in HBM and in pinned host memory
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nproc > gpurun_out/r02_sweep1.txt
timeout 600 python bench.py --steps 40 --warmup 5 --sweep 6x3,6x6,8x4,8x8,12x6,12x12,16x8,4x4 --sweep-device >> gpurun_out/r02_sweep1.txt 2>&1
timeout 600 python bench.py --steps 40 --warmup 5 --sweep 6x3,6x6,8x8,12x6 >> gpurun_out/r02_sweep1.txt 2>&1
cat gpurun_out/r02_sweep1.txt
