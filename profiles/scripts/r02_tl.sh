#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
G=${1:-6}; T=${2:-3}; N=${3:-30}
timeout 300 python profiles/scripts/timeline.py $G $T 64 1 $N > gpurun_out/r02_tl_$G.txt 2>&1
timeout 60 python profiles/scripts/timeline_summary.py gpurun_out/timeline_${G}_1.json | grep -v "^  stream" >> gpurun_out/r02_tl_$G.txt 2>&1
grep -v "^ \+[0-9.]* +" gpurun_out/r02_tl_$G.txt | tail -20
