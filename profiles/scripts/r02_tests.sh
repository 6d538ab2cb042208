#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-t}; shift
timeout 1200 python -m pytest tests -m gpu -q "$@" > gpurun_out/r02_${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_${TAG}_tests.log
tail -40 gpurun_out/r02_${TAG}_tests.log
