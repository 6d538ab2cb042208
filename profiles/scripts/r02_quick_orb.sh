#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz_orb.py -m gpu -q -s > gpurun_out/r02_orb_tests.log 2>&1
echo "pytest rc=$?"; grep -a "^ORB\|ORB C\|passed\|failed\|Error\|error" gpurun_out/r02_orb_tests.log | head -40
