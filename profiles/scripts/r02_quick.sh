#!/bin/bash
# quick look: bench only, optional tag and extra bench flags
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-q}; shift
( time timeout 600 python bench.py --steps 40 --warmup 5 "$@" > gpurun_out/r02_${TAG}_bench.json 2> gpurun_out/r02_${TAG}_bench.err ) 2>&1 | grep real
echo "bench rc=$?"; tail -3 gpurun_out/r02_${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_${TAG}_bench.json"))
print("value",round(d["value"]),d["value_runs"],"e2e",round(d["e2e"]["value"]),d["e2e"]["runs"],"pcie frac",round(d["e2e"]["pcie"]["frac"],3))
print("pageable",d["e2e_pageable"] and round(d["e2e_pageable"]["value"]),"latency",d["latency_single_sequence"])
print({k:round(v,1) for k,v in d["roofline"]["kernel_us_per_step"].items()})
print("frac",{k:round(v["frac"],4) for k,v in d["roofline"]["per_kernel"].items()}, "whole", d["roofline"]["whole_step"])
print("avg ",{k:round(v,1) for k,v in d["align_and_feature_align_kernel_us_per_frame"].items()})
print("slow",{k:round(v,1) for k,v in d["slowest_sequence_of_a_group_step_us"].items()})
print(d["host_phase_thread_seconds"]["value"], "kf/frame", d["keyframes_per_frame"], "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"])
PY
