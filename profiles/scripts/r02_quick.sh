#!/bin/bash
# quick look: bench only (short), optional tag
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-q}
timeout 400 python bench.py --steps 40 --warmup 5 > gpurun_out/r02_${TAG}_bench.json 2> gpurun_out/r02_${TAG}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_${TAG}_bench.json"))
print("value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"pcie frac",round(d["e2e"]["pcie"]["frac"],3))
print({k:round(v,1) for k,v in d["roofline"]["kernel_us_per_step"].items()})
print("avg ",{k:round(v,1) for k,v in d["align_and_feature_align_kernel_us_per_frame"].items()})
print("slow",{k:round(v,1) for k,v in d["slowest_sequence_of_a_group_step_us"].items()})
print(d["host_phase_thread_seconds"]["value"])
PY
