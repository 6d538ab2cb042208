#!/bin/bash
# round-2 final evidence: GPU suite, bench lines of every configuration (+ ORB mode), ncu launch list + ncu --set full of
# every kernel of the path (one context: no overlap), the same for the ORB-mode kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-f}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_${TAG}_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_${TAG}_tests.log
timeout 500 python bench.py --steps 40 --warmup 5 > gpurun_out/r02_${TAG}_bench_C2.json 2> gpurun_out/r02_${TAG}_bench_C2.err; echo "bench C2 rc=$?"
for c in C1 C3 C5; do
  timeout 400 python bench.py --config $c --steps 40 --warmup 5 --no-extras > gpurun_out/r02_${TAG}_bench_$c.json 2> gpurun_out/r02_${TAG}_bench_$c.err; echo "bench $c rc=$?"
done
timeout 400 python bench.py --orb --steps 40 --warmup 5 --no-extras > gpurun_out/r02_${TAG}_bench_C2_orb.json 2> gpurun_out/r02_${TAG}_bench_C2_orb.err; echo "bench C2 orb rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_${TAG}_bench_reference.json 2> gpurun_out/r02_${TAG}_bench_reference.err; echo "reference rc=$?"
B="python bench.py --steps 2 --warmup 3 --groups 1 --threads 1 --repeats 1 --no-extras"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02$TAG.csv $B > gpurun_out/ncu_launch_r02$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"upload|pyr_|fast_cells|fast_select|seq_align|search_seq|seq_post" -s 30 -c 20 -o gpurun_out/prof_r02$TAG $B > gpurun_out/ncu_full_r02$TAG.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
for f in gpurun_out/r02_${TAG}_bench_*.json; do python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1])
    print("$f".split("bench_")[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]) if d.get("e2e") else None, {k:round(v,1) for k,v in d.get("roofline",{}).get("kernel_us_per_step",{}).items()})
except Exception as e:
    print("$f", "parse failed", e)
PY
done
