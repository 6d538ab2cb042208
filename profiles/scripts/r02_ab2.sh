#!/bin/bash
# A/B of library variants (slam-sdvl_b200/_variants/*.so) on one box: FAST / frame tests with the in-tree library first,
# then alternating bench runs (value / e2e / kernel times per run); VARARGS = extra bench flags, ROUNDS = repetitions
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
if [ -n "$TESTS" ]; then timeout 600 python -m pytest tests -m gpu -q -x -k "$TESTS" 2>&1 | tail -3; fi
cp slam-sdvl_b200/libsdvl_b200.so /tmp/orig.so
cp slam-sdvl_b200/libsdvl_b200_host.so /tmp/orig_host.so
for r in $(seq 1 ${ROUNDS:-2}); do
for v in "$@"; do
  so=${v%%@*}; envs=""; if [ "$so" != "$v" ]; then envs="${v#*@}"; fi   # variant@VAR=value: library + one environment setting
  cp slam-sdvl_b200/_variants/$so.so slam-sdvl_b200/libsdvl_b200.so
  # <name>.host.so beside it: a variant of the host library (tracker / class mirror) as well
  if [ -f slam-sdvl_b200/_variants/$so.host.so ]; then cp slam-sdvl_b200/_variants/$so.host.so slam-sdvl_b200/libsdvl_b200_host.so; else cp /tmp/orig_host.so slam-sdvl_b200/libsdvl_b200_host.so; fi
  env $envs timeout 300 python bench.py --steps 40 --warmup 5 --no-extras $VARARGS 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), {k:round(x,1) for k,x in d['roofline']['kernel_us_per_step'].items()})"
done
done
cp /tmp/orig.so slam-sdvl_b200/libsdvl_b200.so
cp /tmp/orig_host.so slam-sdvl_b200/libsdvl_b200_host.so
