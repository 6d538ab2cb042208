#!/bin/bash
# A/B of library variants in slam-sdvl_b200/_variants: alternating runs on one box, value / e2e / kernel times per run
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
cp slam-sdvl_b200/libsdvl_b200.so /tmp/orig.so
for v in "$@"; do
  cp slam-sdvl_b200/_variants/$v.so slam-sdvl_b200/libsdvl_b200.so
  timeout 300 python bench.py --steps 40 --warmup 5 --no-extras $VARARGS 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), {k:round(x,1) for k,x in d['roofline']['kernel_us_per_step'].items()})"
done
cp /tmp/orig.so slam-sdvl_b200/libsdvl_b200.so
