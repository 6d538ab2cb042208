#!/bin/bash
# bench (kernel pass only matters) for library variants in slam-sdvl_b200/_variants/*.so
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
cp slam-sdvl_b200/libsdvl_b200.so /tmp/orig.so
for v in "$@"; do
  cp slam-sdvl_b200/_variants/$v.so slam-sdvl_b200/libsdvl_b200.so
  timeout 300 python bench.py --steps 40 --warmup 5 --no-extras $VARARGS > gpurun_out/r02_var_$v.json 2> gpurun_out/r02_var_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02_var_$v.json"))
print("$v", "value",round(d["value"]), "e2e", round(d["e2e"]["value"]), {k:round(x,1) for k,x in d["roofline"]["kernel_us_per_step"].items()})
PY
done
cp /tmp/orig.so slam-sdvl_b200/libsdvl_b200.so
