#!/bin/bash
# how many concurrent sequences fill one GPU: value / e2e for more sequences per GPU than the 64 of the bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for cfg in "64 6 6" "128 6 6" "128 12 6" "256 12 6" "256 16 8"; do
  set -- $cfg
  timeout 400 python bench.py --steps 40 --warmup 5 --no-extras --repeats 3 --seqs $1 --groups $2 --threads $3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('seqs $1 groups $2 threads $3', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), {k:round(x,1) for k,x in d['roofline']['kernel_us_per_step'].items()})"
done 2>&1 | tee gpurun_out/r02_i_seqs_sweep.txt
