#!/bin/bash
# round 2: BASELINE.json configs[3] as written (64 sequences in total, strong scaling) and weak scaling on N GPUs of one box
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/r02_scale_${N}_gpus.txt
nproc >> gpurun_out/r02_scale_${N}_gpus.txt
run() {  # tag, extra args
  local tag=$1; shift
  if [ "$N" = "1" ]; then
    timeout 400 python bench.py --gpus 1 --steps 40 --warmup 5 --no-extras "$@" > gpurun_out/r02_scale_${N}_${tag}.json 2> gpurun_out/r02_scale_${N}_${tag}.err
  else
    timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 40 --warmup 5 --no-extras "$@" > gpurun_out/r02_scale_${N}_${tag}.json 2> gpurun_out/r02_scale_${N}_${tag}.err
  fi
  echo "$tag rc=$?"; tail -2 gpurun_out/r02_scale_${N}_${tag}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_scale_${N}_${tag}.json").read().strip().splitlines()[-1])
    print("N=$N $tag", d["scaling"], "seqs/gpu", d["config"]["sequences_per_gpu"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],3), "pcie peak/gpu", round(d["e2e"]["pcie"]["h2d_gbs_per_gpu_peak"],1))
except Exception as e:
    print("parse failed", e)
PY
}
run strong --scaling strong
run weak
if [ "$N" = "1" ]; then run strong8 --scaling strong --seqs 8; run strong16 --scaling strong --seqs 16; run strong32 --scaling strong --seqs 32; fi
