N=${1:-8}
for cfg in "4 2" "6 3"; do set -- $cfg
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 60 --warmup 5 --groups $1 --threads $2 > gpurun_out/scale4_${N}_$1x$2.json 2> gpurun_out/scale4_${N}_$1x$2.err || tail -5 gpurun_out/scale4_${N}_$1x$2.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/scale4_${N}_$1x$2.json') if l.startswith('{')][-1])
print("$1x$2", 'value', round(d['value']), 'e2e', round(d['e2e']['value']), {k:round(v,4) for k,v in d['host_phase_thread_seconds']['value'].items() if k in ('marshal_s','gpu_submit_wait_s','replay_s','idle_poll_s')})
PY
done
