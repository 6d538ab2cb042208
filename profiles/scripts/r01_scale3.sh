N=${1:-8}
python bench.py --gpus 1 --steps 60 --warmup 5 > gpurun_out/scale3_1.json 2> gpurun_out/scale3_1.err || tail -5 gpurun_out/scale3_1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 60 --warmup 5 > gpurun_out/scale3_$N.json 2> gpurun_out/scale3_$N.err || tail -5 gpurun_out/scale3_$N.err
python - <<PY
import json
for n in (1, $N):
    d=json.loads([l for l in open(f'gpurun_out/scale3_{n}.json') if l.startswith('{')][-1])
    print(n, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'threads', d['config']['host_threads_per_gpu'], d['e2e']['pcie'])
PY
