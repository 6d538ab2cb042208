# scaling check on one box: N = 1 and N = $1 (default 8), default bench settings
N=${1:-8}
nproc; nvidia-smi -L | wc -l
python bench.py --gpus 1 --steps 40 --warmup 5 > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err || tail -5 gpurun_out/scale_1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err || tail -5 gpurun_out/scale_$N.err
python - <<PY
import json
for n in (1, $N):
    d=json.loads([l for l in open(f'gpurun_out/scale_{n}.json') if l.startswith('{')][-1])
    print(n, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'threads', d['config']['host_threads_per_gpu'], 'clk', d['clocks'])
PY
