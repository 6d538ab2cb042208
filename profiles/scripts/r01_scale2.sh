N=${1:-8}
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
lscpu | grep -E "NUMA|Model name|Socket|Core|Thread|^CPU\(s\)" >> gpurun_out/topo.txt
for f in /sys/bus/pci/devices/*/numa_node; do d=$(dirname $f); if grep -q 0x10de $d/vendor 2>/dev/null && grep -q "^0x0302" $d/class 2>/dev/null; then echo "$d $(cat $f)"; fi; done >> gpurun_out/topo.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/scale2_$N.json 2> gpurun_out/scale2_$N.err || tail -5 gpurun_out/scale2_$N.err
SDVLB_NUMA_BIND=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/scale2_${N}_nobind.json 2> gpurun_out/scale2_${N}_nobind.err || tail -5 gpurun_out/scale2_${N}_nobind.err
python - <<PY
import json
for n in ("$N", "${N}_nobind"):
    d=json.loads([l for l in open(f'gpurun_out/scale2_{n}.json') if l.startswith('{')][-1])
    print(n, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'threads', d['config']['host_threads_per_gpu'], d['config']['numa'])
PY
cat gpurun_out/topo.txt
