# timelines of the pipelined run (HBM-resident frames and pinned host frames) + e2e sweeps
python profiles/scripts/timeline.py 8 4 64 1 > gpurun_out/tl_8_1.txt 2>&1
python profiles/scripts/timeline_summary.py gpurun_out/timeline_8_1.json > gpurun_out/tls_8_1.txt 2>&1
python profiles/scripts/timeline.py 8 4 64 2 > gpurun_out/tl_8_2.txt 2>&1
python profiles/scripts/timeline_summary.py gpurun_out/timeline_8_2.json > gpurun_out/tls_8_2.txt 2>&1
rm -f gpurun_out/timeline_*.json.bak
for pf in 1 2 4; do
python bench.py --steps 60 --warmup 5 --prefetch $pf --sweep 8x4,16x4,16x8,32x8 2>&1 | grep sweep | sed "s/^/e2e pf=$pf /"
done | tee gpurun_out/sweep_e2e.txt
python bench.py --steps 60 --warmup 5 --sweep 8x4,16x4,16x8,32x8 --sweep-device 2>&1 | grep sweep | sed "s/^/dev /" | tee gpurun_out/sweep_dev.txt
python bench.py --steps 60 --warmup 5 --e2e-upload dma --sweep 8x4,16x4 2>&1 | grep sweep | sed "s/^/dma /" | tee -a gpurun_out/sweep_e2e.txt
cat gpurun_out/tls_8_1.txt gpurun_out/tls_8_2.txt
